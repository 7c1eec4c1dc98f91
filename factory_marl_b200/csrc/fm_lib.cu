// fm_lib.cu — CUDA kernels (sm_100a) + the C-ABI of include/fm_api.h.
// Kernels:
//   fm_step_kernel<NA,E,TW> fused env step: 100 x mj_step-equivalent + task logic for E envs per CTA (fm_kernel.cuh,
//                          instantiated per block shape in fm_step_shape_*.cu)
//   fm_ik_kernel<NA>       IK base policy + FSM, fp64, thread per env (fm_ik.cuh)
// No CPU fallback: every entry point launches CUDA work or fails with an error.
#include <cuda_runtime.h>
#include <cstdio>
#include <new>
#include "fm_host_common.h"
#include "fm_ik.cuh"
#include "fm_driver.cuh"
#include "fm_kernel.cuh"
static thread_local std::string g_err;
#define FM_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_err = std::string(#x) + ": " + cudaGetErrorString(e_); return 2; } } while (0)

template <int NA>
__global__ void fm_ik_kernel(const FmIkModel* __restrict__ ik, FmEnvCfg cfg, FmStatePtrs st, float* obs, int obs_dim) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= st.n) return;
  fm_ik_env<NA>(*ik, cfg, st, obs, obs_dim, env);
}

struct fm_handle {
  FmHostModel hm; FmEnvCfg cfg; fm_config user; FmStatePtrs st;
  int E = 4; size_t smem = 0;
  FmDevModel* d_model = nullptr; FmIkModel* d_ik = nullptr; unsigned int* d_cpairs = nullptr; int* d_geom_class = nullptr;
  FmDebugDump* d_dump = nullptr; float* d_ws = nullptr;
  long long launches = 0;
  const int* order = nullptr;      // fm_set_env_order
};

// Block shapes built into the library (one translation unit each, fm_step_shape_*.cu): NA arms, E envs per block, TW warps.
//   E = 14 / TW = 12 (two arms) and E = 7 / TW = 7 (four arms): one block per SM, all warps on the same stretch of code (the
//   hot loop is ~120 KB of SASS against a 32 KB L1.5 instruction cache; independent single-warp blocks thrash it), 4096
//   envs = 2 (4) full waves of 148 blocks.  E = 4 and E = 1: small batches, tests.
static int dispatch(fm_handle* h, const FmStepArgs& ar, cudaStream_t s) {
  if (!h->st.f) { g_err = "fm_bind_state was not called"; return 1; }
  FmLaunch L{h->d_model, h->d_cpairs, h->d_geom_class, h->cfg, h->st};
  cudaError_t e;
  // the kernel variant with the warp-cooperative exact solve of arm + contact islands; fm_config.pgs_outer = 1: Gauss-Seidel sweeps only
  // (the variant without that phase: 10 % faster per sub-step, but an arm collision then costs 32 one-lane sweeps per sub-step
  // that do not converge, and its block sets the launch time)
  const bool bign = h->user.pgs_outer != 1;
  if (h->hm.n_arms == 2) e = h->E == FM_BIG2_E ? (bign ? fm_launch_step<2, FM_BIG2_E, FM_BIG2_TW, true>(L, ar, s) : fm_launch_step<2, FM_BIG2_E, FM_BIG2_TW>(L, ar, s)) : (h->E == 1 ? fm_launch_step<2, 1, 1>(L, ar, s) : fm_launch_step<2, 4, 4>(L, ar, s));
  else if (h->hm.n_arms == 6) e = h->E == 1 ? fm_launch_step<6, 1, 1>(L, ar, s) : (bign ? fm_launch_step<6, FM_BIG6_E, FM_BIG6_TW, true>(L, ar, s) : fm_launch_step<6, FM_BIG6_E, FM_BIG6_TW>(L, ar, s));
  else e = h->E == FM_BIG4_E ? (bign ? fm_launch_step<4, FM_BIG4_E, FM_BIG4_TW, true>(L, ar, s) : fm_launch_step<4, FM_BIG4_E, FM_BIG4_TW>(L, ar, s)) : (h->E == 1 ? fm_launch_step<4, 1, 1>(L, ar, s) : fm_launch_step<4, 4, 4>(L, ar, s));
  if (e != cudaSuccess) { g_err = std::string("fm_step_kernel launch: ") + cudaGetErrorString(e); return 2; }
  h->launches++;
  return 0;
}

extern "C" {
const char* fm_last_error(void) { return g_err.c_str(); }

int fm_destroy(fm_handle* h);
static int fm_create_impl(const fm_config* cfg, const void* blob, size_t n, fm_handle* h) {
    fm_build_host_model(blob, n, h->hm);
    if (cfg->n_arms != h->hm.n_arms) throw std::runtime_error("cfg.n_arms does not match the model blob");
    if (h->hm.n_arms != 2 && h->hm.n_arms != 4 && h->hm.n_arms != 6) throw std::runtime_error("num_arms must be 2, 4 or 6 (scene.py:148-158)");
    if (cfg->env_class < 0 || cfg->env_class > 6) throw std::runtime_error("bad env_class");
    h->user = *cfg; h->cfg = fm_make_cfg(*cfg, h->hm.timestep);
    // envs per block: 1 | 4 | the one-block-per-SM shape (14 with two arms, 7 with four, 6 with six); 0 = by batch size
    const int big = h->hm.n_arms == 2 ? FM_BIG2_E : (h->hm.n_arms == 4 ? FM_BIG4_E : FM_BIG6_E);
    int E = cfg->envs_per_block;
    h->E = (E == 1 || E == 4 || E == big) ? E : (cfg->n_envs >= 2048 ? big : 4);
    if (h->hm.n_arms == 6 && h->E == 4) h->E = big;       // six arms: the one-env shape and the block of 6 only
    h->st = FmStatePtrs{nullptr, nullptr, nullptr, nullptr, cfg->n_envs, nullptr};
    FM_CUDA(cudaMalloc(&h->d_model, sizeof(FmDevModel)));
    FM_CUDA(cudaMemcpy(h->d_model, &h->hm.dm, sizeof(FmDevModel), cudaMemcpyHostToDevice));
    FM_CUDA(cudaMalloc(&h->d_ik, sizeof(FmIkModel)));
    FM_CUDA(cudaMemcpy(h->d_ik, &h->hm.ik, sizeof(FmIkModel), cudaMemcpyHostToDevice));
    FM_CUDA(cudaMalloc(&h->d_cpairs, h->hm.cpairs.size() * sizeof(unsigned int)));
    FM_CUDA(cudaMemcpy(h->d_cpairs, h->hm.cpairs.data(), h->hm.cpairs.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    FM_CUDA(cudaMalloc(&h->d_geom_class, h->hm.geom_class.size() * sizeof(int)));
    FM_CUDA(cudaMemcpy(h->d_geom_class, h->hm.geom_class.data(), h->hm.geom_class.size() * sizeof(int), cudaMemcpyHostToDevice));
    FM_CUDA(cudaMalloc(&h->d_dump, sizeof(FmDebugDump)));
    const size_t ws_rows = (size_t)cfg->n_envs + 32;   // + padding slots of the last block
    const size_t ws_floats = h->hm.n_arms == 2 ? FmCap<2>::WS : (h->hm.n_arms == 4 ? FmCap<4>::WS : FmCap<6>::WS);
    FM_CUDA(cudaMalloc(&h->d_ws, ws_rows * ws_floats * sizeof(float)));
    FM_CUDA(cudaMemset(h->d_ws, 0, ws_rows * ws_floats * sizeof(float)));
    h->st.ws = h->d_ws;
    return 0;
}
int fm_create(const fm_config* cfg, const void* blob, size_t n, fm_handle** out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device: this library has no CPU path"; return 3; }
  fm_handle* h = new fm_handle();
  int rc;
  try { rc = fm_create_impl(cfg, blob, n, h); } catch (const std::exception& e) { g_err = e.what(); rc = 1; }
  if (rc) { fm_destroy(h); return rc; }          // frees whatever was allocated before the failure
  *out = h;
  return 0;
}
int fm_destroy(fm_handle* h) {
  if (!h) return 0;
  cudaFree(h->d_model); cudaFree(h->d_ik); cudaFree(h->d_cpairs); cudaFree(h->d_geom_class); cudaFree(h->d_dump); cudaFree(h->d_ws);
  delete h; return 0;
}
int fm_obs_dim(const fm_handle* h) { return h->hm.n_arms == 2 ? Lay<2>::obs_dim(h->cfg.env_class) : (h->hm.n_arms == 4 ? Lay<4>::obs_dim(h->cfg.env_class) : Lay<6>::obs_dim(h->cfg.env_class)); }
int fm_act_dim(const fm_handle* h) { return fm_act_dim_of(h->cfg.env_class, h->hm.n_arms); }
int fm_state_rows(const fm_handle* h, int32_t* nf, int32_t* nd, int32_t* ni) {
  if (h->hm.n_arms == 2) { *nf = Lay<2>::NF; *nd = Lay<2>::ND; *ni = Lay<2>::NI; } else if (h->hm.n_arms == 4) { *nf = Lay<4>::NF; *nd = Lay<4>::ND; *ni = Lay<4>::NI; } else { *nf = Lay<6>::NF; *nd = Lay<6>::ND; *ni = Lay<6>::NI; }
  return 0;
}
int fm_bind_state(fm_handle* h, float* f, double* d, int32_t* i, uint64_t* rng) {
  h->st.f = f; h->st.d = d; h->st.i = i; h->st.rng = (unsigned long long*)rng; return 0;
}
int fm_field_offset(const fm_handle* h, const char* name) { return h->hm.n_arms == 2 ? fm_field_offset_t<2>(name) : (h->hm.n_arms == 4 ? fm_field_offset_t<4>(name) : fm_field_offset_t<6>(name)); }
long long fm_launch_count(const fm_handle* h) { return h->launches; }

int fm_set_env_order(fm_handle* h, const int32_t* order) { h->order = order; return 0; }
int fm_reset(fm_handle* h, const uint8_t* mask, float* obs, void* stream) {
  FmStepArgs ar; memset(&ar, 0, sizeof(ar)); ar.mode = 1; ar.order = h->order; ar.reset_mask = mask; ar.obs = obs;
  return dispatch(h, ar, (cudaStream_t)stream);
}
int fm_step(fm_handle* h, const void* actions, float* obs, float* reward, uint8_t* terminated, int32_t* scores, uint8_t* flags,
            float* terminal_obs, float* ep_info, void* stream) {
  if (!actions && fm_act_dim(h) > 0) { g_err = "fm_step: actions is NULL"; return 1; }
  FmStepArgs ar; memset(&ar, 0, sizeof(ar));
  ar.mode = 0; ar.act_f = (const float*)actions; ar.act_i = (const int*)actions; ar.obs = obs; ar.reward = reward; ar.terminated = terminated;
  ar.scores = scores; ar.flags = flags; ar.terminal_obs = terminal_obs; ar.ep_info = ep_info; ar.auto_reset = h->user.auto_reset; ar.order = h->order;
  return dispatch(h, ar, (cudaStream_t)stream);
}
int fm_ik_policy(fm_handle* h, float* obs, void* stream) {
  if (!h->st.f) { g_err = "fm_bind_state was not called"; return 1; }
  int N = h->cfg.n_envs, od = fm_obs_dim(h), tpb = 64, nb = (N + tpb - 1) / tpb;
  if (h->hm.n_arms == 2) fm_ik_kernel<2><<<nb, tpb, 0, (cudaStream_t)stream>>>(h->d_ik, h->cfg, h->st, obs, od);
  else if (h->hm.n_arms == 4) fm_ik_kernel<4><<<nb, tpb, 0, (cudaStream_t)stream>>>(h->d_ik, h->cfg, h->st, obs, od);
  else fm_ik_kernel<6><<<nb, tpb, 0, (cudaStream_t)stream>>>(h->d_ik, h->cfg, h->st, obs, od);
  FM_CUDA(cudaGetLastError());
  h->launches++;
  return 0;
}
static int copy_state_to_host(fm_handle* h, std::vector<float>& f, std::vector<int>& iv) {
  int32_t nf, nd, ni; fm_state_rows(h, &nf, &nd, &ni);
  f.resize((size_t)nf * h->st.n); iv.resize((size_t)ni * h->st.n);
  FM_CUDA(cudaDeviceSynchronize());
  FM_CUDA(cudaMemcpy(f.data(), h->st.f, f.size() * sizeof(float), cudaMemcpyDeviceToHost));
  FM_CUDA(cudaMemcpy(iv.data(), h->st.i, iv.size() * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}
static int check_env(const fm_handle* h, int env, const char* who) {
  if (env < 0 || env >= h->st.n) { g_err = std::string(who) + ": env index out of range"; return 1; }
  if (!h->st.f) { g_err = "fm_bind_state was not called"; return 1; }
  return 0;
}
int fm_get_state(fm_handle* h, fm_state_view* v) {
  if (int rc = check_env(h, v->env, "fm_get_state")) return rc;
  std::vector<float> f; std::vector<int> iv;
  if (int rc = copy_state_to_host(h, f, iv)) return rc;
  if (h->hm.n_arms == 2) fm_view_get<2>(f.data(), h->st.n, v); else if (h->hm.n_arms == 4) fm_view_get<4>(f.data(), h->st.n, v); else fm_view_get<6>(f.data(), h->st.n, v);
  return 0;
}
int fm_set_state(fm_handle* h, const fm_state_view* v) {
  if (int rc = check_env(h, v->env, "fm_set_state")) return rc;
  std::vector<float> f; std::vector<int> iv;
  if (int rc = copy_state_to_host(h, f, iv)) return rc;
  if (h->hm.n_arms == 2) { fm_view_set<2>(f.data(), h->st.n, v); fm_view_clear_warm<2>(iv.data(), h->st.n, v->env); }
  else if (h->hm.n_arms == 4) { fm_view_set<4>(f.data(), h->st.n, v); fm_view_clear_warm<4>(iv.data(), h->st.n, v->env); }
  else { fm_view_set<6>(f.data(), h->st.n, v); fm_view_clear_warm<6>(iv.data(), h->st.n, v->env); }
  FM_CUDA(cudaMemcpy(h->st.f, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
  FM_CUDA(cudaMemcpy(h->st.i, iv.data(), iv.size() * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}
int fm_debug_substeps(fm_handle* h, int n_substeps, fm_debug_view* view, void* stream) {
  if (view) if (int rc = check_env(h, view->env, "fm_debug_substeps")) return rc;
  FmDebugDump hd; memset(&hd, 0, sizeof(hd)); hd.env = view ? view->env : -1;
  FM_CUDA(cudaMemcpy(h->d_dump, &hd, sizeof(hd), cudaMemcpyHostToDevice));
  FmStepArgs ar; memset(&ar, 0, sizeof(ar)); ar.mode = 2; ar.n_substeps = n_substeps; ar.dump = view ? h->d_dump : nullptr; ar.order = h->order;
  if (int rc = dispatch(h, ar, (cudaStream_t)stream)) return rc;
  FM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (view) {
    FM_CUDA(cudaMemcpy(&hd, h->d_dump, sizeof(hd), cudaMemcpyDeviceToHost));
    if (h->hm.n_arms == 2) fm_dump_to_view<2>(hd, view); else if (h->hm.n_arms == 4) fm_dump_to_view<4>(hd, view); else fm_dump_to_view<6>(hd, view);
  }
  return 0;
}
}  // extern "C"
