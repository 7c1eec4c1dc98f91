// fm_emu.cpp — HOST-COMPILED LOGIC-TEST BUILD of the step kernel's phase functions (tests only, never shipped).
// It exports the same C-ABI as the CUDA library (include/fm_api.h) but every pointer is a host pointer and every
// barrier interval is executed as a loop over thread ids.  Purpose: exercise the exact device source
// (factory_marl_b200/csrc/fm_*.cuh) against the fp64 oracle in `-m "not gpu"` tests, where no GPU exists.
// The product (factory_marl_b200/) never loads this library; its ops fail loudly without the CUDA build.
#include <cstdio>
#include <cstdlib>
#include <cstdio>
#include <new>
#include <vector>
#include "../../factory_marl_b200/csrc/fm_host_common.h"
#include "../../factory_marl_b200/csrc/fm_ik.cuh"

static thread_local std::string g_err;
// thread ids of one barrier interval are executed in ascending, descending or a fixed shuffled order (fm_emu_set_order):
// a result that depends on that order reveals an intra-interval read-after-write, i.e. a missing barrier on the GPU.
static int g_order = 0;
// a warp's stages (island_newton_warp): each stage is executed for the 32 lanes one after the other, in the order under test
struct HostLanes {
  template <class G> void stage(G g) {
    if (g_order == 0) for (int l = 0; l < 32; ++l) g(l);
    else if (g_order == 1) for (int l = 31; l >= 0; --l) g(l);
    else for (int k = 0; k < 32; ++k) g((k * 13 + 5) % 32);
  }
  void sync() {}
  template <class G> void run(G g) { stage(g); }
  template <class T, int N> struct Arr { T a[32][N]; T* operator()(int lane) { return a[lane]; } };
};
struct HostRun {
  int T;
  template <class F> void warps(F f) { HostLanes L; for (int w = 0; w < T / 32; ++w) f(g_order == 1 ? T / 32 - 1 - w : w, L); }
  template <class F> void operator()(F f) {
    if (g_order == 0) for (int t = 0; t < T; ++t) f(t);
    else if (g_order == 1) for (int t = T - 1; t >= 0; --t) f(t);
    else for (int k = 0; k < T; ++k) f((k * 37 + 11) % T);
  }
  // two warp groups that run concurrently on the GPU: here one after the other, in either order (they must be independent)
  template <int WA, int WB, class PA, class PB> void split(PA progA, PB progB) {
    HostRun ra{32 * WA}, rb{32 * (WB ? WB : WA)};
    if (g_order == 1) { progB(rb, rb.T); progA(ra, ra.T); }
    else { progA(ra, ra.T); progB(rb, rb.T); }
  }
};

struct fm_handle {
  FmHostModel hm; FmEnvCfg cfg; fm_config user; FmStatePtrs st; int E;
  std::vector<char> smem; std::vector<float> ws;
  const int* order = nullptr;
};

template <int NA, int E, int TW = E>
static void run_blocks(fm_handle* h, const FmStepArgs& ar) {
  // same rule as the library: the variant with the big-island Newton unless pgs_outer = 1 (here for every block shape)
  const bool bign = h->user.pgs_outer != 1;
  HostRun run{32 * TW};
  size_t need = E * (sizeof(EnvS<NA>) + sizeof(TaskS<NA>)) + 64;
  if (h->smem.size() < need) h->smem.resize(need);
  int nblk = (h->cfg.n_envs + E - 1) / E;
  for (int b = 0; b < nblk; ++b) {
    memset(h->smem.data(), 0, need);
    EnvS<NA>* S = (EnvS<NA>*)h->smem.data(); TaskS<NA>* TS = (TaskS<NA>*)(S + E);
    if (bign) fm_block_step<NA, E, TW, true>(run, S, TS, h->hm.dm, h->hm.cpairs.data(), h->hm.geom_class.data(), h->cfg, h->st, ar, b * E);
    else fm_block_step<NA, E, TW, false>(run, S, TS, h->hm.dm, h->hm.cpairs.data(), h->hm.geom_class.data(), h->cfg, h->st, ar, b * E);
  }
}
static void dispatch(fm_handle* h, const FmStepArgs& ar) {
  if (h->hm.n_arms == 2) { if (h->E == 1) run_blocks<2, 1>(h, ar); else if (h->E == 2) run_blocks<2, 2>(h, ar); else if (h->E == 7) run_blocks<2, 7>(h, ar); else if (h->E == 14) run_blocks<2, 14, 12>(h, ar); else run_blocks<2, 4>(h, ar); }
  else if (h->hm.n_arms == 4) { if (h->E == 1) run_blocks<4, 1>(h, ar); else if (h->E == 7) run_blocks<4, 7, 7>(h, ar); else run_blocks<4, 4>(h, ar); }
  else { if (h->E == 1) run_blocks<6, 1>(h, ar); else run_blocks<6, 6, 8>(h, ar); }
}

// randomised equivalence check of the box-on-face fast path against the general separating-axis routine
static unsigned long long lcg(unsigned long long& x) { x = x * 6364136223846793005ULL + 1442695040888963407ULL; return x >> 11; }
static float urand(unsigned long long& x, float lo, float hi) { return lo + (hi - lo) * (float)((double)lcg(x) / 9007199254740992.0); }
extern "C" int fm_emu_check_box_on_face(int n, unsigned long long seed, int* applied) {
  int bad = 0; *applied = 0;
  for (int it = 0; it < n; ++it) {
    f3 hb = (it & 1) ? mk3(0.3f, 100.f, 0.04f) : ((it & 2) ? mk3(1.2f, 1.0f, 0.5f) : mk3(0.29f, 0.29f, 0.02f));
    f3 pb = mk3(urand(seed, -1, 1), urand(seed, -1, 1), urand(seed, 0.5f, 1.5f));
    float h = urand(seed, 0.03f, 0.05f); f3 hc = mk3(h, h, h);
    q4 q = qnormf(mkq(urand(seed, -1, 1), urand(seed, -1, 1), urand(seed, -1, 1), urand(seed, -1, 1)));
    if (it % 5 == 0) q = qnormf(mkq(1.f, urand(seed, -1e-3f, 1e-3f), urand(seed, -1e-3f, 1e-3f), urand(seed, -1, 1)));   // nearly flat
    if (it % 7 == 0) q = qnormf(mkq(urand(seed, -1, 1), 0, 0, urand(seed, -1, 1)));                                     // exactly flat, yawed
    float Rc[9]; q2matf(q, Rc);
    float top = pb.z + hb.z;
    f3 pc = mk3(pb.x + urand(seed, -hb.x, hb.x) * 0.9f, pb.y + urand(seed, -1.f, 1.f) * fminf(hb.y, 1.5f), top + urand(seed, -0.01f, 0.1f));
    float Rid[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int order = 0; order < 1; ++order) {   // the kernel uses the fast path only with the big box as geom1
      RawC a[8], b[8];
      int nf = col_box_on_face(pb, hb, pc, Rc, hc, order == 0, a);
      if (nf < 0) continue;
      (*applied)++;
      int ng = order == 0 ? col_box_box(pb, Rid, hb, pc, Rc, hc, b) : col_box_box(pc, Rc, hc, pb, Rid, hb, b);
      bool ok = nf == ng;
      for (int k = 0; ok && k < nf; ++k)
        ok = fabsf(a[k].pos.x - b[k].pos.x) < 2e-6f && fabsf(a[k].pos.y - b[k].pos.y) < 2e-6f && fabsf(a[k].pos.z - b[k].pos.z) < 2e-6f &&
             fabsf(a[k].dist - b[k].dist) < 2e-6f && fabsf(a[k].n.z - b[k].n.z) < 1e-6f;
      if (!ok) bad++;
    }
  }
  return bad;
}
extern "C" {
void fm_emu_set_order(int o) { g_order = o; }
void fm_emu_counters(long long* o) { o[0] = fm_dbg_sweeps; o[1] = fm_dbg_visits; o[2] = fm_dbg_islands; }
void fm_emu_sat(long long* o) { for (int i = 0; i < 512; ++i) o[i] = fm_dbg_sat[i]; }
void fm_emu_counters2(long long* o) { for (int i = 0; i < 32; ++i) o[i] = fm_dbg_cnt[i]; }
const char* fm_last_error(void) { return g_err.c_str(); }
int fm_create(const fm_config* cfg, const void* blob, size_t n, fm_handle** out) {
  try {
    fm_handle* h = new fm_handle();
    fm_build_host_model(blob, n, h->hm);
    if (cfg->n_arms != h->hm.n_arms) throw std::runtime_error("cfg.n_arms does not match the model blob");
    if (h->hm.n_arms != 2 && h->hm.n_arms != 4 && h->hm.n_arms != 6) throw std::runtime_error("num_arms must be 2, 4 or 6");
    h->user = *cfg; h->cfg = fm_make_cfg(*cfg, h->hm.timestep);
    { int e = cfg->envs_per_block; h->E = (e == 1 || e == 7 || (cfg->n_arms == 2 && (e == 2 || e == 14))) ? e : 4; if (cfg->n_arms == 6 && h->E != 1) h->E = 6; }   // same shapes as the library (+ E = 2)
    h->ws.assign(((size_t)cfg->n_envs + 32) * (cfg->n_arms == 2 ? FmCap<2>::WS : (cfg->n_arms == 4 ? FmCap<4>::WS : FmCap<6>::WS)), 0.f);
    h->st = FmStatePtrs{nullptr, nullptr, nullptr, nullptr, cfg->n_envs, h->ws.data()};
    *out = h; return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
int fm_destroy(fm_handle* h) { delete h; return 0; }
int fm_obs_dim(const fm_handle* h) { return h->hm.n_arms == 2 ? Lay<2>::obs_dim(h->cfg.env_class) : (h->hm.n_arms == 4 ? Lay<4>::obs_dim(h->cfg.env_class) : Lay<6>::obs_dim(h->cfg.env_class)); }
int fm_act_dim(const fm_handle* h) { return fm_act_dim_of(h->cfg.env_class, h->hm.n_arms); }
int fm_state_rows(const fm_handle* h, int32_t* nf, int32_t* nd, int32_t* ni) {
  if (h->hm.n_arms == 2) { *nf = Lay<2>::NF; *nd = Lay<2>::ND; *ni = Lay<2>::NI; } else if (h->hm.n_arms == 4) { *nf = Lay<4>::NF; *nd = Lay<4>::ND; *ni = Lay<4>::NI; } else { *nf = Lay<6>::NF; *nd = Lay<6>::ND; *ni = Lay<6>::NI; }
  return 0;
}
int fm_bind_state(fm_handle* h, float* f, double* d, int32_t* i, uint64_t* rng) { h->st.f = f; h->st.d = d; h->st.i = i; h->st.rng = (unsigned long long*)rng; return 0; }
int fm_field_offset(const fm_handle* h, const char* name) { return h->hm.n_arms == 2 ? fm_field_offset_t<2>(name) : (h->hm.n_arms == 4 ? fm_field_offset_t<4>(name) : fm_field_offset_t<6>(name)); }
int fm_set_env_order(fm_handle* h, const int32_t* order) { h->order = order; return 0; }
int fm_reset(fm_handle* h, const uint8_t* mask, float* obs, void*) {
  FmStepArgs ar; memset(&ar, 0, sizeof(ar)); ar.mode = 1; ar.reset_mask = mask; ar.obs = obs; ar.order = h->order;
  dispatch(h, ar); return 0;
}
int fm_step(fm_handle* h, const void* actions, float* obs, float* reward, uint8_t* terminated, int32_t* scores, uint8_t* flags,
            float* terminal_obs, float* ep_info, void*) {
  FmStepArgs ar; memset(&ar, 0, sizeof(ar));
  ar.mode = 0; ar.act_f = (const float*)actions; ar.act_i = (const int*)actions; ar.obs = obs; ar.reward = reward; ar.terminated = terminated;
  ar.scores = scores; ar.flags = flags; ar.terminal_obs = terminal_obs; ar.ep_info = ep_info; ar.auto_reset = h->user.auto_reset; ar.order = h->order;
  dispatch(h, ar); return 0;
}
int fm_ik_policy(fm_handle* h, float* obs, void*) {
  int N = h->cfg.n_envs, od = fm_obs_dim(h);
  for (int e = 0; e < N; ++e) {
    if (h->hm.n_arms == 2) fm_ik_env<2>(h->hm.ik, h->cfg, h->st, obs, od, e);
    else if (h->hm.n_arms == 4) fm_ik_env<4>(h->hm.ik, h->cfg, h->st, obs, od, e);
    else fm_ik_env<6>(h->hm.ik, h->cfg, h->st, obs, od, e);
  }
  return 0;
}
int fm_get_state(fm_handle* h, fm_state_view* v) { if (v->env < 0 || v->env >= h->st.n) { g_err = "fm_get_state: env index out of range"; return 1; } if (h->hm.n_arms == 2) fm_view_get<2>(h->st.f, h->st.n, v); else if (h->hm.n_arms == 4) fm_view_get<4>(h->st.f, h->st.n, v); else fm_view_get<6>(h->st.f, h->st.n, v); return 0; }
int fm_set_state(fm_handle* h, const fm_state_view* v) {
  if (h->hm.n_arms == 2) { fm_view_set<2>(h->st.f, h->st.n, v); fm_view_clear_warm<2>(h->st.i, h->st.n, v->env); }
  else if (h->hm.n_arms == 4) { fm_view_set<4>(h->st.f, h->st.n, v); fm_view_clear_warm<4>(h->st.i, h->st.n, v->env); }
  else { fm_view_set<6>(h->st.f, h->st.n, v); fm_view_clear_warm<6>(h->st.i, h->st.n, v->env); }
  return 0;
}
int fm_debug_substeps(fm_handle* h, int n_substeps, fm_debug_view* view, void*) {
  FmDebugDump* d = new FmDebugDump(); memset(d, 0, sizeof(*d)); d->env = view ? view->env : -1;
  FmStepArgs ar; memset(&ar, 0, sizeof(ar)); ar.mode = 2; ar.n_substeps = n_substeps; ar.dump = view ? d : nullptr; ar.order = h->order;
  dispatch(h, ar);
  if (view) { if (h->hm.n_arms == 2) fm_dump_to_view<2>(*d, view); else if (h->hm.n_arms == 4) fm_dump_to_view<4>(*d, view); else fm_dump_to_view<6>(*d, view); }
  delete d; return 0;
}
}
