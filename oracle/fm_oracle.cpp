// fm_oracle.cpp — C API of the TEST ORACLE (fp64 CPU restatement).  See fmo_physics.hpp / fmo_env.hpp headers.
// Built by oracle/Makefile into oracle/libfm_oracle.so and loaded through ctypes by tests/ and bench.py's
// cpu_baseline leg only.  The product (factory_marl_b200/) never links or loads this.
#include "fmo_physics.hpp"
#include "fmo_env.hpp"

using namespace fmo;

struct Handle { Model m; Data d; Env* env = nullptr; std::string err; };
static thread_local std::string g_err;

#define TRY try {
#define CATCH } catch (const std::exception& e) { g_err = e.what(); return -1; } return 0;

extern "C" {
const char* fmo_last_error() { return g_err.c_str(); }

void* fmo_create(const void* blob, size_t n) {
  try {
    Handle* h = new Handle();
    load_model(h->m, blob, n);
    init_data(h->m, h->d);
    return h;
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void fmo_destroy(void* hv) { Handle* h = (Handle*)hv; delete h->env; delete h; }
void fmo_sizes(void* hv, int* out) {
  Handle* h = (Handle*)hv; const Model& m = h->m;
  int v[] = {m.nb, m.ng, m.nj, m.nq, m.nv, m.nu, m.n_arms, m.n_cubes, (int)m.clusters.size(), (int)m.cpairs.size(), (int)m.n_geom_pairs};
  for (int i = 0; i < 11; ++i) out[i] = v[i];
}
int fmo_set_state(void* hv, const double* qpos, const double* qvel) {
  Handle* h = (Handle*)hv;
  TRY
  std::copy(qpos, qpos + h->m.nq, h->d.qpos.begin()); std::copy(qvel, qvel + h->m.nv, h->d.qvel.begin());
  CATCH
}
void fmo_set_warmstart(void* hv, const double* qacc) { Handle* h = (Handle*)hv; std::copy(qacc, qacc + h->m.nv, h->d.qacc_warmstart.begin()); }
void fmo_get_state(void* hv, double* qpos, double* qvel) {
  Handle* h = (Handle*)hv;
  std::copy(h->d.qpos.begin(), h->d.qpos.end(), qpos); std::copy(h->d.qvel.begin(), h->d.qvel.end(), qvel);
}
int fmo_step1(void* hv) { Handle* h = (Handle*)hv; TRY step1(h->m, h->d); CATCH }
int fmo_step2(void* hv, const double* ctrl) {
  Handle* h = (Handle*)hv;
  TRY std::copy(ctrl, ctrl + h->m.nu, h->d.ctrl.begin()); step2(h->m, h->d); CATCH
}
// intermediates
void fmo_get_xpos(void* hv, double* xpos, double* xquat) {
  Handle* h = (Handle*)hv;
  for (int b = 0; b < h->m.nb; ++b) {
    xpos[3 * b] = h->d.xpos[b].x; xpos[3 * b + 1] = h->d.xpos[b].y; xpos[3 * b + 2] = h->d.xpos[b].z;
    xquat[4 * b] = h->d.xquat[b].w; xquat[4 * b + 1] = h->d.xquat[b].x; xquat[4 * b + 2] = h->d.xquat[b].y; xquat[4 * b + 3] = h->d.xquat[b].z;
  }
}
void fmo_get_geom(void* hv, double* gpos, double* gmat) {
  Handle* h = (Handle*)hv;
  for (int g = 0; g < h->m.ng; ++g) {
    gpos[3 * g] = h->d.gpos[g].x; gpos[3 * g + 1] = h->d.gpos[g].y; gpos[3 * g + 2] = h->d.gpos[g].z;
    for (int k = 0; k < 9; ++k) gmat[9 * g + k] = h->d.gmat[g].m[k];
  }
}
void fmo_get_dyn(void* hv, double* M, double* bias, double* qacc_smooth, double* qacc, double* qfrc_constraint, double* qfrc_actuator) {
  Handle* h = (Handle*)hv; int nv = h->m.nv;
  if (M) std::copy(h->d.M.begin(), h->d.M.end(), M);
  if (bias) std::copy(h->d.qfrc_bias.begin(), h->d.qfrc_bias.end(), bias);
  if (qacc_smooth) std::copy(h->d.qacc_smooth.begin(), h->d.qacc_smooth.end(), qacc_smooth);
  if (qacc) std::copy(h->d.qacc.begin(), h->d.qacc.end(), qacc);
  if (qfrc_constraint) std::copy(h->d.qfrc_constraint.begin(), h->d.qfrc_constraint.end(), qfrc_constraint);
  if (qfrc_actuator) std::copy(h->d.qfrc_actuator.begin(), h->d.qfrc_actuator.end(), qfrc_actuator);
  (void)nv;
}
int fmo_ncon(void* hv) { return (int)((Handle*)hv)->d.con.size(); }
int fmo_nefc(void* hv) { return ((Handle*)hv)->d.nefc; }
// per contact: geoms[2], data[3 pos, 9 frame, dist, mu, 4 forces, solref0, imp] = 20 doubles
void fmo_get_contacts(void* hv, int* geoms, double* data) {
  Handle* h = (Handle*)hv;
  for (size_t i = 0; i < h->d.con.size(); ++i) {
    const Contact& c = h->d.con[i];
    geoms[2 * i] = c.g1; geoms[2 * i + 1] = c.g2;
    double* o = data + 20 * i;
    o[0] = c.pos.x; o[1] = c.pos.y; o[2] = c.pos.z;
    o[3] = c.n.x; o[4] = c.n.y; o[5] = c.n.z; o[6] = c.t1.x; o[7] = c.t1.y; o[8] = c.t1.z; o[9] = c.t2.x; o[10] = c.t2.y; o[11] = c.t2.z;
    o[12] = c.dist; o[13] = c.mu;
    for (int r = 0; r < 4; ++r) o[14 + r] = c.efc >= 0 && (size_t)(c.efc + r) < h->d.efc_force.size() ? h->d.efc_force[c.efc + r] : 0;
    o[18] = c.solref[0]; o[19] = c.efc >= 0 ? h->d.efc_KBI[3 * c.efc + 2] : 0;
  }
}
void fmo_get_efc(void* hv, int* type, double* J, double* pos, double* aref, double* R, double* force) {
  Handle* h = (Handle*)hv; int ne = h->d.nefc;
  for (int i = 0; i < ne; ++i) { type[i] = h->d.efc_type[i]; pos[i] = h->d.efc_pos[i]; aref[i] = h->d.efc_aref[i]; R[i] = h->d.efc_R[i]; force[i] = h->d.efc_force[i]; }
  if (J) std::copy(h->d.efc_J.begin(), h->d.efc_J.end(), J);
}
void fmo_solver_stats(void* hv, double* out) { Handle* h = (Handle*)hv; out[0] = h->d.solver_iter; out[1] = h->d.solver_grad; out[2] = (double)h->d.flops; }

// ---------------------------------------------------------------- env level (fmo_env.hpp)
int fmo_env_create(void* hv, const int* icfg, const double* dcfg) {
  Handle* h = (Handle*)hv;
  TRY delete h->env; h->env = new Env(h->m, h->d, icfg, dcfg); CATCH
}
int fmo_env_reset(void* hv, float* obs) { Handle* h = (Handle*)hv; TRY h->env->reset(obs); CATCH }
int fmo_env_step(void* hv, const float* action_f, const int* action_i, float* obs, double* reward, int* info) {
  Handle* h = (Handle*)hv; TRY h->env->step(action_f, action_i, obs, reward, info); CATCH
}
int fmo_env_obs_dim(void* hv) { return ((Handle*)hv)->env->obs_dim(); }
void fmo_env_get(void* hv, const char* what, double* out) { ((Handle*)hv)->env->get(what, out); }
void fmo_env_set(void* hv, const char* what, const double* in) { ((Handle*)hv)->env->set(what, in); }
int fmo_ik_solve(void* hv, int arm, const double* target_pos, const double* target_quat, double* q_out, int* steps) {
  Handle* h = (Handle*)hv; return h->env->ik_solve(arm, target_pos, target_quat, q_out, steps) ? 1 : 0;
}
uint64_t fmo_pcg64_next(uint64_t* state4) { return pcg64_next(state4); }
}
