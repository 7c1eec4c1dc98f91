// fm_host_common.h — host-side helpers shared by the CUDA library (fm_lib.cu) and the host-compiled logic-test
// build (tests/kernel_emu): config translation, field-offset table, state-view gather/scatter.  Host only.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include "../../include/fm_api.h"
#include "fm_host_model.h"
#include "fm_driver.cuh"

inline FmEnvCfg fm_make_cfg(const fm_config& c, double h) {
  FmEnvCfg e;
  memset(&e, 0, sizeof(e));
  e.env_class = c.env_class; e.n_envs = c.n_envs;
  e.frame_skip = (int)((1.0 / c.control_frequency) / h);                 // base_env.py:134
  e.dt = h * e.frame_skip;                                               // base_env.py:201
  e.pgs_outer = c.pgs_outer > 0 ? c.pgs_outer : 2;
  e.pgs_inner = c.pgs_inner > 0 ? c.pgs_inner : 4;
  e.sleep_hidden = c.reserved == 0 ? 1 : 0;          // fm_config.reserved: 0 = sleep resting hidden cubes (default), 1 = keep them live
  e.pgs_tol = c.pgs_tol > 0 ? c.pgs_tol : 1e-3f;      // m/s^2: below the fp32 noise floor of the stiff rows
  e.initial_conveyor_speed = c.initial_conveyor_speed; e.conveyor_acceleration = c.conveyor_acceleration;
  e.pt_time = c.pt_time; e.force_threshold = c.force_contact_threshold;
  e.init_spawn_freq = c.spawn_freq * c.n_arms;                            // base_env.py:36
  e.spawn_freq_increase = c.spawn_freq_increase;
  e.w_gripper_cube = c.gripper_to_closest_cube_reward_factor; e.w_cube_bucket = c.closest_cube_to_bucket_reward_factor;
  e.w_action_norm = c.small_action_norm_reward_factor; e.base_reward = c.base_reward;
  { const double hd = e.dt / e.frame_skip, alpha = hd / (hd + c.pt_time); for (int k = 0; k < FM_MAX_FRAME_SKIP; ++k) e.pt1_pow[k] = (float)std::pow(1.0 - alpha, k + 1.0); }
  return e;
}

inline int fm_act_dim_of(int env_class, int n_arms) {
  switch (env_class) {
    case FM_ENV_SINGLE_FULL: case FM_ENV_SINGLE_DELTA: return 8;
    case FM_ENV_ALL_FULL: case FM_ENV_ALL_DELTA: return 8 * n_arms;
    case FM_ENV_PAUSE_TOGGLE: case FM_ENV_BACKUP_TOGGLE: return n_arms;
    default: return 0;
  }
}
inline bool fm_act_is_int(int env_class) { return env_class == FM_ENV_PAUSE_TOGGLE || env_class == FM_ENV_BACKUP_TOGGLE; }
inline bool fm_uses_ik(int env_class) { return env_class != FM_ENV_ALL_FULL; }

template <int NA>
inline int fm_field_offset_t(const char* name) {
  typedef Lay<NA> L;
  std::string s(name);
#define FM_OFF(x) if (s == #x) return L::x;
  FM_OFF(F_BELT_Q) FM_OFF(F_BELT_V) FM_OFF(F_CUBE_Q) FM_OFF(F_CUBE_V) FM_OFF(F_ARM_Q) FM_OFF(F_ARM_V) FM_OFF(F_CTRL_T)
  FM_OFF(F_ACC_WS) FM_OFF(F_PRE_Q) FM_OFF(F_PRE_V) FM_OFF(F_LAST_GD) FM_OFF(F_LAST_BD) FM_OFF(F_IK_ACT) FM_OFF(F_LAST_ARM_ACT)
  FM_OFF(F_IK_LAST_CTRL) FM_OFF(F_IK_MOVE_START) FM_OFF(F_SITE) FM_OFF(F_EP_RET) FM_OFF(F_PREV_F) FM_OFF(F_PREV_AF) FM_OFF(F_CUBE_QLO) FM_OFF(F_ARM_QLO) FM_OFF(NF)
  FM_OFF(D_CONV) FM_OFF(D_SPAWN_FREQ) FM_OFF(D_PLAY_TIME) FM_OFF(ND)
  FM_OFF(I_IN) FM_OFF(I_OUT) FM_OFF(I_NIN) FM_OFF(I_NOUT) FM_OFF(I_STEP) FM_OFF(I_SINCE) FM_OFF(I_FAIL) FM_OFF(I_OUTCNT)
  FM_OFF(I_SCORE) FM_OFF(I_LAST_SCORE) FM_OFF(I_EP_LEN) FM_OFF(I_FSM_STATE) FM_OFF(I_FSM_CNT) FM_OFF(I_FSM_TGT) FM_OFF(I_IGNORE)
  FM_OFF(I_FLAGS) FM_OFF(I_EPISODES) FM_OFF(I_MAXCON) FM_OFF(I_ARMCON) FM_OFF(I_PREV_N) FM_OFF(I_PREV_KEY) FM_OFF(I_PREV_AN) FM_OFF(I_PREV_ACODE) FM_OFF(NI)
#undef FM_OFF
  return -1;
}

// gather / scatter one env between the SoA float array (host copy) and the reference qpos/qvel/ctrl layout
template <int NA>
inline void fm_view_get(const float* f, int N, const fm_state_view* v) {
  typedef Lay<NA> L;
  int e = v->env;
  auto F = [&](int k) { return (double)f[(size_t)k * N + e]; };
  if (v->qpos) {
    v->qpos[0] = F(L::F_BELT_Q);
    for (int k = 0; k < 70; ++k) v->qpos[1 + k] = F(L::F_CUBE_Q + k);
    for (int k = 0; k < 30; ++k) v->qpos[1 + 7 * (k / 3) + k % 3] += F(L::F_CUBE_QLO + k);   // compensated low-order part
    for (int k = 0; k < 9 * NA; ++k) v->qpos[71 + k] = F(L::F_ARM_Q + k) + F(L::F_ARM_QLO + k);
  }
  if (v->qvel) {
    v->qvel[0] = F(L::F_BELT_V);
    for (int k = 0; k < 60; ++k) v->qvel[1 + k] = F(L::F_CUBE_V + k);
    for (int k = 0; k < 9 * NA; ++k) v->qvel[61 + k] = F(L::F_ARM_V + k);
  }
  if (v->ctrl_target) for (int k = 0; k < 1 + 8 * NA; ++k) v->ctrl_target[k] = F(L::F_CTRL_T + k);
}
template <int NA>
inline void fm_view_set(float* f, int N, const fm_state_view* v) {
  typedef Lay<NA> L;
  int e = v->env;
  auto W = [&](int k, double x) { f[(size_t)k * N + e] = (float)x; };
  if (v->qpos) {
    W(L::F_BELT_Q, v->qpos[0]);
    for (int k = 0; k < 70; ++k) { W(L::F_CUBE_Q + k, v->qpos[1 + k]); W(L::F_PRE_Q + k, v->qpos[1 + k]); }
    for (int k = 0; k < 30; ++k) W(L::F_CUBE_QLO + k, (v->qpos[1 + 7 * (k / 3) + k % 3] - (double)(float)v->qpos[1 + 7 * (k / 3) + k % 3]));   // what float drops of the fp64 position
    for (int k = 0; k < 9 * NA; ++k) { W(L::F_ARM_Q + k, v->qpos[71 + k]); W(L::F_ARM_QLO + k, v->qpos[71 + k] - (double)(float)v->qpos[71 + k]); }
  }
  if (v->qvel) {
    W(L::F_BELT_V, v->qvel[0]);
    for (int k = 0; k < 60; ++k) { W(L::F_CUBE_V + k, v->qvel[1 + k]); W(L::F_PRE_V + k, v->qvel[1 + k]); }
    for (int k = 0; k < 9 * NA; ++k) W(L::F_ARM_V + k, v->qvel[61 + k]);
  }
  if (v->ctrl_target) for (int k = 0; k < 1 + 8 * NA; ++k) W(L::F_CTRL_T + k, v->ctrl_target[k]);
}
template <int NA>
inline void fm_view_clear_warm(int* i, int N, int env) { typedef Lay<NA> L; i[(size_t)L::I_PREV_N * N + env] = 0; for (int a = 0; a < NA; ++a) i[(size_t)(L::I_PREV_AN + a) * N + env] = 0; }

template <int NA>
inline void fm_dump_to_view(const FmDebugDump& d, fm_debug_view* v) {
  v->ncon = d.ncon; v->flags = d.flags;
  for (int c = 0; c < d.ncon && c < 64; ++c) {
    v->pair[c][0] = d.c_g[c] >> 12; v->pair[c][1] = d.c_g[c] & 4095;
    for (int k = 0; k < 3; ++k) { v->pos[c][k] = d.c_p[c][k]; v->normal[c][k] = d.c_n[c][k]; }
    for (int k = 0; k < 4; ++k) v->force[c][k] = d.c_f[c][k];
  }
  for (int a = 0; a < NA; ++a) {
    for (int k = 0; k < 45; ++k) v->mass_matrix[a][k] = d.M[a][k];
    for (int k = 0; k < 9; ++k) v->bias[a][k] = d.bias[a][k];
  }
  for (int k = 0; k < 61 + 9 * NA; ++k) { v->qacc_smooth[k] = d.acc_smooth[k]; v->qacc[k] = d.acc[k]; }
}
