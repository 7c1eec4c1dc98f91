// fm_task.cuh — HBM state layout, action map, task logic (spawn / bounds / bucket / hide), observation, reward, reset.
// Restates on device (thread per env for the sequential list logic):
//   FactoryManipulationEnv._process_action / _compose_control variants   src/environments.py:84-127,405-645
//   BaseEnv.step_sim / reset_sim / _get_state                            challenge_env/challenge_env/base_env.py:149-282
//   TaskManager                                                          challenge_env/challenge_env/task_utils.py:47-156
//   _get_reward (score delta / progress reward)                          src/environments.py:129-149,284-383
#pragma once
#include "fm_step.cuh"

// ---- HBM state: SoA [field][env].  f = float fields, d = double fields, i = int fields, rng = 4 x u64 (numpy PCG64)
template <int NA>
struct Lay {
  static const int NV = 61 + 9 * NA, NU = 1 + 8 * NA;
  // float fields
  static const int F_BELT_Q = 0, F_BELT_V = 1, F_CUBE_Q = 2, F_CUBE_V = F_CUBE_Q + 70, F_ARM_Q = F_CUBE_V + 60,
                   F_ARM_V = F_ARM_Q + 9 * NA, F_CTRL_T = F_ARM_V + 9 * NA, F_ACC_WS = F_CTRL_T + NU,
                   F_PRE_Q = F_ACC_WS + 1, F_PRE_V = F_PRE_Q + 70, F_LAST_GD = F_PRE_V + 60, F_LAST_BD = F_LAST_GD + NA,
                   F_IK_ACT = F_LAST_BD + NA, F_LAST_ARM_ACT = F_IK_ACT + 8 * NA, F_IK_LAST_CTRL = F_LAST_ARM_ACT + 8 * NA,
                   F_IK_MOVE_START = F_IK_LAST_CTRL + 8 * NA, F_SITE = F_IK_MOVE_START + 3 * NA, F_EP_RET = F_SITE + 3 * NA,
                   F_PREV_F = F_EP_RET + 1, F_PREV_AF = F_PREV_F, F_CUBE_QLO = F_PREV_AF + FM_AROW * NA, F_ARM_QLO = F_CUBE_QLO + 30, NF = F_ARM_QLO + 9 * NA;
  // double fields
  static const int D_CONV = 0, D_SPAWN_FREQ = 1, D_PLAY_TIME = 2, ND = 3;
  // int fields
  static const int I_IN = 0, I_OUT = 10, I_NIN = 20, I_NOUT = 21, I_STEP = 22, I_SINCE = 23, I_FAIL = 24, I_OUTCNT = 25,
                   I_SCORE = 26, I_LAST_SCORE = 28, I_EP_LEN = 30, I_FSM_STATE = 31, I_FSM_CNT = I_FSM_STATE + NA,
                   I_FSM_TGT = I_FSM_CNT + NA, I_IGNORE = I_FSM_TGT + NA, I_FLAGS = I_IGNORE + NA * NA, I_EPISODES = I_FLAGS + 1,
                   I_MAXCON = I_EPISODES + 1, I_ARMCON = I_MAXCON + 1, I_PREV_N = I_ARMCON + 1, I_PREV_KEY = I_PREV_N + 1, I_PREV_AN = I_PREV_KEY + FmCap<NA>::CON,
                   I_PREV_ACODE = I_PREV_AN + NA, NI = I_PREV_ACODE + FM_AROW * NA;
  FM_HDM static int obs_dim(int env_class) { return NA * 24 + 130 + ((env_class == FM_ENV_PAUSE_TOGGLE || env_class == FM_ENV_BACKUP_TOGGLE) ? 8 * NA : 0); }
};

struct FmStatePtrs { float* f; double* d; int* i; unsigned long long* rng; int n; float* ws; /* library-owned scratch [n][FmCap<NA>::WS] */ };

// block slot -> env index: identity, or the caller's env order (fm_set_env_order: heavy envs first, so that they share blocks)
FM_HD int slot_env(const int* ord, int slot, int n) { return (ord != nullptr && slot < n) ? ord[slot] : slot; }

// per-env task scratch kept in shared memory next to EnvS
template <int NA>
struct TaskS {
  int in_scene[FM_NCUBE], out_scene[FM_NCUBE], n_in, n_out, step_counter, since_spawn, fail, out_counter, score[2], last_score[2], ep_len;
  double conveyor_speed, spawn_freq, play_time;
  float reward, ep_ret;
  int terminated, out_of_reach, force_term, done_prev;
  int stale, index_err;   // see ph_task: dm_control dirty flag at the end of the task step; flagged IndexError case of task_utils.py:110
  int order[FM_NCUBE];
  float last_gd[NA], last_bd[NA];
  float site[NA][3];
  int fsm_state[NA], ignore[NA][NA];
  float act_norm;
};

FM_HD unsigned long long fm_mulhi64(unsigned long long a, unsigned long long b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
#endif
}
// numpy PCG64 next double in [0,1): 128-bit LCG + XSL-RR output (task_utils.py:19,49-51 use default_rng(seed).uniform)
FM_HD double pcg64_uniform(unsigned long long* st /*hi, lo, inc_hi, inc_lo*/) {
  const unsigned long long MH = 0x2360ED051FC65DA4ULL, ML = 0x4385DF649FCCF645ULL;
  unsigned long long hi = st[0], lo = st[1];
  unsigned long long nlo = lo * ML;
  unsigned long long nhi = fm_mulhi64(lo, ML) + hi * ML + lo * MH;
  unsigned long long slo = nlo + st[3];
  unsigned long long shi = nhi + st[2] + (slo < nlo ? 1ULL : 0ULL);
  st[0] = shi; st[1] = slo;
  unsigned long long x = shi ^ slo; unsigned rot = (unsigned)(shi >> 58);
  unsigned long long out = (x >> rot) | (x << ((64 - rot) & 63));
  return (double)(out >> 11) * (1.0 / 9007199254740992.0);
}

// ---- load / store between HBM (SoA) and shared memory
template <int NA, int E>
FM_PH void ph_load(EnvS<NA>* S, TaskS<NA>* TS, const FmStatePtrs& st, int env0, const int* ord, int tid, int T) {
  const int NENV = st.n;
  FM_ASSUME_SHARED(S); FM_ASSUME_SHARED(TS);
  typedef Lay<NA> L;
  int N = st.n;
  for (int w = tid; w < L::NF * E; w += T) {
    int e = w % E, fidx = w / E; int env = slot_env(ord, env0 + e, NENV);
    if (env >= N) env = N - 1;   // padding slot of the last block: steps a copy of the last env, nothing of it is stored
    EnvS<NA>& s = S[e];
    float v = st.f[(size_t)fidx * N + env];
    if (fidx == L::F_BELT_Q) s.belt_q = v;
    else if (fidx == L::F_BELT_V) s.belt_v = v;
    else if (fidx < L::F_CUBE_V) { /* post-edit cube qpos: applied after the first derive */ }
    else if (fidx < L::F_ARM_Q) { }
    else if (fidx < L::F_ARM_V) (&s.arm_q[0][0])[fidx - L::F_ARM_Q] = v;
    else if (fidx < L::F_CTRL_T) (&s.arm_v[0][0])[fidx - L::F_ARM_V] = v;
    else if (fidx < L::F_ACC_WS) s.ctrl_target[fidx - L::F_CTRL_T] = v;
    else if (fidx < L::F_PRE_Q) s.belt_S = v;
    else if (fidx < L::F_PRE_V) (&s.cube_q[0][0])[fidx - L::F_PRE_Q] = v;      // pre-edit cube state first (stale-data quirk)
    else if (fidx < L::F_LAST_GD) (&s.cube_v[0][0])[fidx - L::F_PRE_V] = v;
    else if (fidx < L::F_LAST_BD) TS[e].last_gd[fidx - L::F_LAST_GD] = v;
    else if (fidx < L::F_IK_ACT) TS[e].last_bd[fidx - L::F_LAST_BD] = v;
    else if (fidx == L::F_EP_RET) TS[e].ep_ret = v;
    else if (fidx >= L::F_PREV_F && fidx < L::F_PREV_AF) { }   // contact forces persist in the library scratch (st.ws)
    else if (fidx >= L::F_ARM_QLO) { }   // into the global scratch row below
    else if (fidx >= L::F_CUBE_QLO) (&s.cube_qlo[0][0])[fidx - L::F_CUBE_QLO] = v;
    else if (fidx >= L::F_PREV_AF) (&s.prev_arow_f[0][0])[fidx - L::F_PREV_AF] = v;
  }
  for (int w = tid; w < L::NI * E; w += T) {
    int e = w % E, k = w / E; int env = slot_env(ord, env0 + e, NENV);
    if (env >= N) env = N - 1;   // padding slot of the last block: steps a copy of the last env, nothing of it is stored
    int v = st.i[(size_t)k * N + env];
    TaskS<NA>& t = TS[e];
    if (k == L::I_PREV_N) S[e].prev_n = v;
    else if (k >= L::I_PREV_KEY && k < L::I_PREV_AN) S[e].prev_key[k - L::I_PREV_KEY] = v;
    else if (k >= L::I_PREV_AN && k < L::I_PREV_ACODE) S[e].prev_arow_n[k - L::I_PREV_AN] = v;
    else if (k >= L::I_PREV_ACODE) (&S[e].prev_arow_code[0][0])[k - L::I_PREV_ACODE] = v;
    if (k < 10) t.in_scene[k] = v; else if (k < 20) t.out_scene[k - 10] = v;
    else if (k == L::I_NIN) t.n_in = v; else if (k == L::I_NOUT) t.n_out = v; else if (k == L::I_STEP) t.step_counter = v;
    else if (k == L::I_SINCE) t.since_spawn = v; else if (k == L::I_FAIL) t.fail = v; else if (k == L::I_OUTCNT) t.out_counter = v;
    else if (k == L::I_SCORE) t.score[0] = v; else if (k == L::I_SCORE + 1) t.score[1] = v;
    else if (k == L::I_LAST_SCORE) t.last_score[0] = v; else if (k == L::I_LAST_SCORE + 1) t.last_score[1] = v;
    else if (k == L::I_EP_LEN) t.ep_len = v;
    else if (k >= L::I_FSM_STATE && k < L::I_FSM_CNT) t.fsm_state[k - L::I_FSM_STATE] = v;
    else if (k >= L::I_IGNORE && k < L::I_FLAGS) t.ignore[(k - L::I_IGNORE) / NA][(k - L::I_IGNORE) % NA] = v;
  }
  for (int w = tid; w < E; w += T) {
    int env = slot_env(ord, env0 + w, NENV);
    float* ws = st.ws + (size_t)env * FmCap<NA>::WS;    // the scratch has rows for the padding slots too
    if (env >= N) env = N - 1;
    TS[w].conveyor_speed = st.d[(size_t)L::D_CONV * N + env];
    TS[w].spawn_freq = st.d[(size_t)L::D_SPAWN_FREQ * N + env];
    TS[w].play_time = st.d[(size_t)L::D_PLAY_TIME * N + env];
    S[w].conveyor_speed = (float)TS[w].conveyor_speed;
    S[w].flags = 0; S[w].force_max = 0.f;
    S[w].broad_stale = 1; S[w].nqueue = 0;                 // the broad-phase queue is rebuilt in the first sub-step of every launch
    S[w].as_J = reinterpret_cast<float (*)[27]>(ws);
    S[w].as_B = reinterpret_cast<float (*)[27]>(ws + FmCap<NA>::ARMSIDE * 27);
    S[w].prev_f = reinterpret_cast<float (*)[4]>(ws + FmCap<NA>::ARMSIDE * 54);
    S[w].ct_d0 = ws + FmCap<NA>::ARMSIDE * 54 + FmCap<NA>::CON * 4;
    for (int k = 0; k < 9 * NA; ++k) S[w].arm_qlo(0)[k] = st.f[(size_t)(L::F_ARM_QLO + k) * N + env];
  }
}

// the task manager's edits (spawn / hide) become visible to the integrator only after the first derive
template <int NA, int E>
FM_PH void ph_apply_edits(EnvS<NA>* S, const FmStatePtrs& st, int env0, const int* ord, int tid, int T) {
  const int NENV = st.n;
  FM_ASSUME_SHARED(S);
  typedef Lay<NA> L;
  int N = st.n;
  for (int w = tid; w < 130 * E; w += T) {
    int e = w % E, k = w / E; int env = slot_env(ord, env0 + e, NENV);
    if (env >= N) env = N - 1;   // padding slot of the last block: steps a copy of the last env, nothing of it is stored
    if (k < 70) (&S[e].cube_q[0][0])[k] = st.f[(size_t)(L::F_CUBE_Q + k) * N + env];
    else (&S[e].cube_v[0][0])[k - 70] = st.f[(size_t)(L::F_CUBE_V + k - 70) * N + env];
  }
}

// ---- action map -> ctrl (clipped to actuator_ctrlrange, base_env.py:255-262)
template <int NA, int E>
FM_PH void ph_action(EnvS<NA>* S, TaskS<NA>* TS, const FmDevModel& m, const FmEnvCfg& cfg, const FmStatePtrs& st, const float* act_f,
                     const int* act_i, int env0, const int* ord, int tid, int T) {
  const int NENV = st.n;
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m); FM_ASSUME_SHARED(TS);
  typedef Lay<NA> L;
  int N = st.n;
  for (int w = tid; w < (1 + 8 * NA) * E; w += T) {
    int e = w % E, u = w / E; int env = slot_env(ord, env0 + e, NENV);
    if (env >= N) continue;
    EnvS<NA>& s = S[e];
    if (u == 0) { s.ctrl[0] = clampf((float)TS[e].conveyor_speed, m.ctrl_lo[0], m.ctrl_hi[0]); continue; }
    int a = (u - 1) / 8, k = (u - 1) % 8;
    float lo = m.ctrl_lo[1 + k], hi = m.ctrl_hi[1 + k];
    float ikv = st.f[(size_t)(L::F_IK_ACT + 8 * a + k) * N + env];
    float v = 0.f;
    int cls = cfg.env_class;
    bool rl_arm = cls == FM_ENV_ALL_FULL || cls == FM_ENV_ALL_DELTA || ((cls == FM_ENV_SINGLE_FULL || cls == FM_ENV_SINGLE_DELTA) && a == 0);
    float mapped = 0.f;
    if (rl_arm) {
      int adim = (cls == FM_ENV_ALL_FULL || cls == FM_ENV_ALL_DELTA) ? 8 * NA : 8;
      float t = tanhf(act_f[(size_t)env * adim + 8 * a + k]);
      float uu = (t + 1.0f) * 0.5f;
      mapped = lo + uu * (hi - lo);
    }
    if (cls == FM_ENV_BASE_IK) v = ikv;
    else if (cls == FM_ENV_ALL_FULL || cls == FM_ENV_SINGLE_FULL) v = rl_arm ? mapped : ikv;
    else if (cls == FM_ENV_ALL_DELTA || cls == FM_ENV_SINGLE_DELTA) v = (rl_arm && TS[e].fsm_state[a] != 0) ? ikv + 0.5f * mapped : ikv;
    else if (cls == FM_ENV_PAUSE_TOGGLE) {
      float last = st.f[(size_t)(L::F_LAST_ARM_ACT + 8 * a + k) * N + env];
      v = act_i[(size_t)env * NA + a] == 1 ? ikv : last;
      st.f[(size_t)(L::F_LAST_ARM_ACT + 8 * a + k) * N + env] = v;
    } else if (cls == FM_ENV_BACKUP_TOGGLE) {
      const float dp[8] = {-0.5f, -0.5f, 0.0f, 1.0f, 0.0f, -1.6f, 0.0f, 0.06f};
      v = act_i[(size_t)env * NA + a] == 1 ? ikv : dp[k];
    }
    s.ctrl[u] = clampf(v, lo, hi);
    s.ct_d0[u] = s.ctrl_target[u] - s.ctrl[u];     // same lane reads it back in ph_ctrl
  }
  // action-norm term of the progress reward (environments.py:371-373)
  for (int w = tid; w < E; w += T) {
    int env = slot_env(ord, env0 + w, NENV);
    if (env >= N) continue;
    float an = 0.f;
    int cls = cfg.env_class;
    int adim = (cls == FM_ENV_ALL_FULL || cls == FM_ENV_ALL_DELTA) ? 8 * NA : ((cls == FM_ENV_SINGLE_FULL || cls == FM_ENV_SINGLE_DELTA) ? 8 : 0);
    for (int i = 0; i < adim; ++i) if (i % 8 != 7) { float v = act_f[(size_t)env * adim + i]; an += v * v; }
    TS[w].act_norm = sqrtf(an);
  }
}

// ---- save the physical (pre-edit) cube state for the next launch's first derive
template <int NA, int E>
FM_PH void ph_store_pre(EnvS<NA>* S, const FmStatePtrs& st, int env0, const int* ord, int tid, int T) {
  const int NENV = st.n;
  FM_ASSUME_SHARED(S);
  typedef Lay<NA> L;
  int N = st.n;
  for (int w = tid; w < 130 * E; w += T) {
    int e = w % E, k = w / E; int env = slot_env(ord, env0 + e, NENV);
    if (env >= N) continue;
    float v = k < 70 ? (&S[e].cube_q[0][0])[k] : (&S[e].cube_v[0][0])[k - 70];
    st.f[(size_t)(L::F_PRE_Q + k) * N + env] = v;
  }
}

template <int NA>
FM_HD void task_hide(EnvS<NA>& s, TaskS<NA>& t, int k) {
  t.out_scene[t.n_out++] = k;
  float* q = s.cube_q[k]; float* v = s.cube_v[k];
  q[0] = 5.0f; q[1] = (float)(t.out_counter * 0.2); q[2] = 1.0f; q[3] = 1.f; q[4] = 0.f; q[5] = 0.f; q[6] = 0.f;
  for (int i = 0; i < 6; ++i) v[i] = 0.f;
  for (int i = 0; i < 3; ++i) s.cube_qlo[k][i] = 0.f;
  t.out_counter++;
}
template <int NA>
FM_HD void task_remove_in(TaskS<NA>& t, int idx) { for (int i = idx; i + 1 < t.n_in; ++i) t.in_scene[i] = t.in_scene[i + 1]; t.n_in--; }

template <int NA>
FM_HD void task_reset_state(EnvS<NA>& s, TaskS<NA>& t, const FmEnvCfg& cfg) {
  s.belt_q = 0.f; s.belt_v = 0.f;
  for (int a = 0; a < NA; ++a) for (int d = 0; d < 9; ++d) { s.arm_q[a][d] = 0.f; s.arm_v[a][d] = 0.f; s.arm_qlo(a)[d] = 0.f; }
  for (int k = 0; k < FM_NCUBE; ++k) {
    float* q = s.cube_q[k]; float* v = s.cube_v[k];
    q[0] = 4.0f; q[1] = (float)(0.0 + k * 0.2); q[2] = 1.0f; q[3] = 1.f; q[4] = 0.f; q[5] = 0.f; q[6] = 0.f;
    for (int i = 0; i < 6; ++i) v[i] = 0.f;
    for (int i = 0; i < 3; ++i) s.cube_qlo[k][i] = 0.f;
    t.out_scene[k] = k;
  }
  t.n_out = FM_NCUBE; t.n_in = 0;
  t.step_counter = 0; t.since_spawn = 0; t.fail = 0; t.out_counter = 0; t.score[0] = t.score[1] = 0;
  t.last_score[0] = t.last_score[1] = 0;
  t.conveyor_speed = cfg.initial_conveyor_speed; t.spawn_freq = cfg.init_spawn_freq; t.play_time = 0.0;
  for (int u = 0; u < EnvS<NA>::NU; ++u) s.ctrl_target[u] = 0.f;
  s.belt_S = 0.f;
  t.ep_len = 0; t.ep_ret = 0.f;
  s.prev_n = 0;
  for (int a = 0; a < NA; ++a) s.prev_arow_n[a] = 0;
}

// ---- task manager step + reward + termination (thread per env)
template <int NA, int E>
FM_PH void ph_task(EnvS<NA>* S, TaskS<NA>* TS, const FmDevModel& m, const FmEnvCfg& cfg, const FmStatePtrs& st, int env0, const int* ord, int tid, int T) {
  const int NENV = st.n;
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m); FM_ASSUME_SHARED(TS);
  int N = st.n;
  for (int w = tid; w < E; w += T) {
    int env = slot_env(ord, env0 + w, NENV);
    if (env >= N) continue;
    EnvS<NA>& s = S[w]; TaskS<NA>& t = TS[w];
    // TaskManager.step (task_utils.py:135-144)
    int spawn_steps = (int)(1.0 / (0.1 * t.spawn_freq));
    bool dirty = false;
    t.index_err = 0;
    if (t.step_counter == 0 || t.since_spawn >= spawn_steps) {
      if (t.n_out > 0) {
        int k = t.out_scene[0];
        for (int i = 0; i + 1 < t.n_out; ++i) t.out_scene[i] = t.out_scene[i + 1];
        t.n_out--;
        unsigned long long r4[4];
        for (int i = 0; i < 4; ++i) r4[i] = st.rng[(size_t)i * N + env];
        float* q = s.cube_q[k];
        q[0] = 0.f; q[1] = 1.f; q[2] = 2.f;
        for (int i = 0; i < 3; ++i) s.cube_qlo[k][i] = 0.f;
        // raw uniforms (task_utils.py:49-51) ... which the mj_forward triggered by _check_states' bucket-xpos read normalises in
        // place (mj_kinematics) before _get_state reads them: the observation of the spawn step already shows a unit quaternion
        // (runs/rk5rxnav.zip:_last_obs env 4)
        double qr[4], qn = 0.0;
        for (int i = 0; i < 4; ++i) { qr[i] = pcg64_uniform(r4); qn += qr[i] * qr[i]; }
        qn = sqrt(qn);
        for (int i = 0; i < 4; ++i) q[3 + i] = (float)(qr[i] / qn);
        for (int i = 0; i < 2; ++i) st.rng[(size_t)i * N + env] = r4[i];
        t.in_scene[t.n_in++] = k;
        dirty = true;
      }
      t.since_spawn = 0;
    }
    // _check_states (task_utils.py:79-113).  `dirty` follows dm_control's mjcf.Physics: every write through a binding (spawn,
    // hide) marks the physics dirty, and reading a derived quantity through a binding — `physics.bind(bucket).xpos` at
    // task_utils.py:101 and environments.py:363 — runs mj_forward first [3P].  If the flag survives the task step, the next
    // physics.step() starts from contacts / Jacobians of the pre-edit pose (stale); otherwise from the edited one.
    for (int idx = t.n_in - 1; idx >= 0; --idx) {
      const float* q = s.cube_q[t.in_scene[idx]];
      bool out = fabs((double)q[0]) > 1.2 || (double)q[1] < -1.5 || (double)q[2] < 0.9;
      if (out) { int k = t.in_scene[idx]; task_remove_in<NA>(t, idx); task_hide(s, t, k); t.fail++; dirty = true; }
    }
    if (t.n_in > 0) {
      // ONE position snapshot for both buckets (task_utils.py:97); bucket 1's hit indices are applied to the list as bucket 0's
      // removals left it — the reference's index mismatch, reproduced (an index past the end raises IndexError there: flagged)
      float snap[FM_NCUBE][3]; const int nsnap = t.n_in;
      for (int i = 0; i < nsnap; ++i) for (int c = 0; c < 3; ++c) snap[i][c] = s.cube_q[t.in_scene[i]][c];
      for (int b = 0; b < 2; ++b) {
        dirty = false;                                      // physics.bind(b).xpos
        double bx = m.bucket_pos[b][0], by = m.bucket_pos[b][1], bz = m.bucket_pos[b][2];
        double sx = m.bucket_size[b][0], sy = m.bucket_size[b][1], sz = m.bucket_size[b][2];
        for (int idx = nsnap - 1; idx >= 0; --idx) {
          const float* q = snap[idx];
          bool in_x = fabs((double)q[0] - bx) <= 0.6 * sx, in_y = fabs((double)q[1] - by) <= 0.6 * sy, in_z = (double)q[2] - bz - sz / 2 <= 0.07;
          if (!(in_x && in_y && in_z)) continue;
          if (idx >= t.n_in) { t.index_err = 1; continue; }
          int k = t.in_scene[idx]; task_remove_in<NA>(t, idx); task_hide(s, t, k); t.score[b]++; dirty = true;
        }
      }
    }
    t.step_counter++; t.since_spawn++;
    // step_sim tail (base_env.py:266-272)
    t.play_time += cfg.dt;
    t.conveyor_speed += cfg.conveyor_acceleration * cfg.dt;
    t.spawn_freq *= cfg.spawn_freq_increase;
    t.out_of_reach = t.fail > 0;
    t.force_term = s.force_max > (float)cfg.force_threshold;
    t.terminated = t.out_of_reach || t.force_term;
    // reward (environments.py:129-149, 342-383)
    float score_reward = (float)((t.score[0] + t.score[1]) - (t.last_score[0] + t.last_score[1]));
    float reward = score_reward;
    bool progress = cfg.env_class >= FM_ENV_SINGLE_FULL && cfg.env_class <= FM_ENV_ALL_DELTA;
    t.stale = (dirty && !progress) ? 1 : 0;               // ProgressRewardEnv._get_reward reads bucket xpos every step
    if (progress) {
      double gc = 0.0, bc = 0.0;
      for (int a = 0; a < NA; ++a) {
        f3 gp = ld3(t.site[a]);
        int best = -1; float bd = 1e30f;
        for (int i = 0; i < t.n_in; ++i) {
          int k = t.in_scene[i]; bool ign = false;
          for (int o = 0; o < NA; ++o) ign = ign || t.ignore[a][o] == k;
          if (ign) continue;
          float dd = norm3(ld3(s.cube_q[k]) - gp);
          if (dd < bd) { bd = dd; best = k; }
        }
        if (best >= 0) {
          gc += (double)t.last_gd[a] - (double)bd; t.last_gd[a] = bd;
          float db = norm3(ld3(s.cube_q[best]) - ld3(m.bucket_pos[a % 2]));
          bc += (double)t.last_bd[a] - (double)db; t.last_bd[a] = db;
        }
      }
      double pr = cfg.base_reward + cfg.w_gripper_cube * gc + cfg.w_cube_bucket * bc + cfg.w_action_norm * exp(-(double)t.act_norm);
      reward = score_reward > 0.f ? score_reward : (float)pr;
    }
    t.reward = reward;
    t.last_score[0] = t.score[0]; t.last_score[1] = t.score[1];
    t.ep_len++; t.ep_ret += reward;
    // observation order: in-scene cubes by ascending x (stable insertion sort, task_utils.py:68)
    for (int i = 0; i < t.n_in; ++i) {
      int k = t.in_scene[i]; float x = s.cube_q[k][0];
      int j = i;
      while (j > 0 && s.cube_q[t.order[j - 1]][0] > x) { t.order[j] = t.order[j - 1]; --j; }
      t.order[j] = k;
    }
  }
}

template <int NA>
FM_HD float obs_value(const EnvS<NA>& s, const TaskS<NA>& t, int j) {
  if (j < 24 * NA) {
    int a = j / 24, r = j % 24;
    if (r < 8) return s.arm_q[a][r];
    if (r < 16) return s.arm_v[a][r - 8];
    return s.ctrl_target[1 + 8 * a + (r - 16)];
  }
  j -= 24 * NA;
  if (j < 70) { int i = j / 7, c = j % 7; return i < t.n_in ? s.cube_q[t.order[i]][c] : 0.f; }
  j -= 70;
  int i = j / 6, c = j % 6; return i < t.n_in ? s.cube_v[t.order[i]][c] : 0.f;
}

// write base observation (_get_state + _process_observation, base_env.py:149-175, environments.py:55-82)
template <int NA, int E>
FM_PH void ph_obs(EnvS<NA>* S, TaskS<NA>* TS, float* obs, int obs_dim, int n_envs, int env0, const int* ord, int tid, int T, bool only_done) {
  const int NENV = n_envs;
  FM_ASSUME_SHARED(S); FM_ASSUME_SHARED(TS);
  const int base = 24 * NA + 130;
  for (int w = tid; w < base * E; w += T) {
    int e = w / base, j = w % base; int env = slot_env(ord, env0 + e, NENV);
    if (env >= n_envs || obs == nullptr) continue;
    if (only_done && !TS[e].terminated) continue;
    obs[(size_t)env * obs_dim + j] = obs_value(S[e], TS[e], j);
  }
}

template <int NA, int E>
FM_PH void ph_outputs(EnvS<NA>* S, TaskS<NA>* TS, const FmStatePtrs& st, float* reward, unsigned char* terminated, int* scores,
                      unsigned char* flags, float* ep_info, int env0, const int* ord, int tid, int T) {
  const int NENV = st.n;
  FM_ASSUME_SHARED(S); FM_ASSUME_SHARED(TS);
  typedef Lay<NA> L;
  int N = st.n;
  for (int w = tid; w < E; w += T) {
    int env = slot_env(ord, env0 + w, NENV);
    if (env >= N) continue;
    TaskS<NA>& t = TS[w];
    if (reward) reward[env] = t.reward;
    if (terminated) terminated[env] = (unsigned char)t.terminated;
    if (scores) { scores[2 * env] = t.score[0]; scores[2 * env + 1] = t.score[1]; }
    if (flags) flags[env] = (unsigned char)((t.out_of_reach ? 1 : 0) | (t.force_term ? 2 : 0) | ((S[w].flags & 15) << 2) | (t.index_err ? 64 : 0));
    if (ep_info) { ep_info[4 * env] = t.ep_ret; ep_info[4 * env + 1] = (float)t.ep_len; ep_info[4 * env + 2] = (float)t.score[0]; ep_info[4 * env + 3] = (float)t.score[1]; }
    int fl = S[w].flags;
    if (fl) st.i[(size_t)L::I_FLAGS * N + env] |= fl;
    if (S[w].ncon > st.i[(size_t)L::I_MAXCON * N + env]) st.i[(size_t)L::I_MAXCON * N + env] = S[w].ncon;
    // contacts that touch an arm at the end of this step (0 after a termination: the env restarts): the caller's cue for grouping
    // the envs whose next step will be expensive (fm_set_env_order)
    int na = 0;
    if (!t.terminated)
      for (int c = 0; c < S[w].ncon; ++c) {
        const int cs = S[w].c_side[c];
        na += (side_type(cs & FM_SIDE_MASK) == FM_SIDE_ARM || side_type((cs >> FM_SIDE_BITS) & FM_SIDE_MASK) == FM_SIDE_ARM) ? 1 : 0;
      }
    st.i[(size_t)L::I_ARMCON * N + env] = na;
  }
}

// reset (reset_sim base_env.py:177-198 + TaskManager.reset task_utils.py:146-156) for terminated envs when auto-reset is on
template <int NA, int E>
FM_PH void ph_reset_done(EnvS<NA>* S, TaskS<NA>* TS, const FmEnvCfg& cfg, const FmStatePtrs& st, int env0, const int* ord, int tid, int T, const unsigned char* mask) {
  const int NENV = st.n;
  FM_ASSUME_SHARED(S); FM_ASSUME_SHARED(TS);
  typedef Lay<NA> L;
  int N = st.n;
  for (int w = tid; w < E; w += T) {
    int env = slot_env(ord, env0 + w, NENV);
    if (env >= N) continue;
    bool doit = mask ? (mask[env] != 0) : (TS[w].terminated != 0);
    TS[w].done_prev = doit ? 1 : 0;
    if (!doit) continue;
    task_reset_state(S[w], TS[w], cfg);
    TS[w].n_in = 0;
    st.i[(size_t)L::I_EPISODES * N + env] += 1;
  }
}

template <int NA, int E>
FM_PH void ph_store(EnvS<NA>* S, TaskS<NA>* TS, const FmStatePtrs& st, int env0, const int* ord, int tid, int T) {
  const int NENV = st.n;
  FM_ASSUME_SHARED(S); FM_ASSUME_SHARED(TS);
  typedef Lay<NA> L;
  int N = st.n;
  for (int w = tid; w < L::NF * E; w += T) {
    int e = w % E, fidx = w / E; int env = slot_env(ord, env0 + e, NENV);
    if (env >= N) continue;
    EnvS<NA>& s = S[e];
    float v; bool wr = true;
    if (fidx == L::F_BELT_Q) v = s.belt_q;
    else if (fidx == L::F_BELT_V) v = s.belt_v;
    else if (fidx < L::F_CUBE_V) v = (&s.cube_q[0][0])[fidx - L::F_CUBE_Q];
    else if (fidx < L::F_ARM_Q) v = (&s.cube_v[0][0])[fidx - L::F_CUBE_V];
    else if (fidx < L::F_ARM_V) v = (&s.arm_q[0][0])[fidx - L::F_ARM_Q];
    else if (fidx < L::F_CTRL_T) v = (&s.arm_v[0][0])[fidx - L::F_ARM_V];
    else if (fidx < L::F_ACC_WS) v = s.ctrl_target[fidx - L::F_CTRL_T];
    else if (fidx < L::F_PRE_Q) v = s.belt_S;
    else if (fidx < L::F_PRE_V) { v = (&s.cube_q[0][0])[fidx - L::F_PRE_Q]; wr = TS[e].done_prev != 0 || !TS[e].stale; }   // after a reset / a forward the model is in sync
    else if (fidx < L::F_LAST_GD) { v = (&s.cube_v[0][0])[fidx - L::F_PRE_V]; wr = TS[e].done_prev != 0 || !TS[e].stale; }
    else if (fidx < L::F_LAST_BD) v = TS[e].last_gd[fidx - L::F_LAST_GD];
    else if (fidx < L::F_IK_ACT) v = TS[e].last_bd[fidx - L::F_LAST_BD];
    else if (fidx >= L::F_IK_LAST_CTRL && fidx < L::F_IK_MOVE_START) {   // IKPolicy.reset -> idle_ctrl (ik_policy.py:120-127)
      const float dp[8] = {-0.5f, -0.5f, 0.0f, 1.0f, 0.0f, -1.6f, 0.0f, 0.06f};
      v = dp[(fidx - L::F_IK_LAST_CTRL) % 8]; wr = TS[e].done_prev != 0;
    }
    else if (fidx >= L::F_SITE && fidx < L::F_EP_RET) v = (&TS[e].site[0][0])[fidx - L::F_SITE];
    else if (fidx == L::F_EP_RET) v = TS[e].ep_ret;
    else if (fidx >= L::F_PREV_F && fidx < L::F_PREV_AF) wr = false;
    else if (fidx >= L::F_ARM_QLO) v = s.arm_qlo(0)[fidx - L::F_ARM_QLO];
    else if (fidx >= L::F_CUBE_QLO) v = (&s.cube_qlo[0][0])[fidx - L::F_CUBE_QLO];
    else if (fidx >= L::F_PREV_AF) v = (&s.prev_arow_f[0][0])[fidx - L::F_PREV_AF];
    else wr = false;
    if (wr) st.f[(size_t)fidx * N + env] = v;
  }
  for (int w = tid; w < L::NI * E; w += T) {
    int e = w % E, k = w / E; int env = slot_env(ord, env0 + e, NENV);
    if (env >= N) continue;
    TaskS<NA>& t = TS[e];
    int v; bool wr = true;
    if (k < 10) v = t.in_scene[k]; else if (k < 20) v = t.out_scene[k - 10];
    else if (k == L::I_NIN) v = t.n_in; else if (k == L::I_NOUT) v = t.n_out; else if (k == L::I_STEP) v = t.step_counter;
    else if (k == L::I_SINCE) v = t.since_spawn; else if (k == L::I_FAIL) v = t.fail; else if (k == L::I_OUTCNT) v = t.out_counter;
    else if (k == L::I_SCORE) v = t.score[0]; else if (k == L::I_SCORE + 1) v = t.score[1];
    else if (k == L::I_LAST_SCORE) v = t.last_score[0]; else if (k == L::I_LAST_SCORE + 1) v = t.last_score[1];
    else if (k == L::I_EP_LEN) v = t.ep_len;
    else if (k >= L::I_FSM_STATE && k < L::I_FSM_TGT) { v = 0; wr = t.done_prev != 0; }
    else if (k >= L::I_FSM_TGT && k < L::I_IGNORE) { v = -1; wr = t.done_prev != 0; }
    else if (k == L::I_PREV_N) v = S[e].prev_n;
    else if (k >= L::I_PREV_KEY && k < L::I_PREV_AN) v = S[e].prev_key[k - L::I_PREV_KEY];
    else if (k >= L::I_PREV_AN && k < L::I_PREV_ACODE) v = S[e].prev_arow_n[k - L::I_PREV_AN];
    else if (k >= L::I_PREV_ACODE) v = (&S[e].prev_arow_code[0][0])[k - L::I_PREV_ACODE];
    else wr = false;
    if (wr) st.i[(size_t)k * N + env] = v;
  }
  for (int w = tid; w < E; w += T) {
    int env = slot_env(ord, env0 + w, NENV);
    if (env >= N) continue;
    st.d[(size_t)L::D_CONV * N + env] = TS[w].conveyor_speed;
    st.d[(size_t)L::D_SPAWN_FREQ * N + env] = TS[w].spawn_freq;
    st.d[(size_t)L::D_PLAY_TIME * N + env] = TS[w].play_time;
  }
}

// gripper site positions (between_gripper_plates, gripper.xml:43) after the final kinematics
template <int NA, int E>
FM_PH void ph_site(EnvS<NA>* S, TaskS<NA>* TS, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m); FM_ASSUME_SHARED(TS);
  for (int w = tid; w < NA * E; w += T) {
    int e = w % E, a = w / E;
    f3 p = ld3(S[e].body_x[a][m.site_body]) + rmul(S[e].body_R[a][m.site_body], ld3(m.site_pos));
    st3(TS[e].site[a], p);
  }
}
