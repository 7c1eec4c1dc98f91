// fm_driver.cuh — phase schedule of one env step for a block of E envs (shared by the CUDA kernel and by the
// host-compiled logic-test build).  `Run` executes one barrier interval: on the GPU it calls the body for
// threadIdx.x and then __syncthreads(); in the host test build it loops over all thread ids.
//
// Order of one control step (reference: BaseEnv.step_sim base_env.py:240-282):
//   load state (cubes in their pre-edit pose) -> derive (mj_step1 of the pre-edit state: the reference does not
//   re-run kinematics after TaskManager writes qpos, so the first sub-step sees stale contacts) -> apply edits ->
//   100 x { PT1 filter, actuation, smooth acceleration, PGS, implicitfast, integrate, derive } -> force check ->
//   task logic, reward, observation, optional auto-reset -> store.
#pragma once
#include "fm_task.cuh"

struct FmDebugDump {   // optional per-launch dump of one env's intermediates after the first derive (tests only)
  int env, ncon, nqueue, flags;
  int c_g[FM_MAXCON_ALL];
  float c_p[FM_MAXCON_ALL][3], c_n[FM_MAXCON_ALL][3], c_dist[FM_MAXCON_ALL], c_f[FM_MAXCON_ALL][4];
  float body_x[FM_MAX_ARMS][FM_NARMBODY][3], body_R[FM_MAX_ARMS][FM_NARMBODY][9];
  float M[FM_MAX_ARMS][45], bias[FM_MAX_ARMS][9];
  float acc_smooth[61 + 9 * FM_MAX_ARMS], acc[61 + 9 * FM_MAX_ARMS];
  int pgs_iters;
};

struct FmStepArgs {
  const float* act_f; const int* act_i;
  float* obs; float* reward; unsigned char* terminated; int* scores; unsigned char* flags; float* terminal_obs; float* ep_info;
  const unsigned char* reset_mask;
  int auto_reset;
  int mode;          // 0 = env step, 1 = reset only (mask or all), 2 = raw sub-steps (debug: no action map / task logic)
  int n_substeps;    // mode 2
  FmDebugDump* dump;
  const int* order;  // optional env order (n_envs ints on the device): block slot k serves env order[k]; nullptr = identity
};

template <int NA, int E, class Run>
FM_HD void fm_arm_kinematics(Run& run, EnvS<NA>* S, const FmDevModel& m, int T) {
  run([&](int tid) { ph_arm_local<NA, E>(S, m, tid, T); });
  run([&](int tid) { ph_arm_scan<NA, E>(S, tid, T, 1, true); });
  run([&](int tid) { ph_arm_scan<NA, E>(S, tid, T, 2, false); });
  run([&](int tid) { ph_arm_scan<NA, E>(S, tid, T, 4, true); });
  run([&](int tid) { ph_arm_plates<NA, E>(S, tid, T); });
}

// warps of the factorisation group / of the narrow-phase group when a block of TW warps splits (see fm_derive)
#ifndef FM_SPLIT_WA12
#define FM_SPLIT_WA12 8
#endif
FM_HDM constexpr int fm_split_a(int TW) { return TW >= 12 ? FM_SPLIT_WA12 : (TW >= 7 ? 5 : (TW >= 4 ? 3 : TW)); }

template <int NA, int E, int TW, class Run>
FM_HD void fm_derive(Run& run, EnvS<NA>* S, const FmDevModel& m, const unsigned int* cpairs, const int* geom_class, int sleep_hidden) {
  const int T = 32 * TW;
  run([&](int tid) { ph_clear<NA, E>(S, tid, T); ph_cube_derive<NA, E>(S, m, tid, T, sleep_hidden); ph_arm_local<NA, E>(S, m, tid, T); });
  run([&](int tid) { ph_arm_scan<NA, E>(S, tid, T, 1, true); });
  run([&](int tid) { ph_arm_scan<NA, E>(S, tid, T, 2, false); });
  run([&](int tid) { ph_arm_scan<NA, E>(S, tid, T, 4, true); });
  run([&](int tid) { ph_arm_plates<NA, E>(S, tid, T); });
  run([&](int tid) { ph_arm_vel<NA, E>(S, m, tid, T); });
  run([&](int tid) { ph_arm_body<NA, E>(S, m, tid, T); });
  run([&](int tid) { ph_arm_bwd<NA, E>(S, tid, T); ph_broad<NA, E>(S, m, cpairs, tid, T); });
  // Two independent chains run side by side on disjoint warp groups, each with its own barrier:
  //   A: mass matrix, the ten intervals of the two 9x9 Cholesky factorisations and the explicit inverse;
  //   B: narrow phase, whose longest item (a general box-box SAT, ~1.5 k dependent instructions on one lane) would otherwise
  //      hold every warp of the block at the next barrier.
  // The raw contacts reuse arm scratch that is dead after ph_arm_bwd.  Measured (4096 envs, 10 steps): no split 277 k,
  // {chol, minv} | narrow 293 k, {M, chol, minv} | narrow 300 k env-steps/s; also moving the set-up of the arm-free contacts
  // into B (FM_SPLIT_MODE 3) needs a second copy of that large phase and loses 9 % again.
  constexpr int WA = fm_split_a(TW), WB = TW - WA;
#ifndef FM_SPLIT_MODE
#define FM_SPLIT_MODE 1
#endif
  constexpr bool M_IN_A = (FM_SPLIT_MODE & 1) != 0, SETUP_IN_B = (FM_SPLIT_MODE & 2) != 0;
  if (!M_IN_A) run([&](int tid) { ph_arm_M<NA, E>(S, m, tid, T); });
  run.template split<WA, WB>(
      [&](auto& sub, int TA) {
        if (M_IN_A) sub([&](int tid) { ph_arm_M<NA, E>(S, m, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 0>(S, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 1>(S, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 2>(S, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 3>(S, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 4>(S, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 5>(S, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 6>(S, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 7>(S, tid, TA); });
        sub([&](int tid) { ph_arm_chol_col<NA, E, 8>(S, tid, TA); });
        sub([&](int tid) { ph_arm_minv<NA, E>(S, tid, TA); ph_arm_rows<NA, E>(S, m, tid, TA); });   // the arm-local rows need no contact
#ifndef FM_INTEGRATE_SERIAL
        sub([&](int tid) { ph_arm_winv<NA, E>(S, tid, TA);
#if defined(FM_NARROW_STEAL)
                           if (WB > 0) ph_narrow_steal<NA, E>(S, m, cpairs, tid & 31);   // group A is done: help group B with the narrow phase
#endif
        });                                       // Lc is dead now: (M - hD)^-1 goes there
#endif
      },
      [&](auto& sub, int TB) {
        sub([&](int tid) { ph_narrow<NA, E>(S, m, cpairs, tid, TB); });
        if (SETUP_IN_B) sub([&](int tid) { ph_contact_setup<NA, E>(S, m, geom_class, tid, TB, 1); });
      });
  run([&](int tid) { ph_contact_setup<NA, E>(S, m, geom_class, tid, T, SETUP_IN_B ? 2 : 0); });
}

template <int NA, int E>
FM_HD void fm_dump(EnvS<NA>* S, FmDebugDump* d, int env0, const int* ord, int n_envs, int tid, int stage) {
  if (!d || tid != 0) return;
  int e = -1;
  for (int k = 0; k < E; ++k) if (slot_env(ord, env0 + k, n_envs) == d->env) e = k;
  if (e < 0) return;
  const EnvS<NA>& s = S[e];
  if (stage == 0) {
    d->ncon = s.ncon; d->nqueue = s.nqueue; d->flags = s.flags;
    for (int c = 0; c < s.ncon && c < FmCap<NA>::CON; ++c) {
      d->c_g[c] = ((s.c_key[c] >> 13) << 12) | ((s.c_key[c] >> 4) & 511);
      for (int k = 0; k < 3; ++k) { d->c_p[c][k] = s.c_p[c][k]; d->c_n[c][k] = s.c_n[c][k]; }
    }
    for (int a = 0; a < NA; ++a) {
      for (int b = 0; b < FM_NARMBODY; ++b) { for (int k = 0; k < 3; ++k) d->body_x[a][b][k] = s.body_x[a][b][k]; for (int k = 0; k < 9; ++k) d->body_R[a][b][k] = s.body_R[a][b][k]; }
      for (int k = 0; k < 45; ++k) d->M[a][k] = s.M[a][k];
      for (int k = 0; k < 9; ++k) d->bias[a][k] = s.bias[a][k];
    }
  } else if (stage == 1) {
    for (int k = 0; k < EnvS<NA>::NV; ++k) d->acc_smooth[k] = s.acc[k];
  } else {
    for (int k = 0; k < EnvS<NA>::NV; ++k) d->acc[k] = s.acc[k];
    for (int c = 0; c < s.ncon && c < FmCap<NA>::CON; ++c) for (int r = 0; r < 4; ++r) d->c_f[c][r] = s.c_f[c][r];
  }
}

// TW = warps per block (default: one warp per env of the block)
template <int NA, int E, int TW = E, bool BIGN = false, class Run>
FM_HD void fm_block_step(Run& run, EnvS<NA>* S, TaskS<NA>* TS, const FmDevModel& m, const unsigned int* cpairs, const int* geom_class,
                         const FmEnvCfg& cfg, const FmStatePtrs& st, const FmStepArgs& ar, int env0) {
  const int T = 32 * TW;
  typedef Lay<NA> L;
  const int obs_dim = L::obs_dim(cfg.env_class);
  run([&](int tid) { ph_load<NA, E>(S, TS, st, env0, ar.order, tid, T); });
  if (ar.mode == 1) {   // reset only: masked envs (or all) go to the initial state, the others are written back untouched
    run([&](int tid) { ph_apply_edits<NA, E>(S, st, env0, ar.order, tid, T); });
    run([&](int tid) {
      for (int w = tid; w < E; w += T) {
        int env = slot_env(ar.order, env0 + w, st.n);
        if (env >= st.n) continue;
        bool doit = ar.reset_mask ? ar.reset_mask[env] != 0 : true;
        TS[w].done_prev = doit; TS[w].terminated = doit;
        if (doit) task_reset_state(S[w], TS[w], cfg);
      }
    });
    fm_arm_kinematics<NA, E>(run, S, m, T);
    run([&](int tid) { ph_site<NA, E>(S, TS, m, tid, T); ph_obs<NA, E>(S, TS, ar.obs, obs_dim, st.n, env0, ar.order, tid, T, true); });
    run([&](int tid) { ph_store<NA, E>(S, TS, st, env0, ar.order, tid, T); });
    return;
  }
  if (ar.mode == 0) run([&](int tid) { ph_action<NA, E>(S, TS, m, cfg, st, ar.act_f, ar.act_i, env0, ar.order, tid, T); });
  else run([&](int tid) {   // debug: ctrl_target is used as is (no filter): ctrl := ctrl_target
    for (int w = tid; w < EnvS<NA>::NU * E; w += T) S[w % E].ctrl[w / E] = S[w % E].ctrl_target[w / E];
  });
  fm_derive<NA, E, TW>(run, S, m, cpairs, geom_class, cfg.sleep_hidden);
  run([&](int tid) { ph_apply_edits<NA, E>(S, st, env0, ar.order, tid, T); });
  const double hd = cfg.dt / cfg.frame_skip;
  const int nsub = ar.mode == 0 ? cfg.frame_skip : ar.n_substeps;
  for (int t = 0; t < nsub; ++t) {
    if (t > 0) fm_derive<NA, E, TW>(run, S, m, cpairs, geom_class, cfg.sleep_hidden);
    run([&](int tid) { ph_island_union<NA, E>(S, tid, T); if (ar.mode == 0) ph_ctrl<NA, E>(S, tid, T, t < FM_MAX_FRAME_SKIP ? cfg.pt1_pow[t] : (float)pow(1.0 - hd / (hd + cfg.pt_time), t + 1.0)); });
    run([&](int tid) { ph_island_lists<NA, E>(S, tid, T); ph_smooth<NA, E>(S, m, tid, T); });
    if (t == 0 && ar.dump) run([&](int tid) { fm_dump<NA, E>(S, ar.dump, env0, ar.order, st.n, tid, 0); });
    if (t == 0 && ar.dump) run([&](int tid) { fm_dump<NA, E>(S, ar.dump, env0, ar.order, st.n, tid, 1); });
    run([&](int tid) { ph_check<NA, E>(S, m, tid, T); });
    run([&](int tid) { ph_pgs<NA, E, BIGN>(S, m, tid, T, cfg.pgs_inner, cfg.pgs_tol); });
    // BIGN: the kernel variant with the exact warp solve of arm + contact islands (the default; fm_config.pgs_outer = 1 selects the
    // other one).  Compiled separately: the presence of this phase in the fused loop costs the other phases ~10 % (profiles/README)
    if constexpr (BIGN) run.warps([&](int warp, auto& L) { ph_big_islands<NA, E>(S, m, warp, TW, L, cfg.pgs_inner, cfg.pgs_tol); });
    // belt lane: contact term of the belt dof (same lane as the belt row of ph_integrate), then integration
    run([&](int tid) { ph_belt<NA, E>(S, m, tid, T); if (t == nsub - 1) ph_force_check<NA, E>(S, tid, T); ph_integrate<NA, E>(S, m, tid, T); ph_save_forces<NA, E>(S, tid, T); });
    if (t == 0 && ar.dump) run([&](int tid) { fm_dump<NA, E>(S, ar.dump, env0, ar.order, st.n, tid, 2); });
  }
  if (ar.mode == 2) {   // debug: store the raw state, pre-edit == current
    run([&](int tid) { ph_store_pre<NA, E>(S, st, env0, ar.order, tid, T); for (int w = tid; w < E; w += T) TS[w].done_prev = 0; });
    run([&](int tid) { ph_store<NA, E>(S, TS, st, env0, ar.order, tid, T); });
    return;
  }
  // kinematics of the final state for the gripper sites (reward / IK policy read site_xpos after the last mj_step1)
  run([&](int tid) { ph_store_pre<NA, E>(S, st, env0, ar.order, tid, T); });
  fm_arm_kinematics<NA, E>(run, S, m, T);
  run([&](int tid) { ph_site<NA, E>(S, TS, m, tid, T); });
  run([&](int tid) { ph_task<NA, E>(S, TS, m, cfg, st, env0, ar.order, tid, T); });
  run([&](int tid) { ph_obs<NA, E>(S, TS, ar.terminal_obs, obs_dim, st.n, env0, ar.order, tid, T, true); ph_outputs<NA, E>(S, TS, st, ar.reward, ar.terminated, ar.scores, ar.flags, ar.ep_info, env0, ar.order, tid, T); });
  if (ar.auto_reset) run([&](int tid) { ph_reset_done<NA, E>(S, TS, cfg, st, env0, ar.order, tid, T, nullptr); });
  else run([&](int tid) { for (int w = tid; w < E; w += T) TS[w].done_prev = 0; });
  run([&](int tid) { ph_obs<NA, E>(S, TS, ar.obs, obs_dim, st.n, env0, ar.order, tid, T, false); });
  run([&](int tid) { ph_store<NA, E>(S, TS, st, env0, ar.order, tid, T); });
}
