// fm_ik.cuh — IK base policy for all arms of one env (thread per env; arms are sequential because each arm's target
// is pushed into the others' ignore map, environments.py:120-126).  fp64 throughout: the success criterion of
// dm_control's qpos_from_site_pose (||err|| < 1e-14, [3P]) is unreachable in fp32 (SURVEY.md hard part 4).
// Restates  IKPolicy.act / select_target_object / FSM      challenge_env/challenge_env/ik_policy.py:92-282
//           FactoryManipulationEnv._compose_control        src/environments.py:104-127
//           IKTogglingEnv._process_observation (obs tail)  src/environments.py:560-577
//           dm_control.utils.inverse_kinematics.qpos_from_site_pose(max_steps=10) — third party, restated [3P]
#pragma once
#include "fm_task.cuh"

struct dq4 { double w, x, y, z; };
FM_HD dq4 dqmul(dq4 a, dq4 b) {
  dq4 r; r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z; r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x; r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w; return r;
}
FM_HD dq4 dqnorm(dq4 q) { double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); if (n < 1e-15) { dq4 r = {1, 0, 0, 0}; return r; } dq4 r = {q.w / n, q.x / n, q.y / n, q.z / n}; return r; }
FM_HD void dq2mat(dq4 q, double* R) {
  double w = q.w, x = q.x, y = q.y, z = q.z;
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}

// FK of the 7-joint chain to the gripper site; optional 6x7 Jacobian (rows: pos xyz, rot xyz)
FM_HD void ik_site_fk(const FmIkModel& k, int a, const double* q7, double* spos, dq4* squat, double* J) {
  double pos[3] = {k.base_pos[a][0], k.base_pos[a][1], k.base_pos[a][2]};
  dq4 quat = {k.base_quat[a][0], k.base_quat[a][1], k.base_quat[a][2], k.base_quat[a][3]};
  double anchor[7][3], axis[7][3], R[9];
  for (int b = 1; b <= 8; ++b) {
    dq2mat(quat, R);
    for (int i = 0; i < 3; ++i) pos[i] += R[3 * i] * k.body_pos[b][0] + R[3 * i + 1] * k.body_pos[b][1] + R[3 * i + 2] * k.body_pos[b][2];
    dq4 bq = {k.body_quat[b][0], k.body_quat[b][1], k.body_quat[b][2], k.body_quat[b][3]};
    quat = dqmul(quat, bq);
    if (b <= 7) {
      double s = sin(q7[b - 1] / 2), c = cos(q7[b - 1] / 2);
      dq4 rz = {c, 0, 0, s};
      quat = dqmul(quat, rz);
      quat = dqnorm(quat);
      dq2mat(quat, R);
      for (int i = 0; i < 3; ++i) { anchor[b - 1][i] = pos[i]; axis[b - 1][i] = R[3 * i + 2]; }
    } else quat = dqnorm(quat);
  }
  dq2mat(quat, R);
  for (int i = 0; i < 3; ++i) spos[i] = pos[i] + R[3 * i] * k.site_pos[0] + R[3 * i + 1] * k.site_pos[1] + R[3 * i + 2] * k.site_pos[2];
  if (squat) *squat = quat;
  if (J) for (int j = 0; j < 7; ++j) {
    double r[3] = {spos[0] - anchor[j][0], spos[1] - anchor[j][1], spos[2] - anchor[j][2]};
    J[0 * 7 + j] = axis[j][1] * r[2] - axis[j][2] * r[1]; J[1 * 7 + j] = axis[j][2] * r[0] - axis[j][0] * r[2]; J[2 * 7 + j] = axis[j][0] * r[1] - axis[j][1] * r[0];
    J[3 * 7 + j] = axis[j][0]; J[4 * 7 + j] = axis[j][1]; J[5 * 7 + j] = axis[j][2];
  }
}

FM_HD bool ik_gauss(double* A, double* b, int n) {
  for (int c = 0; c < n; ++c) {
    int p = c; for (int r = c + 1; r < n; ++r) if (fabs(A[r * n + c]) > fabs(A[p * n + c])) p = r;
    if (fabs(A[p * n + c]) < 1e-300) return false;
    if (p != c) { for (int q = 0; q < n; ++q) { double t = A[p * n + q]; A[p * n + q] = A[c * n + q]; A[c * n + q] = t; } double t = b[p]; b[p] = b[c]; b[c] = t; }
    for (int r = c + 1; r < n; ++r) { double f = A[r * n + c] / A[c * n + c]; for (int q = c; q < n; ++q) A[r * n + q] -= f * A[c * n + q]; b[r] -= f * b[c]; }
  }
  for (int r = n - 1; r >= 0; --r) { double s = b[r]; for (int q = r + 1; q < n; ++q) s -= A[r * n + q] * b[q]; b[r] = s / A[r * n + r]; }
  return true;
}

FM_HD bool ik_solve(const FmIkModel& k, int a, const double* tpos, const double* tquat, double* q7) {
  const double tol = 1e-14, reg_thresh = 0.1, reg_strength = 3e-2, max_update = 2.0, progress_thresh = 20.0;
  const double PI = 3.14159265358979323846;
  for (int steps = 0; steps < 10; ++steps) {
    double sp[3], J[42], err[6]; dq4 sq;
    ik_site_fk(k, a, q7, sp, &sq, J);
    err[0] = tpos[0] - sp[0]; err[1] = tpos[1] - sp[1]; err[2] = tpos[2] - sp[2];
    double err_norm = sqrt(err[0] * err[0] + err[1] * err[1] + err[2] * err[2]);
    dq4 tq = {tquat[0], tquat[1], tquat[2], tquat[3]}; dq4 nq = {sq.w, -sq.x, -sq.y, -sq.z};
    dq4 eq = dqmul(tq, nq);
    double ax[3] = {eq.x, eq.y, eq.z}; double sn = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    if (sn < 1e-15) { ax[0] = 1; ax[1] = 0; ax[2] = 0; sn = 0; } else { ax[0] /= sn; ax[1] /= sn; ax[2] /= sn; }
    double speed = 2 * atan2(sn, eq.w);
    if (speed > PI) speed -= 2 * PI;
    err[3] = ax[0] * speed; err[4] = ax[1] * speed; err[5] = ax[2] * speed;
    err_norm += sqrt(err[3] * err[3] + err[4] * err[4] + err[5] * err[5]);
    if (err_norm < tol) return true;
    double upd[7];
    if (err_norm > reg_thresh) {
      double H[49], g[7];
      for (int i = 0; i < 7; ++i) { g[i] = 0; for (int r = 0; r < 6; ++r) g[i] += J[r * 7 + i] * err[r]; for (int c = 0; c < 7; ++c) { double s = 0; for (int r = 0; r < 6; ++r) s += J[r * 7 + i] * J[r * 7 + c]; H[i * 7 + c] = s + (i == c ? reg_strength : 0.0); } }
      ik_gauss(H, g, 7);
      for (int i = 0; i < 7; ++i) upd[i] = g[i];
    } else {
      double G[36], y[6];
      for (int r = 0; r < 6; ++r) { y[r] = err[r]; for (int c = 0; c < 6; ++c) { double s = 0; for (int q = 0; q < 7; ++q) s += J[r * 7 + q] * J[c * 7 + q]; G[r * 6 + c] = s; } }
      if (!ik_gauss(G, y, 6)) return false;
      for (int i = 0; i < 7; ++i) { upd[i] = 0; for (int r = 0; r < 6; ++r) upd[i] += J[r * 7 + i] * y[r]; }
    }
    double un = 0; for (int i = 0; i < 7; ++i) un += upd[i] * upd[i]; un = sqrt(un);
    if (err_norm / un > progress_thresh) return false;
    if (un > max_update) for (int i = 0; i < 7; ++i) upd[i] *= max_update / un;
    for (int i = 0; i < 7; ++i) q7[i] += upd[i];
  }
  return false;
}

template <int NA>
FM_HD void fm_ik_env(const FmIkModel& k, const FmEnvCfg& cfg, const FmStatePtrs& st, float* obs, int obs_dim, int env) {
  typedef Lay<NA> L;
  const size_t N = st.n; const size_t e = env;
  const double default_pose[8] = {-0.5, -0.5, 0.0, 1.0, 0.0, -1.6, 0.0, 0.06};
  const double dgq[4] = {0, 0, 1, 0};
  const double pt_comp = cfg.pt_time * cfg.dt * 15.0;
  const int release_wait = (int)(0.5 / cfg.dt), grasp_wait = (int)(1.0 / cfg.dt), move_steps = (int)(1.0 / cfg.dt), timeout_steps = (int)(3.0 / cfg.dt);
  int n_in = st.i[L::I_NIN * N + e];
  int in_scene[FM_NCUBE];
  for (int i = 0; i < FM_NCUBE; ++i) in_scene[i] = st.i[(L::I_IN + i) * N + e];
  auto cube_q = [&](int c, int j) { return (double)st.f[(L::F_CUBE_Q + 7 * c + j) * N + e]; };
  auto cube_v = [&](int c, int j) { return (double)st.f[(L::F_CUBE_V + 6 * c + j) * N + e]; };
  for (int a = 0; a < NA; ++a) {
    int state = st.i[(L::I_FSM_STATE + a) * N + e], counter = st.i[(L::I_FSM_CNT + a) * N + e], target = st.i[(L::I_FSM_TGT + a) * N + e];
    double out[8]; bool done = false, idle = false;
    auto ignored = [&](int c) { for (int o = 0; o < NA; ++o) if (st.i[(L::I_IGNORE + a * NA + o) * N + e] == c) return true; return false; };
    // ---- select_target_object (ik_policy.py:92-118)
    {
      bool keep = false;
      if (target >= 0) {
        double dx = cube_q(target, 0) - k.player_site[a][0], dy = cube_q(target, 1) - k.player_site[a][1], dz = cube_q(target, 2) - k.player_site[a][2];
        bool cand = false; for (int i = 0; i < n_in; ++i) cand = cand || (in_scene[i] == target && !ignored(target));
        keep = sqrt(dx * dx + dy * dy + dz * dz) < 1.0 && cand;
      }
      if (!keep) {
        int best = -1; double bd = 1e300;
        for (int i = 0; i < n_in; ++i) {
          int c = in_scene[i]; if (ignored(c)) continue;
          double dx = cube_q(c, 0) - k.player_site[a][0], dy = cube_q(c, 1) - 0.2 - k.player_site[a][1], dz = cube_q(c, 2) - k.player_site[a][2];
          double dd = sqrt(dx * dx + dy * dy + dz * dz);
          if (dd < bd) { bd = dd; best = c; }
        }
        target = (best >= 0 && bd < 0.8) ? best : -1;
      }
    }
    double q7[7];
    for (int j = 0; j < 7; ++j) q7[j] = (double)st.f[(L::F_ARM_Q + 9 * a + j) * N + e];
    if (target < 0) idle = true;
    double tpos[3] = {0, 0, 0}, tquat[4] = {1, 0, 0, 0}; bool close_gripper = false, compensate = false;
    if (!idle) {
      double op[3] = {cube_q(target, 0), cube_q(target, 1), cube_q(target, 2)};
      dq4 orot = dqnorm(dq4{cube_q(target, 3), cube_q(target, 4), cube_q(target, 5), cube_q(target, 6)});
      dq4 dg = {dgq[0], dgq[1], dgq[2], dgq[3]};
      double bestd = 1e300; dq4 tq = {1, 0, 0, 0};
      for (int g = 0; g < 24; ++g) {
        dq4 gq = {k.cube_group[g][0], k.cube_group[g][1], k.cube_group[g][2], k.cube_group[g][3]};
        dq4 s = dqmul(orot, gq); dq4 si = {s.w, -s.x, -s.y, -s.z};
        dq4 r = dqmul(dg, si);
        double mag = 2 * atan2(sqrt(r.x * r.x + r.y * r.y + r.z * r.z), fabs(r.w));
        if (mag < bestd) { bestd = mag; tq = s; }
      }
      tquat[0] = tq.w; tquat[1] = tq.x; tquat[2] = tq.y; tquat[3] = tq.z;
      double gp[3]; ik_site_fk(k, a, q7, gp, nullptr, nullptr);     // site_xpos of the current physics state
      double pre_grasp[3] = {op[0], op[1], op[2] + 0.15}, grasp[3] = {op[0], op[1], op[2] + 0.04};
      auto dist3 = [](const double* x, const double* y) { return sqrt((x[0] - y[0]) * (x[0] - y[0]) + (x[1] - y[1]) * (x[1] - y[1]) + (x[2] - y[2]) * (x[2] - y[2])); };
      bool near_object = dist3(gp, op) < 0.04;
      double pre_release[3] = {k.bucket_pos[a % 2][0], k.bucket_pos[a % 2][1], 1.3};
      double ms[3]; for (int j = 0; j < 3; ++j) ms[j] = (double)st.f[(L::F_IK_MOVE_START + 3 * a + j) * N + e];
      auto set_state = [&](int s) { counter = 0; state = s; };
      auto set_ms = [&]() { for (int j = 0; j < 3; ++j) { ms[j] = gp[j]; st.f[(L::F_IK_MOVE_START + 3 * a + j) * N + e] = (float)gp[j]; } };
      switch (state) {
        case 0: { double s2 = 0; for (int j = 0; j < 7; ++j) { double d = q7[j] - default_pose[j]; s2 += d * d; } if (sqrt(s2) < 0.1) set_state(1); } break;
        case 1: if (dist3(gp, pre_grasp) < 0.05) set_state(2); break;
        case 2: if (near_object) set_state(3); break;
        case 3: if (near_object && counter > grasp_wait) { set_ms(); set_state(4); } else if (!near_object) set_state(0); break;
        case 4: if (!near_object) set_state(0); else if (fabs(gp[2] - (ms[2] + 0.18)) < 0.05) { set_ms(); set_state(5); } break;
        case 5: if (!near_object) set_state(0); else if (dist3(gp, pre_release) < 0.1) set_state(6); break;
        case 6: if (counter > release_wait) set_state(0); break;
      }
      double t = (double)counter / move_steps;
      switch (state) {
        case 0: idle = true; break;
        case 1: for (int j = 0; j < 3; ++j) tpos[j] = pre_grasp[j]; compensate = true; break;
        case 2: for (int j = 0; j < 3; ++j) tpos[j] = grasp[j]; compensate = true; break;
        case 3: close_gripper = true; for (int j = 0; j < 3; ++j) tpos[j] = grasp[j]; compensate = true; break;
        case 4: { close_gripper = true; double end[3] = {ms[0], ms[1], ms[2] + 0.18}; for (int j = 0; j < 3; ++j) tpos[j] = ms[j] + (end[j] - ms[j]) * t; for (int j = 0; j < 4; ++j) tquat[j] = dgq[j]; } break;
        case 5: close_gripper = true; for (int j = 0; j < 3; ++j) tpos[j] = ms[j] + (pre_release[j] - ms[j]) * t; for (int j = 0; j < 4; ++j) tquat[j] = dgq[j]; break;
        case 6: close_gripper = false; for (int j = 0; j < 3; ++j) tpos[j] = pre_release[j]; for (int j = 0; j < 4; ++j) tquat[j] = dgq[j]; break;
      }
      if (!idle) {
        counter++;
        if (counter > timeout_steps) idle = true;
      }
      if (!idle) {
        if (compensate) { tpos[0] += cube_v(target, 0) * pt_comp; tpos[1] += cube_v(target, 1) * pt_comp; }
        bool ok = ik_solve(k, a, tpos, tquat, q7);
        if (!ok) { for (int j = 0; j < 8; ++j) out[j] = (double)st.f[(L::F_IK_LAST_CTRL + 8 * a + j) * N + e]; }
        else {
          for (int j = 0; j < 7; ++j) out[j] = q7[j];
          out[7] = close_gripper ? 0.0 : 2.0;
          for (int j = 0; j < 8; ++j) st.f[(L::F_IK_LAST_CTRL + 8 * a + j) * N + e] = (float)out[j];
        }
        done = true;
      }
    }
    if (idle && !done) {   // idle_ctrl (ik_policy.py:123-127)
      counter = 0; state = 0; target = -1;
      for (int j = 0; j < 8; ++j) { out[j] = default_pose[j]; st.f[(L::F_IK_LAST_CTRL + 8 * a + j) * N + e] = (float)default_pose[j]; }
    }
    st.i[(L::I_FSM_STATE + a) * N + e] = state; st.i[(L::I_FSM_CNT + a) * N + e] = counter; st.i[(L::I_FSM_TGT + a) * N + e] = target;
    for (int j = 0; j < 8; ++j) {
      double v = fmin(fmax(out[j], k.ctrl_lo[j]), k.ctrl_hi[j]);
      st.f[(L::F_IK_ACT + 8 * a + j) * N + e] = (float)v;
      if (obs && obs_dim > 24 * NA + 130) obs[e * obs_dim + 24 * NA + 130 + 8 * a + j] = (float)v;
    }
    for (int o = 0; o < NA; ++o) if (o != a) st.i[(L::I_IGNORE + o * NA + a) * N + e] = target;
  }
}
