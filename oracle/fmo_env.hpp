// fmo_env.hpp — TEST ORACLE: literal restatement of the reference's Python task logic around physics.step().
//   BaseEnv.step_sim / _step_sim_unscaled / reset_sim / _get_state     challenge_env/challenge_env/base_env.py:149-282
//   TaskManager                                                        challenge_env/challenge_env/task_utils.py:47-156
//   IKPolicy (FSM + dm_control qpos_from_site_pose, max_steps=10)      challenge_env/challenge_env/ik_policy.py:92-282
//   FactoryManipulationEnv and its 6 subclasses                        src/environments.py:25-645
// dm_control's IK (third-party, absent here) is restated from its published algorithm [3P] — PARITY UNPINNED.
// Only tests/, smoke() and bench.py's cpu_baseline leg may use this.
#pragma once
#include "fmo_physics.hpp"

namespace fmo {

// numpy PCG64 (XSL-RR 128/64): state = state*MULT + inc, output of the new state.  state4 = {state_hi, state_lo, inc_hi, inc_lo}
static inline uint64_t pcg64_next(uint64_t* s) {
  typedef unsigned __int128 u128;
  const u128 MULT = ((u128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
  u128 st = ((u128)s[0] << 64) | s[1], inc = ((u128)s[2] << 64) | s[3];
  st = st * MULT + inc;
  s[0] = (uint64_t)(st >> 64); s[1] = (uint64_t)st;
  uint64_t x = s[0] ^ s[1]; unsigned rot = (unsigned)(s[0] >> 58);
  return (x >> rot) | (x << ((64 - rot) & 63));
}
static inline double pcg64_double(uint64_t* s) { return (double)(pcg64_next(s) >> 11) * (1.0 / 9007199254740992.0); }

enum EnvClass { ENV_BASE_IK = 0, ENV_SINGLE_FULL = 1, ENV_SINGLE_DELTA = 2, ENV_ALL_FULL = 3, ENV_ALL_DELTA = 4, ENV_PAUSE_TOGGLE = 5, ENV_BACKUP_TOGGLE = 6 };
enum FsmState { IDLE = 0, GO_TO_GRASP, GRASP_APPROACH, GRASP_CLOSE, POST_GRASP, GO_TO_RELEASE, RELEASE };

struct IKPolicyState {
  int state = IDLE, state_counter = 0, target = -1;
  double last_ctrl[8]; double move_start[3] = {0, 0, 0};
  std::vector<int> ignore;          // per owner: -2 = no entry, -1 = None, k = cube
};

struct Env {
  const Model& m; Data& d;
  int env_class, n_arms, n_obj;
  // BaseEnv kwargs (base_env.py:26-34)
  double initial_conveyor_speed, conveyor_acceleration, pt_time, force_threshold, control_frequency, init_spawn_freq, spawn_freq_increase;
  double w_gripper_cube, w_cube_bucket, w_action_norm, base_reward;
  int frame_skip; double dt;
  // state
  std::vector<double> ctrl_target; double conveyor_speed = 0, spawn_freq = 0, play_time = 0;
  std::vector<int> in_scene, out_of_scene; int step_counter = 0, steps_since_spawn = 0, failure_counter = 0, out_counter = 0; int scores[2] = {0, 0};
  uint64_t rng[4];
  std::vector<IKPolicyState> ik;
  int last_score[2] = {0, 0};
  std::vector<double> last_gripper_dist, last_bucket_dist, last_arm_actions, ik_actions;
  bool force_terminate = false;
  bool dirty = false, index_error = false;   // dm_control mjcf.Physics dirty flag (see forward_if_dirty); flagged IndexError case
  // constants of IKPolicy (ik_policy.py:54-72)
  double default_pose[8] = {-0.5, -0.5, 0.0, 1.0, 0.0, -1.6, 0.0, 0.06};
  double default_gripper_quat[4] = {0, 0, 1, 0};
  double pt_compensation; int release_wait, grasp_wait, move_steps, timeout_steps;
  long long ik_calls = 0, ik_success = 0;

  Env(const Model& m_, Data& d_, const int* icfg, const double* dcfg) : m(m_), d(d_) {
    env_class = icfg[0]; n_arms = m.n_arms; n_obj = m.n_cubes;
    for (int i = 0; i < 4; ++i) rng[i] = ((uint64_t)(uint32_t)icfg[1 + 2 * i + 1] << 32) | (uint32_t)icfg[1 + 2 * i];
    initial_conveyor_speed = dcfg[0]; conveyor_acceleration = dcfg[1]; pt_time = dcfg[2]; force_threshold = dcfg[3];
    control_frequency = dcfg[4]; init_spawn_freq = dcfg[5] * n_arms /* base_env.py:36 */; spawn_freq_increase = dcfg[6];
    w_gripper_cube = dcfg[7]; w_cube_bucket = dcfg[8]; w_action_norm = dcfg[9]; base_reward = dcfg[10];
    frame_skip = (int)((1 / control_frequency) / m.h);        // base_env.py:134
    dt = m.h * frame_skip;                                   // base_env.py:201
    pt_compensation = pt_time * dt * 15.0;
    release_wait = (int)(0.5 / dt); grasp_wait = (int)(1.0 / dt); move_steps = (int)(1.0 / dt); timeout_steps = (int)(3.0 / dt);
    ik.resize(n_arms);
    for (auto& p : ik) { p.ignore.assign(n_arms, -2); std::copy(default_pose, default_pose + 8, p.last_ctrl); }
    ctrl_target.assign(m.nu, 0);
    last_gripper_dist.assign(n_arms, 0); last_bucket_dist.assign(n_arms, 0);
    last_arm_actions.assign(8 * n_arms, 0); ik_actions.assign(8 * n_arms, 0);
  }
  bool is_toggle() const { return env_class == ENV_PAUSE_TOGGLE || env_class == ENV_BACKUP_TOGGLE; }
  bool is_progress() const { return env_class >= ENV_SINGLE_FULL && env_class <= ENV_ALL_DELTA; }
  int obs_dim() const { return n_arms * 24 + 13 * n_obj + (is_toggle() ? 8 * n_arms : 0); }
  int act_dim() const { return (env_class == ENV_SINGLE_FULL || env_class == ENV_SINGLE_DELTA) ? 8 : (env_class == ENV_ALL_FULL || env_class == ENV_ALL_DELTA) ? 8 * n_arms : 0; }

  // ---- helpers on the physics state
  double* cube_qpos(int k) { return &d.qpos[m.jnt_qposadr[m.body_jnt[m.cube_body[k]]]]; }
  double* cube_qvel(int k) { return &d.qvel[m.jnt_dofadr[m.body_jnt[m.cube_body[k]]]]; }
  V3 site_xpos(int a) const { int b = m.site_body[a]; return d.xpos[b] + mul(d.xmat[b], {m.site_pos[3 * a], m.site_pos[3 * a + 1], m.site_pos[3 * a + 2]}); }
  V3 bucket_pos(int k) const { return d.gpos[m.bucket_geom[k]]; }
  double lo(int k) const { return m.act_ctrlrange[2 * (1 + k)]; }      // actuator_ctrlrange[1:9]
  double hi(int k) const { return m.act_ctrlrange[2 * (1 + k) + 1]; }

  // ---- TaskManager (task_utils.py)
  void tm_reset() {
    in_scene.clear(); out_of_scene.clear();
    for (int k = 0; k < n_obj; ++k) {
      double* q = cube_qpos(k); double* v = cube_qvel(k);
      q[0] = 4.0; q[1] = 0.0 + k * 0.2; q[2] = 1.0; q[3] = 1; q[4] = q[5] = q[6] = 0;
      for (int i = 0; i < 6; ++i) v[i] = 0;
      out_of_scene.push_back(k);
    }
    step_counter = 0; steps_since_spawn = 0; failure_counter = 0; out_counter = 0; scores[0] = scores[1] = 0;
  }
  void tm_hide(int k) {
    out_of_scene.push_back(k);
    double* q = cube_qpos(k); double* v = cube_qvel(k);
    q[0] = 4.0 + 1.0; q[1] = out_counter * 0.2; q[2] = 1.0; q[3] = 1; q[4] = q[5] = q[6] = 0;
    for (int i = 0; i < 6; ++i) v[i] = 0;
    out_counter++;
    dirty = true;
  }
  void tm_spawn() {
    if (out_of_scene.empty()) return;
    int k = out_of_scene.front(); out_of_scene.erase(out_of_scene.begin());
    double* q = cube_qpos(k);
    q[0] = 0.0; q[1] = 1.0; q[2] = 2.0;
    for (int i = 0; i < 4; ++i) q[3 + i] = pcg64_double(rng);     // raw uniforms, unnormalised (task_utils.py:49-51)
    in_scene.push_back(k);
    dirty = true;
  }
  // dm_control's mjcf.Physics marks itself dirty on every write through a binding (qpos, qvel, ...) and runs mj_forward the next
  // time a DERIVED quantity is read through a binding [3P: dm_control/mjcf/physics.py, Binding.__getattr__ /
  // SynchronizingArrayWrapper].  The only such reads on this path are `physics.bind(bucket).xpos` in _check_states
  // (task_utils.py:101) and in ProgressRewardEnv._get_reward (environments.py:363).  A write that no such read follows leaves
  // the next physics.step() with contacts / Jacobians of the pre-edit pose (mj_step2 runs first).  Pinned by the reference's
  // own free-fall observations: a cube spawned from rest shows exactly 100 n sub-steps of gravity (tests/ref_cases.py).
  void forward_if_dirty() { if (dirty) { step1(m, d); dirty = false; } }
  void tm_check_states() {
    if (in_scene.empty()) return;
    for (int idx = (int)in_scene.size() - 1; idx >= 0; --idx) {
      double* q = cube_qpos(in_scene[idx]);
      bool out = std::fabs(q[0]) > 1.2 || q[1] < -1.5 || q[2] < 0.9;
      if (out) { int k = in_scene[idx]; in_scene.erase(in_scene.begin() + idx); tm_hide(k); failure_counter++; }
    }
    if (in_scene.empty()) return;
    // task_utils.py:97: ONE position snapshot for both buckets; the hit indices of bucket 1 are then applied to the list as
    // bucket 0's removals left it (the reference's index mismatch, reproduced; an index past the end raises IndexError there —
    // here it is skipped and flagged)
    std::vector<double> snap;
    for (int k : in_scene) { double* q = cube_qpos(k); snap.push_back(q[0]); snap.push_back(q[1]); snap.push_back(q[2]); }
    for (int b = 0; b < 2; ++b) {
      forward_if_dirty();                                    // physics.bind(b).xpos
      V3 bp = bucket_pos(b); const double* bs = &m.geom_size[3 * m.bucket_geom[b]];
      std::vector<int> hit;
      for (int idx = 0; idx < (int)snap.size() / 3; ++idx) {
        const double* q = &snap[3 * idx];
        bool in_x = std::fabs(q[0] - bp.x) <= 0.6 * bs[0], in_y = std::fabs(q[1] - bp.y) <= 0.6 * bs[1];
        bool in_z = q[2] - bp.z - bs[2] / 2 <= 0.07;
        if (in_x && in_y && in_z) hit.push_back(idx);
      }
      for (int i = (int)hit.size() - 1; i >= 0; --i) {
        if (hit[i] >= (int)in_scene.size()) { index_error = true; continue; }
        int k = in_scene[hit[i]]; in_scene.erase(in_scene.begin() + hit[i]); tm_hide(k); scores[b]++;
      }
    }
  }
  int spawn_steps() const { return (int)(1.0 / (0.1 * spawn_freq)); }
  void tm_step() {
    if (step_counter == 0 || steps_since_spawn >= spawn_steps()) { tm_spawn(); steps_since_spawn = 0; }
    tm_check_states();
    step_counter++; steps_since_spawn++;
  }

  // ---- BaseEnv
  void get_state_obs(float* obs) {   // _get_state + _process_observation (base part)
    int o = 0;
    for (int a = 0; a < n_arms; ++a) {
      for (int k = 0; k < 8; ++k) obs[o++] = (float)d.qpos[m.jnt_qposadr[m.arm_jnt[9 * a + k]]];
      for (int k = 0; k < 8; ++k) obs[o++] = (float)d.qvel[m.jnt_dofadr[m.arm_jnt[9 * a + k]]];
      for (int k = 0; k < 8; ++k) obs[o++] = (float)ctrl_target[1 + 8 * a + k];
    }
    std::vector<int> order(in_scene.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cube_qpos(in_scene[a])[0] < cube_qpos(in_scene[b])[0]; });
    for (int i = 0; i < n_obj; ++i)
      for (int k = 0; k < 7; ++k) obs[o++] = i < (int)order.size() ? (float)cube_qpos(in_scene[order[i]])[k] : 0.0f;
    for (int i = 0; i < n_obj; ++i)
      for (int k = 0; k < 6; ++k) obs[o++] = i < (int)order.size() ? (float)cube_qvel(in_scene[order[i]])[k] : 0.0f;
  }
  void reset_sim() {
    // physics.reset(): qpos0, zero velocities/controls
    init_data(m, d);
    tm_reset();
    conveyor_speed = initial_conveyor_speed; spawn_freq = init_spawn_freq;
    std::fill(ctrl_target.begin(), ctrl_target.end(), 0.0);
    step1(m, d);                      // after_reset -> mj_forward
    dirty = false;
    play_time = 0;
  }
  bool step_sim_unscaled(const std::vector<double>& ctrl) {
    double alpha = m.h / (m.h + pt_time);
    for (int t = 0; t < frame_skip; ++t) {
      for (int u = 0; u < m.nu; ++u) ctrl_target[u] += (ctrl[u] - ctrl_target[u]) * alpha;
      ctrl_target[0] = -conveyor_speed;
      d.ctrl = ctrl_target;
      step2(m, d);
      if (t == frame_skip - 1) force_check_last_solve();
      step1(m, d);
    }
    return force_terminate;
  }
  // Force termination (base_env.py:226-236).  Declared clarification: evaluated on the contacts
  // and forces of the last sub-step's constraint solve (the reference pairs fresh contacts with stale efc memory).
  void force_check_last_solve() {
    force_terminate = false;
    double mx = 0; bool any = false;
    for (const Contact& c : d.con) {
      if (m.geom_arm[c.g1] < 0 && m.geom_arm[c.g2] < 0) continue;
      any = true;
      const double* f = &d.efc_force[c.efc];
      double fn = f[0] + f[1] + f[2] + f[3], ft1 = c.mu * (f[0] - f[1]), ft2 = c.mu * (f[2] - f[3]);   // [3P] mju_decodePyramid
      mx = std::max(mx, std::max(std::fabs(fn), std::max(std::fabs(ft1), std::fabs(ft2))));
    }
    if (any && mx > force_threshold) force_terminate = true;
  }
  bool step_sim(const std::vector<double>& arm_actions, bool& out_of_reach) {
    std::vector<double> ctrl(m.nu);
    ctrl[0] = conveyor_speed;
    for (int i = 0; i < 8 * n_arms; ++i) ctrl[1 + i] = arm_actions[i];
    for (int u = 0; u < m.nu; ++u) ctrl[u] = std::min(std::max(ctrl[u], m.act_ctrlrange[2 * u]), m.act_ctrlrange[2 * u + 1]);
    bool ft = step_sim_unscaled(ctrl);
    tm_step();
    play_time += dt;
    conveyor_speed += conveyor_acceleration * dt;
    spawn_freq *= spawn_freq_increase;
    out_of_reach = failure_counter > 0;
    return out_of_reach || ft;
  }

  // ---- IK (dm_control.utils.inverse_kinematics.qpos_from_site_pose) [3P]
  // FK of the arm chain to the gripper site for joint vector q7 (plate joints do not move the site).
  void arm_site_fk(int a, const double* q7, V3& spos, Q4& squat, double* J /*6x7 row-major*/) const {
    int chain[16], nchain = 0;
    for (int b = m.site_body[a]; b > 0; b = m.body_parent[b]) chain[nchain++] = b;
    V3 pos = {0, 0, 0}; Q4 quat = {1, 0, 0, 0};
    V3 anchor[7], axis[7]; int nj = 0;
    for (int i = nchain - 1; i >= 0; --i) {
      int b = chain[i];
      pos = pos + mul(q2mat(quat), {m.body_pos[3 * b], m.body_pos[3 * b + 1], m.body_pos[3 * b + 2]});
      quat = qmul(quat, {m.body_quat[4 * b], m.body_quat[4 * b + 1], m.body_quat[4 * b + 2], m.body_quat[4 * b + 3]});
      int j = m.body_jnt[b];
      if (j >= 0) {
        int k = -1; for (int t = 0; t < 7; ++t) if (m.arm_jnt[9 * a + t] == j) k = t;
        V3 ax = {m.jnt_axis[3 * j], m.jnt_axis[3 * j + 1], m.jnt_axis[3 * j + 2]};
        double s = std::sin(q7[k] / 2), c = std::cos(q7[k] / 2);
        quat = qmul(quat, {c, s * ax.x, s * ax.y, s * ax.z});
        anchor[nj] = pos; axis[nj] = mul(q2mat(quat), ax); nj++;
      }
      quat = qnormalize(quat);
    }
    spos = pos + mul(q2mat(quat), {m.site_pos[3 * a], m.site_pos[3 * a + 1], m.site_pos[3 * a + 2]});
    squat = quat;
    if (J) for (int k = 0; k < 7; ++k) {
      V3 jp = cross(axis[k], spos - anchor[k]);
      J[0 * 7 + k] = jp.x; J[1 * 7 + k] = jp.y; J[2 * 7 + k] = jp.z; J[3 * 7 + k] = axis[k].x; J[4 * 7 + k] = axis[k].y; J[5 * 7 + k] = axis[k].z;
    }
  }
  static bool gauss_solve(double* A, double* b, int n) {   // partial pivoting, in place
    for (int c = 0; c < n; ++c) {
      int p = c; for (int r = c + 1; r < n; ++r) if (std::fabs(A[r * n + c]) > std::fabs(A[p * n + c])) p = r;
      if (std::fabs(A[p * n + c]) < 1e-300) return false;
      if (p != c) { for (int k = 0; k < n; ++k) std::swap(A[p * n + k], A[c * n + k]); std::swap(b[p], b[c]); }
      for (int r = c + 1; r < n; ++r) { double f = A[r * n + c] / A[c * n + c]; for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k]; b[r] -= f * b[c]; }
    }
    for (int r = n - 1; r >= 0; --r) { double s = b[r]; for (int k = r + 1; k < n; ++k) s -= A[r * n + k] * b[k]; b[r] = s / A[r * n + r]; }
    return true;
  }
  bool ik_solve(int a, const double* tpos, const double* tquat, double* q7, int* steps_out) {
    const double tol = 1e-14, reg_thresh = 0.1, reg_strength = 3e-2, max_update = 2.0, progress_thresh = 20.0;
    const int max_steps = 10;
    bool success = false; int steps = 0;
    for (steps = 0; steps < max_steps; ++steps) {
      V3 sp; Q4 sq; double J[42];
      arm_site_fk(a, q7, sp, sq, J);
      double err[6];
      err[0] = tpos[0] - sp.x; err[1] = tpos[1] - sp.y; err[2] = tpos[2] - sp.z;
      double err_norm = std::sqrt(err[0] * err[0] + err[1] * err[1] + err[2] * err[2]);
      Q4 eq = qmul({tquat[0], tquat[1], tquat[2], tquat[3]}, {sq.w, -sq.x, -sq.y, -sq.z});
      V3 axv = {eq.x, eq.y, eq.z}; double sn = norm(axv);
      if (sn < MINVAL) axv = {1, 0, 0}; else axv = axv * (1.0 / sn);
      if (sn < MINVAL) sn = 0;
      double speed = 2 * std::atan2(sn, eq.w);
      if (speed > M_PI) speed -= 2 * M_PI;
      err[3] = axv.x * speed; err[4] = axv.y * speed; err[5] = axv.z * speed;
      err_norm += std::sqrt(err[3] * err[3] + err[4] * err[4] + err[5] * err[5]);
      if (err_norm < tol) { success = true; break; }
      double upd[7];
      if (err_norm > reg_thresh) {   // (J^T J + lambda I) x = J^T e
        double H[49], g[7];
        for (int i = 0; i < 7; ++i) { g[i] = 0; for (int r = 0; r < 6; ++r) g[i] += J[r * 7 + i] * err[r]; for (int k = 0; k < 7; ++k) { double s = 0; for (int r = 0; r < 6; ++r) s += J[r * 7 + i] * J[r * 7 + k]; H[i * 7 + k] = s + (i == k ? reg_strength : 0); } }
        gauss_solve(H, g, 7);
        for (int i = 0; i < 7; ++i) upd[i] = g[i];
      } else {                        // lstsq(J^T J, J^T e, rcond=-1)  ==  minimum-norm  J^T (J J^T)^-1 e
        double G[36], y[6];
        for (int r = 0; r < 6; ++r) { y[r] = err[r]; for (int c = 0; c < 6; ++c) { double s = 0; for (int k = 0; k < 7; ++k) s += J[r * 7 + k] * J[c * 7 + k]; G[r * 6 + c] = s; } }
        if (!gauss_solve(G, y, 6)) break;
        for (int i = 0; i < 7; ++i) { upd[i] = 0; for (int r = 0; r < 6; ++r) upd[i] += J[r * 7 + i] * y[r]; }
      }
      double un = 0; for (int i = 0; i < 7; ++i) un += upd[i] * upd[i]; un = std::sqrt(un);
      if (err_norm / un > progress_thresh) break;
      if (un > max_update) for (int i = 0; i < 7; ++i) upd[i] *= max_update / un;
      for (int i = 0; i < 7; ++i) q7[i] += upd[i];
    }
    if (steps_out) *steps_out = steps;
    ik_calls++; ik_success += success;
    return success;
  }

  // ---- IKPolicy (ik_policy.py)
  bool ignored_by(const IKPolicyState& p, int cube) const { for (int v : p.ignore) if (v == cube) return true; return false; }
  void idle_ctrl(IKPolicyState& p, double* out) {
    p.state_counter = 0; p.state = IDLE; p.target = -1;
    std::copy(default_pose, default_pose + 8, p.last_ctrl);
    std::copy(p.last_ctrl, p.last_ctrl + 8, out);
  }
  int select_target(int a) {
    IKPolicyState& p = ik[a];
    std::vector<int> cand;
    for (int k : in_scene) if (!ignored_by(p, k)) cand.push_back(k);
    V3 base = {m.arm_base_pos[3 * a], m.arm_base_pos[3 * a + 1], m.arm_base_pos[3 * a + 2]};
    if (p.target >= 0) {
      double* q = cube_qpos(p.target);
      double dist = norm(v3(q[0], q[1], q[2]) - base);
      if (dist < 1.0 && std::find(cand.begin(), cand.end(), p.target) != cand.end()) return p.target;
    }
    if (cand.empty()) return -1;
    int best = -1; double bd = 1e300;
    for (int k : cand) {
      double* q = cube_qpos(k);
      double dd = norm(v3(q[0], q[1] - 0.2, q[2]) - base);
      if (dd < bd) { bd = dd; best = k; }
    }
    return bd < 0.8 ? best : -1;
  }
  void ik_act(int a, double* out) {
    IKPolicyState& p = ik[a];
    p.target = select_target(a);
    if (p.target < 0) { idle_ctrl(p, out); return; }
    double* oq = cube_qpos(p.target); double* ov = cube_qvel(p.target);
    V3 object_pos = {oq[0], oq[1], oq[2]};
    // nearest cube-symmetric orientation to the default gripper orientation (scipy Rotation, ik_policy.py:157-162)
    Q4 orot = qnormalize({oq[3], oq[4], oq[5], oq[6]});
    Q4 dg = {default_gripper_quat[0], default_gripper_quat[1], default_gripper_quat[2], default_gripper_quat[3]};
    double bestd = 1e300; Q4 tq = {1, 0, 0, 0};
    for (int g = 0; g < 24; ++g) {
      Q4 s = qmul(orot, {m.cube_group[4 * g], m.cube_group[4 * g + 1], m.cube_group[4 * g + 2], m.cube_group[4 * g + 3]});
      Q4 r = qmul(dg, {s.w, -s.x, -s.y, -s.z});
      double mag = 2 * std::atan2(std::sqrt(r.x * r.x + r.y * r.y + r.z * r.z), std::fabs(r.w));
      if (mag < bestd) { bestd = mag; tq = s; }
    }
    double target_quat[4] = {tq.w, tq.x, tq.y, tq.z};
    V3 gp = site_xpos(a);
    V3 pre_grasp = object_pos; pre_grasp.z += 0.15;
    V3 grasp = object_pos; grasp.z += 0.04;
    bool near_object = norm(gp - object_pos) < 0.04;
    V3 pre_release = {bucket_pos(a % 2).x, bucket_pos(a % 2).y, 1.3};
    auto set_state = [&](int s) { p.state_counter = 0; p.state = s; };
    switch (p.state) {
      case IDLE: {
        double s = 0;
        for (int k = 0; k < 7; ++k) { double e = d.qpos[m.jnt_qposadr[m.arm_jnt[9 * a + k]]] - default_pose[k]; s += e * e; }
        if (std::sqrt(s) < 0.1) set_state(GO_TO_GRASP);
      } break;
      case GO_TO_GRASP: if (norm(gp - pre_grasp) < 0.05) set_state(GRASP_APPROACH); break;
      case GRASP_APPROACH: if (near_object) set_state(GRASP_CLOSE); break;
      case GRASP_CLOSE:
        if (near_object && p.state_counter > grasp_wait) { p.move_start[0] = gp.x; p.move_start[1] = gp.y; p.move_start[2] = gp.z; set_state(POST_GRASP); }
        else if (!near_object) set_state(IDLE);
        break;
      case POST_GRASP:
        if (!near_object) set_state(IDLE);
        else if (std::fabs(gp.z - (p.move_start[2] + 0.18)) < 0.05) { p.move_start[0] = gp.x; p.move_start[1] = gp.y; p.move_start[2] = gp.z; set_state(GO_TO_RELEASE); }
        break;
      case GO_TO_RELEASE:
        if (!near_object) set_state(IDLE);
        else if (norm(gp - pre_release) < 0.1) set_state(RELEASE);
        break;
      case RELEASE: if (p.state_counter > release_wait) set_state(IDLE); break;
    }
    bool close_gripper = false, compensate = false;
    V3 target_pos = {0, 0, 0};
    double t = (double)p.state_counter / move_steps;
    V3 ms = {p.move_start[0], p.move_start[1], p.move_start[2]};
    switch (p.state) {
      case IDLE: idle_ctrl(p, out); return;
      case GO_TO_GRASP: target_pos = pre_grasp; compensate = true; break;
      case GRASP_APPROACH: target_pos = grasp; compensate = true; break;
      case GRASP_CLOSE: close_gripper = true; target_pos = grasp; compensate = true; break;
      case POST_GRASP: { close_gripper = true; V3 end = ms; end.z += 0.18; target_pos = ms + (end - ms) * t; std::copy(default_gripper_quat, default_gripper_quat + 4, target_quat); } break;
      case GO_TO_RELEASE: close_gripper = true; target_pos = ms + (pre_release - ms) * t; std::copy(default_gripper_quat, default_gripper_quat + 4, target_quat); break;
      case RELEASE: close_gripper = false; target_pos = pre_release; std::copy(default_gripper_quat, default_gripper_quat + 4, target_quat); break;
    }
    p.state_counter++;
    if (p.state_counter > timeout_steps) { idle_ctrl(p, out); return; }
    if (compensate) { target_pos.x += ov[0] * pt_compensation; target_pos.y += ov[1] * pt_compensation; }
    double q7[7];
    for (int k = 0; k < 7; ++k) q7[k] = d.qpos[m.jnt_qposadr[m.arm_jnt[9 * a + k]]];
    double tp[3] = {target_pos.x, target_pos.y, target_pos.z};
    bool ok = ik_solve(a, tp, target_quat, q7, nullptr);
    if (!ok) { std::copy(p.last_ctrl, p.last_ctrl + 8, out); return; }
    for (int k = 0; k < 7; ++k) out[k] = q7[k];
    out[7] = close_gripper ? 0.0 : 2.0;
    std::copy(out, out + 8, p.last_ctrl);
  }
  // FactoryManipulationEnv._compose_control (environments.py:104-127): sequential arms + mutual ignore
  void base_compose_control(std::vector<double>& acts) {
    acts.assign(8 * n_arms, 0);
    for (int i = 0; i < n_arms; ++i) {
      double c[8]; ik_act(i, c);
      for (int k = 0; k < 8; ++k) acts[8 * i + k] = std::min(std::max(c[k], lo(k)), hi(k));
      for (int j = 0; j < n_arms; ++j) if (j != i) ik[j].ignore[i] = ik[i].target;
    }
  }
  void process_action(const float* a, double* out) {   // environments.py:84-102
    // np.tanh of a float32 array stays float32, as does (a + 1.0) * 0.5; the product with (high - low) promotes to float64
    for (int k = 0; k < 8; ++k) { float t = std::tanh((float)a[k]); float u = (t + 1.0f) * 0.5f; out[k] = lo(k) + (double)u * (hi(k) - lo(k)); }
  }
  void compose_control(const float* af, const int* ai, std::vector<double>& acts) {
    double pa[8];
    switch (env_class) {
      case ENV_BASE_IK: base_compose_control(acts); break;
      case ENV_SINGLE_FULL: base_compose_control(acts); process_action(af, pa); for (int k = 0; k < 8; ++k) acts[k] = pa[k]; break;
      case ENV_SINGLE_DELTA: base_compose_control(acts); if (ik[0].state != IDLE) { process_action(af, pa); for (int k = 0; k < 8; ++k) acts[k] += 0.5 * pa[k]; } break;
      case ENV_ALL_FULL: acts.assign(8 * n_arms, 0); for (int i = 0; i < n_arms; ++i) process_action(af + 8 * i, &acts[8 * i]); break;
      case ENV_ALL_DELTA:
        base_compose_control(acts);
        for (int i = 0; i < n_arms; ++i) if (ik[i].state != IDLE) { process_action(af + 8 * i, pa); for (int k = 0; k < 8; ++k) acts[8 * i + k] += 0.5 * pa[k]; }
        break;
      case ENV_PAUSE_TOGGLE:
        acts.assign(8 * n_arms, 0);
        for (int i = 0; i < n_arms; ++i) for (int k = 0; k < 8; ++k) acts[8 * i + k] = ai[i] == 1 ? ik_actions[8 * i + k] : last_arm_actions[8 * i + k];
        last_arm_actions = acts;
        break;
      case ENV_BACKUP_TOGGLE:
        acts.assign(8 * n_arms, 0);
        for (int i = 0; i < n_arms; ++i) for (int k = 0; k < 8; ++k) acts[8 * i + k] = ai[i] == 1 ? ik_actions[8 * i + k] : default_pose[k];
        break;
    }
  }
  void process_observation(float* obs) {
    get_state_obs(obs);
    if (is_toggle()) {
      base_compose_control(ik_actions);   // IK proposals for the NEXT step (environments.py:574-576)
      int o = n_arms * 24 + 13 * n_obj;
      for (int i = 0; i < 8 * n_arms; ++i) obs[o + i] = (float)ik_actions[i];
    }
  }
  double get_reward(const float* af) {
    double score_reward = (double)((scores[0] + scores[1]) - (last_score[0] + last_score[1]));
    if (!is_progress()) return score_reward;
    forward_if_dirty();   // environments.py:363 reads physics.bind(bucket).xpos every step
    double gripper_cube = 0, bucket_cube = 0;
    std::vector<int> closest(n_arms, -1);
    for (int i = 0; i < n_arms; ++i) {
      V3 gp = site_xpos(i);
      int best = -1; double bd = 1e300;
      for (int k : in_scene) { if (ignored_by(ik[i], k)) continue; double* q = cube_qpos(k); double dd = norm(v3(q[0], q[1], q[2]) - gp); if (dd < bd) { bd = dd; best = k; } }
      if (best >= 0) { gripper_cube += last_gripper_dist[i] - bd; last_gripper_dist[i] = bd; }
      closest[i] = best;
    }
    for (int i = 0; i < n_arms; ++i) {
      if (closest[i] < 0) continue;
      double* q = cube_qpos(closest[i]);
      double dd = norm(v3(q[0], q[1], q[2]) - bucket_pos(i % 2));
      bucket_cube += last_bucket_dist[i] - dd; last_bucket_dist[i] = dd;
    }
    double an = 0; int ad = act_dim();
    for (int i = 0; i < ad; ++i) if (i % 8 != 7) an += (double)af[i] * (double)af[i];
    double action_norm_reward = std::exp(-std::sqrt(an));
    double progress = base_reward + w_gripper_cube * gripper_cube + w_cube_bucket * bucket_cube + w_action_norm * action_norm_reward;
    return score_reward > 0 ? score_reward : progress;
  }

  // ---- gym API
  void reset(float* obs) {
    reset_sim();
    for (auto& p : ik) { double tmp[8]; idle_ctrl(p, tmp); }
    process_observation(obs);
    last_score[0] = scores[0]; last_score[1] = scores[1];
  }
  // info: terminated, out_of_reach, force_terminate, score0, score1, num_obj
  void step(const float* af, const int* ai, float* obs, double* reward, int* info) {
    std::vector<double> acts;
    compose_control(af, ai, acts);
    bool oor = false;
    bool term = step_sim(acts, oor);
    *reward = get_reward(af);
    process_observation(obs);
    last_score[0] = scores[0]; last_score[1] = scores[1];
    info[0] = term; info[1] = oor; info[2] = force_terminate; info[3] = scores[0]; info[4] = scores[1]; info[5] = (int)in_scene.size();
  }
  void get(const char* what, double* out) {
    std::string w(what);
    if (w == "ctrl_target") std::copy(ctrl_target.begin(), ctrl_target.end(), out);
    else if (w == "scalars") { out[0] = conveyor_speed; out[1] = spawn_freq; out[2] = play_time; out[3] = step_counter; out[4] = steps_since_spawn; out[5] = failure_counter; out[6] = out_counter; out[7] = (double)ik_calls; out[8] = (double)ik_success; }
    else if (w == "lists") { out[0] = (double)in_scene.size(); for (size_t i = 0; i < in_scene.size(); ++i) out[1 + i] = in_scene[i]; out[11] = (double)out_of_scene.size(); for (size_t i = 0; i < out_of_scene.size(); ++i) out[12 + i] = out_of_scene[i]; }
    else if (w == "fsm") { for (int a = 0; a < n_arms; ++a) { out[3 * a] = ik[a].state; out[3 * a + 1] = ik[a].state_counter; out[3 * a + 2] = ik[a].target; } }
    else if (w == "ik_actions") std::copy(ik_actions.begin(), ik_actions.end(), out);
    else if (w == "site") { for (int a = 0; a < n_arms; ++a) { V3 s = site_xpos(a); out[3 * a] = s.x; out[3 * a + 1] = s.y; out[3 * a + 2] = s.z; } }
  }
  void set(const char* what, const double* in) {
    std::string w(what);
    if (w == "ctrl_target") std::copy(in, in + m.nu, ctrl_target.begin());
    else if (w == "scalars") { conveyor_speed = in[0]; spawn_freq = in[1]; play_time = in[2]; step_counter = (int)in[3]; steps_since_spawn = (int)in[4]; failure_counter = (int)in[5]; out_counter = (int)in[6]; }
    else if (w == "lists") { in_scene.clear(); out_of_scene.clear(); int n = (int)in[0]; for (int i = 0; i < n; ++i) in_scene.push_back((int)in[1 + i]); n = (int)in[11]; for (int i = 0; i < n; ++i) out_of_scene.push_back((int)in[12 + i]); }
  }
};

}  // namespace fmo
