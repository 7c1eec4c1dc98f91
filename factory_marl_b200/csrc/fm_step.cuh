// fm_step.cuh — the fused "env step" kernel body: 100 x { mj_step-equivalent } + task logic, one CTA per E envs.
//
// Replaces, for the scene of scene.py:109-169, what the reference computes in
//   BaseEnv._step_sim_unscaled   challenge_env/challenge_env/base_env.py:204-238  (PT1 filter + 100 x physics.step())
//   physics.step()               base_env.py:218  (MuJoCo mj_step2 + mj_step1, third party)
// Design (B200): one env's whole state lives in shared memory for the 100 sub-steps (HBM touched once per
// env-step); the CTA runs a sequence of barrier-separated *phases*, each phase mapping threads to the natural
// work items of that stage (arm chains, bodies, mass-matrix entries, cluster pairs, contacts, constraint islands),
// so serial per-arm recursions of all E envs pack into one warp while wide stages use the whole CTA.
// The constraint problem is solved by projected Gauss-Seidel on the dual in contact space (3x3 Delassus blocks,
// exact GS inside a constraint island, islands in parallel, belt coupling lagged), warm-started from qacc.
//
// Every phase is a plain function of (block, tid) with no intra-phase cross-thread dependency, so the same source
// also compiles with g++ for host-side logic tests (tests/kernel_emu); the product only ever launches the CUDA build.
#pragma once
#include "fm_types.h"
#include "fm_math.cuh"
#include "fm_collide.cuh"
#include <string.h>

#if defined(__CUDA_ARCH__)
#define FM_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define FM_ATOMIC_ADDF(p, v) atomicAdd((p), (v))
#define FM_ATOMIC_OR(p, v) atomicOr((p), (v))
#define FM_ATOMIC_MAXI(p, v) atomicMax((p), (v))
#else
template <class T> static inline T fm_host_add(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline int fm_host_or(int* p, int v) { int o = *p; *p = o | v; return o; }
static inline int fm_host_max(int* p, int v) { int o = *p; *p = o > v ? o : v; return o; }
#define FM_ATOMIC_ADD(p, v) fm_host_add((p), (v))
#define FM_ATOMIC_ADDF(p, v) fm_host_add((p), (v))
#define FM_ATOMIC_OR(p, v) fm_host_or((p), (v))
#define FM_ATOMIC_MAXI(p, v) fm_host_max((p), (v))
#endif

#if !defined(__CUDACC__)
static long long fm_dbg_sweeps = 0, fm_dbg_visits = 0, fm_dbg_islands = 0;   // host logic-test build only
static long long fm_dbg_cnt[32] = {0};
static long long fm_dbg_sat[512] = {0};   // general SAT calls by geom1
#define FM_DBG(x) x
#else
#define FM_DBG(x)
#endif
#if defined(__CUDA_ARCH__)
#define FM_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#define FM_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#else
#define FM_ASSUME_SHARED(p)
#define FM_ASSUME_GLOBAL(p)
#endif
#if defined(__CUDA_ARCH__)
#define FM_FDIV(a, b) __fdividef((a), (b))
#else
#define FM_FDIV(a, b) ((a) / (b))
#endif
#define FM_AROW 10   // equality + up to 9 limit rows per arm
// Broad-phase caching: the survivor queue of the 737 (two arms) statically admissible cluster pairs is rebuilt only when some
// cluster centre has moved more than FM_BROAD_MARGIN / 2 since the last rebuild (or a cube fell asleep / woke up / was moved by the
// task manager); the pair tests use radii inflated by FM_BROAD_MARGIN, so the cached queue is a superset of the near pairs of every
// sub-step in between and the narrow phase — which runs every sub-step — produces bit-identical contact sets.
#ifndef FM_BROAD_MARGIN
#define FM_BROAD_MARGIN 0.02f
#endif
// Box-box contacts are twice as soft in the reference as MuJoCo's documented regulariser (R = 2 mu^2 (1-d)/d diagApprox, four
// pyramid rows on each of the four face-clipping points) predicts: every resting cube in runs/*.zip:_last_obs (90 samples, belt
// and table) sits 2.28e-6 m (belt) / 4.2e-7 m (table) deep, exactly 2x the documented 1.14e-6 / 2.04e-7, independent of mass.
// Which mjc_BoxBox / mj_makeImpedance detail causes it cannot be read here (MuJoCo is absent), so the factor is CALIBRATED to
// that reference-held data (tests/test_ref_fixtures.py) and applied to box-box pairs only, in the oracle and here alike.
#define FM_BOXBOX_RSCALE 2.0f

// packed lower-triangular index
FM_HD int tri(int i, int j) { return i * (i + 1) / 2 + j; }

#ifndef FM_ENVS_PAD
#define FM_ENVS_PAD 1
#endif
#define FM_ENVS_PAD6 1
template <int NA>
struct EnvS {
  static const int NV = 61 + 9 * NA, NQ = 71 + 9 * NA, NU = 1 + 8 * NA, NNODE = FM_NCUBE + NA;
  // ---- state (persisted in HBM between env steps)
  float belt_q, belt_v;
  float cube_q[FM_NCUBE][7], cube_v[FM_NCUBE][6];
  float cube_qlo[FM_NCUBE][3];   // low-order part of the cube positions (compensated integration, see ph_integrate)
  float arm_q[NA][9], arm_v[NA][9];
  float ctrl_target[NU], ctrl[NU];
  float conveyor_speed;
  // ---- position-dependent (mj_step1)
  int cube_sleep[FM_NCUBE];
  float cube_R[FM_NCUBE][9], cube_x[FM_NCUBE][3], cube_vw[FM_NCUBE][6];   // cube_vw: (v, R*omega_local)
  float body_x[NA][FM_NARMBODY][3], body_R[NA][FM_NARMBODY][9];
  alignas(8) float bnd[FM_NCUBE + NA * FM_NARMBODY][4];   // bounding spheres (x, y, z, r) of the cube / arm-body clusters
  float M[NA][45], Minv[NA][81], Li[NA][45], Lc[NA][45];
  float bias[NA][9], chol_d[NA][2][9];      // chol_d: diagonal of the Cholesky factors of M (0) and M - hD (1)
  // ---- acceleration stage
  float acc[NV];
  float belt_acc_smooth, belt_S;
  // ---- arm-local rows (joint equality gripper.xml:46-49, joint limits)
  int arow_n[NA], arow_code[NA][FM_AROW];   // code: dof | sign<<8 (sign bit 1 = negative) ; 0x1000 = equality
  float arow_aref[NA][FM_AROW], arow_R[NA][FM_AROW], arow_f[NA][FM_AROW];
  // ---- contacts
  int ncon, nraw, broad_stale, flags, nqueue;   // broad_stale: the cached broad-phase survivor queue must be rebuilt (ph_broad)
  int c_side[FmCap<NA>::CON];
  float c_p[FmCap<NA>::CON][3], c_n[FmCap<NA>::CON][3], c_t1[FmCap<NA>::CON][3], c_mu[FmCap<NA>::CON], c_R[FmCap<NA>::CON], c_W[FmCap<NA>::CON][6];
  float c_aref[FmCap<NA>::CON][4], c_f[FmCap<NA>::CON][4];
  // rarely / once-per-sub-step touched buffers live in a library-owned global scratch row of this env (L2-resident; see
  // FmCap<NA>::WS): arm-side Jacobians (3x9) and Minv J^T of contacts that touch an arm, and the previous solve's contact forces
  float (*as_J)[27]; float (*as_B)[27]; float (*prev_f)[4];
  float* ct_d0;   // ctrl_target - ctrl at the start of the control step (closed-form PT1 filter, ph_ctrl); followed by arm_qlo
  // low-order part of the arm joint positions (compensated integration like cube_qlo; lives in the global scratch row)
  FM_HDM float* arm_qlo(int a) const { return ct_d0 + (1 + 8 * FM_MAX_ARMS) + 9 * a; }
  // bounding spheres (x, y, z, r) of the clusters at the last rebuild of the broad-phase queue (ph_broad); global scratch row
  FM_HDM float* bnd_ref(int cl) const { return ct_d0 + (1 + 8 * FM_MAX_ARMS) + 9 * FM_MAX_ARMS + 4 * cl; }
  int isl_root[NNODE], isl_start[NNODE + 1], need_union;
  unsigned char isl_list[FmCap<NA>::CON], c_node[FmCap<NA>::CON];
  unsigned short queue[FmCap<NA>::QUEUE];
  // ---- per-body scratch of the arm recursions, overlaid with the raw contact buffer.  ph_narrow writes the raw contacts while
  //      ph_arm_M still runs on other warps (fm_derive), so they may only cover what is dead after ph_arm_bwd: the FK scan buffer,
  //      the per-body inertias / wrenches (d_comp0, d_f0, d_n0) and d_om; d_f, d_n (accumulated wrenches) and d_comp (composite
  //      inertias) are read by ph_arm_M and lie behind the raw buffer.  The raw buffer is dead after ph_contact_setup.
  union {
    struct {
      float d_unused[NA][FM_NARMBODY][3];
      union {
        struct { float d_comp0[NA][FM_NARMBODY][10], d_f0[NA][FM_NARMBODY][3], d_n0[NA][FM_NARMBODY][3]; };   // own inertia / wrench (ph_arm_body)
        struct { float d_L[NA][FM_NARMBODY][9], d_p[NA][FM_NARMBODY][3]; };           // FK scan ping-pong buffer (dead after the scan)
      };
      union {
        struct { float d_om[NA][FM_NARMBODY][3], d_pad0[20], d_al[NA][FM_NARMBODY][3], d_ap[NA][FM_NARMBODY][3]; };   // until ph_arm_body
        struct { float d_gap[NA][FM_NARMBODY][3], d_pad1[20], d_f[NA][FM_NARMBODY][3], d_n[NA][FM_NARMBODY][3]; };    // accumulated wrench (ph_arm_bwd)
      };
      float d_comp[NA][FM_NARMBODY][10];                                               // composite inertia (ph_arm_bwd)
    };
    struct {
      int raw_key[FmCap<NA>::CON], raw_g[FmCap<NA>::CON];
      float raw_d[FmCap<NA>::CON][7];
    };
  };
  static_assert(sizeof(int) * 2 * FmCap<NA>::CON + sizeof(float) * 7 * FmCap<NA>::CON <= sizeof(float) * (NA * FM_NARMBODY * (3 + 16 + 3) + 20),
                "the raw contact buffer must end before d_f / d_n / d_comp");
  float force_max;
  // forces of the previous solve, keyed by contact identity (warm start of the dual; persisted in HBM between env steps)
  int prev_n, prev_key[FmCap<NA>::CON];
  int prev_arow_n[NA], prev_arow_code[NA][FM_AROW];
  float prev_arow_f[NA][FM_AROW];
  int c_key[FmCap<NA>::CON];
  int pad_[NA == 6 ? FM_ENVS_PAD6 : FM_ENVS_PAD];   // sizeof(EnvS) / 4 = 2 (mod 32): the envs of a block start on different (even) shared-memory banks,
                           // and the next item of an odd-stride array falls on the odd ones
};
// the envs of a block must start on distinct shared-memory banks: stride = 2 (mod 4) words (checked for both scenes)
static_assert((sizeof(EnvS<2>) / 4) % 4 == 2 && (sizeof(EnvS<4>) / 4) % 4 == 2 && (sizeof(EnvS<6>) / 4) % 4 == 2, "re-pad EnvS (pad_) after changing its members");

// flags
#define FM_FLAG_CON_OVERFLOW 1
#define FM_FLAG_ARMSIDE_OVERFLOW 2
#define FM_FLAG_QUEUE_OVERFLOW 4
#define FM_FLAG_NAN 8

// side descriptor packing: type (2 bits) | idx (7 bits: cube k or arm*16+body) | slot (6 bits) per side = 15 bits; two sides per
// contact word, bit 30 = an arm geom is involved
FM_HD int pack_side(int type, int idx, int slot) { return type | (idx << 2) | (slot << 9); }
FM_HD int side_type(int s) { return s & 3; }
FM_HD int side_idx(int s) { return (s >> 2) & 127; }
FM_HD int side_slot(int s) { return (s >> 9) & 63; }
#define FM_SIDE_BITS 15
#define FM_SIDE_MASK 32767
FM_HD int contact_node(int cs, int q) {   // island node of side q of a packed contact: cube k -> k, arm a -> 10 + a, else -1
  int sd = (cs >> (FM_SIDE_BITS * q)) & FM_SIDE_MASK;
  int ty = side_type(sd), ix = side_idx(sd);
  return ty == FM_SIDE_CUBE ? ix : (ty == FM_SIDE_ARM ? FM_NCUBE + (ix >> 4) : -1);
}

// parent of template body b (0 base .. 10 right plate)
FM_HD int ab_parent(int b) { return b <= 8 ? b - 1 : 8; }
// dof index (0..8) of template body b, or -1
FM_HD int ab_dof(int b) { return (b >= 1 && b <= 7) ? b - 1 : (b == 9 ? 7 : (b == 10 ? 8 : -1)); }
FM_HD int dof_body(int d) { return d < 7 ? d + 1 : d + 2; }

// the cluster's bounding sphere against its value at the last queue rebuild
template <int NA>
FM_HD void broad_check(EnvS<NA>& s, int cl, const float* bb) {
  const float* rf = s.bnd_ref(cl);
  float dx = bb[0] - rf[0], dy = bb[1] - rf[1], dz = bb[2] - rf[2];
  const float thr = 0.5f * FM_BROAD_MARGIN;
  if (dx * dx + dy * dy + dz * dz > thr * thr || (bb[3] < 0.f) != (rf[3] < 0.f)) FM_ATOMIC_OR(&s.broad_stale, 1);
}

// =====================================================================================================================
// phase: cube kinematics (normalise quats in place as mj_kinematics does [3P])
template <int NA, int E>
FM_PH void ph_cube_derive(EnvS<NA>* S, const FmDevModel& m, int tid, int T, int sleep_hidden) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < FM_NCUBE * E; w += T) {
    EnvS<NA>& s = S[w % E]; int k = w / E;
    q4 q = qnormf(mkq(s.cube_q[k][3], s.cube_q[k][4], s.cube_q[k][5], s.cube_q[k][6]));
    s.cube_q[k][3] = q.w; s.cube_q[k][4] = q.x; s.cube_q[k][5] = q.y; s.cube_q[k][6] = q.z;
    q2matf(q, s.cube_R[k]);
    s.cube_x[k][0] = s.cube_q[k][0]; s.cube_x[k][1] = s.cube_q[k][1]; s.cube_x[k][2] = s.cube_q[k][2];
    // Hidden cubes (parked beyond the table at x = 4 / 5, task_utils.py:31,119) that have come to rest on the floor are put
    // to sleep when cfg.sleep_hidden: no collision, no constraint rows, no integration, until the task manager moves them.
    // They are invisible to observation and reward except through the velocity a cube keeps when it is spawned
    // (SURVEY.md A.3), which this preserves to < 1e-5 m/s.
    {
      const float* vv = s.cube_v[k];
      float v2 = vv[0] * vv[0] + vv[1] * vv[1] + vv[2] * vv[2], w2 = vv[3] * vv[3] + vv[4] * vv[4] + vv[5] * vv[5];
      float hh = m.cube_half[k];
      s.cube_sleep[k] = (sleep_hidden && s.cube_q[k][0] > 3.0f && s.cube_q[k][2] < hh * 1.01f && v2 < 1e-10f && w2 < 1e-7f) ? 1 : 0;
    }
    s.bnd[k][0] = s.cube_q[k][0]; s.bnd[k][1] = s.cube_q[k][1]; s.bnd[k][2] = s.cube_q[k][2];
    s.bnd[k][3] = s.cube_sleep[k] ? -1.f : m.cube_half[k] * 1.7320509f * 1.000001f;
    broad_check(s, k, s.bnd[k]);
    f3 ww = rmul(s.cube_R[k], mk3(s.cube_v[k][3], s.cube_v[k][4], s.cube_v[k][5]));
    s.cube_vw[k][0] = s.cube_v[k][0]; s.cube_vw[k][1] = s.cube_v[k][1]; s.cube_vw[k][2] = s.cube_v[k][2];
    s.cube_vw[k][3] = ww.x; s.cube_vw[k][4] = ww.y; s.cube_vw[k][5] = ww.z;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Arm kinematics + Newton-Euler forward pass, lane-parallel (work item = (env, arm, body)):
//   local transforms -> prefix products by a 3-round scan over the 8-body chain (+ the two plates) -> angular velocity,
//   angular and linear bias accelerations as prefix sums of per-link terms.  World-aligned axes, link-origin reference
//   points.  Every function below is one barrier interval (cheap with one warp per env).
template <int NA>
FM_HD void xf_compose(const float* Ra, const float* xa, const float* Rb, const float* xb, float* Ro, float* xo) {
  float R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = Ra[3 * i] * Rb[j] + Ra[3 * i + 1] * Rb[3 + j] + Ra[3 * i + 2] * Rb[6 + j];
  f3 x = ld3(xa) + rmul(Ra, ld3(xb));
  for (int i = 0; i < 9; ++i) Ro[i] = R[i];
  st3(xo, x);
}
template <int NA, int E>
FM_PH void ph_arm_local(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < NA * FM_NARMBODY * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ab = w / E; int a = ab / FM_NARMBODY, b = ab % FM_NARMBODY;
    if (b == 0) {   // static base body
      for (int i = 0; i < 9; ++i) s.body_R[a][0][i] = m.arm_base_R[a][i];
      st3(s.body_x[a][0], ld3(m.arm_base_pos[a]));
      continue;
    }
    float R[9];
    for (int i = 0; i < 9; ++i) R[i] = m.ab_C[b][i];
    f3 p = ld3(m.ab_pos[b]);
    int d = ab_dof(b);
    if (m.ab_jtype[b] == 1) {
      float sn, cs; sincosf(s.arm_q[a][d], &sn, &cs);
      for (int i = 0; i < 3; ++i) { float c0 = R[3 * i], c1 = R[3 * i + 1]; R[3 * i] = cs * c0 + sn * c1; R[3 * i + 1] = cs * c1 - sn * c0; }
    } else if (m.ab_jtype[b] == 2) p = p + rcol(R, 0) * s.arm_q[a][d];
    if (b == 1) {   // fold the base frame into the first link
      float xb[3]; st3(xb, p);
      xf_compose<NA>(m.arm_base_R[a], m.arm_base_pos[a], R, xb, s.d_L[a][1], s.d_p[a][1]);
    } else { for (int i = 0; i < 9; ++i) s.d_L[a][b][i] = R[i]; st3(s.d_p[a][b], p); }
  }
}
// one scan round: dst_b = (b - dist >= 1 ? src_{b-dist} o src_b : src_b) for chain bodies 1..8; plates are carried through
template <int NA, int E>
FM_PH void ph_arm_scan(EnvS<NA>* S, int tid, int T, int dist, bool to_body) {
  FM_ASSUME_SHARED(S);
  for (int w = tid; w < NA * (FM_NARMBODY - 1) * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ab = w / E; int a = ab / (FM_NARMBODY - 1), b = 1 + ab % (FM_NARMBODY - 1);
    const float *sR = to_body ? s.d_L[a][b] : s.body_R[a][b], *sx = to_body ? s.d_p[a][b] : s.body_x[a][b];
    float *dR = to_body ? s.body_R[a][b] : s.d_L[a][b], *dx = to_body ? s.body_x[a][b] : s.d_p[a][b];
    if (b <= 8 && b - dist >= 1) {
      const float *pR = to_body ? s.d_L[a][b - dist] : s.body_R[a][b - dist], *px = to_body ? s.d_p[a][b - dist] : s.body_x[a][b - dist];
      xf_compose<NA>(pR, px, sR, sx, dR, dx);
    } else { for (int i = 0; i < 9; ++i) dR[i] = sR[i]; for (int i = 0; i < 3; ++i) dx[i] = sx[i]; }
  }
}
// plates: world frame = gripper_base frame o local; reads the plates' local transforms from d_L and writes body_R/body_x.
// After the rounds (dist 1 -> body, 2 -> d_L, 4 -> body) bodies 1..8 are final in body_R and the plates' locals sit in body_R
// as well (carried), so they are first copied back by the caller's last round; here they are composed in place via registers.
template <int NA, int E>
FM_PH void ph_arm_plates(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  for (int w = tid; w < NA * 2 * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ap = w / E; int a = ap / 2, b = 9 + ap % 2;
    float R[9], x[3];
    xf_compose<NA>(s.body_R[a][8], s.body_x[a][8], s.body_R[a][b], s.body_x[a][b], R, x);
    for (int i = 0; i < 9; ++i) s.body_R[a][b][i] = R[i];
    for (int i = 0; i < 3; ++i) s.body_x[a][b][i] = x[i];
  }
}
// chain membership: dof-carrying ancestors-or-self of body b are bodies 1..min(b,7) plus b itself for a plate
FM_HD int chain_hi(int b) { return b < 7 ? b : 7; }
// Newton-Euler forward pass in ONE interval: lane = (env, arm, body) walks its own ancestor chain with running angular
// velocity, angular bias acceleration and origin bias acceleration (no cross-lane dependency; <= 8 links).
template <int NA, int E>
FM_PH void ph_arm_vel(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < NA * FM_NARMBODY * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ab = w / E; int a = ab / FM_NARMBODY, b = ab % FM_NARMBODY;
    f3 om = mk3(0, 0, 0), al = mk3(0, 0, 0), ap = mk3(0, 0, m.g);
    int hi = b <= 8 ? b : 8;
    f3 xp = ld3(s.body_x[a][0]);
    for (int i = 1; i <= hi; ++i) {
      f3 xi = ld3(s.body_x[a][i]);
      f3 r = xi - xp; xp = xi;
      ap = ap + cross3(al, r) + cross3(om, cross3(om, r));
      if (i <= 7) { f3 t = rcol(s.body_R[a][i], 2) * s.arm_v[a][i - 1]; al = al + cross3(om, t); om = om + t; }
    }
    if (b >= 9) {
      f3 r = ld3(s.body_x[a][b]) - xp;
      ap = ap + cross3(al, r) + cross3(om, cross3(om, r)) + 2.0f * cross3(om, rcol(s.body_R[a][b], 0) * s.arm_v[a][b - 2]);
    }
    st3(s.d_om[a][b], om); st3(s.d_al[a][b], al); st3(s.d_ap[a][b], ap);
    // bounding sphere of the body's geom cluster (broad phase) and its displacement since the last queue rebuild
    { f3 cc = ld3(s.body_x[a][b]) + rmul(s.body_R[a][b], ld3(m.arm_cl[b])); float* bb = s.bnd[FM_NCUBE + ab]; bb[0] = cc.x; bb[1] = cc.y; bb[2] = cc.z; bb[3] = m.arm_cl[b][3]; broad_check(s, FM_NCUBE + ab, bb); }
  }
}
// final kinematics only (gripper sites after the last sub-step): same scan without the velocity passes
// phase: per-body quantities (parallel): COM offset, world inertia, composite init, NE body wrench, cluster centre
template <int NA, int E>
FM_PH void ph_arm_body(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < NA * FM_NARMBODY * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ab = w / E; int a = ab / FM_NARMBODY, b = ab % FM_NARMBODY;
    const float* R = s.body_R[a][b];
    if (b == 0) {   // one lane per env: the stale flag is final (cubes in ph_cube_derive, arm clusters in ph_arm_vel)
#ifdef FM_BROAD_ALWAYS
      if (a == 0) s.nqueue = 0;
#else
      if (a == 0 && s.broad_stale) s.nqueue = 0;
#endif
      continue;
    }
    float ms = m.ab_mass[b];
    f3 c = rmul(R, ld3(m.ab_ipos[b]));
    // Iw = R Il R^T
    const float* Il = m.ab_I[b];
    float T1[9];   // R * Il
    for (int i = 0; i < 3; ++i) {
      float r0 = R[3 * i], r1 = R[3 * i + 1], r2 = R[3 * i + 2];
      T1[3 * i] = r0 * Il[0] + r1 * Il[3] + r2 * Il[4];
      T1[3 * i + 1] = r0 * Il[3] + r1 * Il[1] + r2 * Il[5];
      T1[3 * i + 2] = r0 * Il[4] + r1 * Il[5] + r2 * Il[2];
    }
    float Iw[6];
    Iw[0] = T1[0] * R[0] + T1[1] * R[1] + T1[2] * R[2];
    Iw[1] = T1[3] * R[3] + T1[4] * R[4] + T1[5] * R[5];
    Iw[2] = T1[6] * R[6] + T1[7] * R[7] + T1[8] * R[8];
    Iw[3] = T1[0] * R[3] + T1[1] * R[4] + T1[2] * R[5];
    Iw[4] = T1[0] * R[6] + T1[1] * R[7] + T1[2] * R[8];
    Iw[5] = T1[3] * R[6] + T1[4] * R[7] + T1[5] * R[8];
    float cc = dot3(c, c);
    float* cp = s.d_comp0[a][b];
    cp[0] = ms; cp[1] = ms * c.x; cp[2] = ms * c.y; cp[3] = ms * c.z;
    cp[4] = Iw[0] + ms * (cc - c.x * c.x); cp[5] = Iw[1] + ms * (cc - c.y * c.y); cp[6] = Iw[2] + ms * (cc - c.z * c.z);
    cp[7] = Iw[3] - ms * c.x * c.y; cp[8] = Iw[4] - ms * c.x * c.z; cp[9] = Iw[5] - ms * c.y * c.z;
    // Newton-Euler body wrench about the body origin (bias: qacc = 0, gravity folded into the base acceleration)
    f3 om = ld3(s.d_om[a][b]), al = ld3(s.d_al[a][b]), ap = ld3(s.d_ap[a][b]);
    f3 ac = ap + cross3(al, c) + cross3(om, cross3(om, c));
    f3 f = ac * ms;
    f3 n = symmul6(Iw, al) + cross3(om, symmul6(Iw, om)) + cross3(c, f);
    st3(s.d_f0[a][b], f); st3(s.d_n0[a][b], n);
  }
}

// phase: composite inertias and accumulated Newton-Euler wrenches as direct sums over each body's descendants
// (lane = (env, arm, body); subtree of chain body b is b+1..10, of the gripper base the two plates).  Reads the per-body
// values written by ph_arm_body (d_comp0, d_f0, d_n0) and writes the accumulated ones (d_comp, d_f, d_n): one interval.
template <int NA, int E>
FM_PH void ph_arm_bwd(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  for (int w = tid; w < NA * (FM_NARMBODY - 1) * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ab = w / E; int a = ab / (FM_NARMBODY - 1), b = 1 + ab % (FM_NARMBODY - 1);
    float cp[10];
    for (int k = 0; k < 10; ++k) cp[k] = s.d_comp0[a][b][k];
    f3 F = ld3(s.d_f0[a][b]), Nn = ld3(s.d_n0[a][b]);
    f3 xb = ld3(s.body_x[a][b]);
    int d_hi = b <= 8 ? FM_NARMBODY - 1 : b;      // plates are leaves
    for (int d = b + 1; d <= d_hi; ++d) {
      f3 r = ld3(s.body_x[a][d]) - xb;
      const float* cc = s.d_comp0[a][d];
      float ms = cc[0]; f3 hh = mk3(cc[1], cc[2], cc[3]);
      float dg = ms * dot3(r, r) + 2.0f * dot3(r, hh);
      cp[0] += ms; cp[1] += hh.x + ms * r.x; cp[2] += hh.y + ms * r.y; cp[3] += hh.z + ms * r.z;
      cp[4] += cc[4] + dg - (ms * r.x * r.x + 2.0f * r.x * hh.x);
      cp[5] += cc[5] + dg - (ms * r.y * r.y + 2.0f * r.y * hh.y);
      cp[6] += cc[6] + dg - (ms * r.z * r.z + 2.0f * r.z * hh.z);
      cp[7] += cc[7] - (ms * r.x * r.y + r.x * hh.y + hh.x * r.y);
      cp[8] += cc[8] - (ms * r.x * r.z + r.x * hh.z + hh.x * r.z);
      cp[9] += cc[9] - (ms * r.y * r.z + r.y * hh.z + hh.y * r.z);
      f3 f = ld3(s.d_f0[a][d]);
      F = F + f; Nn = Nn + ld3(s.d_n0[a][d]) + cross3(r, f);
    }
    for (int k = 0; k < 10; ++k) s.d_comp[a][b][k] = cp[k];
    st3(s.d_f[a][b], F); st3(s.d_n[a][b], Nn);
  }
}

// phase: mass matrix entries M_ij (parallel over 45 pairs per arm) from composite inertias (CRBA)
template <int NA, int E>
FM_PH void ph_arm_M(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < NA * 45 * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ap = w / E; int a = ap / 45, pr = ap % 45;
    int i = 0; while ((i + 1) * (i + 2) / 2 <= pr) ++i;
    int j = pr - i * (i + 1) / 2;
    float val = 0.f;
    if (!(i == 8 && j == 7)) {        // the two plates are siblings: no coupling
      int bi = dof_body(i), bj = dof_body(j);
      const float* cp = s.d_comp[a][bi];
      f3 hh = mk3(cp[1], cp[2], cp[3]);
      f3 n, f;
      if (i < 7) { f3 ax = rcol(s.body_R[a][bi], 2); n = symmul6(cp + 4, ax); f = cross3(ax, hh); }
      else { f3 ax = rcol(s.body_R[a][bi], 0); f = ax * cp[0]; n = cross3(hh, ax); }
      if (j < 7) {
        f3 r = ld3(s.body_x[a][bi]) - ld3(s.body_x[a][bj]);
        val = dot3(rcol(s.body_R[a][bj], 2), n + cross3(r, f));
      } else {
        val = dot3(rcol(s.body_R[a][bj], 0), f);
      }
    }
    s.M[a][pr] = val;
    // work copies for the two Cholesky factorisations: M and M - h*D (D = d(actuator force)/d(qvel), implicitfast [3P])
    s.Lc[a][pr] = val;
    float add = 0.f;
    if (i == j && i < 7) add = m.h * m.kd;
    else if (i >= 7 && j >= 7) { float gd = m.h * m.g_kd; add = gd * (i == 7 ? m.ten0 : m.ten1) * (j == 7 ? m.ten0 : m.ten1); }   // (7,7), (8,7), (8,8)
    s.Li[a][pr] = val + add;
    if (pr < 9) {   // bias force of dof `pr` from the accumulated wrench of its body
      int bb = dof_body(pr);
      const float* Rb = s.body_R[a][bb];
      s.bias[a][pr] = pr < 7 ? dot3(rcol(Rb, 2), ld3(s.d_n[a][bb])) : dot3(rcol(Rb, 0), ld3(s.d_f[a][bb]));
    }
  }
}

// phase: column k of the two Cholesky factors of every arm, lane = (env, arm, matrix, row i >= k), in place on Lc / Li.
// L[i][k] = (A[i][k] - sum_{j<k} L[i][j] L[k][j]) / L[k][k]; the pivot is recomputed by every lane (k MACs), so an interval
// has no intra-interval dependency.
template <int NA, int E, int K>
FM_PH void ph_arm_chol_col(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  constexpr int ROWS = 9 - K;
  for (int w = tid; w < NA * 2 * ROWS * E; w += T) {
    EnvS<NA>& s = S[w % E]; int r = w / E; int i = K + r % ROWS, mw = r / ROWS; int a = mw >> 1, which = mw & 1;
    float* L = which ? s.Li[a] : s.Lc[a];
    float sk = L[tri(K, K)], si = L[tri(i, K)];
#pragma unroll
    for (int j = 0; j < K; ++j) { float lkj = L[tri(K, j)]; sk -= lkj * lkj; si -= L[tri(i, j)] * lkj; }
    float rp = rsqrtf(fmaxf(sk, 1e-20f));
    // the diagonal lives in chol_d (the unfactored A[k][k] stays in L so that lanes i > k of this interval still read it)
    if (i == K) s.chol_d[a][which][K] = sk * rp; else L[tri(i, K)] = si * rp;
  }
}

FM_HD void chol9_solve(const float* L, const float* dg, float* x) {
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    float sum = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) sum -= L[tri(i, k)] * x[k];
    x[i] = sum / dg[i];
  }
#pragma unroll
  for (int i = 8; i >= 0; --i) {
    float sum = x[i];
#pragma unroll
    for (int k = i + 1; k < 9; ++k) sum -= L[tri(k, i)] * x[k];
    x[i] = sum / dg[i];
  }
}

// phase: explicit inverse, one column per thread
template <int NA, int E>
FM_PH void ph_arm_minv(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  for (int w = tid; w < NA * 9 * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ac = w / E; int a = ac / 9, c = ac % 9;
    float x[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = i == c ? 1.f : 0.f;
    chol9_solve(s.Lc[a], s.chol_d[a][0], x);
#pragma unroll
    for (int i = 0; i < 9; ++i) s.Minv[a][9 * i + c] = x[i];
  }
}

// phase: explicit inverse of M - hD (implicitfast), one column per thread, from its Cholesky factor in Li.  Stored as a packed
// symmetric matrix in Lc, whose Cholesky factor of M is dead once ph_arm_minv has run.  ph_integrate then needs two short
// matrix-vector products per dof instead of a serial 9x9 triangular solve on one lane per arm (that lane set the length of the
// integration interval: 9.5 % of the sub-step spent waiting for it, profiles/r2d).
template <int NA, int E>
FM_PH void ph_arm_winv(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  for (int w = tid; w < NA * 9 * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ac = w / E; int a = ac / 9, c = ac % 9;
    float x[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = i == c ? 1.f : 0.f;
    chol9_solve(s.Li[a], s.chol_d[a][1], x);
#pragma unroll
    for (int i = 0; i < 9; ++i) if (i >= c) s.Lc[a][tri(i, c)] = x[i];
  }
}

// =====================================================================================================================
// broad phase: one cluster pair per work item; survivors are queued for the narrow phase.  Every moving cluster has a
// bounding sphere in shared memory; big geoms (floor, static boxes, belt) have conservative world AABBs in the model.
template <int NA>
FM_HD void cluster_box(const EnvS<NA>& s, const FmDevModel& m, int kind, int idx, f3& p, const float*& R, f3& hs, float* Rid) {
  if (kind == FM_CL_SBOX) { p = ld3(m.sbox_pos[idx]); R = m.sbox_R[idx]; hs = ld3(m.sbox_half[idx]); }
  else {  // belt: slides along +y, identity orientation
    p = mk3(m.belt_pos[0], m.belt_pos[1] + s.belt_q, m.belt_pos[2]); hs = ld3(m.belt_half);
    Rid[0] = 1; Rid[1] = 0; Rid[2] = 0; Rid[3] = 0; Rid[4] = 1; Rid[5] = 0; Rid[6] = 0; Rid[7] = 0; Rid[8] = 1; R = Rid;
  }
}

struct alignas(8) bnd4 { float x, y, z, r; };   // two 64-bit loads: the env stride is 2 (mod 32) words, not a multiple of 4
template <int NA, int E>
FM_PH void ph_broad(EnvS<NA>* S, const FmDevModel& m, const unsigned int* __restrict__ cpairs, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  const int np = m.n_cluster_pair, s0 = m.cl_sph0;
  // lanes per env; consecutive lanes serve the same env, so that a warp whose envs all keep their cached queue skips the phase
  // (with item w -> env w % E every warp would hold lanes of all E envs and run the pair loop for the few stale ones)
  const int stride = T / E;
  if (tid >= stride * E) return;                         // T need not be a multiple of E: the remainder lanes sit this phase out
#ifdef FM_BROAD_INTERLEAVED
  EnvS<NA>& s = S[tid % E];
  const int l0 = tid / E;
#else
  EnvS<NA>& s = S[tid / stride];
  const int l0 = tid % stride;
#endif
#ifndef FM_BROAD_ALWAYS
  if (!s.broad_stale) return;                            // the cached queue is still a superset of the near pairs
#endif
  const float mg = FM_BROAD_MARGIN;
  for (int cl = l0; cl < FmCap<NA>::NCL; cl += stride) {   // remember where this queue was built
    float* rf = s.bnd_ref(cl); const float* bb = s.bnd[cl];
    rf[0] = bb[0]; rf[1] = bb[1]; rf[2] = bb[2]; rf[3] = bb[3];
  }
  const bnd4* bnd = reinterpret_cast<const bnd4*>(&s.bnd[0][0]) - s0;
  for (int p0 = l0; p0 < np; p0 += 4 * stride) {
    // four pair records are fetched up front so their global-load latencies overlap
    unsigned recs[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { int pi = p0 + u * stride; recs[u] = pi < np ? cpairs[pi] : 0xFFFFFFFFu; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      unsigned rec = recs[u];
      if (rec == 0xFFFFFFFFu) continue;
      FM_DBG(fm_dbg_cnt[31]++;)
      int c1 = rec & 255, c2 = (rec >> 8) & 255, type = rec >> 16;
      bnd4 b2 = bnd[c2];
      bool far;
      if (b2.r < 0.f) far = true;                          // sleeping cube
      else if (type == FM_PAIR_SPH_SPH) {
        bnd4 b1 = bnd[c1];
        float dx = b1.x - b2.x, dy = b1.y - b2.y, dz = b1.z - b2.z, rr = b1.r + b2.r + mg;
        far = b1.r < 0.f || dx * dx + dy * dy + dz * dz > rr * rr;
      } else if (type == FM_PAIR_BOX_SPH) {
        const float* bx = m.big_aabb[c1];
        float dx = fmaxf(fmaxf(bx[0] - b2.x, b2.x - bx[3]), 0.f), dy = fmaxf(fmaxf(bx[1] - b2.y, b2.y - bx[4]), 0.f), dz = fmaxf(fmaxf(bx[2] - b2.z, b2.z - bx[5]), 0.f);
        far = dx * dx + dy * dy + dz * dz > (b2.r + mg) * (b2.r + mg);
      } else if (type == FM_PAIR_PLANE_SPH) far = b2.z - b2.r - mg > 0.f;
      else far = false;
      if (far) continue;
      int slot = FM_ATOMIC_ADD(&s.nqueue, 1);
      if (slot < FmCap<NA>::QUEUE) s.queue[slot] = (unsigned short)(p0 + u * stride); else FM_ATOMIC_OR(&s.flags, FM_FLAG_QUEUE_OVERFLOW);
    }
  }
}

// ---- narrow phase helpers
template <int NA>
FM_HD void emit(EnvS<NA>& s, int g1, int g2, int side1, int side2, bool armgeom, const RawC* rc, int n) {
  if (n <= 0) return;
  int slot = FM_ATOMIC_ADD(&s.nraw, n);
  for (int k = 0; k < n; ++k) {
    int i = slot + k;
    if (i >= FmCap<NA>::CON) { FM_ATOMIC_OR(&s.flags, FM_FLAG_CON_OVERFLOW); break; }
    s.raw_key[i] = (g1 << 13) | (g2 << 4) | k;
    s.raw_g[i] = side1 | (side2 << FM_SIDE_BITS) | (armgeom ? (1 << 30) : 0);
    s.raw_d[i][0] = rc[k].pos.x; s.raw_d[i][1] = rc[k].pos.y; s.raw_d[i][2] = rc[k].pos.z;
    s.raw_d[i][3] = rc[k].n.x; s.raw_d[i][4] = rc[k].n.y; s.raw_d[i][5] = rc[k].n.z; s.raw_d[i][6] = rc[k].dist;
  }
}

struct GeomBox { f3 p; const float* R; f3 hs; int geom, side, cls, axis_aligned; float zlo, top_hi, top_lo; };

template <int NA>
FM_HD void get_box(const EnvS<NA>& s, const FmDevModel& m, int kind, int idx, int sub, GeomBox& g, float* Rid) {
  g.zlo = 0.f; g.top_hi = -1.f; g.top_lo = 0.f;
  if (kind == FM_CL_SBOX) { g.top_hi = m.sbox_top[idx][0]; g.top_lo = m.sbox_top[idx][1]; g.p = ld3(m.sbox_pos[idx]); g.R = m.sbox_R[idx]; g.hs = ld3(m.sbox_half[idx]); g.geom = m.sbox_geom[idx]; g.side = pack_side(FM_SIDE_STATIC, 0, 0); g.cls = m.sbox_class[idx]; g.axis_aligned = m.sbox_aligned[idx]; }
  else if (kind == FM_CL_BELT) { g.top_hi = m.belt_top[0]; g.top_lo = m.belt_top[1]; cluster_box(s, m, kind, idx, g.p, g.R, g.hs, Rid); g.geom = m.belt_geom; g.side = pack_side(FM_SIDE_BELT, 0, 0); g.cls = m.belt_class; g.axis_aligned = 1; }
  else if (kind == FM_CL_CUBE) { g.p = ld3(s.cube_x[idx]); g.R = s.cube_R[idx]; float h = m.cube_half[idx]; g.hs = mk3(h, h, h); g.geom = m.cube_geom[idx]; g.side = pack_side(FM_SIDE_CUBE, idx, 0); g.cls = m.cube_class; g.axis_aligned = 0; g.zlo = s.cube_qlo[idx][2]; }
  else {  // arm body box `sub`
    int a = idx / FM_NARMBODY, b = idx % FM_NARMBODY;
    g.R = s.body_R[a][b]; g.p = ld3(s.body_x[a][b]) + rmul(g.R, ld3(m.abox_pos[sub])); g.hs = ld3(m.abox_half[sub]);
    g.geom = m.arm_first_geom[a] + m.abox_geom[sub]; g.side = pack_side(FM_SIDE_ARM, a * 16 + b, 0); g.cls = m.abox_class; g.axis_aligned = 0;
  }
}

template <int NA>
FM_HD void pair_box_box(EnvS<NA>& s, const GeomBox& x, const GeomBox& y, bool armgeom) {
  const GeomBox& g1 = x.geom < y.geom ? x : y; const GeomBox& g2 = x.geom < y.geom ? y : x;
  if (g1.axis_aligned && !g2.axis_aligned) {      // small box on the top face of a big axis-aligned box (belt, table): fast path.
    RawC rf[4];                                   // Only with the big box as geom1: on exact face-axis ties the general routine
    int nf = col_box_on_face(g1.p, g1.hs, g2.p, g2.R, g2.hs, true, rf, g2.zlo, g1.top_hi, g1.top_lo);   // prefers box 1's face, which must be the big one.
    if (nf >= 0) { FM_DBG(fm_dbg_cnt[27]++;) emit(s, g1.geom, g2.geom, g1.side, g2.side, armgeom, rf, nf); return; }
  }
  if (far_sphere_box(g2.p, norm3(g2.hs), g1.p, g1.R, g1.hs) || far_sphere_box(g1.p, norm3(g1.hs), g2.p, g2.R, g2.hs)) { FM_DBG(fm_dbg_cnt[28]++;) return; }
  FM_DBG(fm_dbg_cnt[29]++; fm_dbg_cnt[30] += armgeom; fm_dbg_sat[g1.geom & 511]++;)
  FM_DBG(if (g1.geom == 2) { float r = norm3(g2.hs); float top = g1.p.z + g1.hs.z; int why = g2.p.z < top ? 0 : (fabsf(g2.p.x - g1.p.x) + r > g1.hs.x ? 1 : (fabsf(g2.p.y - g1.p.y) + r > g1.hs.y ? 2 : 3)); fm_dbg_sat[200 + why]++; if (why == 1) fm_dbg_sat[210 + (int)fminf(9.f, fabsf(g2.p.x - g1.p.x) * 20.f)]++; })
  RawC rc[8];
  int n = col_box_box(g1.p, g1.R, g1.hs, g2.p, g2.R, g2.hs, rc);
  emit(s, g1.geom, g2.geom, g1.side, g2.side, armgeom, rc, n);
}

// Narrow phase.  Work item = (queued cluster pair, sub-index j): when the second cluster is an arm body, j selects ONE of
// its geoms (spheres first, then boxes) which is tested against the whole first cluster; otherwise only j == 0 does work.
// This bounds the longest serial chain of the phase by one body's geom count instead of the product of two.
#define FM_NSUB 13
template <int NA>
FM_HD void narrow_arm_geom_vs(EnvS<NA>& s, const FmDevModel& m, int a2, int b2, int j, int k1, int i1) {
  RawC rc[8]; float RidA[9], RidB[9];
  int s0 = m.body_sph_start[b2], ns = m.body_sph_start[b2 + 1] - s0, x0 = m.body_box_start[b2], nb = m.body_box_start[b2 + 1] - x0;
  if (j >= ns + nb) return;
  bool is_sph = j < ns;
  int side2 = b2 == 0 ? pack_side(FM_SIDE_STATIC, 0, 0) : pack_side(FM_SIDE_ARM, a2 * 16 + b2, 0);
  const float* R2 = s.body_R[a2][b2]; f3 p2 = ld3(s.body_x[a2][b2]);
  f3 sq = mk3(0, 0, 0); float sr = 0.f; int sg = 0; GeomBox g2;
  if (is_sph) { int si = s0 + j; sq = p2 + rmul(R2, ld3(m.sph_pos[si])); sr = m.sph_r[si]; sg = m.arm_first_geom[a2] + m.sph_geom[si]; }
  else get_box(s, m, FM_CL_ARM, a2 * FM_NARMBODY + b2, x0 + (j - ns), g2, RidB);
  if (k1 == FM_CL_ARM || k1 == FM_CL_CUBE) {
    // mid phase: this one geom against the first cluster's bounding sphere (conservative; most near cluster pairs of
    // neighbouring links have no near geom pair at all)
    const float* bb = s.bnd[k1 == FM_CL_CUBE ? i1 : FM_NCUBE + i1];
    f3 gc = is_sph ? sq : g2.p; float gr = is_sph ? sr : norm3(g2.hs);
    f3 dd = gc - mk3(bb[0], bb[1], bb[2]); float rr = gr + bb[3];
    if (dot3(dd, dd) > rr * rr) return;
  }
  if (k1 == FM_CL_PLANE) {
    if (is_sph) { int n = col_plane_sphere(sq, sr, rc); emit(s, 0, sg, pack_side(FM_SIDE_STATIC, 0, 0), side2, true, rc, n); }
    else { int n = col_plane_box(g2.p, g2.R, g2.hs, rc); emit(s, 0, g2.geom, pack_side(FM_SIDE_STATIC, 0, 0), g2.side, true, rc, n); }
  } else if (k1 != FM_CL_ARM) {     // static box / belt / cube
    GeomBox x; get_box(s, m, k1, i1, 0, x, RidA);
    if (is_sph) { int n = col_sphere_box(sq, sr, x.p, x.R, x.hs, rc); emit(s, sg, x.geom, side2, x.side, true, rc, n); }   // sphere is geom1 (lower type) [3P]
    else pair_box_box(s, x, g2, true);
  } else {                          // arm body vs arm body
    int a1 = i1 / FM_NARMBODY, b1 = i1 % FM_NARMBODY;
    int t0 = m.body_sph_start[b1], t1 = m.body_sph_start[b1 + 1], y0 = m.body_box_start[b1], y1 = m.body_box_start[b1 + 1];
    int side1 = b1 == 0 ? pack_side(FM_SIDE_STATIC, 0, 0) : pack_side(FM_SIDE_ARM, a1 * 16 + b1, 0);
    const float* R1 = s.body_R[a1][b1]; f3 p1 = ld3(s.body_x[a1][b1]);
    for (int si = t0; si < t1; ++si) {
      f3 sp = p1 + rmul(R1, ld3(m.sph_pos[si])); float r1 = m.sph_r[si]; int g1 = m.arm_first_geom[a1] + m.sph_geom[si];
      if (is_sph) {
        if (g1 < sg) { int n = col_sphere_sphere(sp, r1, sq, sr, rc); emit(s, g1, sg, side1, side2, true, rc, n); }
        else { int n = col_sphere_sphere(sq, sr, sp, r1, rc); emit(s, sg, g1, side2, side1, true, rc, n); }
      } else { int n = col_sphere_box(sp, r1, g2.p, g2.R, g2.hs, rc); emit(s, g1, g2.geom, side1, g2.side, true, rc, n); }
    }
    for (int bi = y0; bi < y1; ++bi) {
      GeomBox g1; get_box(s, m, k1, i1, bi, g1, RidA);
      if (is_sph) { int n = col_sphere_box(sq, sr, g1.p, g1.R, g1.hs, rc); emit(s, sg, g1.geom, side2, g1.side, true, rc, n); }
      else pair_box_box(s, g1, g2, true);
    }
  }
}

// one narrow-phase work item w = (env, queued pair, sub-index j)
template <int NA, int E>
FM_HD void narrow_item(EnvS<NA>* S, const FmDevModel& m, const unsigned int* __restrict__ cpairs, int w) {
  EnvS<NA>& s = S[w % E]; int rest = w / E; int qi = rest / FM_NSUB, j = rest % FM_NSUB;
  int nq = s.nqueue < FmCap<NA>::QUEUE ? s.nqueue : FmCap<NA>::QUEUE;
  if (qi >= nq) return;
  unsigned rec = cpairs[s.queue[qi]];
  int c1 = rec & 255, c2 = (rec >> 8) & 255;
  int k1 = m.cl_kind[c1], k2 = m.cl_kind[c2], i1 = m.cl_idx[c1], i2 = m.cl_idx[c2];
  if (k2 == FM_CL_ARM) { narrow_arm_geom_vs(s, m, i2 / FM_NARMBODY, i2 % FM_NARMBODY, j, k1, i1); return; }
  if (j != 0) return;
  RawC rc[8]; float RidA[9], RidB[9];
  if (k1 == FM_CL_PLANE) {
    if (k2 == FM_CL_CUBE) { float h = m.cube_half[i2]; int n = col_plane_box(ld3(s.cube_x[i2]), s.cube_R[i2], mk3(h, h, h), rc); emit(s, 0, m.cube_geom[i2], pack_side(FM_SIDE_STATIC, 0, 0), pack_side(FM_SIDE_CUBE, i2, 0), false, rc, n); }
    else if (k2 == FM_CL_BELT) { GeomBox b; get_box(s, m, k2, i2, 0, b, RidB); int n = col_plane_box(b.p, b.R, b.hs, rc); emit(s, 0, b.geom, pack_side(FM_SIDE_STATIC, 0, 0), b.side, false, rc, n); }
  } else {
    GeomBox x, y; get_box(s, m, k1, i1, 0, x, RidA); get_box(s, m, k2, i2, 0, y, RidB);
    pair_box_box(s, x, y, false);
  }
}
template <int NA, int E>
FM_HD int narrow_total(EnvS<NA>* S) {
  int mq = 0;
  for (int e = 0; e < E; ++e) mq = S[e].nqueue > mq ? S[e].nqueue : mq;
  if (mq > FmCap<NA>::QUEUE) mq = FmCap<NA>::QUEUE;
  return mq * FM_NSUB * E;
}
// the block's narrow-phase work counter lives in the first env's padding word (reset by ph_clear)
template <int NA> FM_HD int* narrow_counter(EnvS<NA>* S) { return &S[0].pad_[0]; }

// Narrow phase.  Optional (-DFM_NARROW_STEAL): the items are handed out in chunks of 32 from a block-wide counter, so that BOTH
// warp groups of fm_derive work on them (group A joins when its factorisation intervals are done).  Measured 5 % SLOWER than the
// static loop on the bench workload (profiles/README r2l): the interval is as long as its one general box-box SAT item, not as
// its item count.  Contact emission is order-independent either way (atomic slots, ranked afterwards).
template <int NA, int E>
FM_PH void ph_narrow_steal(EnvS<NA>* S, const FmDevModel& m, const unsigned int* __restrict__ cpairs, int lane) {
#if defined(__CUDA_ARCH__)
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  const int total = narrow_total<NA, E>(S);
  int* cnt = narrow_counter<NA>(S);
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(cnt, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= total) break;
    if (base + lane < total) narrow_item<NA, E>(S, m, cpairs, base + lane);
    __syncwarp();
  }
#endif
}
template <int NA, int E>
FM_PH void ph_narrow(EnvS<NA>* S, const FmDevModel& m, const unsigned int* __restrict__ cpairs, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
#if defined(__CUDA_ARCH__) && defined(FM_NARROW_STEAL)
  ph_narrow_steal<NA, E>(S, m, cpairs, tid & 31);
#else
  const int total = narrow_total<NA, E>(S);
  for (int w = tid; w < total; w += T) narrow_item<NA, E>(S, m, cpairs, w);
#endif
}

// =====================================================================================================================
// constraint setup
FM_HD void impedance_f(const float* solref, const float* solimp, float pos, float h, float& K, float& B, float& imp) {
  float dmin = clampf(solimp[0], 0.0001f, 0.9999f), dmax = clampf(solimp[1], 0.0001f, 0.9999f);
  float width = fmaxf(solimp[2], 0.f), mid = clampf(solimp[3], 0.0001f, 0.9999f), power = fmaxf(solimp[4], 1.f);
  if (dmin == dmax || width <= 1e-15f) imp = 0.5f * (dmin + dmax);
  else {
    float x = fabsf(pos) / width, y;
    if (x >= 1.f) y = 1.f;
    else if (x <= 0.f) y = 0.f;
    else if (power == 1.f) y = x;
    else if (power == 2.f) y = x <= mid ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
    else if (x <= mid) y = powf(x, power) / powf(mid, power - 1.f);
    else y = 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
    imp = dmin + y * (dmax - dmin);
  }
  float tc = fmaxf(solref[0], 2.f * h), dr = solref[1];
  K = 1.f / fmaxf(1e-15f, dmax * dmax * tc * tc * dr * dr);
  B = 2.f / fmaxf(1e-15f, dmax * tc);
}

// per-side working set of a contact (contact-space Jacobian base rows n, t1, t2).
// Cubes have isotropic inertia (uniform box with equal sides, scene.py:9-23), so their angular acceleration is kept in the
// WORLD frame during the solve: J a = d . (a_lin + alpha x r) and Minv J^T lam = (F/m, (r x F)/I) with F = sum_q d_q lam_q.
struct SideW {
  int type, idx;
  float sgn;
  f3 r;                // cube: contact point relative to the cube centre
  float invm, invI;
  float jb[3];         // belt: scalar rows
  const float* J; const float* B;   // arm: 3x9 each
};

template <int NA>
FM_HD void side_prepare(const EnvS<NA>& s, const FmDevModel& m, int sd, float sgn, f3 p, f3 n, f3 t1, f3 t2, SideW& o) {
  o.type = side_type(sd); o.idx = side_idx(sd); o.sgn = sgn;
  if (o.type == FM_SIDE_CUBE) {
    int k = o.idx;
    o.r = p - ld3(s.cube_x[k]);
    o.invm = m.cube_invm[k]; o.invI = m.cube_invI[k];
  } else if (o.type == FM_SIDE_BELT) {
    o.jb[0] = n.y; o.jb[1] = t1.y; o.jb[2] = t2.y;
  } else if (o.type == FM_SIDE_ARM) {
    int sl = side_slot(sd);
    o.J = s.as_J[sl]; o.B = s.as_B[sl];
  }
}
// u += sgn * Jbase * vec ; cube block of `cube` = [linear(3), angular(3) in the WORLD frame]
template <int NA>
FM_HD void side_dot(const SideW& o, f3 n, f3 t1, f3 t2, const float* belt, const float* cube, const float* arm, float* u) {
  if (o.type == FM_SIDE_CUBE) {
    const float* v = cube + 6 * o.idx;
    f3 pv = mk3(v[0], v[1], v[2]) + cross3(mk3(v[3], v[4], v[5]), o.r);
    u[0] += o.sgn * dot3(n, pv); u[1] += o.sgn * dot3(t1, pv); u[2] += o.sgn * dot3(t2, pv);
  } else if (o.type == FM_SIDE_BELT) {
    float v = belt[0];
    u[0] += o.sgn * o.jb[0] * v; u[1] += o.sgn * o.jb[1] * v; u[2] += o.sgn * o.jb[2] * v;
  } else if (o.type == FM_SIDE_ARM) {
    const float* v = arm + 9 * (o.idx >> 4);
    for (int d = 0; d < 3; ++d) { float acc = 0.f; for (int k = 0; k < 9; ++k) acc += o.J[9 * d + k] * v[k]; u[d] += o.sgn * acc; }
  }
}
// acc += Minv * Jbase^T * (sgn * lam)
template <int NA>
FM_HD void side_apply(const SideW& o, f3 n, f3 t1, f3 t2, float* cube, float* arm, const float* lam) {
  if (o.type == FM_SIDE_CUBE) {
    float* v = cube + 6 * o.idx;
    f3 F = (n * lam[0] + t1 * lam[1] + t2 * lam[2]) * o.sgn;
    f3 fl = F * o.invm, fa = cross3(o.r, F) * o.invI;
    v[0] += fl.x; v[1] += fl.y; v[2] += fl.z; v[3] += fa.x; v[4] += fa.y; v[5] += fa.z;
  } else if (o.type == FM_SIDE_ARM) {
    float* v = arm + 9 * (o.idx >> 4);
    for (int k = 0; k < 9; ++k) v[k] += o.sgn * (o.B[k] * lam[0] + o.B[9 + k] * lam[1] + o.B[18 + k] * lam[2]);
  }
}

// warm start of contact c: forces of the previous solve for the same contact (geom pair + index); binary search on sorted keys
template <int NA>
FM_HD void warm_start_forces(EnvS<NA>& s, int c, int key) {
  int lo = 0, hi = s.prev_n - 1, found = -1;
  while (lo <= hi) { int mid = (lo + hi) >> 1; int pk = s.prev_key[mid]; if (pk == key) { found = mid; break; } if (pk < key) lo = mid + 1; else hi = mid - 1; }
  for (int r = 0; r < 4; ++r) s.c_f[c][r] = found >= 0 ? s.prev_f[found][r] : 0.f;
}

// which = 1: contacts without an arm geom (they need nothing of the arm factorisation, so fm_derive runs them beside it),
// which = 2: contacts of arm geoms (arm-side Jacobians and M^-1 J^T), which = 0: all
template <int NA, int E>
FM_HDN void ph_contact_setup(EnvS<NA>* S, const FmDevModel& m, const int* __restrict__ geom_class, int tid, int T, int which) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < FmCap<NA>::CON * E; w += T) {
    EnvS<NA>& s = S[w % E]; int ri = w / E;
    int nraw = s.nraw < FmCap<NA>::CON ? s.nraw : FmCap<NA>::CON;
    if (ri >= nraw) continue;
    if (which != 0 && (((s.raw_g[ri] >> 30) & 1) != 0) != (which == 2)) continue;
    int key = s.raw_key[ri];
    // rank = sorted position; arm-side slots are handed out in the same order (deterministic, independent of atomics)
    int rank = 0, slot_base = 0;
    for (int k = 0; k < nraw; ++k) rank += s.raw_key[k] < key ? 1 : 0;
    int c = rank;
    int g1 = key >> 13, g2 = (key >> 4) & 511;
    int rg = s.raw_g[ri];
    if ((rg >> 30) & 1)     // only contacts of arm geoms need a slot for their arm-side Jacobian
      for (int k = 0; k < nraw; ++k) {
        int g = s.raw_g[k];
        if (((g >> 30) & 1) && s.raw_key[k] < key)
          slot_base += (side_type(g & FM_SIDE_MASK) == FM_SIDE_ARM ? 1 : 0) + (side_type((g >> FM_SIDE_BITS) & FM_SIDE_MASK) == FM_SIDE_ARM ? 1 : 0);
      }
    int sd1 = rg & FM_SIDE_MASK, sd2 = (rg >> FM_SIDE_BITS) & FM_SIDE_MASK; bool armgeom = (rg >> 30) & 1;
    f3 p = mk3(s.raw_d[ri][0], s.raw_d[ri][1], s.raw_d[ri][2]), n = mk3(s.raw_d[ri][3], s.raw_d[ri][4], s.raw_d[ri][5]);
    float dist = s.raw_d[ri][6];
    f3 t1, t2; make_frame3(n, t1, t2);
    const int gc1 = geom_class[g1], gc2 = geom_class[g2];
    const float* pp = m.pair_param[gc1 & 15][gc2 & 15];
    float mu = pp[0];
    // arm sides: Jacobian rows and Minv J^T
    int sds[2] = {sd1, sd2};
    float invw = 0.f;
    for (int sdi = 0; sdi < 2; ++sdi) {
      int sd = sds[sdi];
      int ty = side_type(sd), ix = side_idx(sd);
      if (ty == FM_SIDE_CUBE) invw += m.cube_invw[ix];
      else if (ty == FM_SIDE_BELT) invw += m.belt_invw;
      else if (ty == FM_SIDE_ARM) {
        int a = ix >> 4, b = ix & 15;
        invw += m.ab_invw[b];
        int sl = slot_base++;
        if (sl >= FmCap<NA>::ARMSIDE) { FM_ATOMIC_OR(&s.flags, FM_FLAG_ARMSIDE_OVERFLOW); sds[sdi] = pack_side(FM_SIDE_STATIC, 0, 0); continue; }
        sds[sdi] = pack_side(ty, ix, sl);
        float* J = s.as_J[sl]; float* Bv = s.as_B[sl];
        f3 dirs[3] = {n, t1, t2};
        for (int d = 0; d < 9; ++d) {
          int bd = dof_body(d);
          bool on = d < 7 ? (bd <= b) : (bd == b);
          f3 colv = mk3(0, 0, 0);
          if (on) colv = d < 7 ? cross3(rcol(s.body_R[a][bd], 2), p - ld3(s.body_x[a][bd])) : rcol(s.body_R[a][bd], 0);
          for (int q = 0; q < 3; ++q) J[9 * q + d] = dot3(dirs[q], colv);
        }
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 9; ++i) { float acc = 0.f; for (int k = 0; k < 9; ++k) acc += s.Minv[a][9 * i + k] * J[9 * q + k]; Bv[9 * q + i] = acc; }
      }
    }
    s.c_side[c] = sds[0] | (sds[1] << FM_SIDE_BITS) | (armgeom ? (1 << 30) : 0);
    {
      int cs2 = s.c_side[c];
      int n0 = contact_node(cs2, 0), n1 = contact_node(cs2, 1);
      s.c_node[c] = (unsigned char)(n0 >= 0 ? n0 : (n1 >= 0 ? n1 : 255));
      if (n0 >= 0 && n1 >= 0 && n0 != n1) FM_ATOMIC_OR(&s.need_union, 1);
    }
    st3(s.c_p[c], p); st3(s.c_n[c], n); st3(s.c_t1[c], t1); s.c_mu[c] = mu;
    float K, Bd, imp; impedance_f(pp + 1, pp + 3, dist, m.h, K, Bd, imp);
    float dA = invw * (1.f + mu * mu);
    float R0 = fmaxf(1e-15f, (1.f - imp) * dA / imp);
    float Rpy = 2.f * mu * mu * R0;           // [3P] pyramidal regulariser
    if (gc1 & gc2 & 16) Rpy *= FM_BOXBOX_RSCALE;   // box-box pairs: calibrated to the reference's rest depth (see the define)
    s.c_R[c] = Rpy;
    SideW A, Bs;
    side_prepare(s, m, sds[0], -1.f, p, n, t1, t2, A);
    side_prepare(s, m, sds[1], +1.f, p, n, t1, t2, Bs);
    // contact-space Delassus block W = sum_sides Jb Minv Jb^T  (xx,yy,zz,xy,xz,yz)
    float W[6] = {0, 0, 0, 0, 0, 0};
    const SideW* sw[2] = {&A, &Bs};
    for (int q = 0; q < 2; ++q) {
      const SideW& o = *sw[q];
      if (o.type == FM_SIDE_CUBE) {
        f3 a0 = cross3(o.r, n), a1 = cross3(o.r, t1), a2 = cross3(o.r, t2);
        W[0] += o.invm + o.invI * dot3(a0, a0); W[1] += o.invm + o.invI * dot3(a1, a1); W[2] += o.invm + o.invI * dot3(a2, a2);
        W[3] += o.invI * dot3(a0, a1); W[4] += o.invI * dot3(a0, a2); W[5] += o.invI * dot3(a1, a2);
      } else if (o.type == FM_SIDE_BELT) {
        float im = 1.0f / m.belt_mass;
        W[0] += im * o.jb[0] * o.jb[0]; W[1] += im * o.jb[1] * o.jb[1]; W[2] += im * o.jb[2] * o.jb[2];
        W[3] += im * o.jb[0] * o.jb[1]; W[4] += im * o.jb[0] * o.jb[2]; W[5] += im * o.jb[1] * o.jb[2];
      } else if (o.type == FM_SIDE_ARM) {
        float d00 = 0, d11 = 0, d22 = 0, d01 = 0, d02 = 0, d12 = 0;
        for (int k = 0; k < 9; ++k) {
          d00 += o.J[k] * o.B[k]; d11 += o.J[9 + k] * o.B[9 + k]; d22 += o.J[18 + k] * o.B[18 + k];
          d01 += o.J[k] * o.B[9 + k]; d02 += o.J[k] * o.B[18 + k]; d12 += o.J[9 + k] * o.B[18 + k];
        }
        W[0] += d00; W[1] += d11; W[2] += d22; W[3] += d01; W[4] += d02; W[5] += d12;
      }
    }
    for (int q = 0; q < 6; ++q) s.c_W[c][q] = W[q];
    // velocity and warm-start acceleration in contact space
    float uv[3] = {0, 0, 0};
    side_dot<NA>(A, n, t1, t2, &s.belt_v, &s.cube_vw[0][0], &s.arm_v[0][0], uv);
    side_dot<NA>(Bs, n, t1, t2, &s.belt_v, &s.cube_vw[0][0], &s.arm_v[0][0], uv);
    float kid = K * imp * dist;
    for (int r = 0; r < 4; ++r) {
      int tq = 1 + (r >> 1); float cf = (r & 1) ? -mu : mu;
      float vel = uv[0] + cf * uv[tq];
      float aref = -Bd * vel - kid;
      s.c_aref[c][r] = aref;
    }
    // warm start: forces of the previous solve for the same contact (geom pair + index); binary search on sorted keys.
    // (MuJoCo warm-starts the dual from qacc_warmstart as f = -jar/R; in fp32 that quotient is dominated by rounding.)
    s.c_key[c] = key;
    warm_start_forces(s, c, key);
  }
}

// phase: arm-local rows (thread per arm): joint equality q_l - q_r = 0 and joint limits
template <int NA, int E>
FM_PH void ph_arm_rows(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < NA * E; w += T) {
    EnvS<NA>& s = S[w % E]; int a = w / E;
    int nr = 0;
    {
      float pos = s.arm_q[a][7] - s.arm_q[a][8];
      float K, Bd, imp; impedance_f(m.eq_solref, m.eq_solimp, pos, m.h, K, Bd, imp);
      float dA = m.ab_dofinvw[9] + m.ab_dofinvw[10];
      float R = fmaxf(1e-15f, (1.f - imp) * dA / imp);
      float vel = s.arm_v[a][7] - s.arm_v[a][8];
      float aref = -Bd * vel - K * imp * pos;
      s.arow_code[a][nr] = 0x1000; s.arow_aref[a][nr] = aref; s.arow_R[a][nr] = R;
      s.arow_f[a][nr] = 0.f;
      nr++;
    }
    for (int d = 0; d < 9; ++d) {
      int b = dof_body(d);
      float q = s.arm_q[a][d];
      float dlo = q - m.ab_range[b][0], dhi = m.ab_range[b][1] - q;
      for (int side = 0; side < 2; ++side) {
        float dist = side ? dhi : dlo;
        if (dist < 0.f) {
          float sg = side ? -1.f : 1.f;
          float K, Bd, imp; impedance_f(m.lim_solref, m.lim_solimp, dist, m.h, K, Bd, imp);
          float R = fmaxf(1e-15f, (1.f - imp) * m.ab_dofinvw[b] / imp);
          float aref = -Bd * sg * s.arm_v[a][d] - K * imp * dist;
          s.arow_code[a][nr] = d | (side << 8); s.arow_aref[a][nr] = aref; s.arow_R[a][nr] = R;
          s.arow_f[a][nr] = 0.f;
          nr++;
        }
      }
    }
    s.arow_n[a] = nr;
    for (int r = 0; r < nr; ++r)
      for (int q = 0; q < s.prev_arow_n[a]; ++q) if (s.prev_arow_code[a][q] == s.arow_code[a][r]) s.arow_f[a][r] = s.prev_arow_f[a][q];
  }
}

// phase: constraint islands (thread per env): union-find over cubes and arms, contact lists per island root
// Constraint islands.  Nodes = cubes and arms; a contact belongs to the island of its dynamic side(s).  Almost always no
// contact couples two dynamic nodes (cube-cube, cube-arm, arm-arm), so every node is its own island and the per-island
// contact lists are built fully in parallel (ph_island_lists).  Otherwise one thread per env runs a union-find first.
template <int NA, int E>
FM_PH void ph_island_union(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  const int NN = EnvS<NA>::NNODE;
  for (int w = tid; w < E; w += T) {
    EnvS<NA>& s = S[w];
    int nraw = s.nraw < FmCap<NA>::CON ? s.nraw : FmCap<NA>::CON;
    s.ncon = nraw;
    s.broad_stale = 0;                                   // consumed by ph_broad of this sub-step
    FM_DBG(fm_dbg_cnt[0]++; fm_dbg_cnt[1] += nraw; fm_dbg_cnt[2] += s.nqueue; fm_dbg_cnt[3] += s.need_union & 1; for (int k = 0; k < FM_NCUBE; ++k) fm_dbg_cnt[26] += !s.cube_sleep[k];)
    if (!(s.need_union & 1)) continue;
    int* root = s.isl_root;
    for (int c = 0; c < nraw; ++c) {
      int cs = s.c_side[c];
      int n0 = contact_node(cs, 0), n1 = contact_node(cs, 1);
      if (n0 >= 0 && n1 >= 0) {
        int ra = n0; while (root[ra] != ra) ra = root[ra];
        int rb = n1; while (root[rb] != rb) rb = root[rb];
        if (ra != rb) { if (ra < rb) root[rb] = ra; else root[ra] = rb; }
      }
    }
    for (int i = 0; i < NN; ++i) { int r = i; while (root[r] != r) r = root[r]; root[i] = r; }
  }
}
template <int NA, int E>
FM_PH void ph_island_lists(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  const int NN = EnvS<NA>::NNODE;
  for (int w = tid; w < (FmCap<NA>::CON + NN + 1) * E; w += T) {
    EnvS<NA>& s = S[w % E]; int i = w / E;
    int nraw = s.nraw < FmCap<NA>::CON ? s.nraw : FmCap<NA>::CON;
    if (i < FmCap<NA>::CON) {
      if (i >= nraw) continue;
      int nd = s.c_node[i];
      if (nd == 255) continue;
      int r = s.isl_root[nd];
      int pos = 0;
      for (int c = 0; c < nraw; ++c) { int n2 = s.c_node[c]; if (n2 == 255) continue; int r2 = s.isl_root[n2]; pos += (r2 < r || (r2 == r && c < i)) ? 1 : 0; }
      s.isl_list[pos] = (unsigned char)i;
    } else {
      int node = i - FmCap<NA>::CON;      // 0..NN: start offset of island `node` (NN = end sentinel)
      int cnt = 0;
      for (int c = 0; c < nraw; ++c) { int n2 = s.c_node[c]; if (n2 == 255) continue; cnt += (s.isl_root[n2] < node) ? 1 : 0; }
      s.isl_start[node] = cnt;
    }
  }
}

// phase: actuation + smooth acceleration (mj_fwdActuation, mj_fwdAcceleration), plus the warm start of the constraint solve:
// every lane owns one cube (6 rows of acc) or one arm dof and adds Minv J^T f of the previous solve's forces for the contacts /
// arm rows that touch it, in ascending contact order (fixed order: run-to-run determinism).
template <int NA, int E>
FM_PH void ph_smooth(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  const int per = 1 + FM_NCUBE + 9 * NA;
  for (int w = tid; w < per * E; w += T) {
    EnvS<NA>& s = S[w % E]; int i = w / E;
    const int ncon = s.nraw < FmCap<NA>::CON ? s.nraw : FmCap<NA>::CON;
    if (i == 0) {
      float c = clampf(s.ctrl_target[0], m.ctrl_lo[0], m.ctrl_hi[0]);
      float f = m.belt_kv * (c - s.belt_v) - m.belt_damp * s.belt_v;
      s.belt_acc_smooth = f / m.belt_mass; s.acc[0] = s.belt_acc_smooth + s.belt_S;
    } else if (i <= FM_NCUBE) {
      int k = i - 1;
      float a[6] = {0, 0, -m.g, 0, 0, 0};
      if (!s.cube_sleep[k]) {
        f3 xk = ld3(s.cube_x[k]); float invm = m.cube_invm[k], invI = m.cube_invI[k];
        for (int c = 0; c < ncon; ++c) {
          int cs = s.c_side[c];
          int sd0 = cs & FM_SIDE_MASK, sd1 = (cs >> FM_SIDE_BITS) & FM_SIDE_MASK;
          bool h0 = side_type(sd0) == FM_SIDE_CUBE && side_idx(sd0) == k, h1 = side_type(sd1) == FM_SIDE_CUBE && side_idx(sd1) == k;
          if (!h0 && !h1) continue;
          f3 n = ld3(s.c_n[c]), t1 = ld3(s.c_t1[c]); f3 t2 = cross3(n, t1);
          const float* f = s.c_f[c]; float mu = s.c_mu[c];
          float l0 = f[0] + f[1] + f[2] + f[3], l1 = mu * (f[0] - f[1]), l2 = mu * (f[2] - f[3]);
          f3 F = (n * l0 + t1 * l1 + t2 * l2) * (h0 ? -1.f : 1.f);
          f3 fl = F * invm, fa = cross3(ld3(s.c_p[c]) - xk, F) * invI;
          a[0] += fl.x; a[1] += fl.y; a[2] += fl.z; a[3] += fa.x; a[4] += fa.y; a[5] += fa.z;
        }
      }
      float* o = &s.acc[1 + 6 * k];
      for (int q = 0; q < 6; ++q) o[q] = a[q];
    } else {
      int ad = i - 1 - FM_NCUBE; int a = ad / 9, r = ad % 9;
      // row r of Minv * (tau - bias)
      float tau[9];
      for (int d = 0; d < 7; ++d) {
        float c = clampf(s.ctrl_target[1 + 8 * a + d], m.ctrl_lo[1 + d], m.ctrl_hi[1 + d]);
        tau[d] = m.kp * (c - s.arm_q[a][d]) - m.kd * s.arm_v[a][d] - s.bias[a][d];
      }
      float c = clampf(s.ctrl_target[1 + 8 * a + 7], m.ctrl_lo[8], m.ctrl_hi[8]);
      float len = m.ten0 * s.arm_q[a][7] + m.ten1 * s.arm_q[a][8], lv = m.ten0 * s.arm_v[a][7] + m.ten1 * s.arm_v[a][8];
      float f = clampf(m.g_kp * c - m.g_kb * len - m.g_kd * lv, m.g_fmin, m.g_fmax);
      tau[7] = m.ten0 * f - s.bias[a][7]; tau[8] = m.ten1 * f - s.bias[a][8];
      float acc = 0.f;
      const float* Mi = s.Minv[a];
      for (int k = 0; k < 9; ++k) acc += Mi[9 * r + k] * tau[k];
      // warm start: arm rows, then contacts with a side on this arm
      for (int q = 0; q < s.arow_n[a]; ++q) {
        int code = s.arow_code[a][q]; float fq = s.arow_f[a][q];
        if (code & 0x1000) acc += (Mi[9 * r + 7] - Mi[9 * r + 8]) * fq;
        else { int d = code & 255; float sg = (code >> 8) ? -1.f : 1.f; acc += sg * Mi[9 * r + d] * fq; }
      }
      {
        for (int c = 0; c < ncon; ++c) {
          int cs = s.c_side[c];
          if (!((cs >> 30) & 1)) continue;
          const float* fc = s.c_f[c]; float mu = s.c_mu[c];
          float l0 = fc[0] + fc[1] + fc[2] + fc[3], l1 = mu * (fc[0] - fc[1]), l2 = mu * (fc[2] - fc[3]);
          for (int q = 0; q < 2; ++q) {
            int sd = (cs >> (q * FM_SIDE_BITS)) & FM_SIDE_MASK;
            if (side_type(sd) != FM_SIDE_ARM || (side_idx(sd) >> 4) != a) continue;
            const float* Bv = s.as_B[side_slot(sd)];
            acc += (q ? 1.f : -1.f) * (Bv[r] * l0 + Bv[9 + r] * l1 + Bv[18 + r] * l2);
          }
        }
      }
      s.acc[61 + 9 * a + r] = acc;
    }
  }
}

// One Gauss-Seidel visit of contact c on the current acceleration: the four pyramid rows are updated in turn on the
// contact-space acceleration u = J a (3x3 Delassus block W).  commit = false only evaluates (convergence check from the
// warm start); commit = true stores the forces and applies Minv J^T dlam.  The belt-dof projection of the contact's force is
// left in c_belt(c) either way (lagged belt coupling, summed by ph_belt).  Returns max |df| * A_rr.
// the raw contact buffer is dead after ph_contact_setup: its key words hold the per-contact residual and belt projection from
// ph_check on, its data words (raw_d) are the scratch of the big-island solve
template <int NA> FM_HD float& c_res(EnvS<NA>& s, int c) { return reinterpret_cast<float*>(&s.raw_key[0])[c]; }
template <int NA> FM_HD float& c_belt(EnvS<NA>& s, int c) { return reinterpret_cast<float*>(&s.raw_g[0])[c]; }
template <int NA>
// out of line (like arm_rows_visit, island_pgs_sweeps, island_newton_warp): one copy for the check and the sweeps, and its registers
// stay out of the fused loop's allocation — + 8 % env-steps/s for the four together (profiles/r2_ab/r4fgh_outline.log)
FM_HDN float contact_visit(EnvS<NA>& s, const FmDevModel& m, int c, bool commit) {
  int cs = s.c_side[c];
  f3 p = ld3(s.c_p[c]), n = ld3(s.c_n[c]), t1 = ld3(s.c_t1[c]); f3 t2 = cross3(n, t1);
  SideW A, B;
  side_prepare(s, m, cs & FM_SIDE_MASK, -1.f, p, n, t1, t2, A);
  side_prepare(s, m, (cs >> FM_SIDE_BITS) & FM_SIDE_MASK, +1.f, p, n, t1, t2, B);
  float mu = s.c_mu[c], R = s.c_R[c];
  const float* W = s.c_W[c];
  float f[4] = {s.c_f[c][0], s.c_f[c][1], s.c_f[c][2], s.c_f[c][3]};
  float lam[3] = {0, 0, 0};
  float maxd = 0.f;
  float u[3] = {0, 0, 0};
  side_dot<NA>(A, n, t1, t2, &s.acc[0], &s.acc[1], &s.acc[61], u);
  side_dot<NA>(B, n, t1, t2, &s.acc[0], &s.acc[1], &s.acc[61], u);
  float Fn = f[0] + f[1] + f[2] + f[3], F1 = mu * (f[0] - f[1]), F2 = mu * (f[2] - f[3]);
  for (int r = 0; r < 4; ++r) {
    int tq = 1 + (r >> 1); float cf = (r & 1) ? -mu : mu;
    float Wnt = tq == 1 ? W[3] : W[4], Wtt = tq == 1 ? W[1] : W[2], Wtx = W[5];
    float AR = W[0] + 2.f * cf * Wnt + cf * cf * Wtt + R;
    float res = u[0] + cf * u[tq] - s.c_aref[c][r] + R * f[r];
    float fn = fmaxf(0.f, f[r] - FM_FDIV(res, AR));
    float df = fn - f[r];
    f[r] = fn;
    // contact-space acceleration update: u += W * (1, cf on tangent tq) * df
    u[0] += (W[0] + cf * Wnt) * df;
    u[tq] += (Wnt + cf * Wtt) * df;
    u[3 - tq] += ((tq == 1 ? W[4] : W[3]) + cf * Wtx) * df;
    lam[0] += df; lam[tq] += cf * df;
    maxd = fmaxf(maxd, fabsf(df) * AR);
  }
  if (commit) {
    for (int r = 0; r < 4; ++r) s.c_f[c][r] = f[r];
    side_apply<NA>(A, n, t1, t2, &s.acc[1], &s.acc[61], lam);
    side_apply<NA>(B, n, t1, t2, &s.acc[1], &s.acc[61], lam);
    Fn = f[0] + f[1] + f[2] + f[3]; F1 = mu * (f[0] - f[1]); F2 = mu * (f[2] - f[3]);
  }
  float bs = 0.f;
  if (A.type == FM_SIDE_BELT || B.type == FM_SIDE_BELT) {
    const SideW& o = A.type == FM_SIDE_BELT ? A : B;
    bs = o.sgn * (o.jb[0] * Fn + o.jb[1] * F1 + o.jb[2] * F2);
  }
  c_belt(s, c) = bs;
  return maxd;
}

// arm-local rows of arm a: mode 1 = Gauss-Seidel sweep, mode 2 = evaluate only (convergence check)
template <int NA>
FM_HDN float arm_rows_visit(EnvS<NA>& s, int a, int mode) {
  float* acc = &s.acc[61 + 9 * a];
  float maxd = 0.f;
  const float* Mi = s.Minv[a];
  for (int r = 0; r < s.arow_n[a]; ++r) {
    int code = s.arow_code[a][r];
    float R = s.arow_R[a][r], f = s.arow_f[a][r], df;
    if (code & 0x1000) {
      float AR = Mi[9 * 7 + 7] - 2.f * Mi[9 * 7 + 8] + Mi[9 * 8 + 8] + R;
      float res = acc[7] - acc[8] - s.arow_aref[a][r] + R * f; df = -res / AR;
      maxd = fmaxf(maxd, fabsf(df) * AR);
      if (mode == 1) { s.arow_f[a][r] = f + df; for (int k = 0; k < 9; ++k) acc[k] += (Mi[9 * k + 7] - Mi[9 * k + 8]) * df; }
    } else {
      int d = code & 255; float sg = (code >> 8) ? -1.f : 1.f;
      float AR = Mi[9 * d + d] + R;
      float res = sg * acc[d] - s.arow_aref[a][r] + R * f; float fn = fmaxf(0.f, f - res / AR); df = fn - f;
      maxd = fmaxf(maxd, fabsf(df) * AR);
      if (mode == 1) { s.arow_f[a][r] = fn; for (int k = 0; k < 9; ++k) acc[k] += sg * Mi[9 * k + d] * df; }
    }
  }
  return maxd;
}

// phase: convergence check of the warm start, one lane per contact (evaluation only; nothing another lane reads is written)
template <int NA, int E>
FM_PH void ph_check(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < FmCap<NA>::CON * E; w += T) {
    EnvS<NA>& s = S[w % E]; int c = w / E;
    if (c >= s.ncon) continue;
    c_res(s, c) = contact_visit(s, m, c, false);
  }
}

// Exact solve of an island that is ONE cube against static geoms / the belt (the common case: a cube landing on, tumbling on
// or riding the belt).  The primal problem of MuJoCo's soft constraints has only the cube's 6 accelerations as unknowns:
//   min_a 1/2 (a - a0)^T M (a - a0) + sum_rows s_r(J_r a - aref_r),  s_r(x) = x^2 / (2 R_r) for x < 0, else 0   [3P]
// (M = diag(m, m, m, I, I, I): isotropic cube, angular part in the world frame; the belt dof is held at its lagged acceleration),
// whose optimality condition is the fixed point of the dual PGS above with f_r = max(-x_r, 0) / R_r.  Active-set Newton:
// rows with x < 0 at the current a form the set, (M + sum J^T J / R) a = M a0 + sum J^T aref / R is solved by a 6x6 Cholesky,
// repeat until the set is stable (rows that still flip do so within `tol`).  Up to 8 contacts (32 rows, one bit each).
// Returns false (forces stored, acc made consistent with them) if the set did not settle, and the caller sweeps.
// one pass over the island's contacts at acceleration a: active set, forces f = max(-x, 0) / R (stored), belt projection;
// ASSEMBLE adds the active rows to H (packed lower triangle) and g
template <int NA, bool ASSEMBLE>
FM_HD unsigned newton_cube_pass(EnvS<NA>& s, int c0, int c1, f3 xk, float ab, const float* a, unsigned act, float& chg, float* H, float* g) {
  unsigned nact = 0; chg = 0.f;
  const f3 al = mk3(a[0], a[1], a[2]), aa = mk3(a[3], a[4], a[5]);
  for (int ci = c0; ci < c1; ++ci) {
    const int c = s.isl_list[ci];
    const int cs = s.c_side[c];
    const int sd0 = cs & FM_SIDE_MASK, sd1 = (cs >> FM_SIDE_BITS) & FM_SIDE_MASK;
    const bool cube0 = side_type(sd0) == FM_SIDE_CUBE;
    const float sg = cube0 ? -1.f : 1.f;                                  // the cube is side 0 (-) or side 1 (+)
    const bool belt = side_type(cube0 ? sd1 : sd0) == FM_SIDE_BELT;
    const f3 n = ld3(s.c_n[c]), t1 = ld3(s.c_t1[c]); const f3 t2 = cross3(n, t1);
    const f3 rv = ld3(s.c_p[c]) - xk;
    // base rows (normal, tangent 1, tangent 2) of the cube side: v = sg * (d ; r x d); belt side: the constant -sg * d.y * a_belt
    const f3 dl[3] = {n * sg, t1 * sg, t2 * sg};
    const f3 da[3] = {cross3(rv, dl[0]), cross3(rv, dl[1]), cross3(rv, dl[2])};
    float ub[3], kb[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) { kb[q] = belt ? -dl[q].y * ab : 0.f; ub[q] = dot3(dl[q], al) + dot3(da[q], aa) + kb[q]; }
    const float mu = s.c_mu[c], D = 1.f / s.c_R[c];
    float f[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int tq = 1 + (r >> 1); const float cf = (r & 1) ? -mu : mu;
      const float aref = s.c_aref[c][r];
      const float x = ub[0] + cf * ub[tq] - aref;
      const bool on = x < 0.f;
      const unsigned bit = 1u << (4 * (ci - c0) + r);
      if (on != ((act & bit) != 0)) chg = fmaxf(chg, fabsf(x));
      f[r] = on ? -D * x : 0.f;
      if (on) nact |= bit;
      if (ASSEMBLE && on) {
        const f3 vl = dl[0] + dl[tq] * cf, va = da[0] + da[tq] * cf;
        const float v[6] = {vl.x, vl.y, vl.z, va.x, va.y, va.z};
        const float cst = kb[0] + cf * kb[tq] - aref;
        int h = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float dv = D * v[i];
          g[i] -= dv * cst;
#pragma unroll
          for (int j = 0; j <= i; ++j) H[h++] += dv * v[j];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) s.c_f[c][r] = f[r];
    float bs = 0.f;
    if (belt) {   // belt-dof projection of this contact's force (other side: sign -sg)
      const float Fn = f[0] + f[1] + f[2] + f[3], F1 = mu * (f[0] - f[1]), F2 = mu * (f[2] - f[3]);
      bs = -(dl[0].y * Fn + dl[1].y * F1 + dl[2].y * F2);
    }
    c_belt(s, c) = bs;
  }
  return nact;
}

template <int NA>
FM_HD bool island_newton_cube(EnvS<NA>& s, const FmDevModel& m, int k, int c0, int c1, float tol) {
  float* acc = &s.acc[1 + 6 * k];
  float a[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) a[q] = acc[q];
  const float ab = s.acc[0];
  const float invm = m.cube_invm[k], invI = m.cube_invI[k];
  const float mass = 1.f / invm, In = 1.f / invI;
  const f3 xk = ld3(s.cube_x[k]);
  unsigned act = 0; bool ok = false;
  for (int it = 0; it < 5 && !ok; ++it) {
    FM_DBG(fm_dbg_sat[302]++;)
    float H[21], g[6] = {0.f, 0.f, -mass * m.g, 0.f, 0.f, 0.f}, chg;
#pragma unroll
    for (int q = 0; q < 21; ++q) H[q] = 0.f;
    H[0] = mass; H[2] = mass; H[5] = mass; H[9] = In; H[14] = In; H[20] = In;
    act = newton_cube_pass<NA, true>(s, c0, c1, xk, ab, a, act, chg, H, g);     // active set at a, assembled
    // 6x6 Cholesky solve H a = g (packed lower triangle, fully unrolled: registers)
    float iv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float d = H[tri(j, j)];
#pragma unroll
      for (int q = 0; q < j; ++q) d -= H[tri(j, q)] * H[tri(j, q)];
      iv[j] = rsqrtf(fmaxf(d, 1e-30f));
#pragma unroll
      for (int i = j + 1; i < 6; ++i) {
        float v = H[tri(i, j)];
#pragma unroll
        for (int q = 0; q < j; ++q) v -= H[tri(i, q)] * H[tri(j, q)];
        H[tri(i, j)] = v * iv[j];
      }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) { float v = g[i]; for (int q = 0; q < i; ++q) v -= H[tri(i, q)] * a[q]; a[i] = v * iv[i]; }
#pragma unroll
    for (int i = 5; i >= 0; --i) { float v = a[i]; for (int q = i + 1; q < 6; ++q) v -= H[tri(q, i)] * a[q]; a[i] = v * iv[i]; }
    // verification at the new a (no assembly): stable set, or only rows within tol of the switching point flipped
    unsigned nact = newton_cube_pass<NA, false>(s, c0, c1, xk, ab, a, act, chg, nullptr, nullptr);
    if (nact == act || chg < tol) { ok = true; FM_DBG(fm_dbg_sat[311 + it]++;) }
  }
  if (!ok) {   // not settled (no line search here): back to the warm start, a = a0 + Minv J^T f, and the caller sweeps
    f3 Fs = mk3(0, 0, 0), Ts = mk3(0, 0, 0);
    for (int ci = c0; ci < c1; ++ci) {
      const int c = s.isl_list[ci];
      warm_start_forces(s, c, s.c_key[c]);
      const float sg = side_type(s.c_side[c] & FM_SIDE_MASK) == FM_SIDE_CUBE ? -1.f : 1.f;
      const f3 n = ld3(s.c_n[c]), t1 = ld3(s.c_t1[c]); const f3 t2 = cross3(n, t1);
      const float* f = s.c_f[c]; const float mu = s.c_mu[c];
      const f3 F = (n * (f[0] + f[1] + f[2] + f[3]) + t1 * (mu * (f[0] - f[1])) + t2 * (mu * (f[2] - f[3]))) * sg;
      Fs = Fs + F; Ts = Ts + cross3(ld3(s.c_p[c]) - xk, F);
    }
    a[0] = Fs.x * invm; a[1] = Fs.y * invm; a[2] = Fs.z * invm - m.g; a[3] = Ts.x * invI; a[4] = Ts.y * invI; a[5] = Ts.z * invI;
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) acc[q] = a[q];
  return ok;
}

// Gauss-Seidel sweeps over one island (exact GS inside the island), starting from the warm start
template <int NA>
// out of line like island_newton_warp: rarely taken, and its registers stay out of the fused loop's allocation (+ 4 %)
FM_HDN void island_pgs_sweeps(EnvS<NA>& s, const FmDevModel& m, int node, int c0, int c1, bool has_arm, bool single_row, int inner, float tol) {
  // islands that contain an arm AND contacts (grasps, crashes) are stiff and strongly coupled through the 9x9 arm inertia:
  // they get 8x the sweep budget (spurious force peaks would otherwise trip the 200 N termination, base_env.py:226-236);
  // a lone arm with only its plate equality row is solved exactly by one visit
#ifndef FM_ARM_SWEEP_FACTOR
#define FM_ARM_SWEEP_FACTOR 8
#endif
  const int cap = (has_arm && c1 > c0) ? FM_ARM_SWEEP_FACTOR * inner : ((c1 == c0 && single_row) ? 1 : inner);
  for (int it = 0; it < cap; ++it) {
    float maxd = 0.f;
    FM_DBG(fm_dbg_sweeps++; fm_dbg_visits += c1 - c0;)
    for (int a = 0; a < NA; ++a) if (s.isl_root[FM_NCUBE + a] == node) maxd = fmaxf(maxd, arm_rows_visit(s, a, 1));
    for (int ci = c0; ci < c1; ++ci) maxd = fmaxf(maxd, contact_visit(s, m, s.isl_list[ci], true));
    if (maxd < tol) break;
  }
}

// phase: constraint solve, one lane per constraint island.  Islands whose warm start already satisfies the stopping
// criterion (every row's update below tol, the steady state of resting contacts) are left as they are.  One cube against static
// geoms / the belt: exact active-set Newton on the lane (island_newton_cube).  Islands that couple several nodes or an arm with
// contacts (grasps, crashes, cube on cube) are only MARKED here (bit 8 + node of need_union) and solved in the next interval by a
// whole warp each (ph_big_islands).  What is left (a lone arm's limit rows) is swept here.
#define FM_BIG_BIT(node) (1 << (8 + (node)))
#define FM_BIG_LARGE (1 << 24)   // some marked island of the env has more than FM_SLOT_NU unknowns: it needs the env's whole scratch
#define FM_SLOT_NU 15            // an arm and at most one cube
template <int NA, int E, bool BIGN>
FM_PH void ph_pgs(EnvS<NA>* S, const FmDevModel& m, int tid, int T, int inner, float tol) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  const int NN = EnvS<NA>::NNODE;
  for (int w = tid; w < NN * E; w += T) {
    EnvS<NA>& s = S[w % E]; int node = w / E;
    if (s.isl_root[node] != node) continue;
    int c0 = s.isl_start[node], c1 = s.isl_start[node + 1];
    bool has_arm = false, single_row = true;
    for (int a = 0; a < NA; ++a) if (s.isl_root[FM_NCUBE + a] == node) { has_arm = true; single_row = single_row && s.arow_n[a] == 1; }
    if (c0 == c1 && !has_arm) continue;
    float maxd = 0.f;
    for (int a = 0; a < NA; ++a) if (s.isl_root[FM_NCUBE + a] == node) maxd = fmaxf(maxd, arm_rows_visit(s, a, 2));
    for (int ci = c0; ci < c1; ++ci) maxd = fmaxf(maxd, c_res(s, s.isl_list[ci]));
    FM_DBG(fm_dbg_islands++;)
    FM_DBG(int dbg_kind = has_arm ? (c1 > c0 ? 2 : 1) : 0; fm_dbg_cnt[4 + 4 * dbg_kind]++; fm_dbg_cnt[7 + 4 * dbg_kind] += c1 - c0;)
    // a lone cube is solved exactly by the lane's Newton at no extra cost to the block (lock step: some lane runs it in nearly every
    // sub-step anyway), so its warm start only counts as converged far below the sweeps' stopping criterion: resting cubes then keep
    // |v| < 1e-7 m/s instead of drifting inside the tolerance band
    const bool newton_ok = !has_arm && node < FM_NCUBE && c1 - c0 <= 8;
    if (maxd < (newton_ok ? 0.02f * tol : tol)) continue;
    bool alone = true;                                       // no other node was merged into this island
    if (s.need_union & 1) for (int i = 0; i < NN; ++i) alone = alone && (i == node || s.isl_root[i] != node);
    bool solved = false;
    if (newton_ok && alone) {
#ifndef FM_NO_NEWTON
      solved = island_newton_cube(s, m, node, c0, c1, tol);
#endif
      FM_DBG(fm_dbg_sat[300]++; fm_dbg_sat[301] += solved;)
    }
    if (solved) continue;
    // a warp takes the stiff islands (an arm with contacts: 32 sweeps would not converge them); cube-on-cube islands converge in a few sweeps
    if (BIGN && has_arm && c1 > c0) {
      int nu = 0;
      for (int i = 0; i < NN; ++i) if (s.isl_root[i] == node) nu += i >= FM_NCUBE ? 9 : 6;
      FM_ATOMIC_OR(&s.need_union, FM_BIG_BIT(node) | (nu > FM_SLOT_NU ? FM_BIG_LARGE : 0));
      continue;
    }
    island_pgs_sweeps(s, m, node, c0, c1, has_arm, single_row, inner, tol);
  }
}

// =====================================================================================================================
// Big islands: an arm with contacts (grasp, crash), cube on cube, arm against arm.  Gauss-Seidel on the dual converges slowly on
// them (a grasp needs all 32 sweeps and still leaves 8 m/s^2) and runs on ONE lane while the block waits.  Here a whole
// warp solves the island's primal problem exactly, like island_newton_cube does for a lone cube:
//   min_a 1/2 (a - a0)^T M (a - a0) + sum_rows s_r(J_r a - aref_r),  s_r(x) = x^2 / (2 R_r) for x < 0 (always, for equality rows)
// over the island's <= 24 accelerations (9 per arm, 6 per cube; the belt dof is held at its lagged acceleration).  Active-set
// Newton: the rows with x < 0 at the current a form the set; (M + sum J^T J / R) a = M a0 + sum J^T aref / R by a Jacobi-scaled
// Cholesky factorisation; repeat until the set is stable.  The warp works in stages separated by __syncwarp(); everything that
// crosses a stage lives in shared memory that is dead at this point of the sub-step (the Newton-Euler scratch d_f, d_n, d_comp and
// the raw-contact keys), so the same source runs in the host-compiled test build with a loop over the 32 lanes per stage.
#define FM_BIG_NU 24
#define FM_BIG_AW 8    // words of the active-set bit sets: 4 rows per contact + the arm rows <= 256
// generalised force of arm dof d without constraints: actuator - bias (same expressions as ph_smooth)
template <int NA>
FM_HD float arm_qfrc(const EnvS<NA>& s, const FmDevModel& m, int a, int d) {
  if (d < 7) {
    float c = clampf(s.ctrl_target[1 + 8 * a + d], m.ctrl_lo[1 + d], m.ctrl_hi[1 + d]);
    return m.kp * (c - s.arm_q[a][d]) - m.kd * s.arm_v[a][d] - s.bias[a][d];
  }
  float c = clampf(s.ctrl_target[1 + 8 * a + 7], m.ctrl_lo[8], m.ctrl_hi[8]);
  float len = m.ten0 * s.arm_q[a][7] + m.ten1 * s.arm_q[a][8], lv = m.ten0 * s.arm_v[a][7] + m.ten1 * s.arm_v[a][8];
  float f = clampf(m.g_kp * c - m.g_kb * len - m.g_kd * lv, m.g_fmin, m.g_fmax);
  return (d == 7 ? m.ten0 : m.ten1) * f - s.bias[a][d];
}

// islands of one env that can be solved side by side, and the raw-contact data words each of them gets as staging
template <int NA> struct BigSlots {
  static const int N = NA == 2 ? 2 : 3, SL = FmCap<NA>::CON * 7 / N;
  static_assert(SL >= 12 * (FM_SLOT_NU + 1) && (2 * FM_BIG_AW + 1) * N <= FmCap<NA>::CON, "slot staging does not fit");
};
template <int NA>
struct BigIsl {
  EnvS<NA>& s; const FmDevModel& m;
  int nnode, nodes[4], off[5], NU, c0, nc, nrow_arm, arm_row0[4];
  float *H, *g, *a, *Jc, *Wg; unsigned *act, *prev; int* chg;
  FM_HDM BigIsl(EnvS<NA>& s_, const FmDevModel& m_) : s(s_), m(m_) {}
  // unknown j -> (index into nodes, local dof)
  FM_HDM void unk(int j, int& ni, int& q) const { ni = 0; while (ni + 1 < nnode && off[ni + 1] <= j) ++ni; q = j - off[ni]; }
  FM_HDM int acc_index(int ni, int q) const { int nd = nodes[ni]; return nd < FM_NCUBE ? 1 + 6 * nd + q : 61 + 9 * (nd - FM_NCUBE) + q; }
  FM_HDM int node_slot(int node) const { for (int i = 0; i < nnode; ++i) if (nodes[i] == node) return i; return -1; }
  // column j of the contact's base Jacobian (rows n, t1, t2): sum over the contact's two sides
  FM_HDM void jcol(int c, int j, float* out) const {
    out[0] = out[1] = out[2] = 0.f;
    int ni, q; unk(j, ni, q);
    const int node = nodes[ni];
    const int cs = s.c_side[c];
    const f3 n = ld3(s.c_n[c]), t1 = ld3(s.c_t1[c]); const f3 t2 = cross3(n, t1);
    for (int sd_i = 0; sd_i < 2; ++sd_i) {
      const int sd = (cs >> (FM_SIDE_BITS * sd_i)) & FM_SIDE_MASK;
      const int ty = side_type(sd), ix = side_idx(sd);
      const float sg = sd_i ? 1.f : -1.f;
      if (ty == FM_SIDE_CUBE && ix == node) {
        const f3 r = ld3(s.c_p[c]) - ld3(s.cube_x[ix]);
        const f3 d[3] = {n, t1, t2};
        for (int k = 0; k < 3; ++k) { f3 v = q < 3 ? d[k] : cross3(r, d[k]); out[k] += sg * comp3(v, q % 3); }
      } else if (ty == FM_SIDE_ARM && FM_NCUBE + (ix >> 4) == node) {
        const float* J = s.as_J[side_slot(sd)];
        for (int k = 0; k < 3; ++k) out[k] += sg * J[9 * k + q];
      }
    }
  }
  // contact-space acceleration u = J_c a + (belt part) of contact c at the island accelerations `av`
  FM_HDM void contact_u(int c, const float* av, float* u) const {
    belt_const(c, u);
    const int cs = s.c_side[c];
    const f3 n = ld3(s.c_n[c]), t1 = ld3(s.c_t1[c]); const f3 t2 = cross3(n, t1);
    for (int sd_i = 0; sd_i < 2; ++sd_i) {
      const int sd = (cs >> (FM_SIDE_BITS * sd_i)) & FM_SIDE_MASK;
      const int ty = side_type(sd), ix = side_idx(sd);
      const float sg = sd_i ? 1.f : -1.f;
      if (ty == FM_SIDE_CUBE) {
        const int ni = node_slot(ix); if (ni < 0) continue;
        const float* v = av + off[ni];
        const f3 pv = mk3(v[0], v[1], v[2]) + cross3(mk3(v[3], v[4], v[5]), ld3(s.c_p[c]) - ld3(s.cube_x[ix]));
        u[0] += sg * dot3(n, pv); u[1] += sg * dot3(t1, pv); u[2] += sg * dot3(t2, pv);
      } else if (ty == FM_SIDE_ARM) {
        const int ni = node_slot(FM_NCUBE + (ix >> 4)); if (ni < 0) continue;
        const float* v = av + off[ni]; const float* J = s.as_J[side_slot(sd)];
        for (int d = 0; d < 3; ++d) { float acc = 0.f; for (int k = 0; k < 9; ++k) acc += J[9 * d + k] * v[k]; u[d] += sg * acc; }
      }
    }
  }
  // constant part of the contact-space acceleration: the belt side (dof held at its lagged acceleration)
  FM_HDM void belt_const(int c, float* kb) const {
    kb[0] = kb[1] = kb[2] = 0.f;
    const int cs = s.c_side[c];
    for (int sd_i = 0; sd_i < 2; ++sd_i) {
      const int sd = (cs >> (FM_SIDE_BITS * sd_i)) & FM_SIDE_MASK;
      if (side_type(sd) != FM_SIDE_BELT) continue;
      const f3 n = ld3(s.c_n[c]), t1 = ld3(s.c_t1[c]); const f3 t2 = cross3(n, t1);
      const float sg = sd_i ? 1.f : -1.f, ab = s.acc[0];
      kb[0] += sg * n.y * ab; kb[1] += sg * t1.y * ab; kb[2] += sg * t2.y * ab;
    }
  }
};

// slot < 0: the island has the env's whole scratch; slot >= 0: the env's marked islands (all <= FM_SLOT_NU unknowns) run on different
// warps at the same time, each in its own part of the scratch
// out of line: its register needs stay out of the fused loop's allocation (+ 3 % on the headline, + 11 % in IK-action mode)
template <int NA, class Lanes>
FM_HDN bool island_newton_warp(EnvS<NA>& s, const FmDevModel& m, int root, Lanes& L, float tol, int slot) {
  const int NN = EnvS<NA>::NNODE;
  BigIsl<NA> B(s, m);
  typename Lanes::template Arr<int, 5> eij; typename Lanes::template Arr<float, 5> acc;   // per-lane registers that cross stages
  B.nnode = 0; B.NU = 0; B.nrow_arm = 0;
  for (int i = 0; i < NN; ++i)
    if (s.isl_root[i] == root) {
      if (B.nnode == 4) return false;
      B.nodes[B.nnode] = i; B.off[B.nnode] = B.NU; B.NU += i >= FM_NCUBE ? 9 : 6;
      if (i >= FM_NCUBE) { B.arm_row0[B.nnode] = B.nrow_arm; B.nrow_arm += s.arow_n[i - FM_NCUBE]; } else B.arm_row0[B.nnode] = 0;
      B.nnode++;
    }
  B.off[B.nnode] = B.NU;
  B.c0 = s.isl_start[root]; B.nc = s.isl_start[root + 1] - B.c0;
  const int NU = B.NU, nc = B.nc, nrows = 4 * nc + B.nrow_arm;
  FM_DBG(if (NU > FM_BIG_NU) fm_dbg_sat[335]++; else if (nrows > 32 * FM_BIG_AW) fm_dbg_sat[336]++;)
  if (NU > (slot < 0 ? FM_BIG_NU : FM_SLOT_NU) || nrows > 32 * FM_BIG_AW) return false;
  // scratch.  H: packed lower triangle of the (NU + 1) x (NU + 1) matrix [H g; g^T .] — lane i owns row i, the right-hand side is
  // row NU, so the forward substitution falls out of the factorisation; a: NU accelerations (Jacobi scale factors meanwhile).
  // Both sit in the dead Newton-Euler buffers (d_f, d_n, d_comp: 176 floats per arm); the contact staging (JS floats per Jacobian
  // row) in the dead raw-contact data, the bit sets behind it or (slots) in the previous solve's keys, which are dead between the
  // warm-start look-up of ph_contact_setup and ph_save_forces.
  const int JS = slot < 0 ? FM_BIG_NU : FM_SLOT_NU + 1;
  float* rd = &s.raw_d[0][0];
  float* small;                                               // act, prev (FM_BIG_AW words each), chg
  if (slot < 0) { B.H = &s.d_f[0][0][0]; B.a = B.H + 326; B.Jc = rd; B.Wg = rd + 3 * FM_BIG_NU; small = B.Wg + 9; }
  else {
    B.H = &s.d_f[0][0][0] + 176 * slot; B.a = B.H + 152;
    B.Jc = rd + BigSlots<NA>::SL * slot; B.Wg = nullptr;      // two buffers of (J, W J) columns, 6 x JS each
    small = reinterpret_cast<float*>(&s.prev_key[0]) + (2 * FM_BIG_AW + 1) * slot;
  }
  B.g = B.H + tri(NU, 0);
  B.act = reinterpret_cast<unsigned*>(small); B.prev = B.act + FM_BIG_AW; B.chg = reinterpret_cast<int*>(B.prev + FM_BIG_AW);
  const int nitem = nc + B.nrow_arm;                          // active-set work items: a contact (4 rows) or an arm row
  L.stage([&](int lane) {
    if (lane < NU) { int ni, q; B.unk(lane, ni, q); B.a[lane] = s.acc[B.acc_index(ni, q)]; }
    if (lane < FM_BIG_AW) { B.act[lane] = 0u; B.prev[lane] = 0xFFFFFFFFu; }
    if (lane == 0) *B.chg = 0;
  });
  bool ok = false;
  for (int it = 0; it < 7; ++it) {
    FM_DBG(fm_dbg_sat[331]++;)
    // ---- active set at the current a: one work item per contact (4 rows) or per arm row
    L.stage([&](int lane) {
      for (int item = lane; item < nitem; item += 32) {
        if (item < nc) {
          const int c = s.isl_list[B.c0 + item];
          float u[3]; B.contact_u(c, B.a, u);
          const float mu = s.c_mu[c];
          unsigned bits = 0; float mx = 0.f;
          const bool warm = s.c_f[c][0] > 0.f || s.c_f[c][1] > 0.f || s.c_f[c][2] > 0.f || s.c_f[c][3] > 0.f;   // matched by the warm-start key
          for (int r = 0; r < 4; ++r) {
            const int tq = 1 + (r >> 1); const float cf = (r & 1) ? -mu : mu;
            const float x = u[0] + cf * u[tq] - s.c_aref[c][r];
            // first pass: a row that carried force in the previous solve stays in the set.  At the warm start x = -f R, and R is
            // 1e-6..1e-8: its sign is noise against the sub-step's change of aref, so the evaluated set would be half wrong
            const bool on = (it == 0 && warm) ? s.c_f[c][r] > 0.f : x < 0.f;
            const int bit = 4 * item + r;
            if (on) bits |= 1u << (bit & 31);
            if (on != (((B.prev[bit >> 5] >> (bit & 31)) & 1u) != 0u)) mx = fmaxf(mx, fabsf(x));
          }
          if (bits) FM_ATOMIC_OR(reinterpret_cast<int*>(&B.act[(4 * item) >> 5]), (int)bits);
          if (mx > 0.f) { int mi; memcpy(&mi, &mx, 4); FM_ATOMIC_MAXI(B.chg, mi); }
        } else {
          int rr = item - nc, ni = 0;
          while (!(B.nodes[ni] >= FM_NCUBE && rr >= B.arm_row0[ni] && rr < B.arm_row0[ni] + s.arow_n[B.nodes[ni] - FM_NCUBE])) ++ni;
          const int a = B.nodes[ni] - FM_NCUBE, r = rr - B.arm_row0[ni];
          const int code = s.arow_code[a][r];
          const float* aa = B.a + B.off[ni];
          float x; bool on;
          if (code & 0x1000) { x = aa[7] - aa[8] - s.arow_aref[a][r]; on = true; }
          else { const int d = code & 255; const float sg = (code >> 8) ? -1.f : 1.f; x = sg * aa[d] - s.arow_aref[a][r]; on = x < 0.f || (it == 0 && s.arow_f[a][r] > 0.f); }
          const int bit = 4 * nc + rr;
          if (on) FM_ATOMIC_OR(reinterpret_cast<int*>(&B.act[bit >> 5]), (int)(1u << (bit & 31)));
          if (on != (((B.prev[bit >> 5] >> (bit & 31)) & 1u) != 0u)) { float ax = fabsf(x); int mi; memcpy(&mi, &ax, 4); FM_ATOMIC_MAXI(B.chg, mi); }
        }
      }
    });
    {
      bool same = true;
      for (int w = 0; w < FM_BIG_AW; ++w) same = same && B.act[w] == B.prev[w];
      float cg; memcpy(&cg, B.chg, 4);
      L.sync();                                               // every lane has read act / prev / chg before the next stage rewrites them
      if (it > 0 && (same || cg < tol)) { ok = true; break; }
      if (it == 6) break;
    }
    // ---- rows of [H g]: the nodes' inertia blocks, M a0, and the arm-local rows (one or two non-zeros each); lane i owns row i
    // of H and entry i of the right-hand side row g_i = (M a0)_i (+ the active arm rows' J^T aref / R)
    L.stage([&](int lane) {
      if (lane < FM_BIG_AW) B.prev[lane] = B.act[lane];
      if (lane == 0) *B.chg = 0;
      if (lane > NU) return;
      if (lane == NU) { B.g[NU] = 0.f; return; }
      float* row = B.H + tri(lane, 0);
      int ni, qi; B.unk(lane, ni, qi);
      const int nd = B.nodes[ni], o = B.off[ni];
      for (int j = 0; j <= lane; ++j) {
        float v = 0.f;
        if (j >= o) v = nd < FM_NCUBE ? (j == lane ? (qi < 3 ? 1.f / m.cube_invm[nd] : 1.f / m.cube_invI[nd]) : 0.f) : s.M[nd - FM_NCUBE][tri(qi, j - o)];
        row[j] = v;
      }
      float gi = nd < FM_NCUBE ? (qi == 2 ? -m.g / m.cube_invm[nd] : 0.f) : arm_qfrc(s, m, nd - FM_NCUBE, qi);
      if (nd >= FM_NCUBE) {                                  // arm rows touching this dof
        const int a = nd - FM_NCUBE;
        for (int r = 0; r < s.arow_n[a]; ++r) {
          const int bit = 4 * nc + B.arm_row0[ni] + r;
          if (!((B.act[bit >> 5] >> (bit & 31)) & 1u)) continue;
          const int code = s.arow_code[a][r]; const float D = 1.f / s.arow_R[a][r], ar = s.arow_aref[a][r];
          if (code & 0x1000) {
            if (qi == 7) { row[lane] += D; gi += D * ar; }
            else if (qi == 8) { row[lane] += D; row[lane - 1] -= D; gi -= D * ar; }
          } else if ((code & 255) == qi) { row[lane] += D; gi += ((code >> 8) ? -1.f : 1.f) * D * ar; }
        }
      }
      B.g[lane] = gi;
    });
    // ---- contacts: H += J_c^T W_c J_c, g += J_c^T g_c
    if (slot >= 0) {
      // (NU + 1) <= 16 rows: the packed triangle's <= 136 entries are dealt to the 32 lanes (entry e -> lane e % 32) and accumulated in
      // registers over the contacts; per contact lane i < NU stages column i of J_c and of W_c J_c (lane NU: 0 and g_c) in one of two
      // buffers, so one __syncwarp per contact suffices and the next contact's column work overlaps this one's accumulation
      const int TE = tri(NU + 1, 0);
      constexpr int JSS = FM_SLOT_NU + 1;
      L.run([&](int lane) {
        int* ij = eij(lane); float* ac = acc(lane);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int e = lane + 32 * k;
          int i = NU, j = NU;
          if (e < TE) { i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f); while (tri(i + 1, 0) <= e) ++i; while (tri(i, 0) > e) --i; j = e - tri(i, 0); }
          ij[k] = (i << 8) | j; ac[k] = 0.f;
        }
      });
      for (int ci = 0; ci < nc; ++ci) {
        const int c = s.isl_list[B.c0 + ci];
        float* bj = B.Jc + 6 * JSS * (ci & 1);
        L.stage([&](int lane) {
          if (lane > NU) return;
          float W00 = 0.f, W11 = 0.f, W22 = 0.f, W01 = 0.f, W02 = 0.f, gc[3] = {0, 0, 0}, kb[3];   // the pyramid has no (t1, t2) term
          B.belt_const(c, kb);
          const float mu = s.c_mu[c], D = 1.f / s.c_R[c];
          const unsigned bits = (B.act[(4 * ci) >> 5] >> ((4 * ci) & 31)) & 15u;
          for (int r = 0; r < 4; ++r) {
            if (!((bits >> r) & 1u)) continue;
            const int tq = 1 + (r >> 1); const float cf = (r & 1) ? -mu : mu;
            const float rhs = D * (s.c_aref[c][r] - kb[0] - cf * kb[tq]);
            W00 += D; if (tq == 1) { W11 += D * cf * cf; W01 += D * cf; } else { W22 += D * cf * cf; W02 += D * cf; }
            gc[0] += rhs; gc[tq] += cf * rhs;
          }
          float i0 = 0.f, i1 = 0.f, i2 = 0.f, t0 = gc[0], t1 = gc[1], t2 = gc[2];
          if (lane < NU) {
            float jc[3]; B.jcol(c, lane, jc); i0 = jc[0]; i1 = jc[1]; i2 = jc[2];
            t0 = W00 * i0 + W01 * i1 + W02 * i2; t1 = W01 * i0 + W11 * i1; t2 = W02 * i0 + W22 * i2;
          }
          bj[lane] = i0; bj[JS + lane] = i1; bj[2 * JS + lane] = i2; bj[3 * JS + lane] = t0; bj[4 * JS + lane] = t1; bj[5 * JS + lane] = t2;
        });
        L.run([&](int lane) {
          const int* ij = eij(lane); float* ac = acc(lane);
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const float* bi = bj + (ij[k] >> 8); const float* bq = bj + (ij[k] & 255);
            ac[k] += bi[3 * JSS] * bq[0] + bi[4 * JSS] * bq[JSS] + bi[5 * JSS] * bq[2 * JSS];
          }
        });
      }
      L.stage([&](int lane) {
        const float* ac = acc(lane);
#pragma unroll
        for (int k = 0; k < 5; ++k) { const int e = lane + 32 * k; if (e < TE) B.H[e] += ac[k]; }
      });
    } else
    // whole-scratch islands (rare): one contact after the other, columns staged by the lanes, W_c by lane 31, lane i updates row i
    for (int ci = 0; ci < nc; ++ci) {
      const int c = s.isl_list[B.c0 + ci];
      FM_DBG(fm_dbg_sat[338]++; if ((B.act[(4 * ci) >> 5] >> ((4 * ci) & 31)) & 15u) fm_dbg_sat[339]++;)
      L.stage([&](int lane) {
        if (lane < NU) { float jc[3]; B.jcol(c, lane, jc); B.Jc[lane] = jc[0]; B.Jc[JS + lane] = jc[1]; B.Jc[2 * JS + lane] = jc[2]; }
        if (lane == 31) {      // 3x3 weight of the contact's active pyramid rows and the matching right-hand side
          float W[6] = {0, 0, 0, 0, 0, 0}, gc[3] = {0, 0, 0}, kb[3];
          B.belt_const(c, kb);
          const float mu = s.c_mu[c], D = 1.f / s.c_R[c];
          for (int r = 0; r < 4; ++r) {
            const int bit = 4 * ci + r;
            if (!((B.act[bit >> 5] >> (bit & 31)) & 1u)) continue;
            const int tq = 1 + (r >> 1); const float cf = (r & 1) ? -mu : mu;
            const float rhs = D * (s.c_aref[c][r] - kb[0] - cf * kb[tq]);
            W[0] += D; W[tq] += D * cf * cf; W[2 + tq] += D * cf;      // (0,0) (tq,tq) (0,tq); W = [00, 11, 22, 01, 02, 12]
            gc[0] += rhs; gc[tq] += cf * rhs;
          }
          for (int k = 0; k < 6; ++k) B.Wg[k] = W[k];
          for (int k = 0; k < 3; ++k) B.Wg[6 + k] = gc[k];
        }
      });
      L.stage([&](int lane) {
        if (lane > NU) return;
        float t0, t1, t2;
        if (lane == NU) { t0 = B.Wg[6]; t1 = B.Wg[7]; t2 = B.Wg[8]; }           // g row: J_c^T g_c
        else {
          const float i0 = B.Jc[lane], i1 = B.Jc[JS + lane], i2 = B.Jc[2 * JS + lane];
          if (i0 == 0.f && i1 == 0.f && i2 == 0.f) return;
          const float W00 = B.Wg[0], W11 = B.Wg[1], W22 = B.Wg[2], W01 = B.Wg[3], W02 = B.Wg[4];   // the pyramid has no (t1, t2) term
          t0 = W00 * i0 + W01 * i1 + W02 * i2; t1 = W01 * i0 + W11 * i1; t2 = W02 * i0 + W22 * i2;
        }
        float* row = B.H + tri(lane, 0);
        const int jn = lane == NU ? NU : lane + 1;
        for (int j = 0; j < jn; ++j) row[j] += t0 * B.Jc[j] + t1 * B.Jc[JS + j] + t2 * B.Jc[2 * JS + j];
      });
    }
    // ---- Jacobi scaling (the arm's inertias span three decades): S = diag(H)^-1/2, [H g] <- S [H g] S
    L.stage([&](int lane) { if (lane < NU) B.a[lane] = rsqrtf(fmaxf(B.H[tri(lane, lane)], 1e-30f)); });
    L.stage([&](int lane) {
      if (lane > NU) return;
      float* row = B.H + tri(lane, 0);
      const float si = lane == NU ? 1.f : B.a[lane];
      const int jn = lane == NU ? NU : lane + 1;
      for (int j = 0; j < jn; ++j) row[j] *= si * B.a[j];
    });
    // ---- left-looking Cholesky, one stage per column, lane i = row i (row NU carries the right-hand side: y = L^-1 g).  The
    // factor's diagonal goes to dg (the contact staging is dead now): the unfactored H[k][k] must stay readable in this stage
    float* dg = B.Jc; float* xs = B.Jc + JS;
    for (int k = 0; k < NU; ++k) {
      L.stage([&](int lane) {
        if (lane < k || lane > NU) return;
        const float* rk = B.H + tri(k, 0); float* ri = B.H + tri(lane, 0);
        float piv = rk[k], v = ri[k];
        for (int j = 0; j < k; ++j) { const float lkj = rk[j]; piv -= lkj * lkj; v -= ri[j] * lkj; }
        const float rp = rsqrtf(fmaxf(piv, 1e-30f));       // every lane recomputes the pivot from row k's finished part (columns < k)
        if (lane == k) dg[k] = fmaxf(piv, 1e-30f) * rp; else ri[k] = v * rp;
      });
    }
    // ---- back substitution L^T x = y, column-oriented
    for (int k = NU - 1; k >= 0; --k) {
      L.stage([&](int lane) {
        if (lane > k) return;
        const float xk = B.g[k] / dg[k];
        if (lane == k) xs[k] = xk; else B.g[lane] -= B.H[tri(k, lane)] * xk;
      });
    }
    L.stage([&](int lane) { if (lane < NU) B.a[lane] *= xs[lane]; if (lane < FM_BIG_AW) B.act[lane] = 0u; });
  }
  FM_DBG(fm_dbg_sat[332] += ok; fm_dbg_sat[340 + (nc < 24 ? nc : 24)]++; fm_dbg_sat[370 + NU]++;)
  if (!ok) return false;
  // ---- accept: accelerations, forces f = max(-x, 0) / R of the final a, belt projections
  L.stage([&](int lane) {
    if (lane < NU) { int ni, q; B.unk(lane, ni, q); s.acc[B.acc_index(ni, q)] = B.a[lane]; }
    for (int item = lane; item < nitem; item += 32) {
      if (item < nc) {
        const int c = s.isl_list[B.c0 + item];
        float u[3]; B.contact_u(c, B.a, u);
        const float mu = s.c_mu[c], D = 1.f / s.c_R[c];
        float f[4];
        for (int r = 0; r < 4; ++r) {
          const int tq = 1 + (r >> 1); const float cf = (r & 1) ? -mu : mu;
          const float x = u[0] + cf * u[tq] - s.c_aref[c][r];
          f[r] = x < 0.f ? -D * x : 0.f;
          s.c_f[c][r] = f[r];
        }
        float bs = 0.f;
        const int cs = s.c_side[c];
        for (int sd_i = 0; sd_i < 2; ++sd_i) {
          const int sd = (cs >> (FM_SIDE_BITS * sd_i)) & FM_SIDE_MASK;
          if (side_type(sd) != FM_SIDE_BELT) continue;
          const f3 n = ld3(s.c_n[c]), t1 = ld3(s.c_t1[c]); const f3 t2 = cross3(n, t1);
          const float Fn = f[0] + f[1] + f[2] + f[3], F1 = mu * (f[0] - f[1]), F2 = mu * (f[2] - f[3]);
          bs += (sd_i ? 1.f : -1.f) * (n.y * Fn + t1.y * F1 + t2.y * F2);
        }
        c_belt(s, c) = bs;
      } else {
        int rr = item - nc, ni = 0;
        while (!(B.nodes[ni] >= FM_NCUBE && rr >= B.arm_row0[ni] && rr < B.arm_row0[ni] + s.arow_n[B.nodes[ni] - FM_NCUBE])) ++ni;
        const int a = B.nodes[ni] - FM_NCUBE, r = rr - B.arm_row0[ni];
        const int code = s.arow_code[a][r];
        const float* aa = B.a + B.off[ni];
        const float D = 1.f / s.arow_R[a][r];
        if (code & 0x1000) s.arow_f[a][r] = -D * (aa[7] - aa[8] - s.arow_aref[a][r]);
        else { const int d = code & 255; const float sg = (code >> 8) ? -1.f : 1.f; const float x = sg * aa[d] - s.arow_aref[a][r]; s.arow_f[a][r] = x < 0.f ? -D * x : 0.f; }
      }
    }
  });
  return true;
}

// phase: one warp per marked island (see ph_pgs); an island the Newton cannot take or settle falls back to the sweeps on one lane
FM_HD int popc_u(unsigned v) { int n = 0; for (; v; v &= v - 1) ++n; return n; }
template <int NA, int E, class Lanes>
FM_PH void ph_big_islands(EnvS<NA>* S, const FmDevModel& m, int warp, int nwarps, Lanes& L, int inner, float tol) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  const int NN = EnvS<NA>::NNODE;
  const int NSLOT = BigSlots<NA>::N;
  // work items in a fixed order, item i -> warp i % nwarps.  The marked islands of an env are separate items (different warps, own
  // scratch slot each) when they all are small and few enough; otherwise the env is ONE item and its islands run one after the other.
  int item = 0;
  for (int e = 0; e < E; ++e) {
    EnvS<NA>& s = S[e];
    const int nu = s.need_union;
    const unsigned big = ((unsigned)nu >> 8) & ((1u << NN) - 1u);
    if (!big) continue;
    const int nb = popc_u(big);
    const bool slots = nb <= NSLOT && !(nu & FM_BIG_LARGE), split = slots && nb > 1;   // slots: every marked island is small
    int slot = 0;
    for (int node = 0; node < NN; ++node) {
      if (!((big >> node) & 1u)) continue;
      const bool mine = (item % nwarps) == warp;
      if (split) { ++item; }
      if (mine) {
        FM_DBG(fm_dbg_sat[330]++;)
        const bool ok = island_newton_warp<NA>(s, m, node, L, tol, slots ? slot : -1);
        if (!ok) {
          FM_DBG(fm_dbg_sat[333]++; long long sw0 = fm_dbg_sweeps;)
          L.stage([&](int lane) {
            if (lane != 0) return;
            const int c0 = s.isl_start[node], c1 = s.isl_start[node + 1];
            bool has_arm = false;
            for (int a = 0; a < NA; ++a) has_arm = has_arm || s.isl_root[FM_NCUBE + a] == node;
            island_pgs_sweeps(s, m, node, c0, c1, has_arm, false, inner, tol);
          });
          FM_DBG(fm_dbg_sat[334] += fm_dbg_sweeps - sw0;)
        }
      }
      ++slot;
    }
    if (!split) ++item;
  }
}

// belt acceleration from this solve's contact forces (fixed summation order: run-to-run determinism).  During the solve the
// belt dof is held at its smooth acceleration plus the previous sub-step's contact term (the coupling is ~1e-3 of a cube's).
template <int NA, int E>
FM_PH void ph_belt(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  for (int w = tid; w < E; w += T) {
    EnvS<NA>& s = S[w];
    float sum = 0.f;
    for (int c = 0; c < s.ncon; ++c) sum += c_belt(s, c);
    s.belt_S = sum / m.belt_mass;
    s.acc[0] = s.belt_acc_smooth + s.belt_S;
  }
}

// phase: force-termination statistic on the forces of this solve (base_env.py:226-236; clarification SURVEY.md A.2)
template <int NA, int E>
FM_PH void ph_force_check(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  for (int w = tid; w < FmCap<NA>::CON * E; w += T) {
    EnvS<NA>& s = S[w % E]; int c = w / E;
    if (c >= s.ncon || !((s.c_side[c] >> 30) & 1)) continue;
    const float* f = s.c_f[c]; float mu = s.c_mu[c];
    float v = fmaxf(fabsf(f[0] + f[1] + f[2] + f[3]), fmaxf(fabsf(mu * (f[0] - f[1])), fabsf(mu * (f[2] - f[3]))));
    // float max via int compare (non-negative floats order like ints)
    int vi; memcpy(&vi, &v, 4);
    FM_ATOMIC_MAXI((int*)&s.force_max, vi);
  }
}

// phase: implicitfast velocity update + semi-implicit position integration (mj_implicit, mj_advance) [3P]
template <int NA, int E>
FM_PH void ph_integrate(EnvS<NA>* S, const FmDevModel& m, int tid, int T) {
  FM_ASSUME_SHARED(S); FM_ASSUME_GLOBAL(&m);
  const int per = 1 + FM_NCUBE + 9 * NA;
  for (int w = tid; w < per * E; w += T) {
    EnvS<NA>& s = S[w % E]; int i = w / E;
    float h = m.h;
    if (i == 0) {
      float a = s.acc[0];
      float ai = m.belt_mass * a / (m.belt_mass + h * (m.belt_kv + m.belt_damp));
      s.belt_v += h * ai; s.belt_q += h * s.belt_v;
    } else if (i <= FM_NCUBE) {
      int k = i - 1;
      if (s.cube_sleep[k]) continue;
      float* a = &s.acc[1 + 6 * k]; float* v = s.cube_v[k]; float* q = s.cube_q[k];
      f3 al = rmulT(s.cube_R[k], mk3(a[3], a[4], a[5]));      // world-frame angular acceleration -> body frame (free-joint qvel convention)
      v[0] += h * a[0]; v[1] += h * a[1]; v[2] += h * a[2]; v[3] += h * al.x; v[4] += h * al.y; v[5] += h * al.z;
      // Compensated position update (Fast2Sum): a float near 1.13 m resolves 1.2e-7 m, but a resting cube's penetration is
      // 2.3e-6 m and its per-sub-step travel h*v far below one ulp, so a plain `q += h v` freezes z up to 5e-7 m away from the
      // equilibrium with a residual velocity of 4e-5 m/s.  The increments that do not fit are carried in cube_qlo (part of the
      // persisted state; zeroed whenever the task manager or fm_set_state writes the pose).
      float* ql = s.cube_qlo[k];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float inc = h * v[c] + ql[c];
        float nq = q[c] + inc;
        ql[c] = inc - (nq - q[c]);
        q[c] = nq;
      }
      f3 wv = mk3(v[3], v[4], v[5]);
      float wn = norm3(wv);
      q4 qq = qnormf(mkq(q[3], q[4], q[5], q[6]));
      if (wn > 0.f) {
        float sn, cs; sincosf(0.5f * wn * h, &sn, &cs);
        f3 ax = wv * (sn / wn);
        qq = qmulf(qq, mkq(cs, ax.x, ax.y, ax.z));
      }
      q[3] = qq.w; q[4] = qq.x; q[5] = qq.y; q[6] = qq.z;
    } else {
#ifdef FM_INTEGRATE_SERIAL
      const int ad = i - 1 - FM_NCUBE;
      if (ad % 9 != 0) continue;
      const int a = ad / 9;
      float* acc = &s.acc[61 + 9 * a];
      float y[9];
      for (int r = 0; r < 9; ++r) { float sum = 0.f; for (int k = 0; k < 9; ++k) sum += s.M[a][r >= k ? tri(r, k) : tri(k, r)] * acc[k]; y[r] = sum; }
      chol9_solve(s.Li[a], s.chol_d[a][1], y);
      float* ql = s.arm_qlo(a);
      for (int r = 0; r < 9; ++r) {
        s.arm_v[a][r] += h * y[r];
        float q0 = s.arm_q[a][r], inc = h * s.arm_v[a][r] + ql[r];
        float nq = q0 + inc; ql[r] = inc - (nq - q0); s.arm_q[a][r] = nq;
      }
#else
      // arm dof r: qacc_r = [(M - hD)^-1 M acc]_r; every lane forms y = M acc itself (81 multiply-adds, nine independent sums)
      const int ad = i - 1 - FM_NCUBE; const int a = ad / 9, r = ad % 9;
      const float* acc = &s.acc[61 + 9 * a]; const float* Mp = s.M[a]; const float* Wp = s.Lc[a];
      float qa = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float yk = 0.f;
#pragma unroll
        for (int j = 0; j < 9; ++j) yk += Mp[k >= j ? tri(k, j) : tri(j, k)] * acc[j];
        qa += Wp[r >= k ? tri(r, k) : tri(k, r)] * yk;
      }
      const float vn = s.arm_v[a][r] + h * qa;
      s.arm_v[a][r] = vn;
      // compensated like the cube positions: a joint creeping to its target at < 6e-5 rad/s would otherwise stall 1e-5 rad short
      float* ql = s.arm_qlo(a);
      const float q0 = s.arm_q[a][r], inc = h * vn + ql[r];
      const float nq = q0 + inc;
      ql[r] = inc - (nq - q0);
      s.arm_q[a][r] = nq;
#endif
    }
  }
}

// phase: remember this solve's forces for the next warm start
template <int NA, int E>
FM_PH void ph_save_forces(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  for (int w = tid; w < (FmCap<NA>::CON + NA) * E; w += T) {
    EnvS<NA>& s = S[w % E]; int c = w / E;
    if (c < FmCap<NA>::CON) {
      if (c == 0) s.prev_n = s.ncon;
      if (c < s.ncon) { s.prev_key[c] = s.c_key[c]; for (int r = 0; r < 4; ++r) s.prev_f[c][r] = s.c_f[c][r]; }
    } else {
      int a = c - FmCap<NA>::CON;
      s.prev_arow_n[a] = s.arow_n[a];
      for (int r = 0; r < s.arow_n[a]; ++r) { s.prev_arow_code[a][r] = s.arow_code[a][r]; s.prev_arow_f[a][r] = s.arow_f[a][r]; }
    }
  }
}

// phase: PT1 control filter (base_env.py:209-215).  The command is constant over the control step, so the recursion
// ctrl_target += (ctrl - ctrl_target) * alpha has the closed form ctrl + d0 (1 - alpha)^(k+1) with d0 = ctrl_target - ctrl at the
// start of the step.  Evaluated that way because the recursion stalls in fp32 once its increment falls below half an ulp
// (1.2e-5 short of a target of 1.0: an idle arm then hangs 1e-5 rad off the reference's pose, tests/ref_cases.py).
template <int NA, int E>
FM_PH void ph_ctrl(EnvS<NA>* S, int tid, int T, float pk) {
  FM_ASSUME_SHARED(S);
  const int NU = EnvS<NA>::NU;
  for (int w = tid; w < NU * E; w += T) {
    EnvS<NA>& s = S[w % E]; int u = w / E;
    if (u == 0) { s.ctrl_target[0] = -s.conveyor_speed; }
    else s.ctrl_target[u] = s.ctrl[u] + s.ct_d0[u] * pk;
  }
}
template <int NA, int E>
FM_PH void ph_clear(EnvS<NA>* S, int tid, int T) {
  FM_ASSUME_SHARED(S);
  for (int w = tid; w < E; w += T) { EnvS<NA>& s = S[w]; s.nraw = 0; s.ncon = 0; s.need_union = 0; s.pad_[0] = 0; for (int i = 0; i < EnvS<NA>::NNODE; ++i) s.isl_root[i] = i; }
}
