// fmo_physics.hpp — TEST ORACLE (fp64, CPU).  NOT part of the product path.
//
// A restatement of what one `physics.step()` (base_env.py:218) computes for this model class:
// MuJoCo's documented pipeline mj_step2 + mj_step1 — kinematics, composite-rigid-body
// mass matrix, RNE bias forces, primitive collision (plane/sphere/box), constraint rows
// (joint equality, joint limits, pyramidal frictional contacts) with solref/solimp impedance, an exact
// convex solve of the constraint problem (Newton on the primal, as MuJoCo's default solver), and the
// implicitfast velocity update.
//
// MuJoCo (pinned 3.1.3 / 3.2.0, environment.yml:128) is a third-party dependency that is absent from
// /root/reference and cannot be installed here => PARITY UNPINNED: this restates the published algorithm
// from its documentation; every "[3P]" comment marks a convention taken from MuJoCo's docs/source memory
// that could not be verified against a running MuJoCo.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>
#include <stdexcept>
#include <string>
#include "../include/fm_blob.h"

namespace fmo {

static const double MINVAL = 1e-15;
static const double MINMU = 1e-5;
static const double BOXBOX_RSCALE = 2.0;   // see make_constraint
enum { G_PLANE = 0, G_SPHERE = 2, G_BOX = 6 };
enum { J_FREE = 0, J_SLIDE = 2, J_HINGE = 3 };
enum { C_EQ = 0, C_LIMIT = 1, C_CONTACT = 2 };

// ------------------------------------------------------------------ small vector math
struct V3 { double x, y, z; };
static inline V3 v3(double x, double y, double z) { return {x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(double s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
static inline double comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline void setcomp(V3& a, int i, double v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }

struct Q4 { double w, x, y, z; };
static inline Q4 qmul(Q4 a, Q4 b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
static inline Q4 qnormalize(Q4 q) {
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < MINVAL) return {1, 0, 0, 0};
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
struct M3 { double m[9]; };  // row-major
static inline M3 q2mat(Q4 q) {
  M3 r;
  double w = q.w, x = q.x, y = q.y, z = q.z;
  r.m[0] = w * w + x * x - y * y - z * z; r.m[1] = 2 * (x * y - w * z); r.m[2] = 2 * (x * z + w * y);
  r.m[3] = 2 * (x * y + w * z); r.m[4] = w * w - x * x + y * y - z * z; r.m[5] = 2 * (y * z - w * x);
  r.m[6] = 2 * (x * z - w * y); r.m[7] = 2 * (y * z + w * x); r.m[8] = w * w - x * x - y * y + z * z;
  return r;
}
static inline V3 mul(const M3& R, V3 v) {
  return {R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z,
          R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z};
}
static inline V3 mulT(const M3& R, V3 v) {
  return {R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z,
          R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z};
}
static inline V3 col(const M3& R, int j) { return {R.m[j], R.m[3 + j], R.m[6 + j]}; }
static inline M3 matmul(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
static inline M3 transpose(const M3& A) {
  M3 C;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * j + i];
  return C;
}

// spatial 6-vectors about the world origin: motion [w; v], force [n; f]
struct S6 { V3 a, l; };
static inline S6 operator+(S6 p, S6 q) { return {p.a + q.a, p.l + q.l}; }
static inline S6 operator*(S6 p, double s) { return {p.a * s, p.l * s}; }
static inline S6 crossMotion(S6 v, S6 m) { return {cross(v.a, m.a), cross(v.a, m.l) + cross(v.l, m.a)}; }
static inline S6 crossForce(S6 v, S6 f) { return {cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)}; }
static inline double dotMF(S6 m, S6 f) { return dot(m.a, f.a) + dot(m.l, f.l); }
// spatial inertia about origin: mass m, COM c (world), rotational inertia about COM Ic (world frame)
struct SI { double m; V3 h; double I[6]; };  // h = m*c ; I = inertia about origin: xx,yy,zz,xy,xz,yz
static inline SI si_zero() { SI s; s.m = 0; s.h = {0, 0, 0}; for (int i = 0; i < 6; ++i) s.I[i] = 0; return s; }
static inline SI si_add(const SI& a, const SI& b) {
  SI s; s.m = a.m + b.m; s.h = a.h + b.h; for (int i = 0; i < 6; ++i) s.I[i] = a.I[i] + b.I[i]; return s;
}
static inline V3 symmul(const double I[6], V3 v) {
  return {I[0] * v.x + I[3] * v.y + I[4] * v.z, I[3] * v.x + I[1] * v.y + I[5] * v.z, I[4] * v.x + I[5] * v.y + I[2] * v.z};
}
static inline S6 si_mul(const SI& s, S6 v) {  // force = I * motion
  return {symmul(s.I, v.a) + cross(s.h, v.l), s.m * v.l - cross(s.h, v.a)};
}

// ------------------------------------------------------------------ model
struct Model {
  int n_arms = 0, n_cubes = 0, nb = 0, ng = 0, nj = 0, nq = 0, nv = 0, nu = 0;
  double h = 0.001, gz = -9.81;
  std::vector<int> body_parent, body_jnt, body_weldid;
  std::vector<double> body_pos, body_quat, body_mass, body_ipos, body_iquat, body_inertia, body_invweight0;
  std::vector<int> jnt_type, jnt_body, jnt_limited, jnt_qposadr, jnt_dofadr;
  std::vector<double> jnt_axis, jnt_range, jnt_damping, jnt_solref, jnt_solimp, dof_invweight0;
  std::vector<int> geom_type, geom_body, geom_priority, geom_collide, geom_arm;
  std::vector<double> geom_size, geom_pos, geom_quat, geom_friction, geom_solref, geom_solimp;
  std::vector<int> exclude, cube_body, bucket_geom, arm_jnt, arm_base_body, site_body, act_forcelimited;
  std::vector<double> cube_size, arm_base_pos, site_pos, act_gain, act_bias, act_ctrlrange, act_forcerange, tendon_coef,
      eq_solref, eq_solimp, cube_group;
  int belt_body = 0;
  // derived
  std::vector<int> dof_body, dof_jnt, body_root;           // tree root body per body
  std::vector<int> tree_start, tree_size;                  // dof blocks per kinematic tree
  std::vector<int> dof_tree;
  std::vector<double> geom_rbound;
  struct PairG { int g1, g2; };
  // broad phase (own design; conservative, does not change the contact set): clusters of geoms with a bound
  struct Cluster { int body; int kind; /*0 sphere,1 box geom,2 plane geom*/ double c[3]; double r; int geom; };
  struct ClusterPair { int c1, c2; std::vector<PairG> pairs; };   // g1 has lower type enum / id [3P]
  std::vector<Cluster> clusters;
  std::vector<ClusterPair> cpairs;
  size_t n_geom_pairs = 0;
};

static std::vector<int> geti(const void* b, size_t n, const char* name) {
  size_t c = 0; const int32_t* p = (const int32_t*)fm_blob_find(b, n, name, 0, &c);
  if (!p && c == 0) { /* may be empty */ }
  if (!p) { if (c == 0) return {}; throw std::runtime_error(std::string("blob: missing ") + name); }
  return std::vector<int>(p, p + c);
}
static std::vector<double> getd(const void* b, size_t n, const char* name) {
  size_t c = 0; const double* p = (const double*)fm_blob_find(b, n, name, 1, &c);
  if (!p) throw std::runtime_error(std::string("blob: missing ") + name);
  return std::vector<double>(p, p + c);
}

static inline bool body_pair_filtered(const Model& m, int b1, int b2) {
  int w1 = m.body_weldid[b1], w2 = m.body_weldid[b2];
  if (w1 == w2) return true;                                     // [3P] same weld (incl. static-static)
  int wp1 = m.body_weldid[m.body_parent[w1]], wp2 = m.body_weldid[m.body_parent[w2]];
  if (w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) return true;  // [3P] filterparent
  for (size_t k = 0; k + 1 < m.exclude.size(); k += 2) {
    int e1 = m.exclude[k], e2 = m.exclude[k + 1];
    if ((e1 == b1 && e2 == b2) || (e1 == b2 && e2 == b1)) return true;
  }
  return false;
}

static void load_model(Model& m, const void* blob, size_t n) {
  auto sz = geti(blob, n, "sizes");
  m.nb = sz[0]; m.ng = sz[1]; m.nj = sz[2]; m.nq = sz[3]; m.nv = sz[4]; m.nu = sz[5];
  m.n_arms = geti(blob, n, "n_arms")[0]; m.n_cubes = geti(blob, n, "n_cubes")[0];
  auto opt = getd(blob, n, "opt"); m.h = opt[0]; m.gz = opt[1];
#define GI(x) m.x = geti(blob, n, #x)
#define GD(x) m.x = getd(blob, n, #x)
  GI(body_parent); GI(body_jnt); GI(body_weldid); GD(body_pos); GD(body_quat); GD(body_mass); GD(body_ipos);
  GD(body_iquat); GD(body_inertia); GD(body_invweight0);
  GI(jnt_type); GI(jnt_body); GI(jnt_limited); GI(jnt_qposadr); GI(jnt_dofadr); GD(jnt_axis); GD(jnt_range);
  GD(jnt_damping); GD(jnt_solref); GD(jnt_solimp); GD(dof_invweight0);
  GI(geom_type); GI(geom_body); GI(geom_priority); GI(geom_collide); GI(geom_arm); GD(geom_size); GD(geom_pos);
  GD(geom_quat); GD(geom_friction); GD(geom_solref); GD(geom_solimp);
  GI(exclude); GI(cube_body); GI(bucket_geom); GI(arm_jnt); GI(arm_base_body); GI(site_body); GI(act_forcelimited);
  GD(cube_size); GD(arm_base_pos); GD(site_pos); GD(act_gain); GD(act_bias); GD(act_ctrlrange); GD(act_forcerange);
  GD(tendon_coef); GD(eq_solref); GD(eq_solimp);
#undef GI
#undef GD
  m.cube_group = getd(blob, n, "cube_group_wxyz");
  m.belt_body = geti(blob, n, "belt_body")[0];
  m.dof_body.assign(m.nv, 0); m.dof_jnt.assign(m.nv, 0);
  for (int j = 0; j < m.nj; ++j) {
    int nd = m.jnt_type[j] == J_FREE ? 6 : 1;
    for (int k = 0; k < nd; ++k) { m.dof_body[m.jnt_dofadr[j] + k] = m.jnt_body[j]; m.dof_jnt[m.jnt_dofadr[j] + k] = j; }
  }
  // kinematic trees: root = topmost ancestor below world
  m.body_root.assign(m.nb, 0);
  for (int b = 1; b < m.nb; ++b) m.body_root[b] = m.body_parent[b] == 0 ? b : m.body_root[m.body_parent[b]];
  m.dof_tree.assign(m.nv, -1);
  int cur_root = -1;
  for (int d = 0; d < m.nv; ++d) {
    int r = m.body_root[m.dof_body[d]];
    if (r != cur_root) { m.tree_start.push_back(d); m.tree_size.push_back(0); cur_root = r; }
    m.tree_size.back()++;
    m.dof_tree[d] = (int)m.tree_start.size() - 1;
  }
  m.geom_rbound.assign(m.ng, 0);
  for (int g = 0; g < m.ng; ++g) {
    const double* s = &m.geom_size[3 * g];
    if (m.geom_type[g] == G_SPHERE) m.geom_rbound[g] = s[0];
    else if (m.geom_type[g] == G_BOX) m.geom_rbound[g] = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  }
  // clusters: big geoms (plane, boxes with rbound >= 1) alone; every other body = one sphere-bounded cluster
  std::vector<int> geom_cluster(m.ng, -1);
  for (int b = 0; b < m.nb; ++b) {
    std::vector<int> small;
    for (int g = 0; g < m.ng; ++g) {
      if (m.geom_body[g] != b || !m.geom_collide[g]) continue;
      if (m.geom_type[g] == G_PLANE || m.geom_rbound[g] >= 1.0) {
        Model::Cluster c; c.body = b; c.kind = m.geom_type[g] == G_PLANE ? 2 : 1; c.c[0] = c.c[1] = c.c[2] = 0; c.r = 0; c.geom = g;
        geom_cluster[g] = (int)m.clusters.size(); m.clusters.push_back(c);
      } else small.push_back(g);
    }
    if (small.empty()) continue;
    Model::Cluster c; c.body = b; c.kind = 0; c.geom = -1;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int g : small) for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(lo[k], m.geom_pos[3 * g + k] - m.geom_rbound[g]); hi[k] = std::max(hi[k], m.geom_pos[3 * g + k] + m.geom_rbound[g]);
    }
    for (int k = 0; k < 3; ++k) c.c[k] = 0.5 * (lo[k] + hi[k]);
    c.r = 0;
    for (int g : small) {
      double dx = m.geom_pos[3 * g] - c.c[0], dy = m.geom_pos[3 * g + 1] - c.c[1], dz = m.geom_pos[3 * g + 2] - c.c[2];
      c.r = std::max(c.r, std::sqrt(dx * dx + dy * dy + dz * dz) + m.geom_rbound[g]);
    }
    for (int g : small) geom_cluster[g] = (int)m.clusters.size();
    m.clusters.push_back(c);
  }
  int nc = (int)m.clusters.size();
  std::vector<int> cp_index((size_t)nc * nc, -1);
  for (int g1 = 0; g1 < m.ng; ++g1)
    for (int g2 = g1 + 1; g2 < m.ng; ++g2) {
      if (!m.geom_collide[g1] || !m.geom_collide[g2]) continue;
      if (body_pair_filtered(m, m.geom_body[g1], m.geom_body[g2])) continue;
      int a = g1, b = g2;
      if (m.geom_type[a] > m.geom_type[b]) std::swap(a, b);   // [3P] geom1 has the lower type enum
      int c1 = geom_cluster[g1], c2 = geom_cluster[g2];
      if (cp_index[(size_t)c1 * nc + c2] < 0) { cp_index[(size_t)c1 * nc + c2] = (int)m.cpairs.size(); m.cpairs.push_back({c1, c2, {}}); }
      m.cpairs[cp_index[(size_t)c1 * nc + c2]].pairs.push_back({a, b});
      m.n_geom_pairs++;
    }
}

// ------------------------------------------------------------------ data
struct Contact {
  int g1, g2; V3 pos; V3 n, t1, t2; double dist; double mu; double solref[2]; double solimp[5];
  int efc;  // first of 4 rows
};
struct Data {
  std::vector<double> qpos, qvel, ctrl, qacc, qacc_warmstart;
  std::vector<V3> xpos, xipos, gpos; std::vector<Q4> xquat; std::vector<M3> xmat, ximat, gmat;
  std::vector<S6> cdof, cdof_dot, cvel; std::vector<SI> cinert;
  std::vector<double> M, L, Li;                 // dense nv*nv (block diagonal by tree), cholesky of M, of M-hD
  std::vector<double> qfrc_bias, qfrc_passive, qfrc_actuator, qfrc_smooth, qacc_smooth, qfrc_constraint;
  std::vector<double> act_length, act_velocity, act_force;
  std::vector<Contact> con;
  int nefc = 0;
  std::vector<int> efc_type, efc_id;
  std::vector<double> efc_J, efc_pos, efc_vel, efc_aref, efc_R, efc_D, efc_force, efc_diagApprox, efc_KBI;
  int solver_iter = 0; double solver_grad = 0;
  long long flops = 0;
};

static void init_data(const Model& m, Data& d) {
  d.qpos.assign(m.nq, 0); d.qvel.assign(m.nv, 0); d.ctrl.assign(m.nu, 0); d.qacc.assign(m.nv, 0); d.qacc_warmstart.assign(m.nv, 0);
  for (int j = 0; j < m.nj; ++j) if (m.jnt_type[j] == J_FREE) d.qpos[m.jnt_qposadr[j] + 3] = 1;
  d.xpos.assign(m.nb, {0, 0, 0}); d.xipos = d.xpos; d.xquat.assign(m.nb, {1, 0, 0, 0});
  d.xmat.assign(m.nb, q2mat({1, 0, 0, 0})); d.ximat = d.xmat;
  d.gpos.assign(m.ng, {0, 0, 0}); d.gmat.assign(m.ng, q2mat({1, 0, 0, 0}));
  d.cdof.assign(m.nv, {{0, 0, 0}, {0, 0, 0}}); d.cdof_dot = d.cdof; d.cvel.assign(m.nb, {{0, 0, 0}, {0, 0, 0}});
  d.cinert.assign(m.nb, si_zero());
  d.M.assign((size_t)m.nv * m.nv, 0); d.L = d.M; d.Li = d.M;
  for (auto* v : {&d.qfrc_bias, &d.qfrc_passive, &d.qfrc_actuator, &d.qfrc_smooth, &d.qacc_smooth, &d.qfrc_constraint}) v->assign(m.nv, 0);
  d.act_length.assign(m.nu, 0); d.act_velocity.assign(m.nu, 0); d.act_force.assign(m.nu, 0);
}

// ------------------------------------------------------------------ kinematics (mj_kinematics, mj_comPos) [3P]
static void kinematics(const Model& m, Data& d) {
  for (int b = 1; b < m.nb; ++b) {
    int p = m.body_parent[b], j = m.body_jnt[b];
    V3 bp = {m.body_pos[3 * b], m.body_pos[3 * b + 1], m.body_pos[3 * b + 2]};
    Q4 bq = {m.body_quat[4 * b], m.body_quat[4 * b + 1], m.body_quat[4 * b + 2], m.body_quat[4 * b + 3]};
    if (j >= 0 && m.jnt_type[j] == J_FREE) {
      double* q = &d.qpos[m.jnt_qposadr[j]];
      Q4 qq = qnormalize({q[3], q[4], q[5], q[6]});
      q[3] = qq.w; q[4] = qq.x; q[5] = qq.y; q[6] = qq.z;      // [3P] quats in qpos are normalised in place
      d.xpos[b] = {q[0], q[1], q[2]};
      d.xquat[b] = qq;
    } else {
      d.xpos[b] = d.xpos[p] + mul(d.xmat[p], bp);
      d.xquat[b] = qmul(d.xquat[p], bq);
      if (j >= 0) {
        double q = d.qpos[m.jnt_qposadr[j]];
        V3 ax = {m.jnt_axis[3 * j], m.jnt_axis[3 * j + 1], m.jnt_axis[3 * j + 2]};
        if (m.jnt_type[j] == J_HINGE) {
          double s = std::sin(q / 2), c = std::cos(q / 2);
          d.xquat[b] = qmul(d.xquat[b], {c, s * ax.x, s * ax.y, s * ax.z});
        } else {
          d.xpos[b] = d.xpos[b] + mul(q2mat(d.xquat[b]), ax) * q;
        }
      }
      d.xquat[b] = qnormalize(d.xquat[b]);
    }
    d.xmat[b] = q2mat(d.xquat[b]);
    V3 ip = {m.body_ipos[3 * b], m.body_ipos[3 * b + 1], m.body_ipos[3 * b + 2]};
    Q4 iq = {m.body_iquat[4 * b], m.body_iquat[4 * b + 1], m.body_iquat[4 * b + 2], m.body_iquat[4 * b + 3]};
    d.xipos[b] = d.xpos[b] + mul(d.xmat[b], ip);
    d.ximat[b] = matmul(d.xmat[b], q2mat(iq));
  }
  for (int g = 0; g < m.ng; ++g) {
    int b = m.geom_body[g];
    V3 gp = {m.geom_pos[3 * g], m.geom_pos[3 * g + 1], m.geom_pos[3 * g + 2]};
    Q4 gq = {m.geom_quat[4 * g], m.geom_quat[4 * g + 1], m.geom_quat[4 * g + 2], m.geom_quat[4 * g + 3]};
    d.gpos[g] = d.xpos[b] + mul(d.xmat[b], gp);
    d.gmat[g] = matmul(d.xmat[b], q2mat(gq));
  }
  // motion subspaces about the world origin + body spatial inertias
  for (int j = 0; j < m.nj; ++j) {
    int b = m.jnt_body[j], a = m.jnt_dofadr[j];
    V3 ax = mul(d.xmat[b], {m.jnt_axis[3 * j], m.jnt_axis[3 * j + 1], m.jnt_axis[3 * j + 2]});
    if (m.jnt_type[j] == J_FREE) {
      for (int k = 0; k < 3; ++k) {
        V3 e = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
        d.cdof[a + k] = {{0, 0, 0}, e};
        V3 r = col(d.xmat[b], k);
        d.cdof[a + 3 + k] = {r, cross(d.xpos[b], r)};
      }
    } else if (m.jnt_type[j] == J_HINGE) {
      d.cdof[a] = {ax, cross(d.xpos[b], ax)};
    } else {
      d.cdof[a] = {{0, 0, 0}, ax};
    }
  }
  for (int b = 1; b < m.nb; ++b) {
    double ms = m.body_mass[b];
    SI s = si_zero();
    if (ms > 0) {
      const M3& R = d.ximat[b];
      const double* di = &m.body_inertia[3 * b];
      double Ic[9];
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k)
          Ic[3 * i + k] = R.m[3 * i] * di[0] * R.m[3 * k] + R.m[3 * i + 1] * di[1] * R.m[3 * k + 1] + R.m[3 * i + 2] * di[2] * R.m[3 * k + 2];
      V3 c = d.xipos[b];
      double cc = dot(c, c);
      s.m = ms; s.h = c * ms;
      s.I[0] = Ic[0] + ms * (cc - c.x * c.x); s.I[1] = Ic[4] + ms * (cc - c.y * c.y); s.I[2] = Ic[8] + ms * (cc - c.z * c.z);
      s.I[3] = Ic[1] - ms * c.x * c.y; s.I[4] = Ic[2] - ms * c.x * c.z; s.I[5] = Ic[5] - ms * c.y * c.z;
    }
    d.cinert[b] = s;
  }
}

// ------------------------------------------------------------------ CRBA + factor (mj_crb, mj_factorM) [3P]
static bool chol_block(double* A, int ld, int s, int n) {  // in-place lower cholesky on block [s,s+n)
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j <= i; ++j) {
      double sum = A[(s + i) * ld + s + j];
      for (int k = 0; k < j; ++k) sum -= A[(s + i) * ld + s + k] * A[(s + j) * ld + s + k];
      if (i == j) { if (sum <= 0) return false; A[(s + i) * ld + s + i] = std::sqrt(sum); }
      else A[(s + i) * ld + s + j] = sum / A[(s + j) * ld + s + j];
    }
  }
  return true;
}
static void chol_solve_block(const double* L, int ld, int s, int n, double* x) {  // x[s..s+n) in place
  for (int i = 0; i < n; ++i) {
    double sum = x[s + i];
    for (int k = 0; k < i; ++k) sum -= L[(s + i) * ld + s + k] * x[s + k];
    x[s + i] = sum / L[(s + i) * ld + s + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double sum = x[s + i];
    for (int k = i + 1; k < n; ++k) sum -= L[(s + k) * ld + s + i] * x[s + k];
    x[s + i] = sum / L[(s + i) * ld + s + i];
  }
}
static void solveM(const Model& m, const Data& d, const std::vector<double>& L, double* x) {
  for (size_t t = 0; t < m.tree_start.size(); ++t) chol_solve_block(L.data(), m.nv, m.tree_start[t], m.tree_size[t], x);
}

static void crb(const Model& m, Data& d) {
  std::vector<SI> crbI(d.cinert);
  for (int b = m.nb - 1; b > 0; --b) if (m.body_parent[b] > 0) crbI[m.body_parent[b]] = si_add(crbI[m.body_parent[b]], crbI[b]);
  std::fill(d.M.begin(), d.M.end(), 0.0);
  for (int i = 0; i < m.nv; ++i) {
    int b = m.dof_body[i];
    S6 F = si_mul(crbI[b], d.cdof[i]);
    // ancestors' dofs (incl. own joint's earlier dofs and itself)
    for (int j = i; j >= 0; --j) {
      if (m.dof_tree[j] != m.dof_tree[i]) break;
      // j must be on the path: dof_body[j] is ancestor-or-self of b
      int bj = m.dof_body[j], bb = b; bool anc = false;
      while (bb > 0) { if (bb == bj) { anc = true; break; } bb = m.body_parent[bb]; }
      if (!anc) continue;
      double v = dotMF(d.cdof[j], F);
      d.M[(size_t)i * m.nv + j] = v; d.M[(size_t)j * m.nv + i] = v;
    }
  }
  d.L = d.M;
  for (size_t t = 0; t < m.tree_start.size(); ++t)
    if (!chol_block(d.L.data(), m.nv, m.tree_start[t], m.tree_size[t])) throw std::runtime_error("mass matrix not PD");
}

// ------------------------------------------------------------------ velocity stage: comVel + RNE bias [3P]
static void velocity_stage(const Model& m, Data& d) {
  S6 zero = {{0, 0, 0}, {0, 0, 0}};
  std::vector<S6> cacc(m.nb, zero), cfrc(m.nb, zero);
  d.cvel[0] = zero;
  cacc[0] = {{0, 0, 0}, {0, 0, -m.gz}};   // gravity as upward base acceleration
  for (int b = 1; b < m.nb; ++b) {
    int p = m.body_parent[b], j = m.body_jnt[b];
    S6 v = d.cvel[p], a = cacc[p];
    if (j >= 0) {
      int da = m.jnt_dofadr[j];
      if (m.jnt_type[j] == J_FREE) {
        for (int k = 0; k < 3; ++k) { d.cdof_dot[da + k] = zero; v = v + d.cdof[da + k] * d.qvel[da + k]; }
        for (int k = 3; k < 6; ++k) d.cdof_dot[da + k] = crossMotion(v, d.cdof[da + k]);
        for (int k = 3; k < 6; ++k) v = v + d.cdof[da + k] * d.qvel[da + k];
        for (int k = 0; k < 6; ++k) a = a + d.cdof_dot[da + k] * d.qvel[da + k];
      } else {
        d.cdof_dot[da] = crossMotion(v, d.cdof[da]);
        v = v + d.cdof[da] * d.qvel[da];
        a = a + d.cdof_dot[da] * d.qvel[da];
      }
    }
    d.cvel[b] = v; cacc[b] = a;
    S6 Iv = si_mul(d.cinert[b], v);
    cfrc[b] = si_mul(d.cinert[b], a) + crossForce(v, Iv);
  }
  for (int b = m.nb - 1; b > 0; --b) if (m.body_parent[b] > 0) cfrc[m.body_parent[b]] = cfrc[m.body_parent[b]] + cfrc[b];
  for (int i = 0; i < m.nv; ++i) d.qfrc_bias[i] = dotMF(d.cdof[i], cfrc[m.dof_body[i]]);
  for (int i = 0; i < m.nv; ++i) d.qfrc_passive[i] = -m.jnt_damping[m.dof_jnt[i]] * d.qvel[i];
}

// ------------------------------------------------------------------ collision primitives [3P conventions]
struct RawCon { V3 pos, n; double dist; };

static int plane_sphere(V3 pp, const M3& pm, V3 sp, double r, RawCon* out) {
  V3 n = col(pm, 2);
  double dist = dot(sp - pp, n) - r;
  if (dist > 0) return 0;
  out[0] = {sp - n * (r + dist / 2), n, dist};
  return 1;
}
static int plane_box(V3 pp, const M3& pm, V3 bp, const M3& bm, const double* sz, RawCon* out) {
  V3 n = col(pm, 2);
  double dist = dot(bp - pp, n);
  int cnt = 0;
  for (int i = 0; i < 8; ++i) {
    V3 vec = {(i & 1) ? sz[0] : -sz[0], (i & 2) ? sz[1] : -sz[1], (i & 4) ? sz[2] : -sz[2]};
    V3 corner = mul(bm, vec);
    double ld = dot(n, corner);
    if (dist + ld > 0 || ld > 0) continue;
    double cd = dist + ld;
    out[cnt] = {corner + bp - n * (cd / 2), n, cd};
    if (++cnt >= 4) return 4;
  }
  return cnt;
}
static int sphere_sphere(V3 p1, double r1, V3 p2, double r2, RawCon* out) {
  V3 dd = p2 - p1;
  double len = norm(dd);
  double dist = len - r1 - r2;
  if (dist > 0) return 0;
  V3 n = len < MINVAL ? v3(1, 0, 0) : dd * (1.0 / len);
  out[0] = {p1 + n * (r1 + dist / 2), n, dist};
  return 1;
}
static int sphere_box(V3 sp, double r, V3 bp, const M3& bm, const double* sz, RawCon* out) {
  V3 c = mulT(bm, sp - bp);
  V3 cl = {std::min(std::max(c.x, -sz[0]), sz[0]), std::min(std::max(c.y, -sz[1]), sz[1]), std::min(std::max(c.z, -sz[2]), sz[2])};
  V3 dv = cl - c;   // sphere centre -> closest point on the box
  double dl = norm(dv);
  if (dl > MINVAL) {
    double dist = dl - r;
    if (dist > 0) return 0;
    V3 n = mul(bm, dv * (1.0 / dl));
    out[0] = {sp + n * (r + dist / 2), n, dist};
    return 1;
  }
  // centre inside the box: nearest face
  int k = 0; double best = 1e300;
  for (int i = 0; i < 3; ++i) { double g = sz[i] - std::fabs(comp(c, i)); if (g < best) { best = g; k = i; } }
  double s = comp(c, k) >= 0 ? 1.0 : -1.0;
  V3 nl = {0, 0, 0}; setcomp(nl, k, -s);
  V3 n = mul(bm, nl);
  double dist = -best - r;
  out[0] = {sp + n * (r + dist / 2), n, dist};
  return 1;
}

// box-box: separating-axis search + reference-face clipping (own formulation; up to 8 points).
// Normal points from box 1 to box 2; pos is midway between the surfaces; dist = -depth.
static int box_box(V3 p1, const M3& R1, const double* A, V3 p2, const M3& R2, const double* B, RawCon* out) {
  V3 pw = p2 - p1;
  V3 pp = mulT(R1, pw);
  double R[9], Q[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      R[3 * i + j] = dot(col(R1, i), col(R2, j));
      Q[3 * i + j] = std::fabs(R[3 * i + j]);
    }
  double best = -1e300; int code = -1; bool flip = false; V3 nbest = {0, 0, 0};
  // face axes of box 1
  for (int i = 0; i < 3; ++i) {
    double e = comp(pp, i);
    double s = std::fabs(e) - (A[i] + B[0] * Q[3 * i] + B[1] * Q[3 * i + 1] + B[2] * Q[3 * i + 2]);
    if (s > 0) return 0;
    if (s > best) { best = s; code = i; flip = e < 0; nbest = col(R1, i); }
  }
  // face axes of box 2
  for (int j = 0; j < 3; ++j) {
    double e = dot(col(R2, j), pw);
    double s = std::fabs(e) - (B[j] + A[0] * Q[j] + A[1] * Q[3 + j] + A[2] * Q[6 + j]);
    if (s > 0) return 0;
    if (s > best + 1e-6) { best = s; code = 3 + j; flip = e < 0; nbest = col(R2, j); }
  }
  // edge x edge axes
  const double fudge = 1.05;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      V3 ax = cross(col(R1, i), col(R2, j));
      double l = norm(ax);
      if (l < 1e-4) continue;
      ax = ax * (1.0 / l);
      double e = dot(ax, pw);
      double ra = 0, rb = 0;
      for (int k = 0; k < 3; ++k) { ra += A[k] * std::fabs(dot(ax, col(R1, k))); rb += B[k] * std::fabs(dot(ax, col(R2, k))); }
      double s = std::fabs(e) - (ra + rb);
      if (s > 0) return 0;
      if (s > best * (1.0 / fudge) + 1e-6) { best = s; code = 6 + 3 * i + j; flip = e < 0; nbest = ax; }
    }
  V3 n = flip ? nbest * -1.0 : nbest;   // from box1 towards box2
  double depth = -best;
  if (code >= 6) {
    int i = (code - 6) / 3, j = (code - 6) % 3;
    // support points: on box1 farthest along +n, on box2 farthest along -n, then closest points of the two edges
    V3 pa = p1, pb = p2;
    for (int k = 0; k < 3; ++k) {
      if (k != i) { double sg = dot(n, col(R1, k)) > 0 ? 1.0 : -1.0; pa = pa + col(R1, k) * (sg * A[k]); }
      if (k != j) { double sg = dot(n, col(R2, k)) > 0 ? -1.0 : 1.0; pb = pb + col(R2, k) * (sg * B[k]); }
    }
    V3 ua = col(R1, i), ub = col(R2, j);
    V3 dp = pb - pa;
    double uaub = dot(ua, ub), q1 = dot(ua, dp), q2 = -dot(ub, dp);
    double dn = 1 - uaub * uaub;
    double al = 0, be = 0;
    if (dn > 1e-8) { al = (q1 + uaub * q2) / dn; be = (uaub * q1 + q2) / dn; }
    al = std::min(std::max(al, -A[i]), A[i]); be = std::min(std::max(be, -B[j]), B[j]);
    V3 ca = pa + ua * al, cb = pb + ub * be;
    out[0] = {(ca + cb) * 0.5, n, -depth};
    return 1;
  }
  // face contact: reference box owns the axis
  bool ref1 = code < 3;
  int ax = ref1 ? code : code - 3;
  const M3& Rr = ref1 ? R1 : R2; const M3& Ri = ref1 ? R2 : R1;
  const double* Sr = ref1 ? A : B; const double* Si = ref1 ? B : A;
  V3 pr = ref1 ? p1 : p2, pi = ref1 ? p2 : p1;
  V3 nr = ref1 ? n : n * -1.0;          // outward normal of the reference face (towards the incident box)
  // incident face: most anti-parallel to nr
  int iax = 0; double mn = 1e300; double isg = 1;
  for (int k = 0; k < 3; ++k) {
    double dd = dot(col(Ri, k), nr);
    if (dd < mn) { mn = dd; iax = k; isg = 1; }
    if (-dd < mn) { mn = -dd; iax = k; isg = -1; }
  }
  int i1 = (iax + 1) % 3, i2 = (iax + 2) % 3;
  V3 fc = pi + col(Ri, iax) * (isg * Si[iax]);
  V3 quad[4] = {fc + col(Ri, i1) * Si[i1] + col(Ri, i2) * Si[i2], fc - col(Ri, i1) * Si[i1] + col(Ri, i2) * Si[i2],
                fc - col(Ri, i1) * Si[i1] - col(Ri, i2) * Si[i2], fc + col(Ri, i1) * Si[i1] - col(Ri, i2) * Si[i2]};
  int r1 = (ax + 1) % 3, r2 = (ax + 2) % 3;
  V3 u1 = col(Rr, r1), u2 = col(Rr, r2);
  double rsg = dot(col(Rr, ax), nr) > 0 ? 1.0 : -1.0;
  V3 rc = pr + col(Rr, ax) * (rsg * Sr[ax]);   // centre of the reference face
  // polygon in reference-face coordinates (x,y in plane, z = height above the face along nr)
  double poly[16][3], tmp[16][3]; int np = 4;
  for (int k = 0; k < 4; ++k) { V3 dq = quad[k] - rc; poly[k][0] = dot(dq, u1); poly[k][1] = dot(dq, u2); poly[k][2] = dot(dq, nr); }
  const double lim[2] = {Sr[r1], Sr[r2]};
  for (int c = 0; c < 2; ++c)
    for (int sg = -1; sg <= 1; sg += 2) {
      int nn = 0;
      for (int k = 0; k < np; ++k) {
        const double* a = poly[k]; const double* b = poly[(k + 1) % np];
        double da = lim[c] - sg * a[c], db = lim[c] - sg * b[c];
        if (da >= 0) { tmp[nn][0] = a[0]; tmp[nn][1] = a[1]; tmp[nn][2] = a[2]; nn++; }
        if ((da >= 0) != (db >= 0)) {
          double t = da / (da - db);
          for (int q = 0; q < 3; ++q) tmp[nn][q] = a[q] + t * (b[q] - a[q]);
          nn++;
        }
      }
      np = nn;
      for (int k = 0; k < np; ++k) for (int q = 0; q < 3; ++q) poly[k][q] = tmp[k][q];
      if (np == 0) break;
    }
  int cnt = 0;
  for (int k = 0; k < np && cnt < 8; ++k) {
    double hgt = poly[k][2];
    if (hgt > 0) continue;                 // above the reference face: no penetration
    V3 pt = rc + u1 * poly[k][0] + u2 * poly[k][1] + nr * (hgt * 0.5);   // midway between the surfaces
    out[cnt++] = {pt, n, hgt};
  }
  return cnt;
}

static inline void make_frame(V3 n, V3& t1, V3& t2) {   // [3P] mju_makeFrame
  V3 y = (n.y < 0.5 && n.y > -0.5) ? v3(0, 1, 0) : v3(0, 0, 1);
  y = y - n * dot(n, y);
  y = y * (1.0 / norm(y));
  t1 = y; t2 = cross(n, y);
}

// distance from a point to an oriented box (0 if inside)
static inline double point_box_dist(V3 p, V3 bp, const M3& bm, const double* sz) {
  V3 c = mulT(bm, p - bp);
  double dx = std::max(std::fabs(c.x) - sz[0], 0.0), dy = std::max(std::fabs(c.y) - sz[1], 0.0), dz = std::max(std::fabs(c.z) - sz[2], 0.0);
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}

static void contact_params(const Model& m, int g1, int g2, Contact& c) {   // [3P] mj_contactParam
  int p1 = m.geom_priority[g1], p2 = m.geom_priority[g2];
  const double *f1 = &m.geom_friction[3 * g1], *f2 = &m.geom_friction[3 * g2];
  if (p1 != p2) {
    int g = p1 > p2 ? g1 : g2;
    c.mu = m.geom_friction[3 * g];
    for (int i = 0; i < 2; ++i) c.solref[i] = m.geom_solref[2 * g + i];
    for (int i = 0; i < 5; ++i) c.solimp[i] = m.geom_solimp[5 * g + i];
  } else {
    c.mu = std::max(f1[0], f2[0]);
    double mix = 0.5;   // solmix 1 : 1
    for (int i = 0; i < 2; ++i) c.solref[i] = mix * m.geom_solref[2 * g1 + i] + (1 - mix) * m.geom_solref[2 * g2 + i];
    for (int i = 0; i < 5; ++i) c.solimp[i] = mix * m.geom_solimp[5 * g1 + i] + (1 - mix) * m.geom_solimp[5 * g2 + i];
  }
  c.mu = std::max(c.mu, MINMU);
}

static inline bool cluster_far(const Model& m, const Data& d, const Model::Cluster& a, const Model::Cluster& b) {
  const Model::Cluster* s = &a; const Model::Cluster* o = &b;
  if (s->kind != 0) std::swap(s, o);
  if (s->kind != 0) return false;     // two big geoms: always narrow-test
  V3 cs = d.xpos[s->body] + mul(d.xmat[s->body], {s->c[0], s->c[1], s->c[2]});
  if (o->kind == 0) {
    V3 co = d.xpos[o->body] + mul(d.xmat[o->body], {o->c[0], o->c[1], o->c[2]});
    V3 dd = cs - co; double rr = s->r + o->r;
    return dot(dd, dd) > rr * rr;
  }
  int g = o->geom;
  if (o->kind == 2) return dot(cs - d.gpos[g], col(d.gmat[g], 2)) - s->r > 0;
  return point_box_dist(cs, d.gpos[g], d.gmat[g], &m.geom_size[3 * g]) > s->r;
}

static void collision(const Model& m, Data& d) {
  d.con.clear();
  RawCon raw[8];
  for (const auto& cp : m.cpairs) {
    if (cluster_far(m, d, m.clusters[cp.c1], m.clusters[cp.c2])) continue;
    for (const auto& pr : cp.pairs) {
      int g1 = pr.g1, g2 = pr.g2, t1 = m.geom_type[g1], t2 = m.geom_type[g2];
      const double *s1 = &m.geom_size[3 * g1], *s2 = &m.geom_size[3 * g2];
      int n = 0;
      if (t1 == G_PLANE && t2 == G_SPHERE) n = plane_sphere(d.gpos[g1], d.gmat[g1], d.gpos[g2], s2[0], raw);
      else if (t1 == G_PLANE && t2 == G_BOX) n = plane_box(d.gpos[g1], d.gmat[g1], d.gpos[g2], d.gmat[g2], s2, raw);
      else if (t1 == G_SPHERE && t2 == G_SPHERE) n = sphere_sphere(d.gpos[g1], s1[0], d.gpos[g2], s2[0], raw);
      else if (t1 == G_SPHERE && t2 == G_BOX) n = sphere_box(d.gpos[g1], s1[0], d.gpos[g2], d.gmat[g2], s2, raw);
      else if (t1 == G_BOX && t2 == G_BOX) {
        if (point_box_dist(d.gpos[g2], d.gpos[g1], d.gmat[g1], s1) > m.geom_rbound[g2]) continue;
        if (point_box_dist(d.gpos[g1], d.gpos[g2], d.gmat[g2], s2) > m.geom_rbound[g1]) continue;
        n = box_box(d.gpos[g1], d.gmat[g1], s1, d.gpos[g2], d.gmat[g2], s2, raw);
      }
      for (int k = 0; k < n; ++k) {
        Contact c;
        c.g1 = g1; c.g2 = g2; c.pos = raw[k].pos; c.n = raw[k].n; c.dist = raw[k].dist; c.efc = -1;
        make_frame(c.n, c.t1, c.t2);
        contact_params(m, g1, g2, c);
        d.con.push_back(c);
      }
    }
  }
  // canonical order: by (geom1, geom2), emission order within a pair
  std::stable_sort(d.con.begin(), d.con.end(), [](const Contact& a, const Contact& b) { return a.g1 != b.g1 ? a.g1 < b.g1 : a.g2 < b.g2; });
}

// ------------------------------------------------------------------ constraints (mj_makeConstraint etc.) [3P]
static void jac_point(const Model& m, const Data& d, int body, V3 p, double* jp /*3*nv*/) {
  std::fill(jp, jp + 3 * m.nv, 0.0);
  int b = body;
  while (b > 0) {
    int j = m.body_jnt[b];
    if (j >= 0) {
      int a = m.jnt_dofadr[j], nd = m.jnt_type[j] == J_FREE ? 6 : 1;
      for (int k = 0; k < nd; ++k) {
        const S6& s = d.cdof[a + k];
        V3 v = s.l + cross(s.a, p);   // velocity of the point due to unit dof velocity
        jp[a + k] = v.x; jp[m.nv + a + k] = v.y; jp[2 * m.nv + a + k] = v.z;
      }
    }
    b = m.body_parent[b];
  }
}

static void impedance(const double* solref, const double* solimp, double pos, double h, double& K, double& B, double& imp) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = std::min(std::max(dmin, 0.0001), 0.9999); dmax = std::min(std::max(dmax, 0.0001), 0.9999);
  width = std::max(width, 0.0); mid = std::min(std::max(mid, 0.0001), 0.9999); power = std::max(power, 1.0);
  if (dmin == dmax || width <= MINVAL) imp = 0.5 * (dmin + dmax);
  else {
    double x = std::fabs(pos) / width, y;
    if (x >= 1) y = 1;
    else if (x <= 0) y = 0;
    else if (power == 1) y = x;
    else if (x <= mid) y = std::pow(x, power) / std::pow(mid, power - 1);
    else y = 1 - std::pow(1 - x, power) / std::pow(1 - mid, power - 1);
    imp = dmin + y * (dmax - dmin);
  }
  double tc = std::max(solref[0], 2 * h);   // refsafe
  double dr = solref[1];
  K = 1 / std::max(MINVAL, dmax * dmax * tc * tc * dr * dr);
  B = 2 / std::max(MINVAL, dmax * tc);
}

static void add_row(const Model& m, Data& d, int type, int id, const double* J, double pos, double diagApprox,
                    const double* solref, const double* solimp) {
  d.efc_type.push_back(type); d.efc_id.push_back(id);
  d.efc_J.insert(d.efc_J.end(), J, J + m.nv);
  d.efc_pos.push_back(pos); d.efc_diagApprox.push_back(diagApprox);
  double K, B, imp; impedance(solref, solimp, pos, m.h, K, B, imp);
  d.efc_KBI.push_back(K); d.efc_KBI.push_back(B); d.efc_KBI.push_back(imp);
  d.efc_R.push_back(std::max(MINVAL, (1 - imp) * diagApprox / imp));
  d.nefc++;
}

static void make_constraint(const Model& m, Data& d) {
  d.nefc = 0; d.efc_type.clear(); d.efc_id.clear(); d.efc_J.clear(); d.efc_pos.clear(); d.efc_diagApprox.clear();
  d.efc_KBI.clear(); d.efc_R.clear();
  std::vector<double> J(m.nv), j1(3 * m.nv), j2(3 * m.nv);
  // joint equalities (gripper.xml:46-49): q_left - q_right = 0
  for (int a = 0; a < m.n_arms; ++a) {
    int jl = m.arm_jnt[9 * a + 7], jr = m.arm_jnt[9 * a + 8];
    int dl = m.jnt_dofadr[jl], dr = m.jnt_dofadr[jr];
    std::fill(J.begin(), J.end(), 0.0); J[dl] = 1; J[dr] = -1;
    double pos = d.qpos[m.jnt_qposadr[jl]] - d.qpos[m.jnt_qposadr[jr]];
    add_row(m, d, C_EQ, a, J.data(), pos, m.dof_invweight0[dl] + m.dof_invweight0[dr], m.eq_solref.data(), m.eq_solimp.data());
  }
  // joint limits
  for (int j = 0; j < m.nj; ++j) {
    if (!m.jnt_limited[j] || m.jnt_type[j] == J_FREE) continue;
    double q = d.qpos[m.jnt_qposadr[j]];
    for (int side = 0; side < 2; ++side) {
      double dist = side == 0 ? q - m.jnt_range[2 * j] : m.jnt_range[2 * j + 1] - q;
      if (dist < 0) {
        std::fill(J.begin(), J.end(), 0.0); J[m.jnt_dofadr[j]] = side == 0 ? 1 : -1;
        add_row(m, d, C_LIMIT, j, J.data(), dist, m.dof_invweight0[m.jnt_dofadr[j]], &m.jnt_solref[2 * j], &m.jnt_solimp[5 * j]);
      }
    }
  }
  // contacts: pyramidal, condim 3 -> 4 rows
  for (size_t ci = 0; ci < d.con.size(); ++ci) {
    Contact& c = d.con[ci];
    int b1 = m.geom_body[c.g1], b2 = m.geom_body[c.g2];
    jac_point(m, d, b1, c.pos, j1.data()); jac_point(m, d, b2, c.pos, j2.data());
    std::vector<double> Jn(m.nv), Jt1(m.nv), Jt2(m.nv);
    for (int k = 0; k < m.nv; ++k) {
      V3 dv = {j2[k] - j1[k], j2[m.nv + k] - j1[m.nv + k], j2[2 * m.nv + k] - j1[2 * m.nv + k]};
      Jn[k] = dot(c.n, dv); Jt1[k] = dot(c.t1, dv); Jt2[k] = dot(c.t2, dv);
    }
    double tran = m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2];
    double dA = tran + c.mu * c.mu * tran;
    c.efc = d.nefc;
    for (int r = 0; r < 4; ++r) {
      const std::vector<double>& Jt = r < 2 ? Jt1 : Jt2;
      double sg = (r & 1) ? -1.0 : 1.0;
      for (int k = 0; k < m.nv; ++k) J[k] = Jn[k] + sg * c.mu * Jt[k];
      add_row(m, d, C_CONTACT, (int)ci, J.data(), c.dist, dA, c.solref, c.solimp);
    }
    // [3P] pyramidal regulariser: all rows of the contact use Rpy = 2 mu^2 R(first)
    double Rpy = 2 * c.mu * c.mu * d.efc_R[c.efc];
    // CALIBRATED to reference-held data (runs/*.zip:_last_obs, tests/test_ref_fixtures.py): resting cubes sit 2.28e-6 m deep in
    // the belt and 4.2e-7 m in the table, exactly twice what the documented regulariser gives with four clipping points; the
    // MuJoCo detail behind it (mjc_BoxBox point count or impedance) cannot be read here, so box-box pairs get R x 2.
    if (m.geom_type[c.g1] == G_BOX && m.geom_type[c.g2] == G_BOX) Rpy *= BOXBOX_RSCALE;
    for (int r = 0; r < 4; ++r) d.efc_R[c.efc + r] = Rpy;
  }
  d.efc_D.resize(d.nefc);
  for (int i = 0; i < d.nefc; ++i) d.efc_D[i] = 1 / d.efc_R[i];
  d.efc_vel.assign(d.nefc, 0); d.efc_aref.assign(d.nefc, 0); d.efc_force.assign(d.nefc, 0);
}

static void reference_accel(const Model& m, Data& d) {   // efc_vel, efc_aref (mj_referenceConstraint) [3P]
  for (int i = 0; i < d.nefc; ++i) {
    double v = 0; const double* J = &d.efc_J[(size_t)i * m.nv];
    for (int k = 0; k < m.nv; ++k) v += J[k] * d.qvel[k];
    d.efc_vel[i] = v;
    d.efc_aref[i] = -d.efc_KBI[3 * i + 1] * v - d.efc_KBI[3 * i] * d.efc_KBI[3 * i + 2] * d.efc_pos[i];
  }
}

// transmission: actuator length/velocity (joint or fixed tendon)
static void transmission(const Model& m, Data& d) {
  int belt_j = m.body_jnt[m.belt_body];
  d.act_length[0] = d.qpos[m.jnt_qposadr[belt_j]]; d.act_velocity[0] = d.qvel[m.jnt_dofadr[belt_j]];
  for (int a = 0; a < m.n_arms; ++a) {
    for (int k = 0; k < 7; ++k) {
      int j = m.arm_jnt[9 * a + k];
      d.act_length[1 + 8 * a + k] = d.qpos[m.jnt_qposadr[j]]; d.act_velocity[1 + 8 * a + k] = d.qvel[m.jnt_dofadr[j]];
    }
    int jl = m.arm_jnt[9 * a + 7], jr = m.arm_jnt[9 * a + 8];
    d.act_length[1 + 8 * a + 7] = m.tendon_coef[0] * d.qpos[m.jnt_qposadr[jl]] + m.tendon_coef[1] * d.qpos[m.jnt_qposadr[jr]];
    d.act_velocity[1 + 8 * a + 7] = m.tendon_coef[0] * d.qvel[m.jnt_dofadr[jl]] + m.tendon_coef[1] * d.qvel[m.jnt_dofadr[jr]];
  }
}

// mj_step1: position + velocity dependent quantities for the current (qpos, qvel)
static void step1(const Model& m, Data& d) {
  kinematics(m, d);
  crb(m, d);
  collision(m, d);
  make_constraint(m, d);
  transmission(m, d);
  velocity_stage(m, d);
  reference_accel(m, d);
}

// ------------------------------------------------------------------ exact constraint solve (Newton, primal) [3P]
static void solve_constraints(const Model& m, Data& d) {
  int nv = m.nv, ne = d.nefc;
  std::vector<double>& a = d.qacc;
  if (ne == 0) { a = d.qacc_smooth; std::fill(d.qfrc_constraint.begin(), d.qfrc_constraint.end(), 0.0); d.solver_iter = 0; return; }
  a = d.qacc_warmstart;
  std::vector<double> Ma(nv), jar(ne), grad(nv), search(nv), Mv(nv), jv(ne), H, rhs;
  std::vector<char> active(ne);
  auto mulM = [&](const std::vector<double>& x, std::vector<double>& y) {
    for (size_t t = 0; t < m.tree_start.size(); ++t) {
      int s = m.tree_start[t], n = m.tree_size[t];
      for (int i = 0; i < n; ++i) { double acc = 0; for (int k = 0; k < n; ++k) acc += d.M[(size_t)(s + i) * nv + s + k] * x[s + k]; y[s + i] = acc; }
    }
  };
  auto mulJ = [&](const std::vector<double>& x, std::vector<double>& y) {
    for (int i = 0; i < ne; ++i) { double acc = 0; const double* J = &d.efc_J[(size_t)i * nv]; for (int k = 0; k < nv; ++k) acc += J[k] * x[k]; y[i] = acc; }
  };
  auto cost_of = [&](const std::vector<double>& acc) {
    std::vector<double> ma(nv), jr(ne);
    mulM(acc, ma); mulJ(acc, jr);
    double c = 0;
    for (int k = 0; k < nv; ++k) c += 0.5 * (ma[k] - d.qfrc_smooth[k]) * (acc[k] - d.qacc_smooth[k]);
    for (int i = 0; i < ne; ++i) { double r = jr[i] - d.efc_aref[i]; if (d.efc_type[i] == C_EQ || r < 0) c += 0.5 * d.efc_D[i] * r * r; }
    return c;
  };
  // warm start only if it beats qacc_smooth (as mj_warmstart does)
  if (cost_of(a) > cost_of(d.qacc_smooth)) a = d.qacc_smooth;
  // union-find over trees for block-diagonal Hessian
  int nt = (int)m.tree_start.size();
  std::vector<int> comp(nt), dofs;
  int it = 0; double gnorm = 0, fscale = 0;
  for (int k = 0; k < nv; ++k) fscale += d.qfrc_smooth[k] * d.qfrc_smooth[k];
  fscale = std::sqrt(fscale);
  for (it = 0; it < 100; ++it) {
    mulM(a, Ma); mulJ(a, jar);
    for (int i = 0; i < ne; ++i) { jar[i] -= d.efc_aref[i]; active[i] = d.efc_type[i] == C_EQ || jar[i] < 0; d.efc_force[i] = active[i] ? -d.efc_D[i] * jar[i] : 0; }
    for (int k = 0; k < nv; ++k) grad[k] = Ma[k] - d.qfrc_smooth[k];
    for (int i = 0; i < ne; ++i) if (active[i]) { const double* J = &d.efc_J[(size_t)i * nv]; double f = d.efc_force[i]; for (int k = 0; k < nv; ++k) grad[k] -= J[k] * f; }
    gnorm = 0; for (int k = 0; k < nv; ++k) gnorm += grad[k] * grad[k]; gnorm = std::sqrt(gnorm);
    if (gnorm <= 1e-12 * (1.0 + fscale)) break;
    // components
    for (int t = 0; t < nt; ++t) comp[t] = t;
    auto find = [&](int x) { while (comp[x] != x) { comp[x] = comp[comp[x]]; x = comp[x]; } return x; };
    for (int i = 0; i < ne; ++i) if (active[i]) {
      const double* J = &d.efc_J[(size_t)i * nv]; int first = -1;
      for (int t = 0; t < nt; ++t) {
        bool nz = false; for (int k = 0; k < m.tree_size[t]; ++k) if (J[m.tree_start[t] + k] != 0) { nz = true; break; }
        if (nz) { if (first < 0) first = find(t); else comp[find(t)] = first; }
      }
    }
    for (int k = 0; k < nv; ++k) search[k] = -grad[k];
    for (int c = 0; c < nt; ++c) {
      if (find(c) != c) continue;
      dofs.clear();
      for (int t = 0; t < nt; ++t) if (find(t) == c) for (int k = 0; k < m.tree_size[t]; ++k) dofs.push_back(m.tree_start[t] + k);
      int n = (int)dofs.size();
      H.assign((size_t)n * n, 0);
      for (int i = 0; i < n; ++i) for (int k = 0; k < n; ++k) H[(size_t)i * n + k] = d.M[(size_t)dofs[i] * nv + dofs[k]];
      for (int r = 0; r < ne; ++r) if (active[r]) {
        const double* J = &d.efc_J[(size_t)r * nv];
        if (J[dofs[0]] == 0) { bool any = false; for (int i = 0; i < n; ++i) if (J[dofs[i]] != 0) { any = true; break; } if (!any) continue; }
        for (int i = 0; i < n; ++i) { double ji = J[dofs[i]]; if (ji == 0) continue; double w = d.efc_D[r] * ji; for (int k = 0; k <= i; ++k) H[(size_t)i * n + k] += w * J[dofs[k]]; }
      }
      for (int i = 0; i < n; ++i) for (int k = i + 1; k < n; ++k) H[(size_t)i * n + k] = H[(size_t)k * n + i];
      if (!chol_block(H.data(), n, 0, n)) throw std::runtime_error("newton hessian not PD");
      rhs.resize(n); for (int i = 0; i < n; ++i) rhs[i] = -grad[dofs[i]];
      chol_solve_block(H.data(), n, 0, n, rhs.data());
      for (int i = 0; i < n; ++i) search[dofs[i]] = rhs[i];
      d.flops += (long long)n * n * n / 3 + 2LL * n * n;
    }
    // exact line search on the piecewise-quadratic cost
    mulM(search, Mv); mulJ(search, jv);
    double g0 = 0, h0 = 0;
    for (int k = 0; k < nv; ++k) { g0 += search[k] * (Ma[k] - d.qfrc_smooth[k]); h0 += search[k] * Mv[k]; }
    double A = g0, Bc = h0;
    std::vector<std::pair<double, int>> bp;
    std::vector<char> act(active);
    for (int i = 0; i < ne; ++i) {
      if (act[i]) { A += d.efc_D[i] * jv[i] * jar[i]; Bc += d.efc_D[i] * jv[i] * jv[i]; }
      if (d.efc_type[i] != C_EQ && jv[i] != 0) { double al = -jar[i] / jv[i]; if (al > 0) bp.push_back({al, i}); }
    }
    std::sort(bp.begin(), bp.end());
    double alpha = 0; size_t k = 0;
    while (true) {
      double cand = Bc > 0 ? -A / Bc : 1e300;
      double next = k < bp.size() ? bp[k].first : 1e300;
      if (cand <= next) { alpha = std::max(cand, 0.0); break; }
      int i = bp[k].second;
      double sg = act[i] ? -1.0 : 1.0;
      A += sg * d.efc_D[i] * jv[i] * jar[i]; Bc += sg * d.efc_D[i] * jv[i] * jv[i];
      act[i] = !act[i];
      ++k;
    }
    if (alpha <= 0) break;
    for (int q = 0; q < nv; ++q) a[q] += alpha * search[q];
  }
  d.solver_iter = it; d.solver_grad = gnorm;
  // final forces
  mulJ(a, jar);
  std::fill(d.qfrc_constraint.begin(), d.qfrc_constraint.end(), 0.0);
  for (int i = 0; i < ne; ++i) {
    double r = jar[i] - d.efc_aref[i];
    bool on = d.efc_type[i] == C_EQ || r < 0;
    d.efc_force[i] = on ? -d.efc_D[i] * r : 0;
    if (on) { const double* J = &d.efc_J[(size_t)i * nv]; for (int k = 0; k < nv; ++k) d.qfrc_constraint[k] += J[k] * d.efc_force[i]; }
  }
}

// ------------------------------------------------------------------ mj_step2: actuation, acceleration, constraint, implicitfast
static void step2(const Model& m, Data& d) {
  int nv = m.nv;
  // actuation (mj_fwdActuation) [3P]
  std::fill(d.qfrc_actuator.begin(), d.qfrc_actuator.end(), 0.0);
  std::vector<char> clamped(m.nu, 0);
  for (int u = 0; u < m.nu; ++u) {
    double c = std::min(std::max(d.ctrl[u], m.act_ctrlrange[2 * u]), m.act_ctrlrange[2 * u + 1]);
    double f = m.act_gain[u] * c + m.act_bias[3 * u] + m.act_bias[3 * u + 1] * d.act_length[u] + m.act_bias[3 * u + 2] * d.act_velocity[u];
    if (m.act_forcelimited[u]) {
      if (f <= m.act_forcerange[2 * u]) { f = m.act_forcerange[2 * u]; clamped[u] = 1; }
      else if (f >= m.act_forcerange[2 * u + 1]) { f = m.act_forcerange[2 * u + 1]; clamped[u] = 1; }
    }
    d.act_force[u] = f;
  }
  int belt_d = m.jnt_dofadr[m.body_jnt[m.belt_body]];
  d.qfrc_actuator[belt_d] += d.act_force[0];
  for (int a = 0; a < m.n_arms; ++a) {
    for (int k = 0; k < 7; ++k) d.qfrc_actuator[m.jnt_dofadr[m.arm_jnt[9 * a + k]]] += d.act_force[1 + 8 * a + k];
    d.qfrc_actuator[m.jnt_dofadr[m.arm_jnt[9 * a + 7]]] += m.tendon_coef[0] * d.act_force[1 + 8 * a + 7];
    d.qfrc_actuator[m.jnt_dofadr[m.arm_jnt[9 * a + 8]]] += m.tendon_coef[1] * d.act_force[1 + 8 * a + 7];
  }
  for (int k = 0; k < nv; ++k) d.qfrc_smooth[k] = d.qfrc_passive[k] - d.qfrc_bias[k] + d.qfrc_actuator[k];
  d.qacc_smooth = d.qfrc_smooth;
  solveM(m, d, d.L, d.qacc_smooth.data());
  solve_constraints(m, d);
  d.qacc_warmstart = d.qacc;
  // implicitfast: (M - h*D) qacc = qfrc_smooth + qfrc_constraint, D = d(actuator+passive)/dqvel [3P]
  d.Li = d.M;
  for (int k = 0; k < nv; ++k) d.Li[(size_t)k * nv + k] += m.h * m.jnt_damping[m.dof_jnt[k]];
  d.Li[(size_t)belt_d * nv + belt_d] -= m.h * m.act_bias[2];
  for (int a = 0; a < m.n_arms; ++a) {
    for (int k = 0; k < 7; ++k) { int dd = m.jnt_dofadr[m.arm_jnt[9 * a + k]]; d.Li[(size_t)dd * nv + dd] -= m.h * m.act_bias[3 * (1 + 8 * a + k) + 2]; }
    int u = 1 + 8 * a + 7;
    if (!clamped[u]) {
      int dl = m.jnt_dofadr[m.arm_jnt[9 * a + 7]], dr = m.jnt_dofadr[m.arm_jnt[9 * a + 8]];
      double b2 = m.act_bias[3 * u + 2]; int dd[2] = {dl, dr};
      for (int i = 0; i < 2; ++i) for (int k = 0; k < 2; ++k) d.Li[(size_t)dd[i] * nv + dd[k]] -= m.h * b2 * m.tendon_coef[i] * m.tendon_coef[k];
    }
  }
  for (size_t t = 0; t < m.tree_start.size(); ++t)
    if (!chol_block(d.Li.data(), nv, m.tree_start[t], m.tree_size[t])) throw std::runtime_error("implicit matrix not PD");
  std::vector<double> qa(nv);
  for (int k = 0; k < nv; ++k) qa[k] = d.qfrc_smooth[k] + d.qfrc_constraint[k];
  solveM(m, d, d.Li, qa.data());
  d.qacc = qa;   // [3P] data->qacc holds the implicit acceleration after mj_implicit
  // advance (mj_advance): semi-implicit, positions with the new velocity
  for (int k = 0; k < nv; ++k) d.qvel[k] += m.h * qa[k];
  for (int j = 0; j < m.nj; ++j) {
    int qa_ = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    if (m.jnt_type[j] == J_FREE) {
      for (int k = 0; k < 3; ++k) d.qpos[qa_ + k] += m.h * d.qvel[da + k];
      V3 w = {d.qvel[da + 3], d.qvel[da + 4], d.qvel[da + 5]};
      double ang = norm(w) * m.h;
      Q4 q = qnormalize({d.qpos[qa_ + 3], d.qpos[qa_ + 4], d.qpos[qa_ + 5], d.qpos[qa_ + 6]});
      if (ang > 0) {
        V3 ax = w * (1.0 / norm(w));
        double s = std::sin(ang / 2);
        q = qmul(q, {std::cos(ang / 2), ax.x * s, ax.y * s, ax.z * s});
      }
      d.qpos[qa_ + 3] = q.w; d.qpos[qa_ + 4] = q.x; d.qpos[qa_ + 5] = q.y; d.qpos[qa_ + 6] = q.z;
    } else {
      d.qpos[qa_] += m.h * d.qvel[da];
    }
  }
}

}  // namespace fmo
