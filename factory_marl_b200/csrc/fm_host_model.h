// fm_host_model.h — host-side: model blob (include/fm_blob.h, written by factory_marl_b200/model.py) -> FmDevModel
// plus the broad-phase pair table and the per-geom contact classes.  Host only.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include <cmath>
#include "../../include/fm_blob.h"
#include "fm_types.h"

struct FmHostModel {
  FmDevModel dm;
  FmIkModel ik;
  std::vector<unsigned int> cpairs;     // packed pair records: c1 | c2 << 8 | test type << 16 (big-vs-big pairs resolved here)
  std::vector<int> geom_class;
  int n_arms = 0, nq = 0, nv = 0, nu = 0, ngeom = 0;
  double timestep = 0.001;   // fp64 copy: frame_skip = int((1/control_frequency)/timestep) must be evaluated in double
};

namespace fmh {
inline const double* D(const void* b, size_t n, const char* name, size_t* cnt = nullptr) {
  const double* p = (const double*)fm_blob_find(b, n, name, 1, cnt);
  if (!p) throw std::runtime_error(std::string("model blob: missing float array ") + name);
  return p;
}
inline const int32_t* I(const void* b, size_t n, const char* name, size_t* cnt = nullptr) {
  const int32_t* p = (const int32_t*)fm_blob_find(b, n, name, 0, cnt);
  if (!p) throw std::runtime_error(std::string("model blob: missing int array ") + name);
  return p;
}
}  // namespace fmh

inline void fm_build_host_model(const void* blob, size_t n, FmHostModel& H) {
  using namespace fmh;
  FmDevModel& m = H.dm;
  memset(&m, 0, sizeof(m));
  size_t c = 0;
  const int32_t* sizes = I(blob, n, "sizes");
  H.ngeom = sizes[1]; H.nq = sizes[3]; H.nv = sizes[4]; H.nu = sizes[5];
  m.n_arms = H.n_arms = I(blob, n, "n_arms")[0];
  if (m.n_arms > FM_MAX_ARMS) throw std::runtime_error("num_arms > 6 is not supported by this build");
  if (I(blob, n, "n_cubes")[0] != FM_NCUBE) throw std::runtime_error("max_num_objects must be 10");
  const double* opt = D(blob, n, "opt");
  m.h = (float)opt[0]; m.g = (float)(-opt[1]); H.timestep = opt[0];
  const int32_t* gclass = I(blob, n, "g_geom_class", &c);
  H.geom_class.assign(gclass, gclass + c);
  const double* sb = D(blob, n, "g_sbox", &c);
  m.n_sbox = (int)(c / 15);
  if (m.n_sbox > FM_MAX_SBOX) throw std::runtime_error("too many static boxes");
  const int32_t* sbg = I(blob, n, "g_sbox_geom");
  for (int i = 0; i < m.n_sbox; ++i) {
    for (int k = 0; k < 3; ++k) { m.sbox_pos[i][k] = (float)sb[15 * i + k]; m.sbox_half[i][k] = (float)sb[15 * i + 12 + k]; }
    { const double top = sb[15 * i + 2] + sb[15 * i + 14]; m.sbox_top[i][0] = (float)top; m.sbox_top[i][1] = (float)(top - (double)(float)top); }
    for (int k = 0; k < 9; ++k) m.sbox_R[i][k] = (float)sb[15 * i + 3 + k];
    m.sbox_geom[i] = sbg[i]; m.sbox_class[i] = gclass[sbg[i]];
    { const float* R = m.sbox_R[i]; m.sbox_aligned[i] = (R[0] == 1.f && R[4] == 1.f && R[8] == 1.f && R[1] == 0.f && R[2] == 0.f && R[3] == 0.f && R[5] == 0.f && R[6] == 0.f && R[7] == 0.f) ? 1 : 0; }
  }
  const double* bl = D(blob, n, "g_belt");
  for (int k = 0; k < 3; ++k) { m.belt_pos[k] = (float)bl[k]; m.belt_half[k] = (float)bl[3 + k]; }
  { const double top = bl[2] + bl[5]; m.belt_top[0] = (float)top; m.belt_top[1] = (float)(top - (double)(float)top); }
  m.belt_mass = (float)bl[6]; m.belt_damp = (float)bl[7]; m.belt_kv = (float)bl[8]; m.belt_invw = (float)bl[9];
  m.belt_geom = I(blob, n, "g_belt_geom")[0]; m.belt_class = gclass[m.belt_geom];
  const double* cu = D(blob, n, "g_cube");
  const int32_t* cg = I(blob, n, "g_cube_geom");
  for (int k = 0; k < FM_NCUBE; ++k) {
    m.cube_half[k] = (float)cu[4 * k]; m.cube_mass[k] = (float)cu[4 * k + 1]; m.cube_inertia[k] = (float)cu[4 * k + 2]; m.cube_invw[k] = (float)cu[4 * k + 3];
    m.cube_geom[k] = cg[k]; m.cube_invm[k] = (float)(1.0 / cu[4 * k + 1]); m.cube_invI[k] = (float)(1.0 / cu[4 * k + 2]);
  }
  m.cube_class = gclass[cg[0]]; m.floor_class = gclass[0];
  const double* pp = D(blob, n, "g_pair_param", &c);
  int ncl = 1; while (ncl * ncl * 8 < (int)c) ++ncl;
  if (ncl > FM_NCLASS) throw std::runtime_error("too many contact parameter classes");
  for (int i = 0; i < ncl; ++i) for (int j = 0; j < ncl; ++j) for (int k = 0; k < 8; ++k) m.pair_param[i][j][k] = (float)pp[(i * ncl + j) * 8 + k];
  const double* ab = D(blob, n, "g_arm_body");
  for (int b = 0; b < FM_NARMBODY; ++b) {
    const double* r = ab + 25 * b;
    for (int k = 0; k < 3; ++k) { m.ab_pos[b][k] = (float)r[k]; m.ab_ipos[b][k] = (float)r[8 + k]; }
    for (int k = 0; k < 4; ++k) m.ab_quat[b][k] = (float)r[3 + k];
    m.ab_mass[b] = (float)r[7];
    for (int k = 0; k < 6; ++k) m.ab_I[b][k] = (float)r[11 + k];
    m.ab_jtype[b] = (int)r[17];
    // kernel assumes hinge axis = local z, slide axis = local x (iiwa14.xml:9, gripper.xml:23,40)
    if (m.ab_jtype[b] == 1 && !(r[18] == 0 && r[19] == 0 && r[20] == 1)) throw std::runtime_error("hinge axis must be local z");
    if (m.ab_jtype[b] == 2 && !(r[18] == 1 && r[19] == 0 && r[20] == 0)) throw std::runtime_error("slide axis must be local x");
    m.ab_range[b][0] = (float)r[21]; m.ab_range[b][1] = (float)r[22]; m.ab_invw[b] = (float)r[23]; m.ab_dofinvw[b] = (float)r[24];
  }
  const double* base = D(blob, n, "g_arm_base");
  for (int a = 0; a < m.n_arms; ++a) { for (int k = 0; k < 3; ++k) m.arm_base_pos[a][k] = (float)base[7 * a + k]; for (int k = 0; k < 4; ++k) m.arm_base_quat[a][k] = (float)base[7 * a + 3 + k]; }
  auto quat_to_mat = [](const double* q, float* R) {
    double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double w = q[0] / nq, x = q[1] / nq, y = q[2] / nq, z = q[3] / nq;
    double Rd[9] = {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), w * w - x * x + y * y - z * z,
                    2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
    for (int i = 0; i < 9; ++i) R[i] = (float)Rd[i];
  };
  for (int b = 0; b < FM_NARMBODY; ++b) quat_to_mat(ab + 25 * b + 3, m.ab_C[b]);
  for (int a = 0; a < m.n_arms; ++a) quat_to_mat(base + 7 * a + 3, m.arm_base_R[a]);
  const double* sp = D(blob, n, "g_arm_sph", &c);
  if (c / 6 != FM_NSPH) throw std::runtime_error("unexpected arm sphere count");
  int prevb = -1;
  for (int i = 0; i < FM_NSPH; ++i) {
    int b = (int)sp[6 * i];
    if (b < prevb) throw std::runtime_error("arm spheres must be sorted by body");
    prevb = b;
    m.sph_body[i] = b; m.sph_geom[i] = (int)sp[6 * i + 1];
    for (int k = 0; k < 3; ++k) m.sph_pos[i][k] = (float)sp[6 * i + 2 + k];
    m.sph_r[i] = (float)sp[6 * i + 5];
  }
  for (int b = 0; b <= FM_NARMBODY; ++b) { int cnt = 0; for (int i = 0; i < FM_NSPH; ++i) if (m.sph_body[i] < b) cnt++; m.body_sph_start[b] = cnt; }
  const double* bx = D(blob, n, "g_arm_box", &c);
  if (c / 8 != FM_NABOX) throw std::runtime_error("unexpected arm box count");
  for (int i = 0; i < FM_NABOX; ++i) {
    m.abox_body[i] = (int)bx[8 * i]; m.abox_geom[i] = (int)bx[8 * i + 1];
    for (int k = 0; k < 3; ++k) { m.abox_pos[i][k] = (float)bx[8 * i + 2 + k]; m.abox_half[i][k] = (float)bx[8 * i + 5 + k]; }
  }
  for (int b = 0; b <= FM_NARMBODY; ++b) { int cnt = 0; for (int i = 0; i < FM_NABOX; ++i) if (m.abox_body[i] < b) cnt++; m.body_box_start[b] = cnt; }
  const int32_t* fg = I(blob, n, "g_arm_first_geom");
  for (int a = 0; a < m.n_arms; ++a) m.arm_first_geom[a] = fg[a];
  m.sph_class = gclass[fg[0] + m.sph_geom[0]]; m.abox_class = gclass[fg[0] + m.abox_geom[0]];
  const double* cl = D(blob, n, "g_arm_cluster");
  for (int b = 0; b < FM_NARMBODY; ++b) for (int k = 0; k < 4; ++k) m.arm_cl[b][k] = (float)cl[4 * b + k];
  const double* act = D(blob, n, "g_act");
  m.kp = (float)act[0]; m.kd = (float)act[2]; m.g_kp = (float)act[3]; m.g_kb = (float)act[4]; m.g_kd = (float)act[5];
  m.g_fmin = (float)act[6]; m.g_fmax = (float)act[7]; m.ten0 = (float)act[8]; m.ten1 = (float)act[9];
  if (act[0] != act[1]) throw std::runtime_error("position servo: gain and bias must match");
  const double* cr = D(blob, n, "g_ctrlrange");
  for (int k = 0; k < 9; ++k) { m.ctrl_lo[k] = (float)cr[2 * k]; m.ctrl_hi[k] = (float)cr[2 * k + 1]; }
  const double* eq = D(blob, n, "g_eq"); const double* lim = D(blob, n, "g_limit");
  for (int k = 0; k < 2; ++k) { m.eq_solref[k] = (float)eq[k]; m.lim_solref[k] = (float)lim[k]; }
  for (int k = 0; k < 5; ++k) { m.eq_solimp[k] = (float)eq[2 + k]; m.lim_solimp[k] = (float)lim[2 + k]; }
  const double* site = D(blob, n, "g_site");
  for (int k = 0; k < 3; ++k) m.site_pos[k] = (float)site[k];
  m.site_body = (int)site[3];
  const double* bk = D(blob, n, "g_bucket");
  for (int b = 0; b < 2; ++b) for (int k = 0; k < 3; ++k) { m.bucket_pos[b][k] = (float)bk[6 * b + k]; m.bucket_size[b][k] = (float)bk[6 * b + 3 + k]; }
  // fp64 IK constants
  FmIkModel& k = H.ik;
  memset(&k, 0, sizeof(k));
  k.n_arms = m.n_arms;
  const double* abp = D(blob, n, "arm_base_pos");
  for (int a = 0; a < m.n_arms; ++a) {
    for (int q = 0; q < 3; ++q) { k.base_pos[a][q] = base[7 * a + q]; k.player_site[a][q] = abp[3 * a + q]; }
    for (int q = 0; q < 4; ++q) k.base_quat[a][q] = base[7 * a + 3 + q];
  }
  for (int b = 0; b < 9; ++b) { for (int q = 0; q < 3; ++q) k.body_pos[b][q] = ab[25 * b + q]; for (int q = 0; q < 4; ++q) k.body_quat[b][q] = ab[25 * b + 3 + q]; }
  for (int q = 0; q < 3; ++q) k.site_pos[q] = site[q];
  if ((int)site[3] != 8) throw std::runtime_error("gripper site must be on gripper_base");
  const double* grp = D(blob, n, "cube_group_wxyz");
  for (int g = 0; g < 24; ++g) for (int q = 0; q < 4; ++q) k.cube_group[g][q] = grp[4 * g + q];
  for (int b = 0; b < 2; ++b) for (int q = 0; q < 3; ++q) k.bucket_pos[b][q] = bk[6 * b + q];
  for (int q = 0; q < 8; ++q) { k.ctrl_lo[q] = cr[2 * (q + 1)]; k.ctrl_hi[q] = cr[2 * (q + 1) + 1]; }
  const int32_t* clk = I(blob, n, "g_cluster", &c);
  m.n_cluster = (int)(c / 2);
  if (m.n_cluster > (int)(sizeof(m.cl_kind) / sizeof(int))) throw std::runtime_error("too many clusters");
  for (int i = 0; i < m.n_cluster; ++i) { m.cl_kind[i] = clk[2 * i]; m.cl_idx[i] = clk[2 * i + 1]; }
  // conservative world AABBs of the big geoms (rotated boxes: |R| * half); the belt slides along +y over its whole length
  m.cl_sph0 = m.n_sbox + 2;
  for (int i = 0; i < m.n_sbox; ++i)
    for (int ax = 0; ax < 3; ++ax) {
      float e = fabsf(m.sbox_R[i][3 * ax]) * m.sbox_half[i][0] + fabsf(m.sbox_R[i][3 * ax + 1]) * m.sbox_half[i][1] + fabsf(m.sbox_R[i][3 * ax + 2]) * m.sbox_half[i][2];
      m.big_aabb[1 + i][ax] = m.sbox_pos[i][ax] - e - 1e-6f; m.big_aabb[1 + i][3 + ax] = m.sbox_pos[i][ax] + e + 1e-6f;
    }
  for (int ax = 0; ax < 3; ++ax) {
    float ext = ax == 1 ? 1e9f : m.belt_half[ax] + 1e-6f;
    m.big_aabb[m.n_sbox + 1][ax] = m.belt_pos[ax] - ext; m.big_aabb[m.n_sbox + 1][3 + ax] = m.belt_pos[ax] + ext;
  }
  const int32_t* cp = I(blob, n, "g_cluster_pair", &c);
  H.cpairs.clear();
  for (size_t i = 0; i + 1 < c; i += 2) {
    int c1 = cp[i], c2 = cp[i + 1];
    int k1 = m.cl_kind[c1], k2 = m.cl_kind[c2];
    bool s1 = k1 == FM_CL_CUBE || k1 == FM_CL_ARM, s2 = k2 == FM_CL_CUBE || k2 == FM_CL_ARM;
    unsigned type;
    if (s1 && s2) type = FM_PAIR_SPH_SPH;
    else if (s2 && k1 == FM_CL_PLANE) type = FM_PAIR_PLANE_SPH;
    else if (s2) type = FM_PAIR_BOX_SPH;
    else {   // two big geoms: only the belt moves, along its own long axis -> decided once
      if (k1 == FM_CL_PLANE) { if (m.big_aabb[c2][2] > 0.f) continue; }
      else {
        bool sep = false;
        for (int ax = 0; ax < 3; ++ax) sep = sep || m.big_aabb[c1][ax] > m.big_aabb[c2][3 + ax] || m.big_aabb[c2][ax] > m.big_aabb[c1][3 + ax];
        if (sep) continue;
      }
      type = FM_PAIR_ALWAYS;
    }
    H.cpairs.push_back((unsigned)c1 | ((unsigned)c2 << 8) | (type << 16));
  }
  m.n_cluster_pair = (int)H.cpairs.size();
}
