// fm_math.cuh — small fp32 vector helpers for the step kernel (device; also host-compilable for logic tests).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FM_HD __host__ __device__ __forceinline__
#define FM_D __device__ __forceinline__
#define FM_HDM __host__ __device__
#define FM_HDN static __host__ __device__ __noinline__     /* big phase bodies called from more than one place: one copy */
#else
#define FM_HDM
#define FM_HDN static
#define FM_HD static inline
#define FM_D static inline
#endif

// barrier-interval phase functions (ph_*): with FM_PHASE_NOINLINE each is its own function, so that ptxas allocates registers
// per phase instead of across the whole fused sub-step loop (state crosses phases through shared memory anyway)
#if defined(FM_PHASE_NOINLINE) && defined(__CUDACC__)
#define FM_PH FM_HDN
#else
#define FM_PH FM_HD
#endif

#if !defined(__CUDACC__)
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline void sincosf_host(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
#endif
struct f3 { float x, y, z; };
FM_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
FM_HD f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
FM_HD void st3(float* p, f3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
FM_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
FM_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
FM_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
FM_HD f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
FM_HD float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
FM_HD f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
FM_HD float norm3(f3 a) { return sqrtf(dot3(a, a)); }
FM_HD float comp3(f3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
FM_HD f3 unit3(int i, float s) { return mk3(i == 0 ? s : 0.f, i == 1 ? s : 0.f, i == 2 ? s : 0.f); }
// R row-major 3x3 at pointer
FM_HD f3 rmul(const float* R, f3 v) {
  return mk3(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z);
}
FM_HD f3 rmulT(const float* R, f3 v) {
  return mk3(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z);
}
FM_HD f3 rcol(const float* R, int j) { return mk3(R[j], R[3 + j], R[6 + j]); }

struct q4 { float w, x, y, z; };
FM_HD q4 mkq(float w, float x, float y, float z) { q4 q; q.w = w; q.x = x; q.y = y; q.z = z; return q; }
FM_HD q4 qmulf(q4 a, q4 b) {
  return mkq(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
             a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
FM_HD q4 qnormf(q4 q) {
  float n = sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < 1e-15f) return mkq(1, 0, 0, 0);
  float s = 1.0f / n;
  return mkq(q.w * s, q.x * s, q.y * s, q.z * s);
}
FM_HD void q2matf(q4 q, float* R) {
  float w = q.w, x = q.x, y = q.y, z = q.z;
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
// symmetric 3x3 packed xx,yy,zz,xy,xz,yz times vector
FM_HD f3 symmul6(const float* I, f3 v) {
  return mk3(I[0] * v.x + I[3] * v.y + I[4] * v.z, I[3] * v.x + I[1] * v.y + I[5] * v.z, I[4] * v.x + I[5] * v.y + I[2] * v.z);
}
FM_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
// MuJoCo-style contact frame from a unit normal [3P mju_makeFrame]
FM_HD void make_frame3(f3 n, f3& t1, f3& t2) {
  f3 y = (n.y < 0.5f && n.y > -0.5f) ? mk3(0, 1, 0) : mk3(0, 0, 1);
  y = y - n * dot3(n, y);
  y = y * (1.0f / norm3(y));
  t1 = y; t2 = cross3(n, y);
}
