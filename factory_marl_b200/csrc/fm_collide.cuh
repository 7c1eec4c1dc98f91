// fm_collide.cuh — fp32 primitive collision for the step kernel: plane/sphere/box pairs of this scene.
// Conventions follow MuJoCo's contact definition [3P]: normal from geom1 to geom2, pos midway between the
// surfaces, dist < 0 in penetration (margin 0).  Replaces the narrow phase inside physics.step()
// (reference call site base_env.py:218).  Box-box is an own SAT + face-clipping formulation.
#pragma once
#include "fm_math.cuh"

struct RawC { f3 pos, n; float dist; };

FM_HD int col_plane_sphere(f3 sp, float r, RawC* out) {      // floor: z = 0, normal +z (scene.xml:21)
  float dist = sp.z - r;
  if (dist > 0.f) return 0;
  out[0].pos = mk3(sp.x, sp.y, sp.z - (r + 0.5f * dist)); out[0].n = mk3(0, 0, 1); out[0].dist = dist;
  return 1;
}

FM_HD int col_plane_box(f3 bp, const float* bR, f3 hs, RawC* out) {
  float dist = bp.z;
  int cnt = 0;
  for (int i = 0; i < 8; ++i) {
    f3 vec = mk3((i & 1) ? hs.x : -hs.x, (i & 2) ? hs.y : -hs.y, (i & 4) ? hs.z : -hs.z);
    f3 corner = rmul(bR, vec);
    float ld = corner.z;
    if (dist + ld > 0.f || ld > 0.f) continue;
    float cd = dist + ld;
    out[cnt].pos = mk3(corner.x + bp.x, corner.y + bp.y, corner.z + bp.z - 0.5f * cd);
    out[cnt].n = mk3(0, 0, 1); out[cnt].dist = cd;
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

FM_HD int col_sphere_sphere(f3 p1, float r1, f3 p2, float r2, RawC* out) {
  f3 dd = p2 - p1;
  float l2 = dot3(dd, dd), rr = r1 + r2;
  if (l2 > rr * rr) return 0;
  float len = sqrtf(l2);
  float dist = len - rr;
  if (dist > 0.f) return 0;
  f3 n = len < 1e-15f ? mk3(1, 0, 0) : dd * (1.0f / len);
  out[0].pos = p1 + n * (r1 + 0.5f * dist); out[0].n = n; out[0].dist = dist;
  return 1;
}

FM_HD int col_sphere_box(f3 sp, float r, f3 bp, const float* bR, f3 hs, RawC* out) {
  f3 c = rmulT(bR, sp - bp);
  f3 cl = mk3(clampf(c.x, -hs.x, hs.x), clampf(c.y, -hs.y, hs.y), clampf(c.z, -hs.z, hs.z));
  f3 dv = cl - c;
  float dl2 = dot3(dv, dv);
  if (dl2 > r * r) return 0;
  float dl = sqrtf(dl2);
  if (dl > 1e-15f) {
    float dist = dl - r;
    if (dist > 0.f) return 0;
    f3 n = rmul(bR, dv * (1.0f / dl));
    out[0].pos = sp + n * (r + 0.5f * dist); out[0].n = n; out[0].dist = dist;
    return 1;
  }
  int k = 0; float best = 1e30f;
  for (int i = 0; i < 3; ++i) { float g = comp3(hs, i) - fabsf(comp3(c, i)); if (g < best) { best = g; k = i; } }
  float s = comp3(c, k) >= 0.f ? 1.f : -1.f;
  f3 n = rmul(bR, unit3(k, -s));
  float dist = -best - r;
  out[0].pos = sp + n * (r + 0.5f * dist); out[0].n = n; out[0].dist = dist;
  return 1;
}

// conservative sphere-vs-box bound test
FM_HD bool far_sphere_box(f3 sp, float r, f3 bp, const float* bR, f3 hs) {
  f3 c = rmulT(bR, sp - bp);
  float dx = fmaxf(fabsf(c.x) - hs.x, 0.f), dy = fmaxf(fabsf(c.y) - hs.y, 0.f), dz = fmaxf(fabsf(c.z) - hs.z, 0.f);
  return dx * dx + dy * dy + dz * dz > r * r;
}

// box-box: 15-axis separating-axis search, then reference-face clipping (face axes) or closest points of the
// two edges (edge axes).  Up to 8 contacts.  Same formulation as the oracle, in fp32.
// one out-of-line copy: the routine is large, cold for most pairs and called from three places of the narrow phase
FM_HDN int col_box_box(f3 p1, const float* R1, f3 A, f3 p2, const float* R2, f3 B, RawC* out) {
  f3 pw = p2 - p1;
  f3 pp = rmulT(R1, pw);
  float R[9], Q[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { R[3 * i + j] = dot3(rcol(R1, i), rcol(R2, j)); Q[3 * i + j] = fabsf(R[3 * i + j]); }
  float best = -1e30f; int code = -1; bool flip = false; f3 nbest = mk3(0, 0, 0);
  for (int i = 0; i < 3; ++i) {
    float e = comp3(pp, i);
    float s = fabsf(e) - (comp3(A, i) + B.x * Q[3 * i] + B.y * Q[3 * i + 1] + B.z * Q[3 * i + 2]);
    if (s > 0.f) return 0;
    if (s > best) { best = s; code = i; flip = e < 0.f; nbest = rcol(R1, i); }
  }
  for (int j = 0; j < 3; ++j) {
    float e = dot3(rcol(R2, j), pw);
    float s = fabsf(e) - (comp3(B, j) + A.x * Q[j] + A.y * Q[3 + j] + A.z * Q[6 + j]);
    if (s > 0.f) return 0;
    if (s > best + 1e-6f) { best = s; code = 3 + j; flip = e < 0.f; nbest = rcol(R2, j); }
  }
  const float fudge = 1.05f;
  // edge axes a_i x b_j in closed form (box-1 frame): |a_i x b_j|^2 = 1 - R_ij^2, projections from R, |R| and pp
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      float l2 = 1.f - R[3 * i + j] * R[3 * i + j], e, s;
      if (l2 < 0.02f) {      // nearly parallel edges: 1 - R^2 cancels in fp32, take the axis itself
        f3 ax = cross3(rcol(R1, i), rcol(R2, j));
        float l = norm3(ax);
        if (l < 1e-4f) continue;
        ax = ax * (1.0f / l);
        e = dot3(ax, pw);
        float ra = A.x * fabsf(dot3(ax, rcol(R1, 0))) + A.y * fabsf(dot3(ax, rcol(R1, 1))) + A.z * fabsf(dot3(ax, rcol(R1, 2)));
        float rb = B.x * fabsf(dot3(ax, rcol(R2, 0))) + B.y * fabsf(dot3(ax, rcol(R2, 1))) + B.z * fabsf(dot3(ax, rcol(R2, 2)));
        s = fabsf(e) - (ra + rb);
      } else {
        e = comp3(pp, i2) * R[3 * i1 + j] - comp3(pp, i1) * R[3 * i2 + j];
        float ra = comp3(A, i1) * Q[3 * i2 + j] + comp3(A, i2) * Q[3 * i1 + j];
        float rb = comp3(B, j1) * Q[3 * i + j2] + comp3(B, j2) * Q[3 * i + j1];
        s = (fabsf(e) - (ra + rb)) * rsqrtf(l2);
      }
      if (s > 0.f) return 0;
      if (s > best * (1.0f / fudge) + 1e-6f) { best = s; code = 6 + 3 * i + j; flip = e < 0.f; }
    }
  }
  if (code >= 6) { int i = (code - 6) / 3, j = (code - 6) % 3; f3 ax = cross3(rcol(R1, i), rcol(R2, j)); nbest = ax * (1.0f / norm3(ax)); }
  f3 n = flip ? nbest * -1.0f : nbest;
  float depth = -best;
  if (code >= 6) {
    int i = (code - 6) / 3, j = (code - 6) % 3;
    f3 pa = p1, pb = p2;
    for (int k = 0; k < 3; ++k) {
      if (k != i) { float sg = dot3(n, rcol(R1, k)) > 0.f ? 1.f : -1.f; pa = pa + rcol(R1, k) * (sg * comp3(A, k)); }
      if (k != j) { float sg = dot3(n, rcol(R2, k)) > 0.f ? -1.f : 1.f; pb = pb + rcol(R2, k) * (sg * comp3(B, k)); }
    }
    f3 ua = rcol(R1, i), ub = rcol(R2, j);
    f3 dp = pb - pa;
    float uaub = dot3(ua, ub), q1 = dot3(ua, dp), q2 = -dot3(ub, dp);
    float dn = 1.f - uaub * uaub;
    float al = 0.f, be = 0.f;
    if (dn > 1e-8f) { al = (q1 + uaub * q2) / dn; be = (uaub * q1 + q2) / dn; }
    al = clampf(al, -comp3(A, i), comp3(A, i)); be = clampf(be, -comp3(B, j), comp3(B, j));
    f3 ca = pa + ua * al, cb = pb + ub * be;
    out[0].pos = (ca + cb) * 0.5f; out[0].n = n; out[0].dist = -depth;
    return 1;
  }
  bool ref1 = code < 3;
  int ax = ref1 ? code : code - 3;
  const float* Rr = ref1 ? R1 : R2; const float* Ri = ref1 ? R2 : R1;
  f3 Sr = ref1 ? A : B, Si = ref1 ? B : A;
  f3 pr = ref1 ? p1 : p2, pi = ref1 ? p2 : p1;
  f3 nr = ref1 ? n : n * -1.0f;
  int iax = 0; float mn = 1e30f, isg = 1.f;
  for (int k = 0; k < 3; ++k) {
    float dd = dot3(rcol(Ri, k), nr);
    if (dd < mn) { mn = dd; iax = k; isg = 1.f; }
    if (-dd < mn) { mn = -dd; iax = k; isg = -1.f; }
  }
  int i1 = (iax + 1) % 3, i2 = (iax + 2) % 3;
  f3 fc = pi + rcol(Ri, iax) * (isg * comp3(Si, iax));
  f3 e1 = rcol(Ri, i1) * comp3(Si, i1), e2 = rcol(Ri, i2) * comp3(Si, i2);
  int r1 = (ax + 1) % 3, r2 = (ax + 2) % 3;
  f3 u1 = rcol(Rr, r1), u2 = rcol(Rr, r2);
  float rsg = dot3(rcol(Rr, ax), nr) > 0.f ? 1.f : -1.f;
  f3 rc = pr + rcol(Rr, ax) * (rsg * comp3(Sr, ax));
  float poly[16][3], tmp[16][3]; int np = 4;
  for (int k = 0; k < 4; ++k) {
    f3 q = fc + e1 * ((k == 0 || k == 3) ? 1.f : -1.f) + e2 * ((k < 2) ? 1.f : -1.f);
    f3 dq = q - rc;
    poly[k][0] = dot3(dq, u1); poly[k][1] = dot3(dq, u2); poly[k][2] = dot3(dq, nr);
  }
  const float lim[2] = {comp3(Sr, r1), comp3(Sr, r2)};
  for (int c = 0; c < 2; ++c)
    for (int sg = -1; sg <= 1; sg += 2) {
      bool any_out = false;                       // a side plane that cuts nothing leaves the polygon as it is (most do)
      for (int k = 0; k < np; ++k) any_out = any_out || (lim[c] - sg * poly[k][c] < 0.f);
      if (!any_out) continue;
      int nn = 0;
      for (int k = 0; k < np; ++k) {
        const float* a = poly[k]; const float* b = poly[(k + 1) % np];
        float da = lim[c] - sg * a[c], db = lim[c] - sg * b[c];
        if (da >= 0.f) { tmp[nn][0] = a[0]; tmp[nn][1] = a[1]; tmp[nn][2] = a[2]; nn++; }
        if ((da >= 0.f) != (db >= 0.f)) {
          float t = da / (da - db);
          for (int q = 0; q < 3; ++q) tmp[nn][q] = a[q] + t * (b[q] - a[q]);
          nn++;
        }
      }
      np = nn;
      for (int k = 0; k < np; ++k) for (int q = 0; q < 3; ++q) poly[k][q] = tmp[k][q];
    }
  int cnt = 0;
  for (int k = 0; k < np && cnt < 8; ++k) {
    float hgt = poly[k][2];
    if (hgt > 0.f) continue;
    out[cnt].pos = rc + u1 * poly[k][0] + u2 * poly[k][1] + nr * (hgt * 0.5f);
    out[cnt].n = n; out[cnt].dist = hgt;
    cnt++;
  }
  return cnt;
}

// Fast path of col_box_box for a small box resting on / falling onto the top face of a big AXIS-ALIGNED box (belt, table,
// bucket floor): when the small box's vertices all lie horizontally inside the top face and its centre is above it,
// the separating-axis search of col_box_box always selects the big box's +z face (see tests: fm_emu_check_box_on_face),
// no clipping happens, and the contacts are the vertices of the incident face below the surface, in the same order.
// `big_is_1`: the big box is geom1 (normal +z) or geom2 (normal -z).  Returns -1 when the fast path does not apply.
// `zlo`: low-order part of the small box's z position (compensated integration, EnvS::cube_qlo): the penetration of a resting cube
// (2.3e-6 m) is resolved below the 1.2e-7 m grid of a float near z = 1.13.
// `top_hi`, `top_lo`: the big box's top face height as a float pair (FmDevModel::belt_top / sbox_top); NaN = use pb.z + hb.z.
FM_HD int col_box_on_face(f3 pb, f3 hb, f3 pc, const float* Rc, f3 hc, bool big_is_1, RawC* out, float zlo = 0.f, float top_hi = -1.f, float top_lo = 0.f) {
  float r = norm3(hc);
  float top = top_hi >= 0.f ? top_hi : pb.z + hb.z;
  // horizontal half-extents of the small box (exact projection of its vertices on the big box's x / y axes)
  float ex = fabsf(Rc[0]) * hc.x + fabsf(Rc[1]) * hc.y + fabsf(Rc[2]) * hc.z, ey = fabsf(Rc[3]) * hc.x + fabsf(Rc[4]) * hc.y + fabsf(Rc[5]) * hc.z;
  if (!(fminf(hb.x, hb.y) >= 0.25f) || pc.z < top || fabsf(pc.x - pb.x) + ex * 1.0001f > hb.x || fabsf(pc.y - pb.y) + ey * 1.0001f > hb.y) return -1;
  if (pc.z - r > top) return 0;
  f3 nr = mk3(0, 0, 1);
  int iax = 0; float mn = 1e30f, isg = 1.f;
  for (int k = 0; k < 3; ++k) {
    float dd = dot3(rcol(Rc, k), nr);
    if (dd < mn) { mn = dd; iax = k; isg = 1.f; }
    if (-dd < mn) { mn = -dd; iax = k; isg = -1.f; }
  }
  int i1 = (iax + 1) % 3, i2 = (iax + 2) % 3;
  f3 fv = rcol(Rc, iax) * (isg * comp3(hc, iax));
  f3 fc = pc + fv;
  f3 e1 = rcol(Rc, i1) * comp3(hc, i1), e2 = rcol(Rc, i2) * comp3(hc, i2);
  f3 rc = mk3(pb.x, pb.y, top);
  f3 n = big_is_1 ? nr : nr * -1.0f;
  const float zc = (pc.z - top) + (zlo - top_lo);  // exact difference of two nearby floats, plus what the two float values dropped
  int cnt = 0;
  for (int k = 0; k < 4; ++k) {
    const float s1 = (k == 0 || k == 3) ? 1.f : -1.f, s2 = (k < 2) ? 1.f : -1.f;
    f3 q = fc + e1 * s1 + e2 * s2;
    f3 dq = q - rc;
    float hgt = zc + (fv.z + e1.z * s1 + e2.z * s2);   // height of the vertex above the face from small terms only
    if (hgt > 0.f) continue;
    out[cnt].pos = rc + mk3(1, 0, 0) * dq.x + mk3(0, 1, 0) * dq.y + nr * (hgt * 0.5f);
    out[cnt].n = n; out[cnt].dist = hgt;
    cnt++;
  }
  return cnt;
}
