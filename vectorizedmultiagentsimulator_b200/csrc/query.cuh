// Distance / overlap queries between two entities of one env (ref core.py:1788-1969), on VALUES: shared by
// the query kernels of vmas_b200.cu (entities loaded from the slab and the plan tables) and by the fused step
// kernel's epilogue (entities taken from the registers of the thread that just stepped the env, shapes
// known at compile time — every branch on `shape` below then folds away).
#pragma once
#include "geometry.cuh"
#include "vmas_b200.h"

namespace vmas {

#ifndef VMAS_QUERY_CONSTANTS
#define VMAS_QUERY_CONSTANTS
constexpr float Q_HALF_PI_F = 1.57079632679489661923f;  // fp32(torch.pi / 2)
constexpr float Q_LINE_MIN_DIST_F = (float)(4.0 / 6e2);
#endif

struct EntG {
  int shape;
  V2 p;
  float rot, d0, d1;
  float r_plus_lmd;  // spheres: fp32(radius + LINE_MIN_DIST), the sum taken in double as the reference does
};

DEVI Seg seg_of(const EntG& g) {
  float sn, cs;
  sincosf(g.rot, &sn, &cs);
  return mkseg(g.p, cs, sn, g.d0 / 2.f);
}

DEVI BoxG box_of(const EntG& g) {
  BoxG b;
  b.p = g.p;
  sincosf(g.rot, &b.s, &b.c);
  sincosf(g.rot + Q_HALF_PI_F, &b.s2, &b.c2);
  b.half_l = g.d0 / 2.f;
  b.half_w = g.d1 / 2.f;
  return b;
}

// ref core.py:1788-1820
DEVI float dist_from_point(const EntG& g, V2 pt) {
  if (g.shape == VMAS_SHAPE_SPHERE) return norm2(g.p - pt) - g.d0;
  V2 cp = (g.shape == VMAS_SHAPE_BOX) ? closest_point_box(box_of(g), pt) : closest_point_seg(seg_of(g), pt);
  return norm2(pt - cp) - Q_LINE_MIN_DIST_F;
}

// ref core.py:1933-1964 (box / sphere overlap)
DEVI bool box_sphere_overlap(const EntG& box, const EntG& sph) {
  V2 cp = closest_point_box(box_of(box), sph.p);
  float d_s_cp = norm2(sph.p - cp);
  float d_s_b = norm2(sph.p - box.p);
  float d_cp_b = norm2(box.p - cp);
  return (d_s_b < d_cp_b) || (d_s_cp < sph.r_plus_lmd);
}

// Exact early-out for is_overlapping: true only if the two shapes are separated by clearly more
// than the overlap threshold (sphere radii / LINE_MIN_DIST, ref core.py:1907-1969), so the answer
// is "no" without any closest-point arithmetic.  Conservative tests (bounding circles; for boxes
// the other shape's extent in the box frame) with a 1e-3 margin that dwarfs fp32 rounding.
DEVI bool overlap_impossible(const EntG& ga, const EntG& gb) {
  const float M = 1e-3f;
  const int sa = ga.shape, sb = gb.shape;
  if (sa == VMAS_SHAPE_SPHERE && sb == VMAS_SHAPE_SPHERE) return false;  // one norm: nothing to save
  if (sa == VMAS_SHAPE_BOX && sb == VMAS_SHAPE_BOX) {
    const float ra = 0.5f * norm2(ga.d0, ga.d1), rb = 0.5f * norm2(gb.d0, gb.d1);
    const float lim = ra + rb + Q_LINE_MIN_DIST_F + M;
    const V2 d = ga.p - gb.p;
    return d.x * d.x + d.y * d.y > lim * lim;
  }
  if (sa == VMAS_SHAPE_BOX || sb == VMAS_SHAPE_BOX) {
    const EntG& box = sa == VMAS_SHAPE_BOX ? ga : gb;
    const EntG& other = sa == VMAS_SHAPE_BOX ? gb : ga;
    float sn, cs;
    sincosf(box.rot, &sn, &cs);
    const V2 d = other.p - box.p;
    const float lx = d.x * cs + d.y * sn, ly = d.y * cs - d.x * sn;  // other's centre in the box frame
    float ex, ey, reach;
    if (other.shape == VMAS_SHAPE_SPHERE) {
      ex = ey = 0.f;
      reach = other.d0 + Q_LINE_MIN_DIST_F + M;
    } else {  // line: half-length projected on the box axes
      float so, co;
      sincosf(other.rot, &so, &co);
      const float half = other.d0 / 2.f;
      ex = half * fabsf(co * cs + so * sn);
      ey = half * fabsf(so * cs - co * sn);
      reach = Q_LINE_MIN_DIST_F + M;
    }
    return fabsf(lx) - ex > box.d0 / 2.f + reach || fabsf(ly) - ey > box.d1 / 2.f + reach;
  }
  // line - sphere, line - line: bounding circles
  const float ra = sa == VMAS_SHAPE_LINE ? ga.d0 / 2.f : ga.d0, rb = sb == VMAS_SHAPE_LINE ? gb.d0 / 2.f : gb.d0;
  const float lim = ra + rb + Q_LINE_MIN_DIST_F + M;
  const V2 d = ga.p - gb.p;
  return d.x * d.x + d.y * d.y > lim * lim;
}

// ref core.py:1822-1905
DEVI float pair_distance(const EntG& ga, const EntG& gb) {
  const int sa = ga.shape, sb = gb.shape;
  if (sa == VMAS_SHAPE_SPHERE && sb == VMAS_SHAPE_SPHERE) return dist_from_point(ga, gb.p) - gb.d0;
  if ((sa == VMAS_SHAPE_BOX && sb == VMAS_SHAPE_SPHERE) || (sb == VMAS_SHAPE_BOX && sa == VMAS_SHAPE_SPHERE)) {
    const bool a_is_box = sa == VMAS_SHAPE_BOX;
    const EntG& box = a_is_box ? ga : gb;
    const EntG& sph = a_is_box ? gb : ga;
    float d = dist_from_point(box, sph.p) - sph.d0;
    return box_sphere_overlap(box, sph) ? -1.f : d;
  }
  if ((sa == VMAS_SHAPE_LINE && sb == VMAS_SHAPE_SPHERE) || (sb == VMAS_SHAPE_LINE && sa == VMAS_SHAPE_SPHERE)) {
    const EntG& line = sa == VMAS_SHAPE_LINE ? ga : gb;
    const EntG& sph = sa == VMAS_SHAPE_LINE ? gb : ga;
    return dist_from_point(line, sph.p) - sph.d0;
  }
  Pair c;
  if (sa == VMAS_SHAPE_LINE && sb == VMAS_SHAPE_LINE) {
    c = closest_seg_seg(seg_of(ga), seg_of(gb));
  } else if (sa == VMAS_SHAPE_BOX && sb == VMAS_SHAPE_BOX) {
    c = closest_box_box(box_of(ga), box_of(gb));
  } else {
    const EntG& box = sa == VMAS_SHAPE_BOX ? ga : gb;
    const EntG& line = sa == VMAS_SHAPE_BOX ? gb : ga;
    c = closest_box_seg(box_of(box), seg_of(line));
  }
  return norm2(c.a - c.b) - Q_LINE_MIN_DIST_F;
}

// World.is_overlapping of one pair in one env (ref core.py:1907-1969)
DEVI bool pair_overlap(const EntG& ga, const EntG& gb) {
  const bool box_sphere = (ga.shape == VMAS_SHAPE_BOX && gb.shape == VMAS_SHAPE_SPHERE) ||
                          (gb.shape == VMAS_SHAPE_BOX && ga.shape == VMAS_SHAPE_SPHERE);
  if (overlap_impossible(ga, gb)) return false;
  if (box_sphere) {
    const bool a_is_box = ga.shape == VMAS_SHAPE_BOX;
    return box_sphere_overlap(a_is_box ? ga : gb, a_is_box ? gb : ga);
  }
  return pair_distance(ga, gb) < 0.f;
}


__device__ __noinline__ static float obs_remainder(float v, float m) {  // torch.remainder: sign follows the modulus
  float r = fmodf(v, m);
  if (r != 0.f && (signbit(m) != signbit(r))) r = r + m;
  return r;
}

}  // namespace vmas
