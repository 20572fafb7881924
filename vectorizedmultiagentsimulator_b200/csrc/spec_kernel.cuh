// spec_kernel.cuh — world-specialised substep kernel.
//
// The generic kernels in vmas_b200.cu interpret the plan tables at run time (table loads,
// dynamically indexed shared memory, a switch per work item) and end up latency-bound.  Here the
// world's static structure is a compile-time constant: a generated header (csrc/generated/) holds
// one `struct World_<hash>` per pre-registered world with constexpr entity / item tables, and this
// template unrolls every loop over entities and work items.  After unrolling all indices are
// constants, so an env's whole state (positions, velocities, rotations, force accumulators, cached
// sin/cos) lives in REGISTERS of the one thread that owns the env, shape parameters fold into
// immediates, and the kind dispatch disappears.  The arithmetic is the same device functions
// (geometry.cuh) in the same order as the generic kernels: results are bit-identical (tested).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

#include "geometry.cuh"
#include "query.cuh"
#include "vmas_b200.h"

namespace vmas {

struct EntC {
  int shape, flags, agent;
  float d0, d1, mass, inertia, drag_mult, lin_fric, ang_fric, grav_x, grav_y, max_speed, v_range, max_f, f_range,
      max_t, t_range, circ_r, r_plus_lmd;
};
struct ItemC {
  int kind, a, b, flags, mask_bit;
  float dmin_base, ax, ay, bx, by, dist, fixed_rot, broad_thr;
};
struct CfgC {
  int substeps, has_x_semidim, has_y_semidim, has_world_gravity;
  float sub_dt, x_semidim, y_semidim, collision_force, joint_force, torque_constraint_force, contact_margin,
      gravity_x, gravity_y;
};

// The whole-step kernel's epilogue (see spec_epilogue): the scenario's step program and observation rows,
// compile-time data like the world itself
struct ProgC {
  int op, dst, a, b, arg;
  float imm;
};
struct ObsColC {
  int op, src, src2;  // as the observation gather's column codes: (field << 24) | element offset in the env's row
  float par;
};
struct EpiArgs {
  float* obs_out;  // [rows, B, width] or null
  void* buffers[VMAS_PROG_MAX_BUFFERS];
};
// ... and its prologue (see spec_ingest): the policy agents' continuous holonomic actions
struct ActC {
  int agent;  // row of the agent in the force / torque slab
  float range0, range1, mult0, mult1;
};
struct ActArgs {
  const float* actions[VMAS_MAX_INGEST_AGENTS];  // [B, 2] each, the caller's tensors
  float* u[VMAS_MAX_INGEST_AGENTS];              // [B, 2] each: agent.action.u
  uint8_t* bad_flag;
  float* steps;  // [B] or null
  int clamp;
};

struct SpecArgs {
  VmasState st;
  const float* joint_rot;  // [B, n_joints] or null
  uint32_t* mask;          // [mask_words + 1]
  int batch_dim;
  int use_mask;
  int first_substep;
  int n_substeps;
  // env scheduling (optional, see vmas_b200_build_env_order): thread t steps env order[t]; every env
  // records which of its work items produced a force (bit I & 31) for the next re-ordering
  const int32_t* order;  // [B] permutation of the envs, or null: thread t steps env t
  uint32_t* sig;         // [B] out, or null
};

template <class F, int... I>
DEVI void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
DEVI void static_for(F&& f) {
  static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

constexpr float SPEC_HALF_PI_F = 1.57079632679489661923f;
constexpr float SPEC_FAR_MARGIN = 1e-3f;

DEVI bool spec_far_apart(V2 a, V2 b, float reach) {
  V2 d = a - b;
  float lim = reach + SPEC_FAR_MARGIN;
  return d.x * d.x + d.y * d.y > lim * lim;
}

// Register-resident state of one env.
template <int E>
struct EnvRegs {
  float px[E], py[E], rot[E], vx[E], vy[E], w[E];
  float c[E], s[E], c2[E], s2[E];
  float Fx[E], Fy[E], T[E];
};

template <class W, int EI, int E>
DEVI Seg spec_seg(const EnvRegs<E>& r) {
  return mkseg(mk(r.px[EI], r.py[EI]), r.c[EI], r.s[EI], W::ent[EI].d0 / 2.f);
}
template <class W, int EI, int E>
DEVI BoxG spec_box(const EnvRegs<E>& r) {
  BoxG b;
  b.p = mk(r.px[EI], r.py[EI]);
  b.c = r.c[EI];
  b.s = r.s[EI];
  b.c2 = r.c2[EI];
  b.s2 = r.s2[EI];
  b.half_l = W::ent[EI].d0 / 2.f;
  b.half_w = W::ent[EI].d1 / 2.f;
  return b;
}

// One work item, fully resolved at compile time; accumulates into the env's force registers in
// the reference's order (ref core.py:2191-2199).
// TRACK: also record in `sig` whether the item produced a force (env scheduling; off in the default kernel)
template <class W, int I, bool TRACK, int E>
DEVI void spec_item(EnvRegs<E>& r, const SpecArgs& a, long env, const uint32_t* mask_words, uint32_t& sig) {
  constexpr ItemC it = W::item[I];
  constexpr int A = it.a, B = it.b;
  constexpr EntC ea = W::ent[A], eb = W::ent[B];
  constexpr CfgC cfg = W::cfg;
  if constexpr (it.mask_bit >= 0) {
    if (a.use_mask && !((mask_words[it.mask_bit >> 5] >> (it.mask_bit & 31)) & 1u)) return;
  }
  V2 f = mk(0.f, 0.f);
  float ta = 0.f, tb = 0.f;
  const V2 pa = mk(r.px[A], r.py[A]), pb = mk(r.px[B], r.py[B]);

  if constexpr (it.kind == VMAS_K_JOINT) {
    V2 qa = pa + rot2(mk(it.ax, it.ay), r.c[A], r.s[A]);
    V2 qb = pb + rot2(mk(it.bx, it.by), r.c[B], r.s[B]);
    V2 f_attr = constraint_force(qa, qb, it.dist, cfg.joint_force, cfg.contact_margin, true);
    V2 f_rep = constraint_force(qa, qb, it.dist, cfg.joint_force, cfg.contact_margin, false);
    f = f_attr + f_rep;
    V2 fb = neg(f_attr) + neg(f_rep);
    ta = cross2(qa - pa, f);
    tb = cross2(qb - pb, fb);
    if constexpr (!(it.flags & VMAS_IFLAG_JOINT_ROTATE)) {
      float jr = a.joint_rot ? a.joint_rot[(size_t)env * W::N_JOINTS + I] : it.fixed_rot;
      float delta = r.rot[A] - (r.rot[B] + jr);
      float mag = sqrtf(delta * delta);
      float t = (cfg.torque_constraint_force * sgnf(delta)) * (expf(mag) - 1.f);
      if (mag < 1e-9f) t = 0.f;
      ta = ta + (-t);
      tb = tb + t;
    }
  } else if constexpr (it.kind == VMAS_K_SS) {
    f = constraint_force(pa, pb, it.dmin_base, cfg.collision_force, cfg.contact_margin, false);
  } else if constexpr (it.kind == VMAS_K_LS) {  // a = line, b = sphere
    Seg l = spec_seg<W, A>(r);
    if (!spec_far_apart(l.p, pb, l.half + it.dmin_base)) {
      V2 cp = closest_point_seg(l, pb);
      V2 f_sphere = constraint_force(pb, cp, it.dmin_base, cfg.collision_force, cfg.contact_margin, false);
      f = neg(f_sphere);
      ta = cross2(cp - l.p, f);
    }
  } else if constexpr (it.kind == VMAS_K_LL) {
    Seg l1 = spec_seg<W, A>(r), l2 = spec_seg<W, B>(r);
    if (!spec_far_apart(l1.p, l2.p, l1.half + l2.half + it.dmin_base)) {
      Pair c = closest_seg_seg(l1, l2);
      f = constraint_force(c.a, c.b, it.dmin_base, cfg.collision_force, cfg.contact_margin, false);
      ta = cross2(c.a - l1.p, f);
      tb = cross2(c.b - l2.p, neg(f));
    }
  } else if constexpr (it.kind == VMAS_K_BS) {  // a = box, b = sphere
    BoxG bx = spec_box<W, A>(r);
    V2 d0 = pb - bx.p;
    float lx = d0.x * bx.c + d0.y * bx.s, ly = d0.y * bx.c - d0.x * bx.s;
    if (!(fabsf(lx) > bx.half_l + it.dmin_base + SPEC_FAR_MARGIN ||
          fabsf(ly) > bx.half_w + it.dmin_base + SPEC_FAR_MARGIN)) {
      V2 cp = closest_point_box(bx, pb);
      V2 inner = cp;
      float d = 0.f;
      if constexpr (!(ea.flags & VMAS_F_HOLLOW)) inner = inner_point_box(pb, cp, bx.p, &d);
      V2 f_sphere =
          constraint_force(pb, inner, it.dmin_base + d, cfg.collision_force, cfg.contact_margin, false);
      f = neg(f_sphere);
      ta = cross2(cp - bx.p, f);
    }
  } else if constexpr (it.kind == VMAS_K_BL) {  // a = box, b = line
    BoxG bx = spec_box<W, A>(r);
    Seg l = spec_seg<W, B>(r);
    V2 d0 = l.p - bx.p;
    float lx = d0.x * bx.c + d0.y * bx.s, ly = d0.y * bx.c - d0.x * bx.s;
    float ex = l.half * fabsf(l.c * bx.c + l.s * bx.s), ey = l.half * fabsf(l.s * bx.c - l.c * bx.s);
    if (!(fabsf(lx) - ex > bx.half_l + it.dmin_base + SPEC_FAR_MARGIN ||
          fabsf(ly) - ey > bx.half_w + it.dmin_base + SPEC_FAR_MARGIN)) {
      Pair c = closest_box_seg(bx, l);
      V2 inner = c.a;
      float d = 0.f;
      if constexpr (!(ea.flags & VMAS_F_HOLLOW)) inner = inner_point_box(c.b, c.a, bx.p, &d);
      f = constraint_force(inner, c.b, it.dmin_base + d, cfg.collision_force, cfg.contact_margin, false);
      ta = cross2(c.a - bx.p, f);
      tb = cross2(c.b - l.p, neg(f));
    }
  } else if constexpr (it.kind == VMAS_K_BB) {
    BoxG b1 = spec_box<W, A>(r), b2 = spec_box<W, B>(r);
    if (!spec_far_apart(b1.p, b2.p, ea.circ_r + eb.circ_r + it.dmin_base)) {
      Pair c = closest_box_box(b1, b2);
      V2 in1 = c.a, in2 = c.b;
      float d1 = 0.f, d2 = 0.f;
      if constexpr (!(ea.flags & VMAS_F_HOLLOW)) in1 = inner_point_box(c.b, c.a, b1.p, &d1);
      if constexpr (!(eb.flags & VMAS_F_HOLLOW)) in2 = inner_point_box(c.a, c.b, b2.p, &d2);
      f = constraint_force(in1, in2, (d1 + d2) + it.dmin_base, cfg.collision_force, cfg.contact_margin, false);
      ta = cross2(c.a - b1.p, f);
      tb = cross2(c.b - b2.p, neg(f));
    }
  }

  if constexpr (TRACK && it.kind != VMAS_K_JOINT) {
    if (f.x != 0.f || f.y != 0.f) sig |= 1u << (I & 31);  // this env took the contact branch of item I
  }
  if constexpr (ea.flags & VMAS_F_MOVABLE) {
    r.Fx[A] = r.Fx[A] + f.x;
    r.Fy[A] = r.Fy[A] + f.y;
  }
  if constexpr (ea.flags & VMAS_F_ROTATABLE) r.T[A] = r.T[A] + ta;
  if constexpr (eb.flags & VMAS_F_MOVABLE) {
    r.Fx[B] = r.Fx[B] + (-f.x);
    r.Fy[B] = r.Fy[B] + (-f.y);
  }
  if constexpr (eb.flags & VMAS_F_ROTATABLE) r.T[B] = r.T[B] + tb;
}

// ---------------------------------------------------------------------------------------------
// per-thread row I/O.  One env's slice of a state tensor ("row", NF floats) is contiguous; the
// owning thread moves it with the widest vector access the row size allows (16 bytes when
// NF % 4 == 0), so a sector is touched by at most two requests instead of once per scalar.
// MASK has one bit per column: vector chunks without any wanted column are skipped at compile
// time; a chunk with at least one is moved whole.
// ---------------------------------------------------------------------------------------------
template <int NF>
struct RowVec {
  static constexpr int W = (NF % 4 == 0) ? 4 : (NF % 2 == 0) ? 2 : 1;
};

__host__ __device__ constexpr uint64_t chunk_closure(uint64_t mask, int nf, int w) {
  uint64_t out = 0;
  for (int c = 0; c < nf; c += w) {
    const uint64_t bits = ((w >= 64 ? ~0ull : ((1ull << w) - 1)) << c);
    if (mask & bits) out |= bits;
  }
  return out;
}

template <int NF, uint64_t MASK>
DEVI void row_load(const float* __restrict__ g, float (&dst)[NF > 0 ? NF : 1]) {
  if constexpr (NF > 0) {
    constexpr int W = RowVec<NF>::W;
    static_for<NF / W>([&](auto ci) {
      constexpr int c = decltype(ci)::value * W;
      if constexpr ((MASK >> c) & ((1ull << W) - 1)) {
        if constexpr (W == 4) {
          const float4 v = *reinterpret_cast<const float4*>(g + c);
          dst[c] = v.x; dst[c + 1] = v.y; dst[c + 2] = v.z; dst[c + 3] = v.w;
        } else if constexpr (W == 2) {
          const float2 v = *reinterpret_cast<const float2*>(g + c);
          dst[c] = v.x; dst[c + 1] = v.y;
        } else {
          dst[c] = g[c];
        }
      }
    });
  }
}

template <int NF, uint64_t MASK>
DEVI void row_store(float* __restrict__ g, const float (&src)[NF > 0 ? NF : 1]) {
  if constexpr (NF > 0) {
    constexpr int W = RowVec<NF>::W;
    static_for<NF / W>([&](auto ci) {
      constexpr int c = decltype(ci)::value * W;
      if constexpr ((MASK >> c) & ((1ull << W) - 1)) {
        if constexpr (W == 4) {
          *reinterpret_cast<float4*>(g + c) = make_float4(src[c], src[c + 1], src[c + 2], src[c + 3]);
        } else if constexpr (W == 2) {
          *reinterpret_cast<float2*>(g + c) = make_float2(src[c], src[c + 1]);
        } else {
          g[c] = src[c];
        }
      }
    });
  }
}

// column masks derived from the world's entity flags at compile time
template <class W>
__host__ __device__ constexpr uint64_t ent_cols(int flag_any, int width) {
  uint64_t m = 0;
  for (int e = 0; e < W::E; ++e)
    if (W::ent[e].flags & flag_any) m |= ((width == 2 ? 3ull : 1ull) << (width * e));
  return m;
}
template <class W>
__host__ __device__ constexpr uint64_t agent_cols(int flag_all, int flag_any, int width) {
  uint64_t m = 0;
  for (int e = 0; e < W::E; ++e) {
    const int f = W::ent[e].flags;
    if ((f & VMAS_F_AGENT) && (f & flag_all) == flag_all && (flag_any == 0 || (f & flag_any)))
      m |= ((width == 2 ? 3ull : 1ull) << (width * W::ent[e].agent));
  }
  return m;
}

// Tuning knobs.  SPEC_BLOCK: threads (= envs) per block of the specialised kernels.
// SPEC_MIN_BLOCKS: resident blocks per SM the register allocator must leave room for
// (65536 / (SPEC_BLOCK * SPEC_MIN_BLOCKS) registers per thread at most).
#ifndef SPEC_BLOCK
#define SPEC_BLOCK 64
#endif
#ifndef SPEC_MIN_BLOCKS
#define SPEC_MIN_BLOCKS 1
#endif

// ---------------------------------------------------------------------------------------------
// The per-entity statements of one substep, shared by every specialised mapping (thread per env,
// warp tile): same statements, same order => same bits.
// ---------------------------------------------------------------------------------------------
// sin / cos of the entities whose orientation the geometry needs
template <class W, int E>
DEVI void spec_trig(EnvRegs<E>& r) {
  static_for<E>([&](auto ei) {
    constexpr int e = decltype(ei)::value;
    constexpr EntC en = W::ent[e];
    if constexpr (en.flags & VMAS_F_TRIG) {
      sincosf(r.rot[e], &r.s[e], &r.c[e]);
      if constexpr (en.shape == VMAS_SHAPE_BOX) sincosf(r.rot[e] + SPEC_HALF_PI_F, &r.s2[e], &r.c2[e]);
    }
  });
}

// action force / torque (clamped in place), friction, gravity of every entity -> r.Fx, r.Fy, r.T
// (ref core.py:1995-2004, 2018-2102)
template <class W, int E, int NAX>
DEVI void spec_entity_forces(EnvRegs<E>& r, float (&afx)[NAX], float (&afy)[NAX], float (&atq)[NAX]) {
  constexpr float sub_dt = W::cfg.sub_dt;
  static_for<E>([&](auto ei) {
    constexpr int e = decltype(ei)::value;
    constexpr EntC en = W::ent[e];
    float Fx = 0.f, Fy = 0.f, T = 0.f;
    if constexpr (en.flags & VMAS_F_AGENT) {
      constexpr int ai = en.agent;
      if constexpr (en.flags & VMAS_F_MOVABLE) {
        if constexpr (en.flags & VMAS_F_MAX_F) {
          const float n = norm2(afx[ai], afy[ai]);
          if (n > en.max_f) {
            afx[ai] = (afx[ai] / n) * en.max_f;
            afy[ai] = (afy[ai] / n) * en.max_f;
          }
        }
        if constexpr (en.flags & VMAS_F_F_RANGE) {
          afx[ai] = fminf(fmaxf(afx[ai], -en.f_range), en.f_range);
          afy[ai] = fminf(fmaxf(afy[ai], -en.f_range), en.f_range);
        }
        Fx = Fx + afx[ai];
        Fy = Fy + afy[ai];
      }
      if constexpr (en.flags & VMAS_F_ROTATABLE) {
        if constexpr (en.flags & VMAS_F_MAX_T) {
          const float n = sqrtf(atq[ai] * atq[ai]);
          if (n > en.max_t) atq[ai] = (atq[ai] / n) * en.max_t;
        }
        if constexpr (en.flags & VMAS_F_T_RANGE) atq[ai] = fminf(fmaxf(atq[ai], -en.t_range), en.t_range);
        T = T + atq[ai];
      }
    }
    if constexpr (en.flags & VMAS_F_LIN_FRIC) {
      const float speed = norm2(r.vx[e], r.vy[e]);
      if (speed != 0.f) {
        const float cap = en.lin_fric * en.mass;
        Fx = Fx + (-(r.vx[e] / speed)) * fminf(cap, (fabsf(r.vx[e]) / sub_dt) * en.mass);
        Fy = Fy + (-(r.vy[e] / speed)) * fminf(cap, (fabsf(r.vy[e]) / sub_dt) * en.mass);
      }
    }
    if constexpr (en.flags & VMAS_F_ANG_FRIC) {
      const float speed = sqrtf(r.w[e] * r.w[e]);
      if (speed != 0.f) {
        const float cap = en.ang_fric * en.inertia;
        T = T + (-(r.w[e] / speed)) * fminf(cap, (fabsf(r.w[e]) / sub_dt) * en.inertia);
      }
    }
    if constexpr (en.flags & VMAS_F_MOVABLE) {
      if constexpr (W::cfg.has_world_gravity) {
        Fx = Fx + en.mass * W::cfg.gravity_x;
        Fy = Fy + en.mass * W::cfg.gravity_y;
      }
      if constexpr (en.flags & VMAS_F_GRAVITY) {
        Fx = Fx + en.mass * en.grav_x;
        Fy = Fy + en.mass * en.grav_y;
      }
    }
    r.Fx[e] = Fx;
    r.Fy[e] = Fy;
    r.T[e] = T;
  });
}

// semi-implicit Euler of every entity (ref core.py:2862-2908); `sub` is the substep's index in the step
template <class W, int E>
DEVI void spec_integrate(EnvRegs<E>& r, const int sub) {
  constexpr float sub_dt = W::cfg.sub_dt;
  static_for<E>([&](auto ei) {
    constexpr int e = decltype(ei)::value;
    constexpr EntC en = W::ent[e];
    if constexpr (en.flags & VMAS_F_MOVABLE) {
      if (sub == 0) {
        r.vx[e] = r.vx[e] * en.drag_mult;
        r.vy[e] = r.vy[e] * en.drag_mult;
      }
      r.vx[e] = r.vx[e] + div_pos(r.Fx[e], en.mass) * sub_dt;
      r.vy[e] = r.vy[e] + div_pos(r.Fy[e], en.mass) * sub_dt;
      if constexpr (en.flags & VMAS_F_MAX_SPEED) {
        const float n = norm2(r.vx[e], r.vy[e]);
        if (n > en.max_speed) {
          r.vx[e] = (r.vx[e] / n) * en.max_speed;
          r.vy[e] = (r.vy[e] / n) * en.max_speed;
        }
      }
      if constexpr (en.flags & VMAS_F_V_RANGE) {
        r.vx[e] = fminf(fmaxf(r.vx[e], -en.v_range), en.v_range);
        r.vy[e] = fminf(fmaxf(r.vy[e], -en.v_range), en.v_range);
      }
      r.px[e] = r.px[e] + r.vx[e] * sub_dt;
      r.py[e] = r.py[e] + r.vy[e] * sub_dt;
      if constexpr (W::cfg.has_x_semidim) r.px[e] = fminf(fmaxf(r.px[e], -W::cfg.x_semidim), W::cfg.x_semidim);
      if constexpr (W::cfg.has_y_semidim) r.py[e] = fminf(fmaxf(r.py[e], -W::cfg.y_semidim), W::cfg.y_semidim);
    }
    if constexpr (en.flags & VMAS_F_ROTATABLE) {
      if (sub == 0) r.w[e] = r.w[e] * en.drag_mult;
      r.w[e] = r.w[e] + div_pos(r.T[e], en.inertia) * sub_dt;
      r.rot[e] = r.rot[e] + r.w[e] * sub_dt;
    }
  });
}

// The rows of one env (pos, vel, rot, ang_vel, agent force, agent torque) and which vector chunks
// of them are read / written: derived from the world's entity flags at compile time.
// ALL_FORCE: every movable agent's force columns are stored (the whole-step kernel ingests the actions itself)
template <class W, bool ALL_FORCE = false>
struct SpecRows {
  static constexpr int E = W::E, NA = W::A;
  static constexpr uint64_t ALL_POS = (2 * E >= 64) ? ~0ull : ((1ull << (2 * E)) - 1);
  static constexpr uint64_t MOV2 = ent_cols<W>(VMAS_F_MOVABLE, 2);
  static constexpr uint64_t ROT1 = ent_cols<W>(VMAS_F_ROTATABLE, 1);
  static constexpr uint64_t F_DIRTY =
      ALL_FORCE ? agent_cols<W>(VMAS_F_MOVABLE, 0, 2) : agent_cols<W>(VMAS_F_MOVABLE, VMAS_F_MAX_F | VMAS_F_F_RANGE, 2);
  static constexpr uint64_t T_DIRTY = agent_cols<W>(VMAS_F_ROTATABLE, VMAS_F_MAX_T | VMAS_F_T_RANGE, 1);
  // a vector chunk that will be stored must have been loaded whole (it carries unchanged columns)
  static constexpr uint64_t VEL_IO = chunk_closure(MOV2, 2 * E, RowVec<2 * E>::W);
  static constexpr uint64_t ROT_ST = chunk_closure(ROT1, E, RowVec<E>::W);
  static constexpr uint64_t ROT_LD = ROT_ST | ent_cols<W>(VMAS_F_TRIG | VMAS_F_ROTATABLE, 1);
  static constexpr uint64_t F_ST = chunk_closure(F_DIRTY, 2 * NA, RowVec<2 * NA>::W);
  static constexpr uint64_t T_ST = chunk_closure(T_DIRTY, NA, RowVec<NA>::W);
  static constexpr uint64_t F_LD = F_ST | agent_cols<W>(VMAS_F_MOVABLE, 0, 2);
  static constexpr uint64_t T_LD = T_ST | agent_cols<W>(VMAS_F_ROTATABLE, 0, 1);

  float pos[2 * E], vel[2 * E], rot[E], w[E];
  float f[NA > 0 ? 2 * NA : 1], t[NA > 0 ? NA : 1];

  DEVI void load_pos_rot(const SpecArgs& a, const long env) {
    row_load<2 * E, ALL_POS>(a.st.pos + (size_t)env * 2 * E, pos);
    row_load<E, ROT_LD>(a.st.rot + (size_t)env * E, rot);
  }
  DEVI void load_rest(const SpecArgs& a, const long env) {
    row_load<2 * E, VEL_IO>(a.st.vel + (size_t)env * 2 * E, vel);
    row_load<E, ROT_ST>(a.st.ang_vel + (size_t)env * E, w);
    row_load<2 * NA, F_LD>(a.st.force + (size_t)env * 2 * NA, f);
    row_load<NA, T_LD>(a.st.torque + (size_t)env * NA, t);
  }
  // rows -> registers (pos, rot; trig caches and velocities zeroed)
  DEVI void unpack_pos_rot(EnvRegs<E>& r) const {
    static_for<E>([&](auto ei) {
      constexpr int e = decltype(ei)::value;
      constexpr EntC en = W::ent[e];
      r.px[e] = pos[2 * e];
      r.py[e] = pos[2 * e + 1];
      r.rot[e] = (en.flags & (VMAS_F_TRIG | VMAS_F_ROTATABLE)) ? rot[e] : 0.f;
      r.vx[e] = r.vy[e] = r.w[e] = 0.f;
      r.c[e] = r.s[e] = r.c2[e] = r.s2[e] = 0.f;
    });
  }
  template <int NAX>
  DEVI void unpack_rest(EnvRegs<E>& r, float (&afx)[NAX], float (&afy)[NAX], float (&atq)[NAX]) const {
    static_for<E>([&](auto ei) {
      constexpr int e = decltype(ei)::value;
      constexpr EntC en = W::ent[e];
      if constexpr (en.flags & VMAS_F_MOVABLE) {
        r.vx[e] = vel[2 * e];
        r.vy[e] = vel[2 * e + 1];
      }
      if constexpr (en.flags & VMAS_F_ROTATABLE) r.w[e] = w[e];
      if constexpr (en.flags & VMAS_F_AGENT) {
        if constexpr (en.flags & VMAS_F_MOVABLE) {
          afx[en.agent] = f[2 * en.agent];
          afy[en.agent] = f[2 * en.agent + 1];
        }
        if constexpr (en.flags & VMAS_F_ROTATABLE) atq[en.agent] = t[en.agent];
      }
    });
  }
  // registers -> rows -> global: vector stores of the chunks that hold a changed column
  template <int NAX>
  DEVI void store(const SpecArgs& a, const long env, const EnvRegs<E>& r, const float (&afx)[NAX],
                  const float (&afy)[NAX], const float (&atq)[NAX]) {
    static_for<E>([&](auto ei) {
      constexpr int e = decltype(ei)::value;
      constexpr EntC en = W::ent[e];
      if constexpr (en.flags & VMAS_F_MOVABLE) {
        pos[2 * e] = r.px[e];
        pos[2 * e + 1] = r.py[e];
        vel[2 * e] = r.vx[e];
        vel[2 * e + 1] = r.vy[e];
      }
      if constexpr (en.flags & VMAS_F_ROTATABLE) {
        rot[e] = r.rot[e];
        w[e] = r.w[e];
      }
      if constexpr (en.flags & VMAS_F_AGENT) {
        if constexpr ((en.flags & VMAS_F_MOVABLE) && (ALL_FORCE || (en.flags & (VMAS_F_MAX_F | VMAS_F_F_RANGE)))) {
          f[2 * en.agent] = afx[en.agent];
          f[2 * en.agent + 1] = afy[en.agent];
        }
        if constexpr ((en.flags & VMAS_F_ROTATABLE) && (en.flags & (VMAS_F_MAX_T | VMAS_F_T_RANGE)))
          t[en.agent] = atq[en.agent];
      }
    });
    row_store<2 * E, VEL_IO>(a.st.pos + (size_t)env * 2 * E, pos);
    row_store<2 * E, VEL_IO>(a.st.vel + (size_t)env * 2 * E, vel);
    row_store<E, ROT_ST>(a.st.rot + (size_t)env * E, rot);
    row_store<E, ROT_ST>(a.st.ang_vel + (size_t)env * E, w);
    row_store<2 * NA, F_ST>(a.st.force + (size_t)env * 2 * NA, f);
    row_store<NA, T_ST>(a.st.torque + (size_t)env * NA, t);
  }
};

// ---- the whole-step kernel's epilogue -----------------------------------------------------------------
// What Environment.step does after World.step for a scenario written on a StepProgram and an ObservationPlan
// (vmas_b200_post_step: the reward / done glue and the slab-derived observation columns), executed by the
// thread that just stepped the env, on the registers that still hold its state: the program and the column
// table are constexpr members of `P`, every query is resolved against compile-time shapes.  Same functions,
// same arithmetic, same bits as post_step_kernel.

// one element of an env's state after the step: from the registers where the kernel keeps it, else from the slab
template <class W, int FIELD, int OFF>
DEVI float epi_state(const EnvRegs<W::E>& r, const SpecArgs& a, const long env) {
  constexpr int E = W::E;
  if constexpr (FIELD == VMAS_OBS_POS) {
    return (OFF & 1) ? r.py[OFF / 2] : r.px[OFF / 2];
  } else if constexpr (FIELD == VMAS_OBS_VEL) {
    if constexpr (W::ent[OFF / 2].flags & VMAS_F_MOVABLE)
      return (OFF & 1) ? r.vy[OFF / 2] : r.vx[OFF / 2];
    else
      return a.st.vel[(size_t)env * 2 * E + OFF];
  } else if constexpr (FIELD == VMAS_OBS_ROT) {
    if constexpr (W::ent[OFF].flags & (VMAS_F_TRIG | VMAS_F_ROTATABLE))
      return r.rot[OFF];
    else
      return a.st.rot[(size_t)env * E + OFF];
  } else {
    if constexpr (W::ent[OFF].flags & VMAS_F_ROTATABLE)
      return r.w[OFF];
    else
      return a.st.ang_vel[(size_t)env * E + OFF];
  }
}

template <class W, int EI>
DEVI EntG epi_ent(const EnvRegs<W::E>& r, const SpecArgs& a, const long env) {
  constexpr EntC en = W::ent[EI];
  EntG g;
  g.shape = en.shape;
  g.p = mk(r.px[EI], r.py[EI]);
  g.rot = epi_state<W, VMAS_OBS_ROT, EI>(r, a, env);
  g.d0 = en.d0;
  g.d1 = en.d1;
  g.r_plus_lmd = en.r_plus_lmd;
  return g;
}

// The epilogue's per-env inputs (carried shaping terms, loaded flags) are first touched at the very end of the
// thread's instruction chain; requested when the thread starts, their DRAM round trip is over by then.
DEVI void spec_prefetch(const void* p) {
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

template <class P>
DEVI void spec_epilogue_prefetch(const EpiArgs& e, const long env) {
#ifdef __CUDA_ARCH__
  static_for<P::N_PROG>([&](auto ii) {
    constexpr ProgC in = P::prog[decltype(ii)::value];
    if constexpr (in.op == VMAS_OP_SHAPING || in.op == VMAS_OP_LOAD_F32)
      spec_prefetch(static_cast<const float*>(e.buffers[in.a]) + env);
    else if constexpr (in.op == VMAS_OP_LOAD_BOOL)
      spec_prefetch(static_cast<const uint8_t*>(e.buffers[in.a]) + env);
  });
#endif
}

template <class W, class P>
DEVI void spec_epilogue(const EnvRegs<W::E>& r, const SpecArgs& a, const EpiArgs& e, const long env) {
  // the step program (ref scenarios/balance.py:197-263 as a StepProgram; see vmas_b200_post_step)
  float pr[VMAS_PROG_REGS];
#pragma unroll
  for (int i = 0; i < VMAS_PROG_REGS; ++i) pr[i] = 0.f;
  static_for<P::N_PROG>([&](auto ii) {
    constexpr ProgC in = P::prog[decltype(ii)::value];
    constexpr int ia = in.arg & 0xFFFF, ib = (in.arg >> 16) & 0xFFFF;
    if constexpr (in.op == VMAS_OP_OVERLAP) {
      pr[in.dst] = pair_overlap(epi_ent<W, ia>(r, a, env), epi_ent<W, ib>(r, a, env)) ? 1.f : 0.f;
    } else if constexpr (in.op == VMAS_OP_DISTANCE) {
      pr[in.dst] = pair_distance(epi_ent<W, ia>(r, a, env), epi_ent<W, ib>(r, a, env));
    } else if constexpr (in.op == VMAS_OP_CENTER_DISTANCE) {
      pr[in.dst] = norm2(r.px[ia] - r.px[ib], r.py[ia] - r.py[ib]);
    } else if constexpr (in.op == VMAS_OP_SHAPING) {
      const float d = norm2(r.px[ia] - r.px[ib], r.py[ia] - r.py[ib]);
      const float shaping = d * in.imm;
      float* prev = static_cast<float*>(e.buffers[in.a]) + env;
      pr[in.dst] = *prev - shaping;
      pr[in.dst + 1] = d;
      *prev = shaping;
    } else if constexpr (in.op == VMAS_OP_LOAD_F32) {
      pr[in.dst] = static_cast<const float*>(e.buffers[in.a])[env];
    } else if constexpr (in.op == VMAS_OP_LOAD_BOOL) {
      pr[in.dst] = static_cast<const uint8_t*>(e.buffers[in.a])[env] ? 1.f : 0.f;
    } else if constexpr (in.op == VMAS_OP_CONST) {
      pr[in.dst] = in.imm;
    } else if constexpr (in.op == VMAS_OP_ADD) {
      pr[in.dst] = pr[in.a] + pr[in.b];
    } else if constexpr (in.op == VMAS_OP_SUB) {
      pr[in.dst] = pr[in.a] - pr[in.b];
    } else if constexpr (in.op == VMAS_OP_MUL) {
      pr[in.dst] = pr[in.a] * pr[in.b];
    } else if constexpr (in.op == VMAS_OP_MIN) {
      pr[in.dst] = fminf(pr[in.a], pr[in.b]);
    } else if constexpr (in.op == VMAS_OP_MAX) {
      pr[in.dst] = fmaxf(pr[in.a], pr[in.b]);
    } else if constexpr (in.op == VMAS_OP_NEG) {
      pr[in.dst] = -pr[in.a];
    } else if constexpr (in.op == VMAS_OP_OR) {
      pr[in.dst] = (pr[in.a] != 0.f || pr[in.b] != 0.f) ? 1.f : 0.f;
    } else if constexpr (in.op == VMAS_OP_AND) {
      pr[in.dst] = (pr[in.a] != 0.f && pr[in.b] != 0.f) ? 1.f : 0.f;
    } else if constexpr (in.op == VMAS_OP_NOT) {
      pr[in.dst] = pr[in.a] != 0.f ? 0.f : 1.f;
    } else if constexpr (in.op == VMAS_OP_LT) {
      pr[in.dst] = pr[in.a] < pr[in.b] ? 1.f : 0.f;
    } else if constexpr (in.op == VMAS_OP_LE) {
      pr[in.dst] = pr[in.a] <= pr[in.b] ? 1.f : 0.f;
    } else if constexpr (in.op == VMAS_OP_WHERE) {
      pr[in.dst] = pr[in.a] != 0.f ? pr[in.b] : pr[in.arg & 0xFF];
    } else if constexpr (in.op == VMAS_OP_STORE_F32) {
      static_cast<float*>(e.buffers[in.b])[env] = pr[in.a];
    } else if constexpr (in.op == VMAS_OP_STORE_BOOL) {
      static_cast<uint8_t*>(e.buffers[in.b])[env] = pr[in.a] != 0.f ? 1 : 0;
    }
  });
  // the observation rows (ref scenarios/balance.py:236-262 as an ObservationPlan): [rows, B, width]
  if constexpr (P::OBS_ROWS > 0) {
    constexpr int F = P::OBS_WIDTH;
    constexpr int VEC = F % 4 == 0 ? 4 : 1;
    static_for<P::OBS_ROWS>([&](auto ri) {
      constexpr int row = decltype(ri)::value;
      float* dst = e.obs_out + ((size_t)row * a.batch_dim + env) * F;
      static_for<F / VEC>([&](auto gi) {
        constexpr int c0 = decltype(gi)::value * VEC;
        float v[VEC];
        bool all = true;
        static_for<VEC>([&](auto ki) {
          constexpr int k = decltype(ki)::value;
          constexpr ObsColC col = P::obs[row * F + c0 + k];
          v[k] = 0.f;
          if constexpr (col.op == VMAS_OBS_REG) {
            v[k] = pr[col.src];  // a value the step program (above, same thread) computed
          } else if constexpr (col.op != VMAS_OBS_SKIP) {
            v[k] = epi_state<W, (col.src >> 24) & 3, col.src & 0xFFFFFF>(r, a, env);
            if constexpr (col.op == VMAS_OBS_DIFF)
              v[k] = v[k] - epi_state<W, (col.src2 >> 24) & 3, col.src2 & 0xFFFFFF>(r, a, env);
            if constexpr (col.op == VMAS_OBS_REMAINDER) v[k] = obs_remainder(v[k], col.par);
          } else {
            all = false;
          }
        });
        if constexpr (VEC == 4) {
          if (all) {
            *reinterpret_cast<float4*>(dst + c0) = make_float4(v[0], v[1], v[2], v[3]);
            return;
          }
        }
        static_for<VEC>([&](auto ki) {
          constexpr int k = decltype(ki)::value;
          if constexpr (P::obs[row * F + c0 + k].op != VMAS_OBS_SKIP) dst[c0 + k] = v[k];
        });
      });
    });
  }
}

// One env, all of `a.n_substeps` substeps, state in the calling thread's registers.  `P` (not void): the
// step's epilogue runs behind the last substep (the whole-step kernel).
template <class W, bool TRACK = false, class P = void>
DEVI void spec_env_step(const SpecArgs& a, const long env, const uint32_t (&mask_words)[W::MASK_WORDS > 0 ? W::MASK_WORDS : 1],
                        const EpiArgs* epi = nullptr) {
  constexpr int E = W::E, NA = W::A, NI = W::NI;
  if (env >= a.batch_dim) return;
  SpecRows<W> rows;
  rows.load_pos_rot(a, env);
  rows.load_rest(a, env);
  if constexpr (!std::is_void_v<P>) {
    if (a.first_substep + a.n_substeps == W::cfg.substeps) spec_epilogue_prefetch<P>(*epi, env);
  }
  EnvRegs<E> r;
  float afx[NA > 0 ? NA : 1], afy[NA > 0 ? NA : 1], atq[NA > 0 ? NA : 1];
  rows.unpack_pos_rot(r);
  rows.unpack_rest(r, afx, afy, atq);
  uint32_t sig = 0;
  for (int sub = a.first_substep; sub < a.first_substep + a.n_substeps; ++sub) {
    spec_trig<W>(r);
    spec_entity_forces<W>(r, afx, afy, atq);
    // joints and contacts, in accumulation order
    static_for<NI>([&](auto ii) { spec_item<W, decltype(ii)::value, TRACK>(r, a, env, mask_words, sig); });
    spec_integrate<W>(r, sub);
  }
  rows.store(a, env, r, afx, afy, atq);
  if constexpr (!std::is_void_v<P>) {
    if (a.first_substep + a.n_substeps == W::cfg.substeps) spec_epilogue<W, P>(r, a, *epi, env);
  }
  if constexpr (TRACK) {
    if (a.sig) a.sig[env] = (a.first_substep == 0 ? 0u : a.sig[env]) | sig;  // OR over the substeps of a step
  }
}

#ifdef __CUDACC__
// SCHED: the env-scheduling variant (thread t steps env order[t], signatures recorded); the default
// kernel carries none of that code
template <class W, bool SCHED = false>
__global__ void __launch_bounds__(W::BLOCK, W::MIN_BLOCKS) step_spec_kernel(const SpecArgs a) {
  constexpr int MW = W::MASK_WORDS;
  const long tid = (long)blockIdx.x * W::BLOCK + threadIdx.x;

  // the block's copy of the broad-phase mask (ref core.py:2797-2801); the last block to have copied
  // it clears it for the next substep (no memset node: CUDA-graph safe)
  uint32_t mask_words[MW > 0 ? MW : 1];
  if constexpr (MW > 0) {
    __shared__ uint32_t s_mask[MW];
    if (a.use_mask) {
      for (int w = threadIdx.x; w < MW; w += W::BLOCK) s_mask[w] = a.mask[w];
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        unsigned done = atomicAdd(&a.mask[MW], 1u);
        if (done == gridDim.x - 1) {
          for (int w = 0; w < MW; ++w) a.mask[w] = 0u;
          a.mask[MW] = 0u;
        }
      }
#pragma unroll
      for (int w = 0; w < MW; ++w) mask_words[w] = s_mask[w];
    }
  }
  if (tid >= a.batch_dim) return;
  if constexpr (SCHED) {
    // env scheduling: with an order table, neighbouring threads step envs with the same contact pattern
    const long env = a.order ? (long)a.order[tid] : tid;
    spec_env_step<W, true>(a, env, mask_words);
  } else {
    spec_env_step<W, false>(a, tid, mask_words);
  }
}

// The whole-step kernel: step_spec_kernel with the epilogue P behind the last substep.
template <class W, class P>
__global__ void __launch_bounds__(W::BLOCK, W::MIN_BLOCKS) step_fused_kernel(const SpecArgs a, const EpiArgs e) {
  constexpr int MW = W::MASK_WORDS;
  const long tid = (long)blockIdx.x * W::BLOCK + threadIdx.x;
  uint32_t mask_words[MW > 0 ? MW : 1];
  if constexpr (MW > 0) {
    __shared__ uint32_t s_mask[MW];
    if (a.use_mask) {
      for (int w = threadIdx.x; w < MW; w += W::BLOCK) s_mask[w] = a.mask[w];
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        unsigned done = atomicAdd(&a.mask[MW], 1u);
        if (done == gridDim.x - 1) {
          for (int w = 0; w < MW; ++w) a.mask[w] = 0u;
          a.mask[MW] = 0u;
        }
      }
#pragma unroll
      for (int w = 0; w < MW; ++w) mask_words[w] = s_mask[w];
    }
  }
  if (tid >= a.batch_dim) return;
  spec_env_step<W, false, P>(a, tid, mask_words, &e);
}

// ---- the whole Environment.step as ONE kernel -----------------------------------------------------------
// step_fused_kernel plus what the ingest launch in front of it does (vmas_b200_ingest_actions_broad_phase):
// every thread decodes its env's actions (continuous, holonomic: ref environment.py:616-655, 707 and
// dynamics/holonomic.py:14-15) straight into the force registers, counts the step, and tests its env's
// masked pairs for the batch-wide broad phase (ref core.py:2797-2801); the mask is complete once every block
// has contributed — a grid-wide barrier, which needs all blocks resident (cooperative launch; the launcher
// says no for batches beyond that and the caller keeps the separate ingest launch).
//   a.mask: per substep [MASK_WORDS] bits, [MASK_WORDS] arrivals at the barrier, [MASK_WORDS + 1] blocks done
//   with the mask (substeps x (MASK_WORDS + 2) words, zero between launches)
template <class W>
__host__ __device__ constexpr int spec_first_masked() {
  for (int i = 0; i < W::NI; ++i)
    if (W::item[i].mask_bit >= 0) return i;
  return W::NI;
}

DEVI unsigned ld_acquire_u32(const uint32_t* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <class W, class P>
__global__ void __launch_bounds__(W::BLOCK, W::MIN_BLOCKS) step_env_kernel(const SpecArgs a, const EpiArgs e, const ActArgs act) {
  constexpr int MW = W::MASK_WORDS, E = W::E, NA = W::A, NI = W::NI;
  const long tid = (long)blockIdx.x * W::BLOCK + threadIdx.x;
  const bool live = tid < a.batch_dim;
  const long env = live ? tid : (long)a.batch_dim - 1;  // (the tail threads shadow the last env and store nothing)
  SpecRows<W, true> rows;
  rows.load_pos_rot(a, env);
  rows.load_rest(a, env);
  spec_epilogue_prefetch<P>(e, env);
#ifdef __CUDA_ARCH__
  if (act.steps) spec_prefetch(act.steps + env);
#endif
  EnvRegs<E> r;
  float afx[NA > 0 ? NA : 1], afy[NA > 0 ? NA : 1], atq[NA > 0 ? NA : 1];
  rows.unpack_pos_rot(r);
  rows.unpack_rest(r, afx, afy, atq);
  // the actions
  bool bad = false;
  static_for<P::N_ACT>([&](auto ki) {
    constexpr int k = decltype(ki)::value;
    constexpr ActC ac = P::act[k];
    float2 v = reinterpret_cast<const float2*>(act.actions[k])[env];
    if (act.clamp) {  // torch.clamp keeps NaN
      v.x = fminf(fmaxf(v.x, -ac.range0), ac.range0);
      v.y = fminf(fmaxf(v.y, -ac.range1), ac.range1);
    }
    bad |= (v.x != v.x) || (fabsf(v.x) > ac.range0) || (v.y != v.y) || (fabsf(v.y) > ac.range1);
    const float2 u = make_float2(v.x * ac.mult0, v.y * ac.mult1);
    if (live) reinterpret_cast<float2*>(act.u[k])[env] = u;
    afx[ac.agent] = u.x;
    afy[ac.agent] = u.y;
  });
  if (live) {
    if (bad && act.bad_flag) *act.bad_flag = 1;
    if (act.steps) act.steps[env] = act.steps[env] + 1.f;
  }
  // The batch-wide broad phase of every substep, inside the kernel.  ARRIVE: the block's pairs-in-range bits go
  // to the global mask of the substep and the block checks in at its barrier; WAIT only where the first masked
  // work item is due — the trigonometry, the per-entity forces and the unmasked items in front of it (sphere
  // pairs) run while the other blocks arrive.  Substep s uses a.mask + s * (MW + 2): [MW] bits, arrivals, blocks
  // done with the mask (the last of them clears the region for the next step).
  uint32_t mask_words[MW > 0 ? MW : 1];
  [[maybe_unused]] __shared__ uint32_t s_mask[MW > 0 ? MW : 1];
  static_assert(MW <= W::BLOCK, "more mask words than threads in a block");
  uint32_t sig = 0;
  for (int sub = 0; sub < W::cfg.substeps; ++sub) {
    [[maybe_unused]] uint32_t* gmask = a.mask + sub * (MW + 2);
    if constexpr (MW > 0) {
      if (a.use_mask) {
        if (sub > 0) __syncthreads();  // (every thread has taken the previous substep's mask out of s_mask)
        if (threadIdx.x < MW) s_mask[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t bits[MW];
#pragma unroll
        for (int w = 0; w < MW; ++w) bits[w] = 0u;
        static_for<NI>([&](auto ii) {
          constexpr ItemC it = W::item[decltype(ii)::value];
          if constexpr (it.mask_bit >= 0) {
            const bool near = live && norm2(r.px[it.a] - r.px[it.b], r.py[it.a] - r.py[it.b]) <= it.broad_thr;
            if (__any_sync(0xffffffffu, near)) bits[it.mask_bit >> 5] |= 1u << (it.mask_bit & 31);
          }
        });
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
          for (int w = 0; w < MW; ++w)
            if (bits[w]) atomicOr(&s_mask[w], bits[w]);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          for (int w = 0; w < MW; ++w) {  // (bits only ever get set: a stale read costs one redundant atomic)
            const uint32_t b = s_mask[w];
            if (b && (ld_acquire_u32(&gmask[w]) & b) != b) atomicOr(&gmask[w], b);
          }
          __threadfence();
          atomicAdd(&gmask[MW], 1u);
        }
      }
    }
    spec_trig<W>(r);
    spec_entity_forces<W>(r, afx, afy, atq);
    static_for<NI>([&](auto ii) {
      constexpr int I = decltype(ii)::value;
      if constexpr (MW > 0 && I == spec_first_masked<W>()) {
        if (a.use_mask) {  // WAIT
          if (threadIdx.x == 0) {
            while (ld_acquire_u32(&gmask[MW]) < gridDim.x) __nanosleep(20);
            for (int w = 0; w < MW; ++w) s_mask[w] = ld_acquire_u32(&gmask[w]);
            __threadfence();
            const unsigned done = atomicAdd(&gmask[MW + 1], 1u);
            if (done == gridDim.x - 1) {  // every block has its copy: clear for the next step
              for (int w = 0; w < MW + 2; ++w) gmask[w] = 0u;
            }
          }
          __syncthreads();
#pragma unroll
          for (int w = 0; w < MW; ++w) mask_words[w] = s_mask[w];
        }
      }
      spec_item<W, I, false>(r, a, env, mask_words, sig);
    });
    spec_integrate<W>(r, sub);
  }
  if (!live) return;
  rows.store(a, env, r, afx, afy, atq);
  spec_epilogue<W, P>(r, a, e, env);
}

// cudaErrorCooperativeLaunchTooLarge: the batch does not fit the GPU at once (masked worlds only)
template <class W, class P>
static cudaError_t launch_env(const SpecArgs& a, const EpiArgs& e, const ActArgs& act, cudaStream_t stream) {
  const long blocks = ((long)a.batch_dim + W::BLOCK - 1) / W::BLOCK;
  if (W::MASK_WORDS > 0 && a.use_mask) {
    static long capacity[64] = {0};
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err != cudaSuccess) return err;
    long cap = device < 64 ? capacity[device] : 0;
    if (cap == 0) {
      int per_sm = 0, sms = 0;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_env_kernel<W, P>, W::BLOCK, 0);
      if (err != cudaSuccess) return err;
      err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      if (err != cudaSuccess) return err;
      cap = (long)per_sm * sms;
      if (device < 64) capacity[device] = cap;
    }
    if (blocks > cap) return cudaErrorCooperativeLaunchTooLarge;
    void* args[] = {const_cast<SpecArgs*>(&a), const_cast<EpiArgs*>(&e), const_cast<ActArgs*>(&act)};
    return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(step_env_kernel<W, P>), dim3((unsigned)blocks),
                                       dim3(W::BLOCK), args, 0, stream);
  }
  step_env_kernel<W, P><<<(unsigned)blocks, W::BLOCK, 0, stream>>>(a, e, act);
  return cudaGetLastError();
}

template <class W, class P>
static cudaError_t launch_fused(const SpecArgs& a, const EpiArgs& e, cudaStream_t stream) {
  const long blocks = ((long)a.batch_dim + W::BLOCK - 1) / W::BLOCK;
  step_fused_kernel<W, P><<<(unsigned)blocks, W::BLOCK, 0, stream>>>(a, e);
  return cudaGetLastError();
}

// host-side launcher used by the registry in generated/specializations.cuh
template <class W>
static cudaError_t launch_spec(const SpecArgs& a, cudaStream_t stream) {
  const long blocks = ((long)a.batch_dim + W::BLOCK - 1) / W::BLOCK;
  if (a.order || a.sig)
    step_spec_kernel<W, true><<<(unsigned)blocks, W::BLOCK, 0, stream>>>(a);
  else
    step_spec_kernel<W, false><<<(unsigned)blocks, W::BLOCK, 0, stream>>>(a);
  return cudaGetLastError();
}
#endif  // __CUDACC__

#ifdef __CUDACC__
struct SpecEntry {
  uint64_t hash;
  const char* name;
  int n_entities, n_items;
  cudaError_t (*launch)(const SpecArgs&, cudaStream_t);       // one thread per env (step_spec_kernel)
  cudaError_t (*launch_tile)(const SpecArgs&, cudaStream_t);  // a warp owns a tile of 32 envs, compacted narrow phase
  bool has_tile;                                              // (step_tile_kernel; not for worlds with joints)
};
#endif

}  // namespace vmas
