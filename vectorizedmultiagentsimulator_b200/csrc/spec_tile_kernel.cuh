// spec_tile_kernel.cuh — warp-tile variant of the world-specialised substep kernel: the narrow
// phase runs COMPACTED.
//
// step_spec_kernel gives one thread one env and walks the env's work items in order.  The cheap
// part of an item (is the pair anywhere near contact?) is uniform, but the expensive part (closest
// points of boxes / segments, the soft-plus contact force with its IEEE divisions, expf, log1pf) is
// needed by a different subset of items in every env: a warp executes it once per item that is near
// in ANY of its 32 envs, with ~8 of 32 lanes active (ncu, profiles/r2a_*: 60 % of the balance
// kernel's warp-instructions run at 5-13 lanes).
//
// Here a warp owns a tile of 32 envs and splits the substep in three phases:
//   P1  lane = env     load positions / rotations, sin / cos, the far test of every item (unrolled,
//                      constexpr parameters, as in step_spec_kernel); the env's positions and trig
//                      go to shared memory as [row][lane]; every NEAR (item, env) pair is appended
//                      to the queue of the item's kind (warp ballot + prefix popcount; the queues
//                      live in shared memory, entry = item << 5 | lane)
//   P2  lane = entry   the queues are drained kind by kind, 32 entries at a time: closest points +
//                      contact force of one (item, env) pair per lane, all lanes busy with the same
//                      code; parameters of the (now dynamic) item come from a constant-memory table
//                      derived from the constexpr world; results (fx, fy[, ta][, tb]) go to the
//                      item's result rows in shared memory, column = env
//   P3  lane = env     velocities / forces are loaded, per-entity forces, then the items' results
//                      are added IN ITEM ORDER (only near items: a far item contributes an exact
//                      zero, and the accumulators are never -0, so skipping it changes no bit),
//                      integration, write-back
// Only __syncwarp() separates the phases: warps are independent.
//
// Same device functions (geometry.cuh), same operand values, same per-entity summation order as
// step_spec_kernel => bit-identical results (CPU: tests/hostsim, GPU: tests/test_cabi_gpu.py).
// Worlds with joints are not tiled (they keep step_spec_kernel).
#pragma once
#include "spec_kernel.cuh"

namespace vmas {

constexpr int TILE_LANES = 32;
#ifndef TILE_MAX_WARPS
#define TILE_MAX_WARPS 4
#endif
// resident warps per SM the register allocator must leave room for (65536 / (32 * TILE_MIN_WARPS)
// registers per thread at most)
#ifndef TILE_MIN_WARPS
#define TILE_MIN_WARPS 16
#endif
constexpr int TILE_N_KINDS = 7;  // VMAS_K_JOINT .. VMAS_K_BB

// dynamic view of one work item for P2 (constant memory)
struct TileItem {
  int a, b;            // entities
  int trig_a, trig_b;  // first trig row (c, s, c2, s2) of the entity, or -1
  int row;             // first result row: fx, fy, then ta (if has_ta), tb (if has_tb)
  int has_ta, has_tb;
  int solid_a, solid_b;
  float dmin, a_h0, a_h1, b_h0, b_h1, circ_a, circ_b;
};

template <class W>
struct TileLayout {
  static constexpr int E = W::E, NI = W::NI;
  __host__ __device__ static constexpr bool is_trig(int e) { return (W::ent[e].flags & VMAS_F_TRIG) != 0; }
  __host__ __device__ static constexpr int trig_slot(int e) {
    int n = 0;
    for (int k = 0; k < e; ++k)
      if (is_trig(k)) ++n;
    return n;
  }
  static constexpr int N_TRIG = trig_slot(E);
  static constexpr int ROW_TRIG = 2 * E;                 // rows 0 .. 2E-1: px, py of every entity
  static constexpr int ROWS_STATE = 2 * E + 4 * N_TRIG;  // then c, s, c2, s2 of every trig entity
  __host__ __device__ static constexpr int trig_row(int e) { return is_trig(e) ? ROW_TRIG + 4 * trig_slot(e) : -1; }
  // torque rows exist only where the item kind can produce a torque: sphere-sphere items produce
  // none, and the sphere side (b) of line-sphere / box-sphere items none either.  spec_item adds a
  // literal +0 there, which changes no bit (the accumulators are never -0: they start at +0).
  __host__ __device__ static constexpr bool has_ta(int i) {
    return (W::ent[W::item[i].a].flags & VMAS_F_ROTATABLE) != 0 && W::item[i].kind != VMAS_K_SS;
  }
  __host__ __device__ static constexpr bool has_tb(int i) {
    const int k = W::item[i].kind;
    return (W::ent[W::item[i].b].flags & VMAS_F_ROTATABLE) != 0 && (k == VMAS_K_LL || k == VMAS_K_BL || k == VMAS_K_BB);
  }
  __host__ __device__ static constexpr int res_width(int i) { return 2 + (has_ta(i) ? 1 : 0) + (has_tb(i) ? 1 : 0); }
  __host__ __device__ static constexpr int res_row(int i) {
    int n = 0;
    for (int k = 0; k < i; ++k) n += res_width(k);
    return n;
  }
  static constexpr int ROW_RES = ROWS_STATE;
  static constexpr int ROWS_RES = res_row(NI);
  static constexpr int ROWS = ROWS_STATE + ROWS_RES;
  // queues: one per kind, capacity = (items of the kind) * 32 entries of 16 bits
  __host__ __device__ static constexpr int kind_count(int k) {
    int n = 0;
    for (int i = 0; i < NI; ++i)
      if (W::item[i].kind == k) ++n;
    return n;
  }
  __host__ __device__ static constexpr int kind_base(int k) {  // in entries
    int n = 0;
    for (int q = 0; q < k; ++q) n += kind_count(q) * TILE_LANES;
    return n;
  }
  static constexpr int QUEUE_ENTRIES = NI * TILE_LANES;
  static constexpr int FLOATS = ROWS * TILE_LANES + (QUEUE_ENTRIES + 1) / 2;  // per warp
  static constexpr size_t BYTES = (size_t)FLOATS * sizeof(float);
  // warps per block: as many as fit in the static shared-memory limit (48 KB)
  static constexpr int WARPS = (4 * BYTES <= 48 * 1024 && TILE_MAX_WARPS >= 4)   ? 4
                               : (2 * BYTES <= 48 * 1024 && TILE_MAX_WARPS >= 2) ? 2
                                                                                 : 1;
  static constexpr int MIN_BLOCKS = (TILE_MIN_WARPS / WARPS) > 0 ? TILE_MIN_WARPS / WARPS : 1;
  static constexpr bool SUPPORTED = W::N_JOINTS == 0 && NI > 0 && NI <= 64 && BYTES <= 48 * 1024 && NI < 2048;
};

template <class W>
struct TileTable {
  TileItem it[W::NI > 0 ? W::NI : 1];
};

template <class W>
__host__ __device__ constexpr TileTable<W> make_tile_table() {
  using L = TileLayout<W>;
  TileTable<W> t{};
  for (int i = 0; i < W::NI; ++i) {
    const ItemC it = W::item[i];
    const EntC ea = W::ent[it.a], eb = W::ent[it.b];
    TileItem& d = t.it[i];
    d.a = it.a;
    d.b = it.b;
    d.trig_a = L::trig_row(it.a);
    d.trig_b = L::trig_row(it.b);
    d.row = L::ROW_RES + L::res_row(i);
    d.has_ta = L::has_ta(i) ? 1 : 0;
    d.has_tb = L::has_tb(i) ? 1 : 0;
    d.solid_a = (ea.flags & VMAS_F_HOLLOW) ? 0 : 1;
    d.solid_b = (eb.flags & VMAS_F_HOLLOW) ? 0 : 1;
    d.dmin = it.dmin_base;
    d.a_h0 = ea.d0 / 2.f;
    d.a_h1 = ea.d1 / 2.f;
    d.b_h0 = eb.d0 / 2.f;
    d.b_h1 = eb.d1 / 2.f;
    d.circ_a = ea.circ_r;
    d.circ_b = eb.circ_r;
  }
  return t;
}

#ifdef __CUDACC__
template <class W>
__constant__ TileTable<W> c_tile_table = make_tile_table<W>();
#define TILE_TABLE(W) c_tile_table<W>
#else
template <class W>
inline const TileTable<W> h_tile_table = make_tile_table<W>();
#define TILE_TABLE(W) h_tile_table<W>
#endif

template <class W>
struct Tile {
  using L = TileLayout<W>;
  static constexpr int E = W::E, NA = W::A, NI = W::NI;
  using NearMask = uint64_t;

  // what one lane (= env) keeps in registers between the phases
  struct Lane {
    SpecRows<W> rows;
    EnvRegs<E> r;
    float afx[NA > 0 ? NA : 1], afy[NA > 0 ? NA : 1], atq[NA > 0 ? NA : 1];
    NearMask near;
  };

  DEVI static uint16_t* queue(float* sm) { return reinterpret_cast<uint16_t*>(sm + L::ROWS * TILE_LANES); }

  // ---- P1a: far test of item I (the condition under which spec_item evaluates the narrow phase) --
  template <int I>
  DEVI static bool near_item(const EnvRegs<E>& r) {
    constexpr ItemC it = W::item[I];
    constexpr int A = it.a, B = it.b;
    constexpr EntC ea = W::ent[A], eb = W::ent[B];
    const V2 pa = mk(r.px[A], r.py[A]), pb = mk(r.px[B], r.py[B]);
    if constexpr (it.kind == VMAS_K_SS) {
      // the first early-out of constraint_force(): beyond it the force is an exact zero
      const V2 delta = pa - pb;
      const float s = __fmaf_rn(delta.y, delta.y, __fmul_rn(delta.x, delta.x));
      return !(s > it.dmin_base * it.dmin_base * 1.000002f);
    } else if constexpr (it.kind == VMAS_K_LS) {
      Seg l = spec_seg<W, A>(r);
      return !spec_far_apart(l.p, pb, l.half + it.dmin_base);
    } else if constexpr (it.kind == VMAS_K_LL) {
      Seg l1 = spec_seg<W, A>(r), l2 = spec_seg<W, B>(r);
      return !spec_far_apart(l1.p, l2.p, l1.half + l2.half + it.dmin_base);
    } else if constexpr (it.kind == VMAS_K_BS) {
      BoxG bx = spec_box<W, A>(r);
      V2 d0 = pb - bx.p;
      float lx = d0.x * bx.c + d0.y * bx.s, ly = d0.y * bx.c - d0.x * bx.s;
      return !(fabsf(lx) > bx.half_l + it.dmin_base + SPEC_FAR_MARGIN ||
               fabsf(ly) > bx.half_w + it.dmin_base + SPEC_FAR_MARGIN);
    } else if constexpr (it.kind == VMAS_K_BL) {
      BoxG bx = spec_box<W, A>(r);
      Seg l = spec_seg<W, B>(r);
      V2 d0 = l.p - bx.p;
      float lx = d0.x * bx.c + d0.y * bx.s, ly = d0.y * bx.c - d0.x * bx.s;
      float ex = l.half * fabsf(l.c * bx.c + l.s * bx.s), ey = l.half * fabsf(l.s * bx.c - l.c * bx.s);
      return !(fabsf(lx) - ex > bx.half_l + it.dmin_base + SPEC_FAR_MARGIN ||
               fabsf(ly) - ey > bx.half_w + it.dmin_base + SPEC_FAR_MARGIN);
    } else if constexpr (it.kind == VMAS_K_BB) {
      return !spec_far_apart(pa, pb, ea.circ_r + eb.circ_r + it.dmin_base);
    } else {
      return false;
    }
  }

  // ---- P1: trig, state rows -> shared memory, near mask of this env -------------------------------
  DEVI static void p1(float* sm, const int lane, Lane& ln, const SpecArgs& a, const uint32_t* mask_words) {
    EnvRegs<E>& r = ln.r;
    spec_trig<W>(r);
    static_for<E>([&](auto ei) {
      constexpr int e = decltype(ei)::value;
      sm[(2 * e) * TILE_LANES + lane] = r.px[e];
      sm[(2 * e + 1) * TILE_LANES + lane] = r.py[e];
      if constexpr (L::is_trig(e)) {
        constexpr int t = L::trig_row(e);
        sm[(t + 0) * TILE_LANES + lane] = r.c[e];
        sm[(t + 1) * TILE_LANES + lane] = r.s[e];
        if constexpr (W::ent[e].shape == VMAS_SHAPE_BOX) {
          sm[(t + 2) * TILE_LANES + lane] = r.c2[e];
          sm[(t + 3) * TILE_LANES + lane] = r.s2[e];
        }
      }
    });
    NearMask near = 0;
    static_for<NI>([&](auto ii) {
      constexpr int I = decltype(ii)::value;
      constexpr ItemC it = W::item[I];
      bool active = true;
      if constexpr (it.mask_bit >= 0) {  // batch-wide broad phase (ref core.py:2797-2801): warp-uniform
        if (a.use_mask && !((mask_words[it.mask_bit >> 5] >> (it.mask_bit & 31)) & 1u)) active = false;
      }
      if (active && near_item<I>(r)) near |= (NearMask)1 << I;
    });
    ln.near = near;
  }

  // ---- P2: the narrow phase of one queue entry (kind K is compile-time, the item is not) --------
  template <int K>
  DEVI static void narrow(float* sm, const unsigned entry) {
    constexpr CfgC cfg = W::cfg;
    const int lane = entry & 31u, i = entry >> 5;
    const TileItem& t = TILE_TABLE(W).it[i];
    auto row = [&](int rw) -> float& { return sm[rw * TILE_LANES + lane]; };
    const V2 pa = mk(row(2 * t.a), row(2 * t.a + 1)), pb = mk(row(2 * t.b), row(2 * t.b + 1));
    auto seg_a = [&]() { return mkseg(pa, row(t.trig_a), row(t.trig_a + 1), t.a_h0); };
    auto seg_b = [&]() { return mkseg(pb, row(t.trig_b), row(t.trig_b + 1), t.b_h0); };
    auto box_of = [&](V2 p, int tr, float h0, float h1) {
      BoxG b;
      b.p = p;
      b.c = row(tr);
      b.s = row(tr + 1);
      b.c2 = row(tr + 2);
      b.s2 = row(tr + 3);
      b.half_l = h0;
      b.half_w = h1;
      return b;
    };
    V2 f = mk(0.f, 0.f);
    float ta = 0.f, tb = 0.f;
    if constexpr (K == VMAS_K_SS) {
      f = constraint_force(pa, pb, t.dmin, cfg.collision_force, cfg.contact_margin, false);
    } else if constexpr (K == VMAS_K_LS) {  // a = line, b = sphere
      Seg l = seg_a();
      V2 cp = closest_point_seg(l, pb);
      V2 f_sphere = constraint_force(pb, cp, t.dmin, cfg.collision_force, cfg.contact_margin, false);
      f = neg(f_sphere);
      ta = cross2(cp - l.p, f);
    } else if constexpr (K == VMAS_K_LL) {
      Seg l1 = seg_a(), l2 = seg_b();
      Pair c = closest_seg_seg(l1, l2);
      f = constraint_force(c.a, c.b, t.dmin, cfg.collision_force, cfg.contact_margin, false);
      ta = cross2(c.a - l1.p, f);
      tb = cross2(c.b - l2.p, neg(f));
    } else if constexpr (K == VMAS_K_BS) {  // a = box, b = sphere
      BoxG bx = box_of(pa, t.trig_a, t.a_h0, t.a_h1);
      V2 cp = closest_point_box(bx, pb);
      V2 inner = cp;
      float d = 0.f;
      if (t.solid_a) inner = inner_point_box(pb, cp, bx.p, &d);
      V2 f_sphere = constraint_force(pb, inner, t.dmin + d, cfg.collision_force, cfg.contact_margin, false);
      f = neg(f_sphere);
      ta = cross2(cp - bx.p, f);
    } else if constexpr (K == VMAS_K_BL) {  // a = box, b = line
      BoxG bx = box_of(pa, t.trig_a, t.a_h0, t.a_h1);
      Seg l = seg_b();
      Pair c = closest_box_seg(bx, l);
      V2 inner = c.a;
      float d = 0.f;
      if (t.solid_a) inner = inner_point_box(c.b, c.a, bx.p, &d);
      f = constraint_force(inner, c.b, t.dmin + d, cfg.collision_force, cfg.contact_margin, false);
      ta = cross2(c.a - bx.p, f);
      tb = cross2(c.b - l.p, neg(f));
    } else if constexpr (K == VMAS_K_BB) {
      BoxG b1 = box_of(pa, t.trig_a, t.a_h0, t.a_h1), b2 = box_of(pb, t.trig_b, t.b_h0, t.b_h1);
      Pair c = closest_box_box(b1, b2);
      V2 in1 = c.a, in2 = c.b;
      float d1 = 0.f, d2 = 0.f;
      if (t.solid_a) in1 = inner_point_box(c.b, c.a, b1.p, &d1);
      if (t.solid_b) in2 = inner_point_box(c.a, c.b, b2.p, &d2);
      f = constraint_force(in1, in2, (d1 + d2) + t.dmin, cfg.collision_force, cfg.contact_margin, false);
      ta = cross2(c.a - b1.p, f);
      tb = cross2(c.b - b2.p, neg(f));
    }
    row(t.row) = f.x;
    row(t.row + 1) = f.y;
    if (t.has_ta) row(t.row + 2) = ta;
    if (t.has_tb) row(t.row + 2 + t.has_ta) = tb;
  }

  // ---- P3: per-entity forces, the near items' results in item order, integration -------------------
  DEVI static void p3(const float* sm, const int lane, Lane& ln, const int sub) {
    EnvRegs<E>& r = ln.r;
    spec_entity_forces<W>(r, ln.afx, ln.afy, ln.atq);
    const NearMask near = ln.near;
    static_for<NI>([&](auto ii) {
      constexpr int I = decltype(ii)::value;
      constexpr ItemC it = W::item[I];
      constexpr int A = it.a, B = it.b;
      constexpr EntC ea = W::ent[A], eb = W::ent[B];
      constexpr int R0 = L::ROW_RES + L::res_row(I);
      if ((near >> I) & 1u) {
        const float fx = sm[R0 * TILE_LANES + lane], fy = sm[(R0 + 1) * TILE_LANES + lane];
        if constexpr (ea.flags & VMAS_F_MOVABLE) {
          r.Fx[A] = r.Fx[A] + fx;
          r.Fy[A] = r.Fy[A] + fy;
        }
        if constexpr (L::has_ta(I)) r.T[A] = r.T[A] + sm[(R0 + 2) * TILE_LANES + lane];
        if constexpr (eb.flags & VMAS_F_MOVABLE) {
          r.Fx[B] = r.Fx[B] + (-fx);
          r.Fy[B] = r.Fy[B] + (-fy);
        }
        if constexpr (L::has_tb(I)) r.T[B] = r.T[B] + sm[(R0 + 2 + (L::has_ta(I) ? 1 : 0)) * TILE_LANES + lane];
      }
    });
    spec_integrate<W>(r, sub);
  }
};

#ifdef __CUDACC__
// One warp = one tile of 32 envs.  `sm` is the warp's private slice of shared memory.
template <class W>
DEVI void tile_warp_step(float* sm, const SpecArgs& a, const long env, const int lane,
                         const uint32_t (&mask_words)[W::MASK_WORDS > 0 ? W::MASK_WORDS : 1]) {
  using T = Tile<W>;
  using L = TileLayout<W>;
  constexpr unsigned FULL = 0xffffffffu;
  const bool valid = env < a.batch_dim;
  const long env_c = valid ? env : (long)a.batch_dim - 1;  // the tail tile: idle lanes shadow the last env
  typename T::Lane ln;
  ln.rows.load_pos_rot(a, env_c);
  ln.rows.unpack_pos_rot(ln.r);
  uint16_t* q = T::queue(sm);

  for (int sub = a.first_substep; sub < a.first_substep + a.n_substeps; ++sub) {
    T::p1(sm, lane, ln, a, mask_words);
    if (!valid) ln.near = 0;
    // ---- queues: every near (item, env) pair, item-major within its kind ----------------------------
    int cnt[TILE_N_KINDS];
#pragma unroll
    for (int k = 0; k < TILE_N_KINDS; ++k) cnt[k] = 0;
    const unsigned lt = (1u << lane) - 1u;
    static_for<W::NI>([&](auto ii) {
      constexpr int I = decltype(ii)::value;
      constexpr int K = W::item[I].kind;
      const bool near = (ln.near >> I) & 1u;
      const unsigned m = __ballot_sync(FULL, near);
      if (m) {
        if (near) q[L::kind_base(K) + cnt[K] + __popc(m & lt)] = (uint16_t)((I << 5) | lane);
        cnt[K] += __popc(m);
      }
    });
    __syncwarp();
    // ---- P2: drain the queues, one kind after the other -----------------------------------------------
    static_for<TILE_N_KINDS>([&](auto ki) {
      constexpr int K = decltype(ki)::value;
      if constexpr (L::kind_count(K) > 0 && K != VMAS_K_JOINT) {
        for (int idx = lane; idx < cnt[K]; idx += TILE_LANES) T::template narrow<K>(sm, q[L::kind_base(K) + idx]);
      }
    });
    __syncwarp();
    if (sub == a.first_substep) {
      ln.rows.load_rest(a, env_c);
      ln.rows.unpack_rest(ln.r, ln.afx, ln.afy, ln.atq);
    }
    T::p3(sm, lane, ln, sub);
    __syncwarp();  // the next substep's P1 overwrites the state rows other lanes' P2 entries read
  }
  if (valid) ln.rows.store(a, env, ln.r, ln.afx, ln.afy, ln.atq);
}

template <class W>
__global__ void __launch_bounds__(TileLayout<W>::WARPS * TILE_LANES, TileLayout<W>::MIN_BLOCKS) step_tile_kernel(const SpecArgs a) {
  using L = TileLayout<W>;
  constexpr int MW = W::MASK_WORDS;
  __shared__ float s_tile[L::WARPS * L::FLOATS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long env = ((long)blockIdx.x * L::WARPS + warp) * TILE_LANES + lane;

  uint32_t mask_words[MW > 0 ? MW : 1];
  if constexpr (MW > 0) {
    __shared__ uint32_t s_mask[MW];
    if (a.use_mask) {
      for (int w = threadIdx.x; w < MW; w += L::WARPS * TILE_LANES) s_mask[w] = a.mask[w];
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
  // whole warps beyond the batch have nothing to do (the last partial warp stays: its idle lanes
  // take part in the ballots)
  if (((long)blockIdx.x * L::WARPS + warp) * TILE_LANES >= a.batch_dim) return;
  tile_warp_step<W>(s_tile + warp * L::FLOATS, a, env, lane, mask_words);
}

template <class W>
static cudaError_t launch_tile(const SpecArgs& a, cudaStream_t stream) {
  using L = TileLayout<W>;
  if constexpr (L::SUPPORTED) {
    constexpr int per_block = L::WARPS * TILE_LANES;
    const long blocks = ((long)a.batch_dim + per_block - 1) / per_block;
    step_tile_kernel<W><<<(unsigned)blocks, per_block, 0, stream>>>(a);
    return cudaGetLastError();
  } else {
    return cudaErrorNotSupported;
  }
}
#endif  // __CUDACC__

}  // namespace vmas
