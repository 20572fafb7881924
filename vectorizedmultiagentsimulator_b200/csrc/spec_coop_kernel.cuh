// spec_coop_kernel.cuh — cooperative variant of the world-specialised substep kernel, for SMALL batches.
//
// step_spec_kernel gives one thread one env: ~3 k dependent instructions per thread.  At 32768 envs
// that is 7 warps per SM walking a long chain — latency-bound far below the memory roofline.  Here a
// block of COOP_WARPS warps shares a tile of 32 envs: the LANE is the env, the WARP is the unit of
// work — an entity (its forces, its integration) or a work item (far test + narrow phase).  A
// warp's code is still the fully specialised, unrolled code of that entity / item (the index is
// warp-uniform, so there is no divergence between the entities or items of different warps), the
// env's state lives in shared memory as [row][lane] (conflict-free), and the dependency chain per
// thread shrinks from "everything" to "one entity + a few items + one entity".
//
// Same arithmetic, same order: the items call the very spec_item() of spec_kernel.cuh on a register
// view of the two entities they touch, starting from zeroed accumulators (0 + f == f), and the
// owning warp of an entity adds the items' contributions in ascending item order — the order the
// thread-per-env kernel and the reference use (ref core.py:2191-2199).  Results are bit-identical
// (tests/test_cabi_gpu.py; on the CPU: tests/hostsim).
//
// Every phase is a function of (shared memory, warp, lane): the kernel is those functions with
// __syncthreads() between them, and tests/hostsim runs the same functions in loops on the CPU.
#pragma once
#include "spec_kernel.cuh"

namespace vmas {

constexpr int COOP_LANES = 32;  // envs per block
#ifndef COOP_WARPS
#define COOP_WARPS 8
#endif

// rows of the shared-memory tile, each COOP_LANES floats
template <class W>
struct CoopRows {
  enum { PX, PY, ROT, VX, VY, WV, C, S, C2, S2, FX, FY, TQ, AFX, AFY, ATQ, PER_ENTITY };
  enum { FAX, FAY, TA, FBX, FBY, TB, PER_ITEM };
  static constexpr int ENTITY_ROWS = PER_ENTITY * W::E;
  static constexpr int ROWS = ENTITY_ROWS + PER_ITEM * W::NI;
  static constexpr size_t BYTES = (size_t)ROWS * COOP_LANES * sizeof(float);
  DEVI static float& ent(float* sm, int field, int e, int lane) { return sm[(field * W::E + e) * COOP_LANES + lane]; }
  DEVI static float& item(float* sm, int field, int i, int lane) {
    return sm[(ENTITY_ROWS + field * W::NI + i) * COOP_LANES + lane];
  }
};

template <class W>
struct Coop {
  using R = CoopRows<W>;
  static constexpr int E = W::E, NA = W::A, NI = W::NI;

  // entity e -> the warp that owns it; item i -> the warp that evaluates it
  DEVI static bool owns_entity(int warp, int e) { return e % COOP_WARPS == warp; }
  DEVI static bool owns_item(int warp, int i) { return i % COOP_WARPS == warp; }

  // ---- global -> shared: the owner of an entity loads that entity's state of its env -------------
  DEVI static void load(float* sm, int warp, int lane, long env, const SpecArgs& a) {
    static_for<E>([&](auto ei) {
      constexpr int e = decltype(ei)::value;
      constexpr EntC en = W::ent[e];
      if (!owns_entity(warp, e)) return;
      const float2 p = reinterpret_cast<const float2*>(a.st.pos)[(size_t)env * E + e];
      R::ent(sm, R::PX, e, lane) = p.x;
      R::ent(sm, R::PY, e, lane) = p.y;
      float rot = 0.f, vx = 0.f, vy = 0.f, w = 0.f;
      if constexpr (en.flags & (VMAS_F_TRIG | VMAS_F_ROTATABLE)) rot = a.st.rot[(size_t)env * E + e];
      if constexpr (en.flags & VMAS_F_MOVABLE) {
        const float2 v = reinterpret_cast<const float2*>(a.st.vel)[(size_t)env * E + e];
        vx = v.x;
        vy = v.y;
      }
      if constexpr (en.flags & VMAS_F_ROTATABLE) w = a.st.ang_vel[(size_t)env * E + e];
      R::ent(sm, R::ROT, e, lane) = rot;
      R::ent(sm, R::VX, e, lane) = vx;
      R::ent(sm, R::VY, e, lane) = vy;
      R::ent(sm, R::WV, e, lane) = w;
      R::ent(sm, R::C, e, lane) = 0.f;
      R::ent(sm, R::S, e, lane) = 0.f;
      R::ent(sm, R::C2, e, lane) = 0.f;
      R::ent(sm, R::S2, e, lane) = 0.f;
      if constexpr (en.flags & VMAS_F_AGENT) {
        if constexpr (en.flags & VMAS_F_MOVABLE) {
          const float2 f = reinterpret_cast<const float2*>(a.st.force)[(size_t)env * NA + en.agent];
          R::ent(sm, R::AFX, e, lane) = f.x;
          R::ent(sm, R::AFY, e, lane) = f.y;
        }
        if constexpr (en.flags & VMAS_F_ROTATABLE) R::ent(sm, R::ATQ, e, lane) = a.st.torque[(size_t)env * NA + en.agent];
      }
    });
  }

  // ---- per-entity forces of one substep (ref core.py:1995-2004): same statements as spec_env_step --
  DEVI static void forces(float* sm, int warp, int lane) {
    constexpr float sub_dt = W::cfg.sub_dt;
    static_for<E>([&](auto ei) {
      constexpr int e = decltype(ei)::value;
      constexpr EntC en = W::ent[e];
      if (!owns_entity(warp, e)) return;
      if constexpr (en.flags & VMAS_F_TRIG) {
        const float rot = R::ent(sm, R::ROT, e, lane);
        float s, c;
        sincosf(rot, &s, &c);
        R::ent(sm, R::S, e, lane) = s;
        R::ent(sm, R::C, e, lane) = c;
        if constexpr (en.shape == VMAS_SHAPE_BOX) {
          float s2, c2;
          sincosf(rot + SPEC_HALF_PI_F, &s2, &c2);
          R::ent(sm, R::S2, e, lane) = s2;
          R::ent(sm, R::C2, e, lane) = c2;
        }
      }
      float Fx = 0.f, Fy = 0.f, T = 0.f;
      if constexpr (en.flags & VMAS_F_AGENT) {
        if constexpr (en.flags & VMAS_F_MOVABLE) {
          float afx = R::ent(sm, R::AFX, e, lane), afy = R::ent(sm, R::AFY, e, lane);
          if constexpr (en.flags & VMAS_F_MAX_F) {
            const float n = norm2(afx, afy);
            if (n > en.max_f) {
              afx = (afx / n) * en.max_f;
              afy = (afy / n) * en.max_f;
            }
          }
          if constexpr (en.flags & VMAS_F_F_RANGE) {
            afx = fminf(fmaxf(afx, -en.f_range), en.f_range);
            afy = fminf(fmaxf(afy, -en.f_range), en.f_range);
          }
          R::ent(sm, R::AFX, e, lane) = afx;
          R::ent(sm, R::AFY, e, lane) = afy;
          Fx = Fx + afx;
          Fy = Fy + afy;
        }
        if constexpr (en.flags & VMAS_F_ROTATABLE) {
          float atq = R::ent(sm, R::ATQ, e, lane);
          if constexpr (en.flags & VMAS_F_MAX_T) {
            const float n = sqrtf(atq * atq);
            if (n > en.max_t) atq = (atq / n) * en.max_t;
          }
          if constexpr (en.flags & VMAS_F_T_RANGE) atq = fminf(fmaxf(atq, -en.t_range), en.t_range);
          R::ent(sm, R::ATQ, e, lane) = atq;
          T = T + atq;
        }
      }
      if constexpr (en.flags & VMAS_F_LIN_FRIC) {
        const float vx = R::ent(sm, R::VX, e, lane), vy = R::ent(sm, R::VY, e, lane);
        const float speed = norm2(vx, vy);
        if (speed != 0.f) {
          const float cap = en.lin_fric * en.mass;
          Fx = Fx + (-(vx / speed)) * fminf(cap, (fabsf(vx) / sub_dt) * en.mass);
          Fy = Fy + (-(vy / speed)) * fminf(cap, (fabsf(vy) / sub_dt) * en.mass);
        }
      }
      if constexpr (en.flags & VMAS_F_ANG_FRIC) {
        const float w = R::ent(sm, R::WV, e, lane);
        const float speed = sqrtf(w * w);
        if (speed != 0.f) {
          const float cap = en.ang_fric * en.inertia;
          T = T + (-(w / speed)) * fminf(cap, (fabsf(w) / sub_dt) * en.inertia);
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
      R::ent(sm, R::FX, e, lane) = Fx;
      R::ent(sm, R::FY, e, lane) = Fy;
      R::ent(sm, R::TQ, e, lane) = T;
    });
  }

  // ---- work items: spec_item() on a register view of the two entities, from zeroed accumulators ----
  template <int EI>
  DEVI static void view(EnvRegs<E>& r, const float* sm, int lane) {
    r.px[EI] = sm[(R::PX * E + EI) * COOP_LANES + lane];
    r.py[EI] = sm[(R::PY * E + EI) * COOP_LANES + lane];
    r.rot[EI] = sm[(R::ROT * E + EI) * COOP_LANES + lane];
    r.c[EI] = sm[(R::C * E + EI) * COOP_LANES + lane];
    r.s[EI] = sm[(R::S * E + EI) * COOP_LANES + lane];
    r.c2[EI] = sm[(R::C2 * E + EI) * COOP_LANES + lane];
    r.s2[EI] = sm[(R::S2 * E + EI) * COOP_LANES + lane];
    r.Fx[EI] = 0.f;
    r.Fy[EI] = 0.f;
    r.T[EI] = 0.f;
  }

  DEVI static void items(float* sm, int warp, int lane, long env, const SpecArgs& a, const uint32_t* mask_words) {
    static_for<NI>([&](auto ii) {
      constexpr int I = decltype(ii)::value;
      constexpr ItemC it = W::item[I];
      if (!owns_item(warp, I)) return;
      EnvRegs<E> r;
      view<it.a>(r, sm, lane);
      view<it.b>(r, sm, lane);
      spec_item<W, I>(r, a, env, mask_words);
      R::item(sm, R::FAX, I, lane) = r.Fx[it.a];
      R::item(sm, R::FAY, I, lane) = r.Fy[it.a];
      R::item(sm, R::TA, I, lane) = r.T[it.a];
      R::item(sm, R::FBX, I, lane) = r.Fx[it.b];
      R::item(sm, R::FBY, I, lane) = r.Fy[it.b];
      R::item(sm, R::TB, I, lane) = r.T[it.b];
    });
  }

  // ---- ordered accumulation (ascending item order) + semi-implicit Euler (ref core.py:2862-2908) --
  DEVI static void integrate(float* sm, int warp, int lane, int sub) {
    constexpr float sub_dt = W::cfg.sub_dt;
    static_for<E>([&](auto ei) {
      constexpr int e = decltype(ei)::value;
      constexpr EntC en = W::ent[e];
      if (!owns_entity(warp, e)) return;
      if constexpr ((en.flags & (VMAS_F_MOVABLE | VMAS_F_ROTATABLE)) == 0) return;
      float Fx = R::ent(sm, R::FX, e, lane), Fy = R::ent(sm, R::FY, e, lane), T = R::ent(sm, R::TQ, e, lane);
      constexpr bool movable = (en.flags & VMAS_F_MOVABLE) != 0, rotatable = (en.flags & VMAS_F_ROTATABLE) != 0;
      static_for<NI>([&](auto ii) {
        constexpr int I = decltype(ii)::value;
        constexpr int ia = W::item[I].a, ib = W::item[I].b;
        if constexpr (ia == e) {
          if constexpr (movable) {
            Fx = Fx + R::item(sm, R::FAX, I, lane);
            Fy = Fy + R::item(sm, R::FAY, I, lane);
          }
          if constexpr (rotatable) T = T + R::item(sm, R::TA, I, lane);
        } else if constexpr (ib == e) {
          if constexpr (movable) {
            Fx = Fx + R::item(sm, R::FBX, I, lane);
            Fy = Fy + R::item(sm, R::FBY, I, lane);
          }
          if constexpr (rotatable) T = T + R::item(sm, R::TB, I, lane);
        }
      });
      if constexpr (en.flags & VMAS_F_MOVABLE) {
        float vx = R::ent(sm, R::VX, e, lane), vy = R::ent(sm, R::VY, e, lane);
        float px = R::ent(sm, R::PX, e, lane), py = R::ent(sm, R::PY, e, lane);
        if (sub == 0) {
          vx = vx * en.drag_mult;
          vy = vy * en.drag_mult;
        }
        vx = vx + div_pos(Fx, en.mass) * sub_dt;
        vy = vy + div_pos(Fy, en.mass) * sub_dt;
        if constexpr (en.flags & VMAS_F_MAX_SPEED) {
          const float n = norm2(vx, vy);
          if (n > en.max_speed) {
            vx = (vx / n) * en.max_speed;
            vy = (vy / n) * en.max_speed;
          }
        }
        if constexpr (en.flags & VMAS_F_V_RANGE) {
          vx = fminf(fmaxf(vx, -en.v_range), en.v_range);
          vy = fminf(fmaxf(vy, -en.v_range), en.v_range);
        }
        px = px + vx * sub_dt;
        py = py + vy * sub_dt;
        if constexpr (W::cfg.has_x_semidim) px = fminf(fmaxf(px, -W::cfg.x_semidim), W::cfg.x_semidim);
        if constexpr (W::cfg.has_y_semidim) py = fminf(fmaxf(py, -W::cfg.y_semidim), W::cfg.y_semidim);
        R::ent(sm, R::VX, e, lane) = vx;
        R::ent(sm, R::VY, e, lane) = vy;
        R::ent(sm, R::PX, e, lane) = px;
        R::ent(sm, R::PY, e, lane) = py;
      }
      if constexpr (en.flags & VMAS_F_ROTATABLE) {
        float w = R::ent(sm, R::WV, e, lane), rot = R::ent(sm, R::ROT, e, lane);
        if (sub == 0) w = w * en.drag_mult;
        w = w + div_pos(T, en.inertia) * sub_dt;
        rot = rot + w * sub_dt;
        R::ent(sm, R::WV, e, lane) = w;
        R::ent(sm, R::ROT, e, lane) = rot;
      }
    });
  }

  // ---- shared -> global: what spec_env_step writes back (changed columns) --------------------------
  DEVI static void store(const float* sm, int warp, int lane, long env, const SpecArgs& a) {
    static_for<E>([&](auto ei) {
      constexpr int e = decltype(ei)::value;
      constexpr EntC en = W::ent[e];
      if (!owns_entity(warp, e)) return;
      const auto row = [&](int field) { return sm[(field * E + e) * COOP_LANES + lane]; };
      if constexpr (en.flags & VMAS_F_MOVABLE) {
        reinterpret_cast<float2*>(a.st.pos)[(size_t)env * E + e] = make_float2(row(R::PX), row(R::PY));
        reinterpret_cast<float2*>(a.st.vel)[(size_t)env * E + e] = make_float2(row(R::VX), row(R::VY));
      }
      if constexpr (en.flags & VMAS_F_ROTATABLE) {
        a.st.rot[(size_t)env * E + e] = row(R::ROT);
        a.st.ang_vel[(size_t)env * E + e] = row(R::WV);
      }
      if constexpr (en.flags & VMAS_F_AGENT) {
        if constexpr ((en.flags & VMAS_F_MOVABLE) && (en.flags & (VMAS_F_MAX_F | VMAS_F_F_RANGE)))
          reinterpret_cast<float2*>(a.st.force)[(size_t)env * NA + en.agent] = make_float2(row(R::AFX), row(R::AFY));
        if constexpr ((en.flags & VMAS_F_ROTATABLE) && (en.flags & (VMAS_F_MAX_T | VMAS_F_T_RANGE)))
          a.st.torque[(size_t)env * NA + en.agent] = row(R::ATQ);
      }
    });
  }
};

#ifdef __CUDACC__
template <class W>
__global__ void __launch_bounds__(COOP_LANES* COOP_WARPS) step_coop_kernel(const SpecArgs a) {
  constexpr int MW = W::MASK_WORDS;
  extern __shared__ float coop_sm[];
  const int lane = threadIdx.x, warp = threadIdx.y;
  const long env = (long)blockIdx.x * COOP_LANES + lane;

  uint32_t mask_words[MW > 0 ? MW : 1];
  if constexpr (MW > 0) {
    __shared__ uint32_t s_mask[MW];
    if (a.use_mask) {
      const int tid = warp * COOP_LANES + lane;
      for (int w = tid; w < MW; w += COOP_LANES * COOP_WARPS) s_mask[w] = a.mask[w];
      __syncthreads();
      if (tid == 0) {  // the last block to have copied the mask clears it
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
  const bool live = env < a.batch_dim;  // a lane beyond the batch only keeps the barriers company
  if (live) Coop<W>::load(coop_sm, warp, lane, env, a);
  for (int sub = a.first_substep; sub < a.first_substep + a.n_substeps; ++sub) {
    if (live) Coop<W>::forces(coop_sm, warp, lane);
    __syncthreads();
    if (live) Coop<W>::items(coop_sm, warp, lane, env, a, mask_words);
    __syncthreads();
    if (live) Coop<W>::integrate(coop_sm, warp, lane, sub);
    // no barrier here: the next forces() only touches the rows of the warp's own entities, and
    // the barrier that follows it orders this substep's reads of the item rows before the next
    // substep's writes to them
  }
  if (live) Coop<W>::store(coop_sm, warp, lane, env, a);
}

template <class W>
static cudaError_t launch_coop(const SpecArgs& a, cudaStream_t stream) {
  constexpr size_t smem = CoopRows<W>::BYTES;
  if (smem > 48 * 1024) {  // opt in to a large dynamic tile once per device
    static bool configured[64] = {false};
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    if (device >= 64 || !configured[device]) {
      e = cudaFuncSetAttribute(step_coop_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      if (device < 64) configured[device] = true;
    }
  }
  const long blocks = ((long)a.batch_dim + COOP_LANES - 1) / COOP_LANES;
  step_coop_kernel<W><<<(unsigned)blocks, dim3(COOP_LANES, COOP_WARPS), smem, stream>>>(a);
  return cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace vmas
