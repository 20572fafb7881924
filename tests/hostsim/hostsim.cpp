// hostsim.cpp — TEST INFRASTRUCTURE: runs the device functions of the specialised substep kernels on
// the CPU (g++, shim/cuda_runtime.h), one "thread" after the other, so the tile kernel's
// phase structure (who owns which entity / item, shared-memory rows, accumulation order) can be
// checked bit for bit against the thread-per-env formulation without a GPU.  Not product code.
#include <stdlib.h>

#include <vector>

#include "generated/specializations.cuh"

using namespace vmas;

template <class W>
static void run_thread_per_env(const SpecArgs& a, const uint32_t* mask) {
  uint32_t mask_words[W::MASK_WORDS > 0 ? W::MASK_WORDS : 1] = {0};
  if (a.use_mask)
    for (int w = 0; w < W::MASK_WORDS; ++w) mask_words[w] = mask[w];
  if (a.sig)
    for (long env = 0; env < a.batch_dim; ++env) spec_env_step<W, true>(a, env, mask_words);
  else
    for (long env = 0; env < a.batch_dim; ++env) spec_env_step<W, false>(a, env, mask_words);
}

// the body of tile_warp_step (csrc/spec_tile_kernel.cuh) for one tile of 32 envs: the lanes of the warp one
// after the other, every __syncwarp() turned into "finish the loop over the lanes"; the ballot-based
// appends (contact ring, geometry queues) are restated as plain loops in lane order, with the same
// ring discipline (a full round of the force as soon as 32 contacts wait)
template <class W>
static void run_tile(const SpecArgs& a, const uint32_t* mask) {
  using T = Tile<W>;
  using L = TileLayout<W>;
  if constexpr (L::SUPPORTED) {
    uint32_t mask_words[W::MASK_WORDS > 0 ? W::MASK_WORDS : 1] = {0};
    if (a.use_mask)
      for (int w = 0; w < W::MASK_WORDS; ++w) mask_words[w] = mask[w];
    std::vector<float> sm(L::FLOATS);
    std::vector<typename T::Lane> lanes(TILE_LANES);
    const long tiles = ((long)a.batch_dim + TILE_LANES - 1) / TILE_LANES;
    for (long tile = 0; tile < tiles; ++tile) {
      // poison the tile: a row that is read before its owner wrote it must show up as NaN
      for (float& v : sm) v = NAN;
      auto env_of = [&](int lane) { return tile * TILE_LANES + lane; };
      auto valid = [&](int lane) { return env_of(lane) < a.batch_dim; };
      auto env_c = [&](int lane) { return valid(lane) ? env_of(lane) : (long)a.batch_dim - 1; };
      for (int lane = 0; lane < TILE_LANES; ++lane) {
        lanes[lane].rows.load_pos_rot(a, env_c(lane));
        lanes[lane].rows.unpack_pos_rot(lanes[lane].r);
      }
      uint16_t* q = T::queue(sm.data());
      for (int sub = a.first_substep; sub < a.first_substep + a.n_substeps; ++sub) {
        uint64_t any = 0;
        for (int lane = 0; lane < TILE_LANES; ++lane) {
          T::p1(sm.data(), lane, lanes[lane], a, mask_words);
          if (!valid(lane)) lanes[lane].near = lanes[lane].hits = 0;
          any |= lanes[lane].near;
        }
        int head = 0, tail = 0;
        auto flush_full_rounds = [&]() {
          while (tail - head >= TILE_LANES) {
            for (int l = 0; l < TILE_LANES; ++l) T::contact_force(sm.data(), head + l);
            head += TILE_LANES;
          }
        };
        // direct kinds
        static_for<W::NI>([&](auto ii) {
          constexpr int I = decltype(ii)::value;
          if constexpr (tile_kind_is_direct(W::item[I].kind)) {
            if ((any >> I) & 1u) {
              int appended = 0;
              for (int lane = 0; lane < TILE_LANES; ++lane) {
                TileContact c;
                bool has = false;
                if ((lanes[lane].near >> I) & 1u) has = T::template direct_contact<I>(lane, lanes[lane].r, c);
                if (has) {
                  lanes[lane].hits |= (uint64_t)1 << I;
                  if (tail + appended - head >= TILE_RING) abort();  // the ring can never overflow
                  T::ring_put(sm.data(), tail + appended++, c);
                }
              }
              tail += appended;
              flush_full_rounds();
            }
          }
        });
        // geometry kinds
        if constexpr (L::HAS_GEOM) {
          int cnt[TILE_N_KINDS] = {0};
          for (int i = 0; i < W::NI; ++i) {
            const int k = W::item[i].kind;
            if (!tile_kind_is_geom(k)) continue;
            for (int lane = 0; lane < TILE_LANES; ++lane)
              if ((lanes[lane].near >> i) & 1u) q[L::geom_base(k) + cnt[k]++] = (uint16_t)((i << 5) | lane);
          }
          static_for<TILE_N_KINDS>([&](auto ki) {
            constexpr int K = decltype(ki)::value;
            if constexpr (tile_kind_is_geom(K) && L::kind_count(K) > 0) {
              for (int base = 0; base < cnt[K]; base += TILE_LANES) {
                int appended = 0;
                for (int l = 0; l < TILE_LANES && base + l < cnt[K]; ++l) {
                  TileContact c;
                  if (T::template geom_contact<K>(sm.data(), q[L::geom_base(K) + base + l], c)) {
                    if (tail + appended - head >= TILE_RING) abort();
                    T::ring_put(sm.data(), tail + appended++, c);
                  }
                }
                tail += appended;
                flush_full_rounds();
              }
            }
          });
        }
        for (int slot = head; slot < tail; ++slot) T::contact_force(sm.data(), slot);  // the last, partial round
        for (int lane = 0; lane < TILE_LANES; ++lane) {
          if (sub == a.first_substep) {
            lanes[lane].rows.load_rest(a, env_c(lane));
            lanes[lane].rows.unpack_rest(lanes[lane].r, lanes[lane].afx, lanes[lane].afy, lanes[lane].atq);
          }
          T::p3(sm.data(), lane, lanes[lane], sub);
        }
      }
      for (int lane = 0; lane < TILE_LANES; ++lane)
        if (valid(lane)) lanes[lane].rows.store(a, env_of(lane), lanes[lane].r, lanes[lane].afx, lanes[lane].afy, lanes[lane].atq);
    }
  }
}

template <class W>
static int tile_supported() {
  return TileLayout<W>::SUPPORTED ? 1 : 0;
}

extern "C" {

int hostsim_num_worlds(void) {
  int n = 0;
#define COUNT(i, W, h) ++n;
  VMAS_FOR_EACH_SPEC_WORLD(COUNT)
#undef COUNT
  return n;
}

// variant 0: one thread per env (spec_env_step); 1: warp tile (Tile<W> phases).  Returns 0, -1 if no
// specialised world has this hash, -2 if the world has no tile kernel.
static uint32_t* g_sig_out = nullptr;
// the next hostsim_step (variant 0) also records every env's contact signature into `sig` (uint32[B])
void hostsim_record_signatures(uint32_t* sig) { g_sig_out = sig; }

int hostsim_step(uint64_t world_hash, int variant, int batch_dim, float* pos, float* vel, float* rot,
                 float* ang_vel, float* force, float* torque, const uint32_t* mask, int use_mask,
                 int first_substep, int n_substeps) {
  SpecArgs a;
  a.st.pos = pos; a.st.vel = vel; a.st.rot = rot; a.st.ang_vel = ang_vel; a.st.force = force; a.st.torque = torque;
  a.joint_rot = nullptr;
  a.order = nullptr;
  a.sig = g_sig_out;
  g_sig_out = nullptr;
  a.mask = nullptr;
  a.batch_dim = batch_dim;
  a.use_mask = use_mask;
  a.first_substep = first_substep;
  a.n_substeps = n_substeps;
#define TRY(i, W, h)                                  \
  if (world_hash == h) {                              \
    if (variant == 0) run_thread_per_env<W>(a, mask); \
    else if (!tile_supported<W>()) return -2;         \
    else run_tile<W>(a, mask);                        \
    return 0;                                         \
  }
  VMAS_FOR_EACH_SPEC_WORLD(TRY)
#undef TRY
  return -1;
}

}  // extern "C"
