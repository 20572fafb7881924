// hostsim.cpp — TEST INFRASTRUCTURE: runs the device functions of the specialised substep kernels on
// the CPU (g++, shim/cuda_runtime.h), one "thread" after the other, so the cooperative kernel's
// phase structure (who owns which entity / item, shared-memory rows, accumulation order) can be
// checked bit for bit against the thread-per-env formulation without a GPU.  Not product code.
#include <vector>

#include "generated/specializations.cuh"

using namespace vmas;

template <class W>
static void run_thread_per_env(const SpecArgs& a, const uint32_t* mask) {
  uint32_t mask_words[W::MASK_WORDS > 0 ? W::MASK_WORDS : 1] = {0};
  if (a.use_mask)
    for (int w = 0; w < W::MASK_WORDS; ++w) mask_words[w] = mask[w];
  for (long env = 0; env < a.batch_dim; ++env) spec_env_step<W>(a, env, mask_words);
}

// the body of step_coop_kernel with every __syncthreads() turned into "finish the loop over threads"
template <class W>
static void run_cooperative(const SpecArgs& a, const uint32_t* mask) {
  uint32_t mask_words[W::MASK_WORDS > 0 ? W::MASK_WORDS : 1] = {0};
  if (a.use_mask)
    for (int w = 0; w < W::MASK_WORDS; ++w) mask_words[w] = mask[w];
  std::vector<float> sm(CoopRows<W>::BYTES / sizeof(float));
  const long blocks = ((long)a.batch_dim + COOP_LANES - 1) / COOP_LANES;
  for (long block = 0; block < blocks; ++block) {
    // poison the tile: a row that is read before its owner wrote it must show up as NaN
    for (float& v : sm) v = NAN;
    auto for_threads = [&](auto&& fn) {
      for (int warp = 0; warp < COOP_WARPS; ++warp)
        for (int lane = 0; lane < COOP_LANES; ++lane) {
          const long env = block * COOP_LANES + lane;
          if (env < a.batch_dim) fn(warp, lane, env);
        }
    };
    for_threads([&](int warp, int lane, long env) { Coop<W>::load(sm.data(), warp, lane, env, a); });
    for (int sub = a.first_substep; sub < a.first_substep + a.n_substeps; ++sub) {
      for_threads([&](int warp, int lane, long) { Coop<W>::forces(sm.data(), warp, lane); });
      for_threads([&](int warp, int lane, long env) { Coop<W>::items(sm.data(), warp, lane, env, a, mask_words); });
      for_threads([&](int warp, int lane, long) { Coop<W>::integrate(sm.data(), warp, lane, sub); });
    }
    for_threads([&](int warp, int lane, long env) { Coop<W>::store(sm.data(), warp, lane, env, a); });
  }
}

extern "C" {

int hostsim_num_worlds(void) {
  int n = 0;
#define COUNT(i, W, h) ++n;
  VMAS_FOR_EACH_SPEC_WORLD(COUNT)
#undef COUNT
  return n;
}

// variant 0: one thread per env (spec_env_step); 1: cooperative (Coop<W> phases).  Returns 0, or -1
// if no specialised world has this hash.
int hostsim_step(uint64_t world_hash, int variant, int batch_dim, float* pos, float* vel, float* rot,
                 float* ang_vel, float* force, float* torque, const uint32_t* mask, int use_mask,
                 int first_substep, int n_substeps) {
  SpecArgs a;
  a.st.pos = pos; a.st.vel = vel; a.st.rot = rot; a.st.ang_vel = ang_vel; a.st.force = force; a.st.torque = torque;
  a.joint_rot = nullptr;
  a.mask = nullptr;
  a.batch_dim = batch_dim;
  a.use_mask = use_mask;
  a.first_substep = first_substep;
  a.n_substeps = n_substeps;
#define TRY(i, W, h)                                  \
  if (world_hash == h) {                              \
    if (variant == 0) run_thread_per_env<W>(a, mask); \
    else run_cooperative<W>(a, mask);                 \
    return 0;                                         \
  }
  VMAS_FOR_EACH_SPEC_WORLD(TRY)
#undef TRY
  return -1;
}

}  // extern "C"
