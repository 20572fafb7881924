// reset.cuh — device-side episode reset: masked state zeroing and the batched
// rejection-sampling respawn of entities.  Included by vmas_b200.cu (uses its fail / CUDA_OK).
//
// The reference resets with a python `while True` per entity (utils.py:241-319: two uniform_
// draws, torch.cdist against the occupied positions, torch.any -> one host sync per iteration)
// and zeroes the state with one indexed assignment per entity and field (core.py:286-296,
// 1179-1181).  Here one thread owns one env: it places the requested entities one after the
// other, redrawing a position until it keeps `min_dist` from everything already placed — the
// same per-env procedure (the reference only re-draws the envs that still overlap), without
// host round trips, for all envs, one env, or the envs flagged in a device mask.
//
// Random numbers: Philox4x32-10 (Salmon et al., SC'11 — the generator family torch uses on
// CUDA), counter-based: counter = (env (+ the shard's offset in a multi-GPU job), episode number of
// that env, spawn call number, draw slot << 26 | attempt block), key = seed.  A position therefore depends only on (seed, env, how often that
// env has been reset, which spawn call of the reset, which entity, which attempt) and not on
// which other envs are reset in the same launch: a masked reset of many envs equals resetting
// them one at a time, bit for bit.  oracle/reset.py restates the same procedure in numpy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vmas_b200.h"

namespace vmas {

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  Philox4 r;
  r.x = c0; r.y = c1; r.z = c2; r.w = c3;
  return r;
}

// The sampler's arithmetic is written with the explicitly rounded intrinsics: the result (and with it
// every accept / reject decision) must not depend on the build's contraction or division / square-root
// flags — oracle/reset.py reproduces it bit for bit, also for the VMAS_B200_ARITH=fast library.

// 24 random bits -> [0, 1), then lo + u * span with separate roundings
__device__ __forceinline__ float uniform_in(uint32_t bits, float lo, float span) {
  const float u = __fmul_rn((float)(bits >> 8), 5.9604644775390625e-8f);  // * 2^-24, exact
  return __fadd_rn(lo, __fmul_rn(u, span));
}

// |p - q| < min_dist with sqrt(dx*dx + dy*dy), every operation rounded on its own (IEEE square root)
__device__ __forceinline__ bool too_close(float2 p, float2 q, float min_dist) {
  const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) < min_dist;
}

__device__ __forceinline__ bool env_selected(long env, int32_t env_index, const uint8_t* env_mask) {
  if (env_index >= 0) return env == env_index;
  return env_mask == nullptr || env_mask[env] != 0;
}

struct ResetStateArgs {
  VmasState st;
  int32_t* reset_count;
  const uint8_t* env_mask;
  int32_t env_index, batch_dim, n_entities, n_agents;
};

// thread = (env, entity): zero the entity's state in that env (ref core.py:286-296); the thread
// of entity 0 bumps the env's episode counter.
__global__ void __launch_bounds__(256) reset_state_kernel(const ResetStateArgs a) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n_sel = a.env_index >= 0 ? 1 : a.batch_dim;
  if (t >= n_sel * a.n_entities) return;
  const long env = a.env_index >= 0 ? a.env_index : t / a.n_entities;
  const int e = (int)(t % a.n_entities);
  if (!env_selected(env, a.env_index, a.env_mask)) return;
  const size_t row = (size_t)env * a.n_entities + e;
  reinterpret_cast<float2*>(a.st.pos)[row] = make_float2(0.f, 0.f);
  reinterpret_cast<float2*>(a.st.vel)[row] = make_float2(0.f, 0.f);
  a.st.rot[row] = 0.f;
  a.st.ang_vel[row] = 0.f;
  if (e < a.n_agents && a.st.force && a.st.torque) {  // agent rows [B, A]: reuse the first A threads of the env
    const size_t arow = (size_t)env * a.n_agents + e;
    reinterpret_cast<float2*>(a.st.force)[arow] = make_float2(0.f, 0.f);
    a.st.torque[arow] = 0.f;
  }
  if (e == 0 && a.reset_count) a.reset_count[env] += 1;
}

struct SpawnArgs {
  VmasSpawn sp;
  float* pos;  // slab [B, E, 2]
  int32_t batch_dim, n_entities;
};

// thread = env.  Sequential rejection sampling of sp.n_spawn positions (ref utils.py:241-319).
__global__ void __launch_bounds__(128) spawn_entities_kernel(const SpawnArgs a) {
  const VmasSpawn& sp = a.sp;
  const long env = sp.env_index >= 0 ? sp.env_index : (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= a.batch_dim || (sp.env_index >= 0 && (blockIdx.x != 0 || threadIdx.x != 0))) return;
  if (!env_selected(env, sp.env_index, sp.env_mask)) return;

  float2* row = reinterpret_cast<float2*>(a.pos) + (size_t)env * a.n_entities;
  const float2* extra =
      sp.occupied ? reinterpret_cast<const float2*>(sp.occupied + (size_t)env * sp.occupied_env_stride) : nullptr;
  const uint32_t k0 = (uint32_t)sp.seed, k1 = (uint32_t)(sp.seed >> 32);
  const uint32_t episode = sp.reset_count ? (uint32_t)sp.reset_count[env] : 0u;
  const float span_x = __fsub_rn(sp.x_hi, sp.x_lo), span_y = __fsub_rn(sp.y_hi, sp.y_lo);

  float2 placed[VMAS_MAX_SPAWN];
  bool exhausted = false;
  for (int i = 0; i < sp.n_spawn; ++i) {
    const uint32_t slot = (uint32_t)i << 26;  // n_spawn <= 64, max_tries <= 2^27
    float2 p = make_float2(0.f, 0.f);
    Philox4 r = {0u, 0u, 0u, 0u};
    for (int tries = 0;; ++tries) {
      if ((tries & 1) == 0) r = philox4x32_10((uint32_t)env + sp.env_offset, episode, sp.stream_id, slot | (uint32_t)(tries >> 1), k0, k1);
      p.x = uniform_in((tries & 1) ? r.z : r.x, sp.x_lo, span_x);
      p.y = uniform_in((tries & 1) ? r.w : r.y, sp.y_lo, span_y);
      bool ok = true;
      for (int j = 0; ok && j < sp.n_occupied_entities; ++j) ok = !too_close(p, row[sp.occupied_entity[j]], sp.min_dist);
      for (int j = 0; ok && j < sp.n_occupied; ++j) ok = !too_close(p, extra[j], sp.min_dist);
      for (int j = 0; ok && j < i; ++j) ok = !too_close(p, placed[j], sp.min_dist);
      if (ok) break;
      if (tries + 1 >= sp.max_tries) {  // keep the last proposal, tell the host
        exhausted = true;
        break;
      }
    }
    placed[i] = p;
    if (sp.entity[i] >= 0) row[sp.entity[i]] = p;
    if (sp.out) reinterpret_cast<float2*>(sp.out)[(size_t)env * sp.n_spawn + i] = p;
  }
  if (exhausted && sp.status) atomicAdd(sp.status, 1);
}

}  // namespace vmas

extern "C" {

int vmas_b200_reset_state(const VmasWorldConfig* cfg, const VmasState* st, int32_t env_index,
                          const uint8_t* env_mask, int32_t* reset_count, void* cuda_stream) {
  using namespace vmas;
  if (!cfg || !st || !st->pos || !st->vel || !st->rot || !st->ang_vel) return fail("null argument%s");
  if (cfg->batch_dim <= 0 || cfg->n_entities <= 0) return fail("empty world%s");
  if (env_index >= cfg->batch_dim) return fail("env_index out of range%s");
  if (cfg->n_agents > cfg->n_entities) return fail("more agents than entities%s");
  ResetStateArgs a;
  a.st = *st;
  a.reset_count = reset_count;
  a.env_mask = env_index >= 0 ? nullptr : env_mask;
  a.env_index = env_index < 0 ? -1 : env_index;
  a.batch_dim = cfg->batch_dim;
  a.n_entities = cfg->n_entities;
  a.n_agents = cfg->n_agents;
  const long n = (env_index >= 0 ? 1L : (long)cfg->batch_dim) * cfg->n_entities;
  const int threads = 256;
  reset_state_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(a);
  CUDA_OK(cudaGetLastError());
  return 1;
}

int vmas_b200_spawn_entities(const VmasWorldConfig* cfg, const VmasState* st, const VmasSpawn* spawn,
                             void* cuda_stream) {
  using namespace vmas;
  if (!cfg || !st || !st->pos || !spawn) return fail("null argument%s");
  if (cfg->batch_dim <= 0 || cfg->n_entities <= 0) return fail("empty world%s");
  if (spawn->n_spawn <= 0 || spawn->n_spawn > VMAS_MAX_SPAWN) return fail("n_spawn must be in [1, VMAS_MAX_SPAWN]%s");
  if (spawn->n_occupied_entities < 0 || spawn->n_occupied_entities > VMAS_MAX_SPAWN)
    return fail("n_occupied_entities must be in [0, VMAS_MAX_SPAWN]%s");
  if (spawn->n_occupied < 0 || (spawn->n_occupied > 0 && !spawn->occupied)) return fail("occupied points missing%s");
  if (spawn->env_index >= cfg->batch_dim) return fail("env_index out of range%s");
  if (spawn->max_tries <= 0 || spawn->max_tries > (1 << 27)) return fail("max_tries must be in [1, 2^27]%s");
  if (!(spawn->x_hi >= spawn->x_lo) || !(spawn->y_hi >= spawn->y_lo)) return fail("empty spawn bounds%s");
  bool writes = spawn->out != nullptr;
  for (int i = 0; i < spawn->n_spawn; ++i) {
    if (spawn->entity[i] >= cfg->n_entities) return fail("spawn entity out of range%s");
    writes = writes || spawn->entity[i] >= 0;
  }
  if (!writes) return fail("nothing to write: no slab entity and no `out`%s");
  for (int i = 0; i < spawn->n_occupied_entities; ++i)
    if (spawn->occupied_entity[i] < 0 || spawn->occupied_entity[i] >= cfg->n_entities)
      return fail("occupied entity out of range%s");
  SpawnArgs a;
  a.sp = *spawn;
  if (a.sp.env_index >= 0) a.sp.env_mask = nullptr;
  if (a.sp.env_index < 0) a.sp.env_index = -1;
  a.pos = st->pos;
  a.batch_dim = cfg->batch_dim;
  a.n_entities = cfg->n_entities;
  const int threads = 128;
  const long n = a.sp.env_index >= 0 ? 1L : (long)cfg->batch_dim;
  spawn_entities_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(a);
  CUDA_OK(cudaGetLastError());
  return 1;
}

}  // extern "C"
