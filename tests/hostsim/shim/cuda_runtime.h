/* Host stand-in for <cuda_runtime.h> — TEST INFRASTRUCTURE (tests/hostsim), never shipped.
 * Lets g++ compile the device functions of csrc/geometry.cuh, spec_kernel.cuh and
 * spec_coop_kernel.cuh so their control flow and indexing can be exercised on the CPU.
 * Arithmetic: IEEE fp32 with contraction disabled (-ffp-contract=off), like nvcc -fmad=false;
 * libm's sincosf / expf / log1pf are not bit-equal to CUDA's, so CPU runs are compared with each
 * other bit for bit and with the golden vectors to a tolerance. */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
