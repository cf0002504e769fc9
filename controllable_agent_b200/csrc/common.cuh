// Shared device helpers for libfb_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FB_WARP 32
#define FB_FULL_MASK 0xffffffffu

__host__ __device__ __forceinline__ int fb_round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ __forceinline__ int fb_ceil_div(int x, int m) { return (x + m - 1) / m; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FB_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FB_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FB_FULL_MASK, v, o));
  return v;
}

// 128-bit streaming load through the read-only path without L1 allocation (replay rows are read once)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based so graph replays stay reproducible ----------
struct Philox {
  uint32_t key[2];
  __device__ Philox(uint64_t seed) { key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); }
  __device__ static __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // 4 x 32 random bits for counter (ctr, stream, sub)
  __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint32_t stream, uint32_t sub) const {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), stream, sub};
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c[0], c[1], c[2], c[3]);
  }
};
// (0,1] uniform from 32 bits
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// two independent N(0,1) from two 32-bit words (Box-Muller)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float r = sqrtf(-2.0f * logf(u01(a)));
  float s, c;
  sincosf(6.28318530717958647692f * u01(b), &s, &c);
  return make_float2(r * c, r * s);
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start while
// its predecessor in the stream is still running, once every CTA of the predecessor has executed fb_pdl_trigger() (or
// exited); it must call fb_pdl_wait() before touching anything the predecessor reads or writes.  Used to overlap the
// prologue of k_gemm_tc (barrier init, TMEM allocation) with the tail of the kernel before it.
__device__ __forceinline__ void fb_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Small descriptor tables travel BY VALUE as kernel parameters (constant bank): a grouped kernel then finds its problem without
// a chain of dependent global loads (one L2 round trip per probed descriptor, microseconds on the step's critical path).
template <typename D, int N>
struct DescTable { int n; int pad[3]; D d[N]; };
