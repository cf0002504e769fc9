// Data-parallel exchange over NVLink peer memory, by the library's own kernels (no NCCL on the step's path).
//
// One process per GPU; rank r steps B/R rows of the global batch against its replay shard.  Every rank owns
// an ARENA (one cudaMalloc, opened by the peers through CUDA IPC, or plain pointers when several engines share a device in the
// tests) with one fixed layout:
//     [flags | gradient fb | gradient actor | parameters fb | parameters actor | global exchange block]
// and the three exchanges of a step are kernels that load / store peer arenas directly:
//   * k_p2p_scatter_rows   this rank's [F1|F2|tF1|tF2|B|tB|discount] rows -> rows [rank*B, +B) of every rank's global block
//                          (replaces the all-gather between FB_FWD and FB_LOSS);
//   * k_p2p_adam           reduce-scatter + Adam + all-gather in one pass: rank r sums slice r of the flat gradient over the R arenas
//                          (its own from HBM, R-1 over NVLink), applies Adam to slice r of the parameters (moments of slice r live on
//                          rank r only) and stores the new parameter slice into every arena (replaces all-reduce + k_adam; the bytes
//                          on NVLink per rank are 2 (R-1)/R of a segment instead of an all-reduce's 2 (R-1)/R in AND out plus an extra
//                          HBM round trip of the summed gradient);
//   * k_p2p_finish         after every rank's slices have landed: target-network lerp towards the new parameters (local) and the
//                          clear of the local gradient (no peer reads it any more).
// Cross-GPU ordering uses monotonic epoch flags: a rank SIGNALS barrier `id` by storing the barrier's next epoch into slot
// flags[id][rank] of every arena (st.release.sys after __threadfence_system), and WAITS by spinning on its own arena's slots
// (ld.acquire.sys on local memory) until every rank's slot has reached the epoch.  Epochs only grow, so graph replays need no
// reset; a rank can never run more than one barrier ahead of its slowest peer, which is what makes the single global block and
// the in-place parameter stores race-free (see the barrier list in DESIGN.md).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

#define P2P_MAX_WORLD 8
#define P2P_NUM_BARRIERS 8
#define P2P_FLAG_BYTES 4096
#define P2P_SPIN_TIMEOUT_CYCLES (8000000000ll)    // ~4 s: a dead peer makes this rank give up the wait and record an error code (fb_p2p_status)
                                                  // instead of hanging the device; the step's results are then invalid

enum { P2P_BAR_STEP = 0, P2P_BAR_ROWS, P2P_BAR_GRAD_FB, P2P_BAR_PARAM_FB, P2P_BAR_GRAD_ACTOR, P2P_BAR_PARAM_ACTOR };

// first bytes of an arena
struct P2pFlags {
  unsigned long long slot[P2P_NUM_BARRIERS][P2P_MAX_WORLD];   // slot[id][q]: last epoch of barrier id signalled by rank q (written by q)
  unsigned long long epoch[P2P_NUM_BARRIERS];                 // local: epochs of barrier id this rank has signalled so far
  unsigned int ticket[P2P_NUM_BARRIERS];                      // local: CTAs of a multi-CTA producer that have finished
  unsigned int error;
};
static_assert(sizeof(P2pFlags) <= P2P_FLAG_BYTES, "flags must fit their block");

struct P2pPeers { char* base[P2P_MAX_WORLD]; int world, rank; };   // base[q]: arena of rank q as addressable from this rank

__device__ __forceinline__ void p2p_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long p2p_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// one thread per peer: publish this rank's next epoch of barrier `id` into every arena (everything this GRID wrote before is ordered
// in front of it: callers fence first)
__device__ __forceinline__ void p2p_signal(const P2pPeers& P, int id, int q) {
  P2pFlags* mine = reinterpret_cast<P2pFlags*>(P.base[P.rank]);
  const unsigned long long e = mine->epoch[id] + 1ull;
  if (q < P.world) p2p_st_release_sys(&reinterpret_cast<P2pFlags*>(P.base[q])->slot[id][P.rank], e);
  __syncwarp();
  if (q == 0) mine->epoch[id] = e;
}
// one thread per peer: wait until rank q has signalled the epoch this rank has reached on barrier `id`
__device__ __forceinline__ void p2p_wait(const P2pPeers& P, int id, int q) {
  P2pFlags* mine = reinterpret_cast<P2pFlags*>(P.base[P.rank]);
  if (q >= P.world) return;
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&mine->epoch[id]);
  const long long t0 = clock64();
  while (p2p_ld_acquire_sys(&mine->slot[id][q]) < e) {
    if (clock64() - t0 > P2P_SPIN_TIMEOUT_CYCLES) { atomicCAS(&mine->error, 0u, 0xDEAD0000u + ((unsigned)id << 8) + (unsigned)q); break; }
  }
}

// signal and / or wait as a kernel of its own (1 CTA of 32 threads): stream order puts it behind the producer / in front of the consumer
// No early launch_dependents here: the kernel behind a waiting barrier must not become resident (and hold an SM's shared memory)
// while this one spins — with several ranks on one device (tests) that would starve the very peer it waits for.
__global__ void k_p2p_barrier(const __grid_constant__ P2pPeers P, int id, int do_signal, int do_wait) {
  fb_pdl_wait();
  const int q = threadIdx.x;
  if (do_signal) { __threadfence_system(); p2p_signal(P, id, q); }
  __syncwarp();
  if (do_wait) p2p_wait(P, id, q);
}

// the last CTA of a multi-CTA producer signals: every CTA fences its peer stores, then takes a ticket
__device__ __forceinline__ void p2p_signal_last_cta(const P2pPeers& P, int id) {
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int s_last;
  P2pFlags* mine = reinterpret_cast<P2pFlags*>(P.base[P.rank]);
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&mine->ticket[id], 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    if (s_last) mine->ticket[id] = 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence_system();
    p2p_signal(P, id, threadIdx.x);
  }
}

// ---- exchange 1: this rank's rows of the exchange block into every arena -----------------------------------------------------
__global__ void __launch_bounds__(256) k_p2p_scatter_rows(const __grid_constant__ P2pPeers P, const float4* __restrict__ src, size_t off_blk,
                                                          size_t n4_local) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4_local; i += stride) {
    const float4 v = src[i];
#pragma unroll
    for (int q = 0; q < P2P_MAX_WORLD; ++q)
      if (q < P.world) reinterpret_cast<float4*>(P.base[q] + off_blk)[(size_t)P.rank * n4_local + i] = v;
  }
  p2p_signal_last_cta(P, P2P_BAR_ROWS);
}

// ---- exchange 2 / 3: reduce-scatter + Adam + all-gather over the arenas ------------------------------------------------------
// Slice of rank r: float4 range [r * slice4, min(n4, (r + 1) * slice4)).  m, v: this rank's moment segments (only its slice is live).
// Elements [0, split4) use lr_a, the rest lr_b.  Signals `bar_param` when every store has been issued.
struct P2pAdamParams {
  size_t off_grad, off_param;   // byte offsets of the segment inside an arena
  float4* m; float4* v;
  size_t n4, split4, slice4;
  DevScalars* sc;
  int which, bar_param;
  float beta1, beta2, eps;
};

__device__ __forceinline__ float4 p2p_ld_peer(const float4* p) {   // peer memory: never through a (possibly stale) L1 line
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__global__ void __launch_bounds__(256) k_p2p_adam(const __grid_constant__ P2pPeers P, const __grid_constant__ P2pAdamParams A) {
  fb_pdl_trigger();
  fb_pdl_wait();
  DevScalars* sc = A.sc;
  const long long t = (A.which == 0 ? sc->step_fb : sc->step_actor) + 1;
  const float bc1 = A.which == 0 ? sc->bc1_fb : sc->bc1_actor;
  const float bc2s = A.which == 0 ? sc->bc2s_fb : sc->bc2s_actor;
  const float lr_a = A.which == 0 ? sc->lr_forward : sc->lr_actor;
  const float lr_b = A.which == 0 ? sc->lr_backward : sc->lr_actor;
  const float gs = sc->grad_scale;
  const size_t lo = (size_t)P.rank * A.slice4;
  const size_t hi = lo + A.slice4 < A.n4 ? lo + A.slice4 : A.n4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    // gradient of the global batch: the R per-rank partial gradients, summed in rank order (every rank computes the same bits)
    float4 g[P2P_MAX_WORLD];
#pragma unroll
    for (int q = 0; q < P2P_MAX_WORLD; ++q)
      if (q < P.world) g[q] = p2p_ld_peer(reinterpret_cast<const float4*>(P.base[q] + A.off_grad) + i);
    float4 gg = g[0];
#pragma unroll
    for (int q = 1; q < P2P_MAX_WORLD; ++q)
      if (q < P.world) { gg.x += g[q].x; gg.y += g[q].y; gg.z += g[q].z; gg.w += g[q].w; }
    float4* pl = reinterpret_cast<float4*>(P.base[P.rank] + A.off_param) + i;
    float4 pp = *pl, mm = A.m[i], vv = A.v[i];
    const float step_size = (i < A.split4 ? lr_a : lr_b) / bc1;
    float* pa = reinterpret_cast<float*>(&pp); float* ga = reinterpret_cast<float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm); float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * gs;
      ma[j] = ma[j] * A.beta1 + gr * (1.f - A.beta1);
      va[j] = va[j] * A.beta2 + (gr * gr) * (1.f - A.beta2);
      const float denom = sqrtf(va[j]) / bc2s + A.eps;
      pa[j] = pa[j] - step_size * (ma[j] / denom);
    }
    A.m[i] = mm; A.v[i] = vv;
#pragma unroll
    for (int q = 0; q < P2P_MAX_WORLD; ++q)
      if (q < P.world) reinterpret_cast<float4*>(P.base[q] + A.off_param)[i] = pp;
  }
  // Adam step count + next step's bias corrections: published by the last CTA (as k_adam does), which also signals the peers
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int s_last;
  P2pFlags* mine = reinterpret_cast<P2pFlags*>(P.base[P.rank]);
  if (threadIdx.x == 0) {
    const unsigned int tk = atomicAdd(&mine->ticket[A.bar_param], 1u);
    s_last = (tk == gridDim.x - 1) ? 1u : 0u;
    if (s_last) {
      mine->ticket[A.bar_param] = 0u;
      float n1, n2;
      adam_bias_corrections(A.beta1, A.beta2, t + 1, &n1, &n2);
      if (A.which == 0) { sc->step_fb = t; sc->bc1_fb = n1; sc->bc2s_fb = n2; }
      else { sc->step_actor = t; sc->bc1_actor = n1; sc->bc2s_actor = n2; }
    }
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence_system();
    p2p_signal(P, A.bar_param, threadIdx.x);
  }
}

// after the wait on bar_param: every slice of the new parameters is in this arena.  Target lerp (fb only) + gradient clear, local.
__global__ void __launch_bounds__(256) k_p2p_finish(const float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ target, size_t n4,
                                                    const DevScalars* __restrict__ sc) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const float tau = sc->tau;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (target) {
      const float4 pp = p2p_ld_peer(p + i);   // written by peers over NVLink during this step: not through L1
      float4 tt = target[i];
      tt.x = tau * pp.x + (1.f - tau) * tt.x; tt.y = tau * pp.y + (1.f - tau) * tt.y;
      tt.z = tau * pp.z + (1.f - tau) * tt.z; tt.w = tau * pp.w + (1.f - tau) * tt.w;
      target[i] = tt;
    }
  }
}
