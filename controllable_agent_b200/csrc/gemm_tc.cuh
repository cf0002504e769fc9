// Grouped "NT" GEMM on the 5th-generation tensor cores with fp32-grade accuracy (3xTF32), for the wide layers of the
// MLP stacks:    C[M,N] = epi( A[M,K] . B[N,K]^T  (+ A2[M,K2] . B2[N,K2]^T)  + bias )      A, B row-major, K contiguous
//
//   forward  Y  = X . W^T   : A = X,  B = W              (nn.Linear, fb_modules.py:76)
//   backward dX = dY . W    : A = dY, B = W^T          (transposed copy staged by the plan right before the launch)
//   backward dW = dY^T . X  : A = dY^T, B = X^T        (both staged transposed: K = batch becomes the contiguous dimension)
//
// Why tensor cores here: the parity gate is 1e-3 on gradients against an fp32 reference, which plain TF32 (10-bit
// mantissa) cannot hold, but the split  x = hi + lo  (hi = x with the low 13 mantissa bits cleared = exactly what
// kind::tf32 reads, lo = x - hi, exact) with three MMA chains  hi.hi + lo.hi + hi.lo  reproduces fp32 products to ~2^-21
// and still runs several times faster than the CUDA-core FFMA loop.  The fp32 SIMT kernel (gemm_simt.cuh) remains for the
// tiny / shared-output products.  Operands TMA cannot address directly (mn-major, or a leading dimension that is not a
// multiple of 4 floats) are first staged into an aligned K-major copy by k_transpose_grouped.
//
// CTA = one 128 x BN output tile (BN = 128 or 64), 6 warps:
//   warp 0      TMA producer      raw fp32 tiles (BK = 32 floats = one 128-byte swizzle atom) -> smem ring, 3 stages
//   warp 1      MMA issuer        one thread: 3 chains x 4 tcgen05.mma (M128 x BN x K8, kind::tf32) per stage, accumulators in TMEM
//                                 (the tensor core's fp32 accumulate truncates, a bias that grows with the number of sequential
//                                 accumulations: the hi.hi chain is therefore spread round-robin over three accumulators and the
//                                 two small correction chains go to a fourth; the epilogue adds the four in IEEE fp32)
//   warps 2..5  lo-part builders  lo = x - trunc_tf32(x) for the landed A and B tiles (elementwise, layout agnostic),
//                                 then the epilogue: tcgen05.ld, bias / ReLU / ReLU-mask / tanh'-mask, global store
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_simt.cuh"  // GF_* epilogue flags

#define TC_BM 128
#define TC_BK 32
#define TC_MAX_STAGES 6
#define TC_THREADS 192
#define TC_RING_BYTES 196608           // shared-memory ring; a stage is [A raw 16K | B raw BN*128 | A lo 16K | B lo BN*128]
#define TC_SMEM_BYTES (TC_RING_BYTES + 1024)

struct __align__(64) TcGemmDesc {
  CUtensorMap mapA, mapB, mapA2, mapB2;  // box = 32 floats x 128 rows (A) / BN rows (B), SWIZZLE_128B
  float* C;
  const float* bias;
  const float* mask;
  int M, N, K, K2;
  int ldc, ldmask, flags, bn;
  int tiles_m, tiles_n, work_begin, work_count;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(tc_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint64_t tc_umma_desc(uint32_t smem_addr) {  // K-major, SWIZZLE_128B, SBO = 1024 B, version 1
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc(const TcGemmDesc* __restrict__ descs, int nprob) {
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_raw[TC_MAX_STAGES];    // TMA -> lo builders
  __shared__ __align__(8) uint64_t bar_ready[TC_MAX_STAGES];  // lo builders -> MMA issuer
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];  // MMA issuer (tcgen05.commit) -> TMA producer
  __shared__ __align__(8) uint64_t bar_accum;             // all MMAs retired -> epilogue
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // locate this CTA's problem and tile
  const int w = blockIdx.x;
  int p = 0;
  while (p + 1 < nprob && descs[p + 1].work_begin <= w) ++p;
  const TcGemmDesc* __restrict__ d = &descs[p];
  const int local = w - d->work_begin;
  const int tm = local / d->tiles_n, tn = local - tm * d->tiles_n;
  const int bn = d->bn;
  const int m0 = tm * TC_BM, n0 = tn * bn;
  const int nk1 = (d->K + TC_BK - 1) / TC_BK;
  const int nk = nk1 + (d->K2 + TC_BK - 1) / TC_BK;
  // ring geometry: narrower B tiles leave room for more stages (3 at BN = 128, 4 at BN = 64 / 32)
  const uint32_t half_bytes = 16384u + (uint32_t)bn * 128u;   // raw (or lo) part of a stage: A tile then B tile
  const uint32_t stage_bytes = 2u * half_bytes;
  const int nst = min(TC_MAX_STAGES, (int)(TC_RING_BYTES / stage_bytes));

  const uint32_t smem_base = (tc_smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = tc_smem_raw + (smem_base - tc_smem_u32(tc_smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_MAX_STAGES; ++s) {
      tc_mbar_init(&bar_raw[s], 1);
      tc_mbar_init(&bar_ready[s], 128);
      tc_mbar_init(&bar_empty[s], 1);
    }
    tc_mbar_init(&bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_smem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      const uint32_t tx_bytes = (uint32_t)(TC_BM + bn) * 128u;
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % nst;
        if (kb >= nst) tc_mbar_wait(&bar_empty[s], ((kb / nst) - 1) & 1);
        tc_mbar_expect_tx(&bar_raw[s], tx_bytes);
        const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
        if (kb < nk1) {
          tc_tma_load_2d(st, &d->mapA, &bar_raw[s], kb * TC_BK, m0);
          tc_tma_load_2d(st + 16384u, &d->mapB, &bar_raw[s], kb * TC_BK, n0);
        } else {
          tc_tma_load_2d(st, &d->mapA2, &bar_raw[s], (kb - nk1) * TC_BK, m0);
          tc_tma_load_2d(st + 16384u, &d->mapB2, &bar_raw[s], (kb - nk1) * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % nst;
        tc_mbar_wait(&bar_ready[s], (kb / nst) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
#pragma unroll
        for (int chain = 0; chain < 3; ++chain) {   // (A raw, B raw), (A lo, B raw), (A raw, B lo)
          const uint32_t a = st + (chain == 1 ? half_bytes : 0u);
          const uint32_t b = st + 16384u + (chain == 2 ? half_bytes : 0u);
          // accumulators: columns [j*bn, (j+1)*bn): j = kb % 3 for hi.hi, j = 3 for the two correction chains
          const uint32_t acc = tmem_base + (uint32_t)(chain == 0 ? (kb % 3) : 3) * (uint32_t)bn;
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            const bool first = (ks == 0) && (chain == 0 ? kb < 3 : (kb == 0 && chain == 1));
            tc_mma_tf32(acc, tc_umma_desc(a + ks * 32u), tc_umma_desc(b + ks * 32u), idesc, first ? 0u : 1u);
          }
        }
        tc_mma_commit(&bar_empty[s]);
      }
      tc_mma_commit(&bar_accum);
    }
  } else {
    // ===== lo-part builders (128 threads), then epilogue =====
    const int t = threadIdx.x - 64;
    const int nvec = (TC_BM + bn) * 8;   // float4s of the raw A and B tiles (contiguous: A 16 KB then B)
    for (int kb = 0; kb < nk; ++kb) {
      const int s = kb % nst;
      tc_mbar_wait(&bar_raw[s], (kb / nst) & 1);
      float4* raw = reinterpret_cast<float4*>(smem_gen + (size_t)s * stage_bytes);
      float4* lo = raw + half_bytes / 16;
#pragma unroll 4
      for (int i = t; i < nvec; i += 128) {
        const float4 x = raw[i];
        float4 l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
        l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
        l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
        lo[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core (async proxy)
      tc_mbar_arrive(&bar_ready[s]);
    }
    tc_mbar_wait(&bar_accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < d->M;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int flags = d->flags, N = d->N, ldc = d->ldc, ldmask = d->ldmask;
    const float* __restrict__ bias = d->bias;
    const float* __restrict__ mask = d->mask;
    float* __restrict__ C = d->C;
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
    const int n_hh = nk < 3 ? nk : 3;   // hi.hi accumulators that were written
    for (int cb = 0; cb < bn; cb += 16) {
      float v[16], u[16];
      tc_tmem_ld16(lane_addr + (uint32_t)(3 * bn + cb), v);   // correction chains first (smallest terms)
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int a = n_hh - 1; a >= 0; --a) {
        tc_tmem_ld16(lane_addr + (uint32_t)(a * bn + cb), u);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += u[j];
      }
      const int col0 = n0 + cb;
      if (!row_ok || col0 >= N) continue;
      float mk[16];
      if (flags & (GF_MASK_RELU | GF_MASK_TANH)) {   // the saved activation of this row: 4 x 128-bit loads when aligned
        const float* mp = mask + (size_t)row * ldmask + col0;
        if (((ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15u) == 0) && col0 + 15 < N) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(mp + j));
            mk[j] = t4.x; mk[j + 1] = t4.y; mk[j + 2] = t4.z; mk[j + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) mk[j] = (col0 + j < N) ? __ldg(mp + j) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = col0 + j;
        if (col < N) {
          float x = v[j];
          if (bias) x += __ldg(bias + col);
          if (flags & GF_RELU) x = fmaxf(x, 0.f);
          if (flags & GF_MASK_RELU) x = (mk[j] > 0.f) ? x : 0.f;
          if (flags & GF_MASK_TANH) x *= (1.f - mk[j] * mk[j]);
          v[j] = x;
        }
      }
      float* cp = C + (size_t)row * ldc + col0;
      if (vec_ok && col0 + 15 < N) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < N) cp[j] = v[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// operand staging for the tensor-core GEMM: out = in^T (transpose = 1) or an aligned copy of in (transpose = 0), with a
// 16-byte aligned base and leading dimension so that TMA can address it.  32 x 32 tiles through shared memory.
struct TransposeDesc { const float* in; float* out; int rows, cols, ld_in, ld_out, transpose, cta_begin, ctas_x; };

__global__ void __launch_bounds__(256) k_transpose_grouped(const TransposeDesc* __restrict__ descs, int nprob) {
  __shared__ float tile[32][33];
  int p = 0;
  while (p + 1 < nprob && descs[p + 1].cta_begin <= (int)blockIdx.x) ++p;
  const TransposeDesc d = descs[p];
  const int local = blockIdx.x - d.cta_begin;
  const int bx = local % d.ctas_x, by = local / d.ctas_x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (!d.transpose) {
    for (int j = ty; j < 32; j += 8) {
      const int r = by * 32 + j, c = bx * 32 + tx;
      if (r < d.rows && c < d.cols) d.out[(size_t)r * d.ld_out + c] = d.in[(size_t)r * d.ld_in + c];
    }
    return;
  }
  for (int j = ty; j < 32; j += 8) {
    const int r = by * 32 + j, c = bx * 32 + tx;
    tile[j][tx] = (r < d.rows && c < d.cols) ? d.in[(size_t)r * d.ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = bx * 32 + j, r = by * 32 + tx;
    if (c < d.cols && r < d.rows) d.out[(size_t)c * d.ld_out + r] = tile[tx][j];
  }
}
