// Grouped fp32 SIMT SGEMM for the MLP stacks of the FB-DDPG step (sm_100a CUDA cores).
//
// One launch executes a *group* of independent problems C = epi(op(A) . op(B)^T [+ op(A2) . op(B2)^T] + bias): the
// layers of different networks that sit at the same depth of the step's dependency graph are issued together so that a
// launch fills the 148 SMs even though each problem only has M = batch = 1024 rows.
//
// op(A)[m,k]: a_kmajor ? A[m*lda+k] : A[k*lda+m]      op(B)[n,k]: b_kmajor ? B[n*ldb+k] : B[k*ldb+n]
//   forward  Y  = X . W^T      : A = X  (k-major), B = W  [N,K] (k-major)       (nn.Linear, fb_modules.py:76)
//   backward dX = dY . W       : A = dY (k-major), B = W  as [k=N_out][n=K_in]  (mn-major)
//   backward dW = dY^T . X     : A = dY as [k=batch][m=N_out] (mn-major), B = X as [k=batch][n=K_in] (mn-major)
//
// Kernel structure: 256 threads, CTA tile BM x BN in {128x128, 128x64, 64x64}, BK = 16, per-thread micro-tile
// (BM/16) x (BN/16) in registers.  Operand tiles are staged global -> shared with cp.async (LDGSTS, zero-fill for
// ragged edges) in a 3-stage ring: one __syncthreads per k-tile, no register staging.  Both tiles live in shared
// memory as S[k][m] so that the FFMA loop reads its fragments as LDS.128 over 4 consecutive m; a k-major operand is
// transposed on the way in by 4-byte cp.async scatters, with an XOR swizzle of the m index by the k-chunk so that the
// scatter is bank-conflict free.
#pragma once
#include "common.cuh"

enum {
  GF_RELU = 1,       // C = max(C, 0)
  GF_MASK_RELU = 2,  // C *= (mask > 0)                      (backward through a saved ReLU output)
  GF_MASK_TANH = 4,  // C *= (1 - mask^2)                    (backward through a saved tanh output)
  GF_ATOMIC = 8,     // C += result with fp32 atomics (split-K or several problems sharing one C)
  GF_SHARED_C = 16,  // (host only) several problems of the group accumulate into this C: must stay on the atomic path
  GF_RELU_LAZY_OK = 32  // (host only) with GF_RELU: every consumer of C can apply the ReLU itself (tensor-core A operand, staged
                        // copies, sign masks), so the plan may store the pre-activation instead when that lets it split K
};
enum { GEMM_CFG_BIG = 0, GEMM_CFG_SMALL = 1, GEMM_CFG_WIDE = 2 };  // 128x128 / 64x64 / 128x64 CTA tiles

struct __align__(16) GemmDesc {
  const float* A;
  const float* B;
  const float* A2;    // optional second K segment (same majors / leading dims): C = A.B^T + A2.B2^T
  const float* B2;
  float* C;
  const float* bias;  // [N] or null; added once (by k-split 0)
  const float* mask;  // same indexing as C with ldmask, or null
  int M, N, K, K2;
  int lda, ldb, ldc, ldmask;
  int a_kmajor, b_kmajor;
  int a_vec, b_vec, c_vec;  // 128-bit access legal (pointer and leading dimension 16-byte aligned)
  int flags, cfg;
  int tiles_m, tiles_n, splitk, k_per_split;
  int work_begin, work_count;  // CTA range of this problem inside the group
  float alpha;
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * (128 + 128) * GEMM_BK * 4;  // 49152

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// swizzled position of (k, m) inside a staged tile S[BK][ROWS]: m ^ (8 * ((k >> 2) & 3)) keeps 4-float groups intact
__device__ __forceinline__ int gemm_sw(int k, int m, int rows) { return k * rows + (m ^ (((k >> 2) & 3) << 3)); }

// stage one ROWS x BK operand tile (rows [row0, row0+ROWS) x k [k0, k0+BK), clipped to rows_total / kend) into S[k][m]
template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void gemm_stage_tile(float* __restrict__ S, const float* __restrict__ base, int ld, int rows_total,
                                                int row0, int k0, int kend, bool vec) {
  constexpr int CHUNKS = ROWS * GEMM_BK / 4;
  static_assert(CHUNKS % GEMM_THREADS == 0, "tile chunks must divide over the CTA");
#pragma unroll
  for (int t = 0; t < CHUNKS / GEMM_THREADS; ++t) {
    const int c = threadIdx.x + t * GEMM_THREADS;
    if (KMAJOR) {
      // 4 consecutive k of one row -> 4 scattered 4-byte copies (transpose); any alignment
      const int r = c >> 2, kc = (c & 3) << 2;
      const int row = row0 + r, k = k0 + kc;
      const int nvalid = (row < rows_total) ? max(0, min(4, kend - k)) : 0;
      const float* src = nvalid > 0 ? base + (size_t)row * ld + k : base;
#pragma unroll
      for (int e = 0; e < 4; ++e) cp_async4(S + gemm_sw(kc + e, r, ROWS), e < nvalid ? src + e : base, e < nvalid ? 4 : 0);
    } else {
      constexpr int C4 = ROWS / 4;
      const int kk = c / C4, rc = (c % C4) << 2;
      const int k = k0 + kk, row = row0 + rc;
      float* dst = S + gemm_sw(kk, rc, ROWS);
      const int nvalid = (k < kend) ? max(0, min(4, rows_total - row)) : 0;
      const float* src = nvalid > 0 ? base + (size_t)k * ld + row : base;
      if (vec) {
        cp_async16(dst, src, nvalid * 4);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) cp_async4(dst + e, e < nvalid ? src + e : base, e < nvalid ? 4 : 0);
      }
    }
  }
}

template <int BM, int BN, bool AK, bool BKM>
__device__ __forceinline__ void gemm_tile(const GemmDesc& d, int tm, int tn, int ks, float* __restrict__ smem) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int A_STAGE = BM * GEMM_BK, B_STAGE = BN * GEMM_BK, STAGE = A_STAGE + B_STAGE;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = tm * BM, n0 = tn * BN;
  // `d` lives in shared memory: its fields are re-read where needed instead of being pinned in registers
  const int kbeg = ks * d.k_per_split;
  const int kend = min(d.K, kbeg + d.k_per_split);
  const int nk = (kend - kbeg + GEMM_BK - 1) / GEMM_BK + (d.K2 + GEMM_BK - 1) / GEMM_BK;  // K2 never meets split-K

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // staging cursor: walks the k-tiles of segment 0 (A, B), then of segment 1 (A2, B2)
  const float* sA = d.A;
  const float* sB = d.B;
  int sk = kbeg, skend = kend;
  auto stage = [&](int buf) {
    if (sk >= skend) { sA = d.A2; sB = d.B2; sk = 0; skend = d.K2; }
    float* As = smem + buf * STAGE;
    gemm_stage_tile<BM, AK>(As, sA, d.lda, d.M, m0, sk, skend, d.a_vec != 0);
    gemm_stage_tile<BN, BKM>(As + A_STAGE, sB, d.ldb, d.N, n0, sk, skend, d.b_vec != 0);
    sk += GEMM_BK;
  };

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < nk) stage(s);
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();  // tile kt has landed for every thread; every thread is done reading tile kt-1
    if (kt + GEMM_STAGES - 1 < nk) stage((kt + GEMM_STAGES - 1) % GEMM_STAGES);  // overwrites the buffer of tile kt-1
    cp_async_commit();

    const float* __restrict__ as = smem + (kt % GEMM_STAGES) * STAGE;
    const float* __restrict__ bs = as + A_STAGE;
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      const int sw = ((kk >> 2) & 3) << 3;
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < TM / 4; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(&as[kk * BM + ((g * 64 + ty * 4) ^ sw)]);
        a[g * 4 + 0] = t.x; a[g * 4 + 1] = t.y; a[g * 4 + 2] = t.z; a[g * 4 + 3] = t.w;
      }
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(&bs[kk * BN + ((g * 64 + tx * 4) ^ sw)]);
        b[g * 4 + 0] = t.x; b[g * 4 + 1] = t.y; b[g * 4 + 2] = t.z; b[g * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: thread rows (i>>2)*64 + ty*4 + (i&3), columns (j>>2)*64 + tx*4 + (j&3) ----
  const int M = d.M, N = d.N;
  const float alpha = d.alpha;
  const float* __restrict__ bias = (ks == 0) ? d.bias : nullptr;
  const float* __restrict__ mask = d.mask;
  const int flags = d.flags, ldc = d.ldc, ldmask = d.ldmask;
  const bool cvec = d.c_vec != 0;
  float* __restrict__ C = d.C;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (row >= M) continue;
#pragma unroll
    for (int g = 0; g < TN / 4; ++g) {
      const int col = n0 + g * 64 + tx * 4;
      if (col >= N) continue;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x = alpha * acc[i][g * 4 + e];
        if (col + e < N) {
          if (bias) x += __ldg(bias + col + e);
          if (flags & GF_RELU) x = fmaxf(x, 0.f);
          if (flags & GF_MASK_RELU) x = (__ldg(mask + (size_t)row * ldmask + col + e) > 0.f) ? x : 0.f;
          if (flags & GF_MASK_TANH) { const float t = __ldg(mask + (size_t)row * ldmask + col + e); x *= (1.f - t * t); }
        }
        v[e] = x;
      }
      float* cp = C + (size_t)row * ldc + col;
      if (flags & GF_ATOMIC) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < N) atomicAdd(cp + e, v[e]);
      } else if (cvec && col + 3 < N) {
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < N) cp[e] = v[e];
      }
    }
  }
}

template <int BM, int BN>
__device__ __forceinline__ void gemm_dispatch_major(const GemmDesc& d, int tm, int tn, int ks, float* smem) {
  if (d.a_kmajor) {
    if (d.b_kmajor) gemm_tile<BM, BN, true, true>(d, tm, tn, ks, smem);
    else gemm_tile<BM, BN, true, false>(d, tm, tn, ks, smem);
  } else {
    if (d.b_kmajor) gemm_tile<BM, BN, false, true>(d, tm, tn, ks, smem);
    else gemm_tile<BM, BN, false, false>(d, tm, tn, ks, smem);
  }
}

// grid.x = total work items (tiles x k-splits) of the group
__global__ void __launch_bounds__(GEMM_THREADS, 2) k_gemm_grouped(const GemmDesc* __restrict__ descs, int nprob) {
  fb_pdl_trigger();
  fb_pdl_wait();
  extern __shared__ __align__(16) float gemm_smem[];
  __shared__ GemmDesc sd;
  const int w = blockIdx.x;
  int p = 0;
  while (p + 1 < nprob && descs[p + 1].work_begin <= w) ++p;
  if (threadIdx.x < (int)(sizeof(GemmDesc) / sizeof(int)))
    reinterpret_cast<int*>(&sd)[threadIdx.x] = reinterpret_cast<const int*>(&descs[p])[threadIdx.x];
  __syncthreads();
  const int local = w - sd.work_begin;
  const int per_split = sd.tiles_m * sd.tiles_n;
  const int ks = local / per_split;
  const int t = local - ks * per_split;
  const int tm = t / sd.tiles_n, tn = t - tm * sd.tiles_n;
  if (sd.cfg == GEMM_CFG_BIG) gemm_dispatch_major<128, 128>(sd, tm, tn, ks, gemm_smem);
  else if (sd.cfg == GEMM_CFG_WIDE) gemm_dispatch_major<128, 64>(sd, tm, tn, ks, gemm_smem);
  else gemm_dispatch_major<64, 64>(sd, tm, tn, ks, gemm_smem);
}
