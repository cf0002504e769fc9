// Grouped fp32 SIMT SGEMM for the MLP stacks of the FB-DDPG step (sm_100a CUDA cores).
//
// One launch executes a *group* of independent problems C = epi(alpha * op(A) . op(B)^T + bias): the
// layers of different networks that sit at the same depth of the step's dependency graph are issued
// together so that a launch fills the 148 SMs even though each problem only has M = batch = 1024 rows.
//
// op(A)[m,k]: a_kmajor ? A[m*lda+k] : A[k*lda+m]      op(B)[n,k]: b_kmajor ? B[n*ldb+k] : B[k*ldb+n]
//   forward  Y  = X . W^T      : A = X  (k-major), B = W  [N,K] (k-major)       (nn.Linear, fb_modules.py:76)
//   backward dX = dY . W       : A = dY (k-major), B = W  as [k=N_out][n=K_in]  (mn-major)
//   backward dW = dY^T . X     : A = dY as [k=batch][m=N_out] (mn-major), B = X as [k=batch][n=K_in] (mn-major)
#pragma once
#include "common.cuh"

enum {
  GF_RELU = 1,       // C = max(C, 0)
  GF_MASK_RELU = 2,  // C *= (mask > 0)                      (backward through a saved ReLU output)
  GF_MASK_TANH = 4,  // C *= (1 - mask^2)                    (backward through a saved tanh output)
  GF_ATOMIC = 8      // C += result with fp32 atomics (split-K or several problems sharing one C)
};
enum { GEMM_CFG_BIG = 0, GEMM_CFG_SMALL = 1 };  // 128x128 tile, 8x8 per thread / 64x64 tile, 4x4 per thread

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
constexpr int GEMM_SMEM_BYTES = 2 * GEMM_BK * (128 + 4 + 128 + 4) * 4;

template <int ROWS, bool KMAJOR>
__device__ __forceinline__ float4 gemm_load_slot(const float* __restrict__ base, int ld, int rows_total, int row0,
                                                 int k0, int kend, int slot, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (KMAJOR) {
    const int r = slot >> 2, k = k0 + ((slot & 3) << 2);
    const int row = row0 + r;
    if (row < rows_total && k < kend) {
      const float* p = base + (size_t)row * ld + k;
      if (vec && k + 3 < kend) {
        v = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        v.x = __ldg(p);
        if (k + 1 < kend) v.y = __ldg(p + 1);
        if (k + 2 < kend) v.z = __ldg(p + 2);
        if (k + 3 < kend) v.w = __ldg(p + 3);
      }
    }
  } else {
    constexpr int C4 = ROWS / 4;
    const int k = k0 + slot / C4, row = row0 + (slot % C4) * 4;
    if (k < kend && row < rows_total) {
      const float* p = base + (size_t)k * ld + row;
      if (vec && row + 3 < rows_total) {
        v = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        v.x = __ldg(p);
        if (row + 1 < rows_total) v.y = __ldg(p + 1);
        if (row + 2 < rows_total) v.z = __ldg(p + 2);
        if (row + 3 < rows_total) v.w = __ldg(p + 3);
      }
    }
  }
  return v;
}

template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void gemm_store_slot(float* __restrict__ S, int slot, float4 v) {
  constexpr int LDS = ROWS + 4;
  if (KMAJOR) {
    const int r = slot >> 2, k = (slot & 3) << 2;
    S[(k + 0) * LDS + r] = v.x;
    S[(k + 1) * LDS + r] = v.y;
    S[(k + 2) * LDS + r] = v.z;
    S[(k + 3) * LDS + r] = v.w;
  } else {
    constexpr int C4 = ROWS / 4;
    const int k = slot / C4, c = (slot % C4) * 4;
    *reinterpret_cast<float4*>(&S[k * LDS + c]) = v;
  }
}

// accumulate op(A)[m0.., kbeg..kend) . op(B)[n0.., kbeg..kend)^T into acc (register-prefetch double buffering)
template <int BM, int BN, int TM, int TN, bool AK, bool BKM>
__device__ __forceinline__ void gemm_accumulate(const float* __restrict__ A, const float* __restrict__ B, int lda, int ldb,
                                                int M, int N, int m0, int n0, int kbeg, int kend, bool avec, bool bvec,
                                                float* __restrict__ smem, float (&acc)[TM][TN]) {
  constexpr int BK = GEMM_BK;
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int A_SLOTS = BM * BK / 4 / GEMM_THREADS;
  constexpr int B_SLOTS = BN * BK / 4 / GEMM_THREADS;
  static_assert(A_SLOTS >= 1 && B_SLOTS >= 1, "tile too small for 256 threads");
  static_assert((BM / TM) * (BN / TN) == GEMM_THREADS, "thread tiling must cover the CTA tile");
  float* As = smem;                    // [2][BK][LDA_S]
  float* Bs = smem + 2 * BK * LDA_S;   // [2][BK][LDB_S]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float4 ra[A_SLOTS], rb[B_SLOTS];

  const int nk = (kend - kbeg + BK - 1) / BK;
  if (nk <= 0) return;
#pragma unroll
  for (int i = 0; i < A_SLOTS; ++i) ra[i] = gemm_load_slot<BM, AK>(A, lda, M, m0, kbeg, kend, tid + i * GEMM_THREADS, avec);
#pragma unroll
  for (int i = 0; i < B_SLOTS; ++i) rb[i] = gemm_load_slot<BN, BKM>(B, ldb, N, n0, kbeg, kend, tid + i * GEMM_THREADS, bvec);
#pragma unroll
  for (int i = 0; i < A_SLOTS; ++i) gemm_store_slot<BM, AK>(As, tid + i * GEMM_THREADS, ra[i]);
#pragma unroll
  for (int i = 0; i < B_SLOTS; ++i) gemm_store_slot<BN, BKM>(Bs, tid + i * GEMM_THREADS, rb[i]);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    const bool more = kt + 1 < nk;
    if (more) {
      const int k0 = kbeg + (kt + 1) * BK;
#pragma unroll
      for (int i = 0; i < A_SLOTS; ++i) ra[i] = gemm_load_slot<BM, AK>(A, lda, M, m0, k0, kend, tid + i * GEMM_THREADS, avec);
#pragma unroll
      for (int i = 0; i < B_SLOTS; ++i) rb[i] = gemm_load_slot<BN, BKM>(B, ldb, N, n0, k0, kend, tid + i * GEMM_THREADS, bvec);
    }
    const float* __restrict__ as = As + buf * BK * LDA_S;
    const float* __restrict__ bs = Bs + buf * BK * LDB_S;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      {
        const float4 t = *reinterpret_cast<const float4*>(&as[kk * LDA_S + ty * 4]);
        a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
      }
      if (TM == 8) {
        const float4 t = *reinterpret_cast<const float4*>(&as[kk * LDA_S + BM / 2 + ty * 4]);
        a[TM - 4] = t.x; a[TM - 3] = t.y; a[TM - 2] = t.z; a[TM - 1] = t.w;
      }
      {
        const float4 t = *reinterpret_cast<const float4*>(&bs[kk * LDB_S + tx * 4]);
        b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
      }
      if (TN == 8) {
        const float4 t = *reinterpret_cast<const float4*>(&bs[kk * LDB_S + BN / 2 + tx * 4]);
        b[TN - 4] = t.x; b[TN - 3] = t.y; b[TN - 2] = t.z; b[TN - 1] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      float* an = As + (buf ^ 1) * BK * LDA_S;
      float* bn = Bs + (buf ^ 1) * BK * LDB_S;
#pragma unroll
      for (int i = 0; i < A_SLOTS; ++i) gemm_store_slot<BM, AK>(an, tid + i * GEMM_THREADS, ra[i]);
#pragma unroll
      for (int i = 0; i < B_SLOTS; ++i) gemm_store_slot<BN, BKM>(bn, tid + i * GEMM_THREADS, rb[i]);
    }
    __syncthreads();
  }
}

template <int BM, int BN, int TM, int TN, bool AK, bool BKM>
__device__ __forceinline__ void gemm_tile(const GemmDesc& d, int tm, int tn, int ks, float* __restrict__ smem) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = tm * BM, n0 = tn * BN;
  const int M = d.M, N = d.N;
  const bool avec = d.a_vec != 0, bvec = d.b_vec != 0;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  {
    const int kbeg = ks * d.k_per_split;
    const int kend = min(d.K, kbeg + d.k_per_split);
    gemm_accumulate<BM, BN, TM, TN, AK, BKM>(d.A, d.B, d.lda, d.ldb, M, N, m0, n0, kbeg, kend, avec, bvec, smem, acc);
  }
  if (d.K2 > 0)  // second segment (never combined with split-K)
    gemm_accumulate<BM, BN, TM, TN, AK, BKM>(d.A2, d.B2, d.lda, d.ldb, M, N, m0, n0, 0, d.K2, avec, bvec, smem, acc);

  // ---- epilogue ----
  const float alpha = d.alpha;
  const float* __restrict__ bias = (ks == 0) ? d.bias : nullptr;
  const float* __restrict__ mask = d.mask;
  const int flags = d.flags, ldc = d.ldc, ldmask = d.ldmask;
  float* __restrict__ C = d.C;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + ((i < 4) ? (ty * 4 + i) : (BM / 2 + ty * 4 + (i - 4)));
    if (row >= M) continue;
#pragma unroll
    for (int jc = 0; jc < TN / 4; ++jc) {
      const int col = n0 + ((jc == 0) ? tx * 4 : (BN / 2 + tx * 4));
      if (col >= N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = alpha * acc[i][jc * 4 + j];
        if (col + j < N) {
          if (bias) x += __ldg(bias + col + j);
          if (flags & GF_RELU) x = fmaxf(x, 0.f);
          if (flags & GF_MASK_RELU) x = (__ldg(mask + (size_t)row * ldmask + col + j) > 0.f) ? x : 0.f;
          if (flags & GF_MASK_TANH) {
            const float t = __ldg(mask + (size_t)row * ldmask + col + j);
            x *= (1.f - t * t);
          }
        }
        v[j] = x;
      }
      float* cp = C + (size_t)row * ldc + col;
      if (flags & GF_ATOMIC) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < N) atomicAdd(cp + j, v[j]);
      } else if (d.c_vec && col + 3 < N) {
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < N) cp[j] = v[j];
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void gemm_dispatch_major(const GemmDesc& d, int tm, int tn, int ks, float* smem) {
  if (d.a_kmajor) {
    if (d.b_kmajor) gemm_tile<BM, BN, TM, TN, true, true>(d, tm, tn, ks, smem);
    else gemm_tile<BM, BN, TM, TN, true, false>(d, tm, tn, ks, smem);
  } else {
    if (d.b_kmajor) gemm_tile<BM, BN, TM, TN, false, true>(d, tm, tn, ks, smem);
    else gemm_tile<BM, BN, TM, TN, false, false>(d, tm, tn, ks, smem);
  }
}

// grid.x = total work items (tiles x k-splits) of the group
__global__ void __launch_bounds__(GEMM_THREADS, 2) k_gemm_grouped(const GemmDesc* __restrict__ descs, int nprob) {
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
  if (sd.cfg == GEMM_CFG_BIG) gemm_dispatch_major<128, 128, 8, 8>(sd, tm, tn, ks, gemm_smem);
  else gemm_dispatch_major<64, 64, 4, 4>(sd, tm, tn, ks, gemm_smem);
}
