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
// Where the operand bytes go (profiles/r1c_gemm_tc_breakdown.txt, profiles/r2_gemm_tc_ts.txt): with both operands in shared memory
// ("SS" form) a k-block of a 128 x 128 tile moves ~176 KB through the SM's 128 B/clk shared-memory port (TMA writes, the lo
// builders' reads and writes, 6 operand reads of the three MMA chains) against an MMA floor of 768 clk = 98 KB, and every lo plane
// that TMA loads next to its raw tile also costs L2 -> SM bandwidth (~42 B/clk/SM chip-wide).  The A operand therefore takes the
// "TS" form: the builder warps read the landed raw A tile ONCE from shared memory, split it in registers (hi = the tf32 bits the
// tensor core would read, lo = x - hi, exact) and write both halves into TENSOR MEMORY (tcgen05.st), where the MMAs read them at
// no shared-memory cost; A needs no lo plane in shared memory or in HBM at all.  B stays in shared memory: its lo plane is
// pre-split once per step by the staging launch for operands that pass through one anyway (weights, transposed operands:
// TC_B_PRE) and built next to the raw tile by the builder warps otherwise.
//
// Persistent CTAs (one per SM), each looping over 128 x BN output tiles of the whole group, 6 warps:
//   warp 0      TMA producer      raw fp32 tiles (BK = 32 floats = one 128-byte swizzle atom) (+ the pre-split B lo tile) -> smem
//                                 ring; runs ahead into the next tile while the current one drains
//   warp 1      MMA issuer        one thread: 3 chains x 4 tcgen05.mma (M128 x BN x K8, kind::tf32, A from TMEM) per stage,
//                                 accumulators in TMEM (the tensor core's fp32 accumulate truncates, a bias that grows with the
//                                 number of sequential accumulations: the hi.hi chain alternates between two accumulators and the
//                                 two small correction chains go to a third; the epilogue adds the three in IEEE fp32)
//   warps 2..5  operand builders  thread = one row of the A tile: swizzled 128-byte row -> registers -> (ReLU) -> hi / lo ->
//                                 TMEM slot of the k-block; B lo = x - trunc_tf32(x) in shared memory when not pre-split;
//                                 then the epilogue: tcgen05.ld, transpose through a padded smem scratch so that every global
//                                 access (C, ReLU / tanh' mask) is a coalesced 128-bit one, bias / ReLU / masks
// TMEM (512 columns): [0, 3 bn) accumulators hi0 | hi1 | corrections, [3 ring_bn, ...) A slots of 64 columns (hi 32 | lo 32).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_simt.cuh"  // GF_* epilogue flags

#define TC_BM 128
#ifndef TC_BK
#define TC_BK 32                       // k-block = one swizzle atom row: 32 floats (SWIZZLE_128B); -DTC_BK=16 builds the SWIZZLE_64B variant
#endif                                 // (half-size stages, 6-deep ring: measured SLOWER, profiles/r2_gemm_tc_bk16_experiment.txt)
#define TC_ROW_BYTES (TC_BK * 4)
#define TC_A_TILE_BYTES (TC_BM * TC_ROW_BYTES)
#define TC_KB_PER_32 (32 / TC_BK)      // k-blocks per 32 floats of K: the unit the host-side cost model and the accumulator rotation count in
#define TC_MAX_STAGES 6
#ifndef TC_EPI_GROUPS
#define TC_EPI_GROUPS 3                // epilogue warp groups of k_gemm_tc (4 warps each; group 0 also builds the lo parts): 1 -> 2 -> 3 groups
                                       // measured 975 -> 1022 -> 1038 steps/s; a fourth does not fit next to the 192 KB ring
#endif
#define TC_THREADS (64 + 128 * TC_EPI_GROUPS)   // warp 0 TMA, warp 1 MMA, warps 2..5 builders + epilogue, further warps epilogue helpers
#define TC_EPI_CW 16                   // columns of one epilogue chunk
#define TC_RING_BYTES 196608           // shared-memory ring; a stage is [A raw | B raw (BN rows) | A lo | B lo], rows of TC_ROW_BYTES.
                                       // (3 stages at BN = 128 and TC_BK = 32)
#define TC_SMEM_BYTES (TC_RING_BYTES + 1024)
#define TC_EPI_LD 20                   // padded row of the epilogue scratch (floats): 16 + 4, conflict-free 128-bit row writes
#define TC_EPI_SCR (32 * TC_EPI_LD)    // floats of one epilogue warp's scratch (>= 16 x 33 for the transposed copy)
#define TC_A_SLOT_COLS 64              // TS form: TMEM columns of one k-block of A (hi: 32 columns, lo: 32 columns; lane = tile row)
#define TC_MAX_ASLOTS 4                // A slots in flight (2 at BN = 128: 3 x 128 accumulator columns + 2 x 64 = 512)

#define TC_A_PRE (1 << 8)    // mapAlo / mapA2lo address a pre-split lo plane of A
#define TC_B_PRE (1 << 9)    // same for B
#define TC_A_RELU (1 << 10)  // A holds the pre-activation of a lazily-ReLU'd layer (a split-K producer): the builder warps clamp
                             // the landed raw tile at 0 in shared memory before splitting it
// profiling knobs (fb_gemm_tc_bench only; results are wrong when set): which part of the pipeline bounds a tile
#define TC_DBG_NOBUILD (1 << 16)   // lo builders skip the split (arrive as soon as the raw tiles land)
#define TC_DBG_ONECHAIN (1 << 17)  // the MMA warp issues only the hi.hi chain
#define TC_DBG_NOEPI (1 << 18)     // epilogue skipped

struct __align__(64) TcGemmDesc {
  CUtensorMap mapA, mapB, mapA2, mapB2;          // box = 32 floats x 128 rows (A) / BN rows (B), SWIZZLE_128B
  CUtensorMap mapAlo, mapBlo, mapA2lo, mapB2lo;  // lo planes (same geometry), valid with TC_A_PRE / TC_B_PRE
  float* C;
  const float* bias;
  const float* mask;
  int M, N, K, K2;
  int ldc, ldmask, flags, bn;
  int tiles_m, tiles_n, work_begin, work_count;
  float* CT;      // optional: C^T [N, ldct] written next to C (the K-major operand of a later dW product: K = batch contiguous),
  float* CT_lo;   //           and its pre-split lo plane; requires splitk == 1 and M % 4 == 0
  int ldct, pad0;
  int splitk, kb_per_split;   // splitk > 1: a tile's k-blocks (of both products) are shared by splitk work items that add into a zeroed C (no ReLU)
};

// Launch header passed BY VALUE (kernel parameter = constant bank): everything a role needs to find its tile and size its
// loops without a dependent chain of global loads (a search over the 1 KB descriptors costs one L2 round trip per problem,
// several microseconds of every launch); the descriptors in global memory only supply tensor maps and pointers.
#define TC_MAX_PROBS 16
struct TcProb { int M, N, K, K2, bn, flags, tiles_n, work_begin, splitk, kb_per_split; };
struct TcLaunch { int nprob, total, ring_bn, pad; TcProb p[TC_MAX_PROBS]; };

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
// ---- CTA pairs (cta_group::2): the two CTAs of a cluster share one 256 x BN tile; rank 0 issues the MMAs for both ----
__device__ __forceinline__ uint32_t tc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (CTA-scope) semantics on purpose: a
// .release.cluster arrive compiles to MEMBAR.ALL.GPU in front of it and the matching .acquire.cluster wait to CCTL.IVALL behind it,
// once per k-block and warp (measured: a pair k-block at 1600 clk whatever the MMA count).  What the barrier orders here is shared /
// tensor memory handed to the async proxy (fence.proxy.async / tcgen05 fences on the writer's side), not global memory.
__device__ __forceinline__ void tc_mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(tc_smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait on a barrier whose arrivals come from both CTAs of the pair (see tc_mbar_arrive_remote for the scope)
__device__ __forceinline__ void tc_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void tc_mma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts_2cta(uint32_t tmem_d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of every MMA issued so far -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_mma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint64_t tc_umma_desc(uint32_t smem_addr) {  // K-major, swizzle = row bytes, SBO = 8 rows, version 1
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((8u * TC_ROW_BYTES) >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(TC_BK == 32 ? 2 : 4) << 61);   // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
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
// the same with the A operand in tensor memory ("TS" form): a_tmem = TMEM address (lane 0, first column) of a 128 x 8 tf32 block,
// row r of the tile in lane r, one 32-bit column per k
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive columns of this thread's TMEM lane (warp w of the CTA owns lanes 32 (w % 4) .. +31)
__device__ __forceinline__ void tc_tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
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

__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float4 tc_lo4(const float4 x) {
  float4 l;
  l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
  l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
  l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
  l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
  return l;
}

// lo = x - trunc_tf32(x) over CNT float4 per thread (128 threads): all loads first, then the stores
template <int CNT>
__device__ __forceinline__ void tc_build_lo(const float4* __restrict__ raw, float4* __restrict__ lo, int t) {
  float4 x[CNT];
#pragma unroll
  for (int i = 0; i < CNT; ++i) x[i] = raw[t + i * 128];
#pragma unroll
  for (int i = 0; i < CNT; ++i) lo[t + i * 128] = tc_lo4(x[i]);
}

// the same for an A tile that still needs its ReLU: raw <- max(raw, 0) in place, lo from the clamped value
template <int CNT>
__device__ __forceinline__ void tc_build_lo_relu(float4* __restrict__ raw, float4* __restrict__ lo, int t) {
  float4 x[CNT];
#pragma unroll
  for (int i = 0; i < CNT; ++i) x[i] = raw[t + i * 128];
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    x[i].x = fmaxf(x[i].x, 0.f); x[i].y = fmaxf(x[i].y, 0.f); x[i].z = fmaxf(x[i].z, 0.f); x[i].w = fmaxf(x[i].w, 0.f);
    raw[t + i * 128] = x[i];
    lo[t + i * 128] = tc_lo4(x[i]);
  }
}

// which problem / tile / k-range of the group is work item w (header only: no global loads)
template <int NCTA = 1>   // NCTA = 2: a work item is a 256-row pair tile (the header's work counts are built for pairs)
__device__ __forceinline__ int tc_locate(const TcLaunch& L, int w, int* m0, int* n0, int* kb0, int* kb1) {
  int p = 0;
  while (p + 1 < L.nprob && L.p[p + 1].work_begin <= w) ++p;
  const TcProb& d = L.p[p];
  int local = w - d.work_begin;
  const int nk = (d.K + TC_BK - 1) / TC_BK + (d.K2 + TC_BK - 1) / TC_BK;
  if (d.splitk > 1) {
    const int split = local % d.splitk;
    local /= d.splitk;
    *kb0 = split * d.kb_per_split;
    *kb1 = min(nk, *kb0 + d.kb_per_split);
  } else {
    *kb0 = 0; *kb1 = nk;
  }
  const int tm = local / d.tiles_n, tn = local - tm * d.tiles_n;
  *m0 = tm * (TC_BM * NCTA); *n0 = tn * d.bn;
  return p;
}

// Work item of (virtual) CTA `vcta` in the round that starts at item `base`: rounds alternate direction (0, 1, .., n-1, then n-1, ..,
// 0), so that with the items ordered longest first the CTAs that drew the long items of one round draw the short ones of the next.
// A plain round-robin stacks long on long: a group of 64 long + 133 short tiles on 148 CTAs ends at long + short instead of
// 2 x short (the LPT schedule; fb_b200.cu gemm_tc() orders the problems and models exactly this deal).
__device__ __forceinline__ int tc_snake(int base, int vcta, int ncta) {
  return base + (((base / ncta) & 1) ? (ncta - 1 - vcta) : vcta);
}

__device__ __forceinline__ void tc_prefetch_map(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// Shared-memory objects of the GEMM roles (static in the owning kernel: k_gemm_tc, or the fused stack kernel of fused.cuh)
struct TcShared {
  uint64_t bar_raw[TC_MAX_STAGES];    // TMA -> lo builders
  uint64_t bar_ready[TC_MAX_STAGES];  // lo builders -> MMA issuer
  uint64_t bar_empty[TC_MAX_STAGES];  // MMA issuer (tcgen05.commit) -> TMA producer
  uint64_t bar_afree[TC_MAX_ASLOTS];  // TS form: MMA issuer (tcgen05.commit) -> builders: the k-block's TMEM A slot may be rewritten
  uint64_t bar_accum;                 // all MMAs of a tile retired -> epilogue
  uint64_t bar_tmem_empty;            // epilogue has read the accumulators -> MMA issuer (next tile)
  uint32_t tmem_base;
  uint32_t pad;
};

// one thread: (re)initialise the pipeline barriers of a launch / of a GEMM item of the fused kernel (reinit: the objects are live)
__device__ __forceinline__ void tc_init_barriers(TcShared* sh, bool reinit, int group_ctas = 1, int epi_groups = 1) {
  if (reinit) {
    for (int s = 0; s < TC_MAX_STAGES; ++s) {
      asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh->bar_raw[s])) : "memory");
      asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh->bar_ready[s])) : "memory");
      asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh->bar_empty[s])) : "memory");
    }
    for (int j = 0; j < TC_MAX_ASLOTS; ++j) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh->bar_afree[j])) : "memory");
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh->bar_accum)) : "memory");
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc_smem_u32(&sh->bar_tmem_empty)) : "memory");
  }
  for (int s = 0; s < TC_MAX_STAGES; ++s) {
    tc_mbar_init(&sh->bar_raw[s], 1);
    tc_mbar_init(&sh->bar_ready[s], 4 * group_ctas);   // one arrival per builder warp of every CTA of the group, on the leader's barrier
    tc_mbar_init(&sh->bar_empty[s], 1);
  }
  for (int j = 0; j < TC_MAX_ASLOTS; ++j) tc_mbar_init(&sh->bar_afree[j], 1);
  tc_mbar_init(&sh->bar_accum, 1);
  tc_mbar_init(&sh->bar_tmem_empty, 4 * epi_groups * group_ctas);   // one arrival per epilogue warp
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// The three roles of one grouped launch, for the CTA that plays virtual CTA `vcta` of `ncta` (warps 0..5 of the block; further
// warps fall through).  Barriers freshly initialised, TMEM allocated (sh->tmem_base), `L` readable by every thread (constant bank
// or shared memory).  smem_base / smem_gen: the 1024-byte aligned operand ring (shared-window address / generic pointer).
// Activations and masks produced by earlier launches are read with plain (coherent) loads: inside the fused kernel they were
// written by other SMs during the same kernel.
// TS = true: A operand through tensor memory (FB_TC_TS=1; measured latency-bound by the number of A slots that fit next to the
// accumulators, profiles/r2_gemm_tc_ts.txt); TS = false: both operands in shared memory.
// NCTA = 2: CTA pairs (clusters of two, tcgen05 cta_group::2).  A work item is a 256 x bn tile; CTA `rank` of the pair loads its own
// 128 rows of A and HALF of the B tile (rows [rank bn/2, (rank+1) bn/2)), so a k-block costs each SM 16 + 8 KB of L2 -> SM and
// shared-memory traffic instead of 16 + 16; rank 0's MMA warp issues M = 256 instructions that read both CTAs' operands and write
// both CTAs' tensor memory; the builders / epilogue warps of both CTAs arrive on rank 0's barriers, tcgen05.commit multicasts the
// completions to both.  `vcta` / `ncta` then count pairs.
// EG: epilogue warp groups.  The epilogue of a tile is instruction-latency-bound with ONE warp per scheduler (TMEM loads, transpose,
// masks, stores: ~5 us per 128 x 128 tile, serial with the mainloop because TMEM holds one tile's accumulators), so k_gemm_tc runs a
// second group (warps 6..9, same TMEM lane quadrants) that takes every other 16-column chunk.  The fused kernel keeps one group.
template <bool TS, int NCTA, int EG = 1>
__device__ __forceinline__ void tc_gemm_roles(const TcGemmDesc* descs, const TcLaunch& L, int vcta, int ncta, TcShared* sh,
                                              float* epi_scratch, uint32_t smem_base, uint8_t* smem_gen, uint32_t rank = 0) {
  const int total = L.total, ring_bn = L.ring_bn;
  uint64_t* bar_raw = sh->bar_raw; uint64_t* bar_ready = sh->bar_ready; uint64_t* bar_empty = sh->bar_empty;
  uint64_t& bar_accum = sh->bar_accum; uint64_t& bar_tmem_empty = sh->bar_tmem_empty;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ring geometry: narrower B tiles leave room for more stages.  SS: a stage is [A raw | B raw | A lo | B lo] (3 stages at BN = 128);
  // TS: [A raw | B raw | B lo] (4 stages at BN = 128, 6 at BN <= 64)
  const uint32_t b_off = TC_A_TILE_BYTES;
  const uint32_t b_bytes = (uint32_t)(ring_bn / NCTA) * TC_ROW_BYTES;   // this CTA's share of the B tile
  const uint32_t half_bytes = TC_A_TILE_BYTES + b_bytes;   // raw part of a stage: A tile then B tile
  const uint32_t blo_off = TS ? half_bytes : half_bytes + b_off;   // the B lo tile inside a stage
  const uint32_t stage_bytes = TS ? half_bytes + b_bytes : 2u * half_bytes;
  const int nst = min(TC_MAX_STAGES, (int)(TC_RING_BYTES / stage_bytes));
  const uint32_t tmem_base = sh->tmem_base;
  // TS: TMEM columns [0, (n_hh + 1) ring_bn) hold the accumulators of the widest tile, A slots of 64 columns follow.  At BN = 128
  // only ONE hi.hi accumulator fits next to four slots (two slots cannot cover the commit -> rebuild -> issue latency)
  const int n_hh_max = TS ? (ring_bn > 64 ? 1 : 2) : 3;
  const uint32_t slot_col0 = (uint32_t)(n_hh_max + 1) * (uint32_t)ring_bn;
  const int nslots = min(TC_MAX_ASLOTS, (int)((512u - slot_col0) / TC_A_SLOT_COLS));
  uint64_t* bar_afree = sh->bar_afree;
  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t kbg = 0;   // k-blocks issued by this CTA so far (ring position), across tiles
      for (int base = 0; base < total; base += ncta) {
        const int w = tc_snake(base, vcta, ncta);
        if (w >= total) continue;
        int m0, n0, kb0, kb1;
        const int pi = tc_locate<NCTA>(L, w, &m0, &n0, &kb0, &kb1);
        const TcGemmDesc* __restrict__ d = &descs[pi];
        const int bn = L.p[pi].bn, flags = L.p[pi].flags;
        const int nk1 = (L.p[pi].K + TC_BK - 1) / TC_BK;
        m0 += (int)rank * TC_BM; n0 += (int)rank * (bn / NCTA);   // this CTA's rows of A and of the B tile
        if (kb0 < nk1) { tc_prefetch_map(&d->mapA); tc_prefetch_map(&d->mapB); }
        const bool a_pre = !TS && (flags & TC_A_PRE);   // TS: the A lo half is built in registers, a pre-split plane is never loaded
        const uint32_t tx_bytes = (uint32_t)TC_A_TILE_BYTES * (a_pre ? 2u : 1u) + (uint32_t)(bn / NCTA) * TC_ROW_BYTES * ((flags & TC_B_PRE) ? 2u : 1u);
        // nothing to build (both lo planes pre-split): the landed bytes complete the MMA issuer's barrier directly, the builder warps
        // sit this tile out (a stage's wake -> arrive hop through them is latency on every k-block); both barriers still see one phase
        // per use so that the parities of a mixed group stay in step
        const bool direct = !TS && NCTA == 1 && a_pre && (flags & TC_B_PRE) && !(flags & TC_A_RELU);
        for (int kb = kb0; kb < kb1; ++kb, ++kbg) {
          const uint32_t s = kbg % (uint32_t)nst;
          if (kbg >= (uint32_t)nst) tc_mbar_wait(&bar_empty[s], ((kbg / (uint32_t)nst) - 1u) & 1u);
          uint64_t* land = direct ? &bar_ready[s] : &bar_raw[s];
          tc_mbar_expect_tx(land, tx_bytes);
          if (direct) {
            for (int e = 1; e < 4; ++e) tc_mbar_arrive(&bar_ready[s]);   // the builder warps' arrivals
            tc_mbar_arrive(&bar_raw[s]);
          }
          const uint32_t st = smem_base + s * stage_bytes;
          const bool second = kb >= nk1;
          const int kc = (second ? kb - nk1 : kb) * TC_BK;
          tc_tma_load_2d(st, second ? &d->mapA2 : &d->mapA, land, kc, m0);
          tc_tma_load_2d(st + b_off, second ? &d->mapB2 : &d->mapB, land, kc, n0);
          if (a_pre) tc_tma_load_2d(st + half_bytes, second ? &d->mapA2lo : &d->mapAlo, land, kc, m0);
          if (flags & TC_B_PRE) tc_tma_load_2d(st + blo_off, second ? &d->mapB2lo : &d->mapBlo, land, kc, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (the pair's leader) =====
      uint32_t kbg = 0, itc = 0;
      for (int base = 0; base < total; base += ncta) {
        const int w = tc_snake(base, vcta, ncta);
        if (w >= total) continue;
        const uint32_t it = itc++;
        int m0, n0, kb0, kb1;
        const int pi = tc_locate<NCTA>(L, w, &m0, &n0, &kb0, &kb1);
        const int bn = L.p[pi].bn;
        const int nk = kb1 - kb0;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((TC_BM * NCTA) >> 4) << 24);
        auto mma_ss = [](uint32_t d_, uint64_t a_, uint64_t b_, uint32_t i_, uint32_t acc_) {
          if constexpr (NCTA == 2) tc_mma_tf32_2cta(d_, a_, b_, i_, acc_); else tc_mma_tf32(d_, a_, b_, i_, acc_);
        };
        auto mma_ts = [](uint32_t d_, uint32_t a_, uint64_t b_, uint32_t i_, uint32_t acc_) {
          if constexpr (NCTA == 2) tc_mma_tf32_ts_2cta(d_, a_, b_, i_, acc_); else tc_mma_tf32_ts(d_, a_, b_, i_, acc_);
        };
        auto commit = [](uint64_t* bar) { if constexpr (NCTA == 2) tc_mma_commit_2cta(bar); else tc_mma_commit(bar); };
        const int nchain = (L.p[pi].flags & TC_DBG_ONECHAIN) ? 1 : 3;
        if (it > 0) {   // the previous tile's accumulators must have been read out
          if constexpr (NCTA == 2) tc_mbar_wait_cluster(&bar_tmem_empty, (it - 1u) & 1u); else
          tc_mbar_wait(&bar_tmem_empty, (it - 1u) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kb = 0; kb < nk; ++kb, ++kbg) {
          const uint32_t s = kbg % (uint32_t)nst;
          if constexpr (NCTA == 2) tc_mbar_wait_cluster(&bar_ready[s], (kbg / (uint32_t)nst) & 1u); else
          tc_mbar_wait(&bar_ready[s], (kbg / (uint32_t)nst) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_base + s * stage_bytes;
          const int k32 = kb / TC_KB_PER_32;
          if constexpr (TS) {
            // accumulators: columns [j*bn, (j+1)*bn): j = (32-float k-step) % n_hh_max for hi.hi, j = n_hh_max for the two correction chains
            const uint32_t a_hi = tmem_base + slot_col0 + (uint32_t)(kbg % (uint32_t)nslots) * TC_A_SLOT_COLS, a_lo = a_hi + 32u;
            const uint32_t b_raw = st + b_off, b_lo = st + blo_off;
            const uint32_t acc_hh = tmem_base + (uint32_t)(k32 % n_hh_max) * (uint32_t)bn, acc_c = tmem_base + (uint32_t)n_hh_max * (uint32_t)bn;
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks)
              mma_ts(acc_hh, a_hi + ks * 8u, tc_umma_desc(b_raw + ks * 32u), idesc, (ks == 0 && k32 < n_hh_max) ? 0u : 1u);
            if (nchain > 1) {
#pragma unroll
              for (int ks = 0; ks < TC_BK / 8; ++ks)
                mma_ts(acc_c, a_lo + ks * 8u, tc_umma_desc(b_raw + ks * 32u), idesc, (ks == 0 && kb == 0) ? 0u : 1u);
#pragma unroll
              for (int ks = 0; ks < TC_BK / 8; ++ks)
                mma_ts(acc_c, a_hi + ks * 8u, tc_umma_desc(b_lo + ks * 32u), idesc, 1u);
            }
            commit(&bar_afree[kbg % (uint32_t)nslots]);
          } else {
#pragma unroll
          for (int chain = 0; chain < 3; ++chain) {   // (A raw, B raw), (A lo, B raw), (A raw, B lo)
            if (chain >= nchain) break;
            const uint32_t a = st + (chain == 1 ? half_bytes : 0u);
            const uint32_t b = st + b_off + (chain == 2 ? half_bytes : 0u);
            // accumulators: columns [j*bn, (j+1)*bn): j = (32-float k-step) % 3 for hi.hi, j = 3 for the two correction chains
            const uint32_t acc = tmem_base + (uint32_t)(chain == 0 ? (k32 % 3) : 3) * (uint32_t)bn;
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {
              const bool first = (ks == 0) && (chain == 0 ? (k32 < 3 && kb % TC_KB_PER_32 == 0) : (kb == 0 && chain == 1));
              mma_ss(acc, tc_umma_desc(a + ks * 32u), tc_umma_desc(b + ks * 32u), idesc, first ? 0u : 1u);
            }
          }
          }
          commit(&bar_empty[s]);
        }
        commit(&bar_accum);
      }
    }
  } else if (warp < 2 + 4 * EG) {
    // ===== lo-part builders (group 0: 128 threads), then epilogue (every group) =====
    const int t = threadIdx.x - 64;
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int grp = (warp - 2) >> 2;        // 0: builders + epilogue, 1: epilogue helpers
    float* scr = epi_scratch + (grp * 4 + q) * TC_EPI_SCR;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t kbg = 0, itc = 0;
    for (int base = 0; base < total; base += ncta) {
      const int w = tc_snake(base, vcta, ncta);
      if (w >= total) continue;
      const uint32_t it = itc++;
      int m0, n0, kb0, kb1;
      const int pi = tc_locate<NCTA>(L, w, &m0, &n0, &kb0, &kb1);
      const TcGemmDesc* __restrict__ d = &descs[pi];
      const int bn = L.p[pi].bn, flags = L.p[pi].flags;
      const int nk = kb1 - kb0;
      const int bnl = bn / NCTA;          // rows of this CTA's share of the B tile
      m0 += (int)rank * TC_BM;            // this CTA's 128 rows of the (pair) tile
      // epilogue pointers: loaded now, used after the mainloop (their latency hides behind it)
      const int M = L.p[pi].M, N = L.p[pi].N, ldc = d->ldc, ldmask = d->ldmask;
      const float* __restrict__ bias = d->bias;
      const float* __restrict__ mask = d->mask;
      float* __restrict__ C = d->C;
      float* __restrict__ CT = d->CT;
      float* __restrict__ CT_lo = d->CT_lo;
      const int ldct = d->ldct;
      const bool build_a = !TS && !(flags & (TC_A_PRE | TC_DBG_NOBUILD)), build_b = !(flags & (TC_B_PRE | TC_DBG_NOBUILD));
      const bool direct = !TS && NCTA == 1 && (flags & TC_A_PRE) && (flags & TC_B_PRE) && !(flags & TC_A_RELU);   // see the producer
      if (direct) kbg += (uint32_t)nk;
      for (int kb = 0; kb < ((grp == 0 && !direct) ? nk : 0); ++kb, ++kbg) {   // (the helper groups only take part in the epilogue)
        const uint32_t s = kbg % (uint32_t)nst;
        tc_mbar_wait(&bar_raw[s], (kbg / (uint32_t)nst) & 1u);
        const float4* raw = reinterpret_cast<const float4*>(smem_gen + (size_t)s * stage_bytes);
        float4* lo = reinterpret_cast<float4*>(smem_gen + (size_t)s * stage_bytes + half_bytes);   // SS: [A lo | B lo]
        if constexpr (TS) {
          // this thread's row of the A tile: eight 16-byte chunks, chunk c of row r stored at c ^ (r & 7) (SWIZZLE_128B); a
          // quarter-warp reads eight different rows, i.e. eight different physical chunks: conflict-free
          const int r = q * 32 + lane;
          const uint8_t* arow = smem_gen + (size_t)s * stage_bytes + (size_t)r * TC_ROW_BYTES;
          uint32_t hi[32], lw[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 x = *reinterpret_cast<const float4*>(arow + ((c ^ (r & 7)) << 4));
            if (flags & TC_A_RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t h = __float_as_uint(xv[j]) & 0xFFFFE000u;
              hi[c * 4 + j] = h;
              lw[c * 4 + j] = (flags & TC_DBG_NOBUILD) ? 0u : __float_as_uint(xv[j] - __uint_as_float(h));
            }
          }
          if (build_b) {                                               // B lo tile next to the raw one: bn rows x TC_ROW_BYTES
            const float4* braw = reinterpret_cast<const float4*>(smem_gen + (size_t)s * stage_bytes + b_off);
            float4* blo = reinterpret_cast<float4*>(smem_gen + (size_t)s * stage_bytes + blo_off);
            if (bnl == 128) tc_build_lo<8>(braw, blo, t);
            else if (bnl == 64) tc_build_lo<4>(braw, blo, t);
            else if (bnl == 32) tc_build_lo<2>(braw, blo, t);
            else tc_build_lo<1>(braw, blo, t);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (the MMA's reads)
          }
          const uint32_t slot = kbg % (uint32_t)nslots;
          if (kbg >= (uint32_t)nslots) {   // the MMAs that read this slot's previous k-block have retired
            tc_mbar_wait(&bar_afree[slot], ((kbg / (uint32_t)nslots) - 1u) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const uint32_t a_hi = lane_addr + slot_col0 + slot * TC_A_SLOT_COLS;
          tc_tmem_st32(a_hi, hi);
          tc_tmem_st32(a_hi + 32u, lw);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();   // ONE arrival per warp (128 arrivals per stage on one barrier serialise; across the pair they cost ~25 clk each)
          if (lane == 0) { if constexpr (NCTA == 2) tc_mbar_arrive_remote(&bar_ready[s], 0); else tc_mbar_arrive(&bar_ready[s]); }
        } else {
        if (build_a) {                                                 // A tile: 128 rows x TC_ROW_BYTES = 32 * TC_BK float4
          if (flags & TC_A_RELU) tc_build_lo_relu<TC_BK / 4>(const_cast<float4*>(raw), lo, t);
          else tc_build_lo<TC_BK / 4>(raw, lo, t);
        }
        if (build_b) {                                                 // B tile: bn rows x TC_ROW_BYTES
          constexpr int B_OFF4 = TC_A_TILE_BYTES / 16;
          if (bnl == 128) tc_build_lo<TC_BK / 4>(raw + B_OFF4, lo + B_OFF4, t);
          else if (bnl == 64) tc_build_lo<TC_BK / 8>(raw + B_OFF4, lo + B_OFF4, t);
          else if (bnl * TC_BK == 32 * 32) tc_build_lo<2>(raw + B_OFF4, lo + B_OFF4, t);
          else tc_build_lo<1>(raw + B_OFF4, lo + B_OFF4, t);
        }
        // generic-proxy stores -> async proxy (the MMAs' operand reads).  The .shared::cta form also for a pair: the unqualified fence
        // compiles to MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC, and what the other SM's MMA reads is this CTA's shared memory
        if (build_a || build_b) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { if constexpr (NCTA == 2) tc_mbar_arrive_remote(&bar_ready[s], 0); else tc_mbar_arrive(&bar_ready[s]); }
        }
      }
      tc_mbar_wait(&bar_accum, it & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      const bool c_vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
      const bool m_vec = ((ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15u) == 0);
      const bool b_vec = (reinterpret_cast<uintptr_t>(bias) & 15u) == 0;
      const int n_hh = min(n_hh_max, (nk + TC_KB_PER_32 - 1) / TC_KB_PER_32);   // hi.hi accumulators that were written (the corrections follow the n_hh_max)
      const int sub = lane >> 2, c4 = (lane & 3) * 4;   // chunk = 32 rows x 16 columns: lane = 4 consecutive columns of 4 different rows
      const bool split = L.p[pi].splitk > 1;   // partial sums: added into the zeroed C; the bias rides on the first k-range
      if (split && kb0 > 0) bias = nullptr;
      for (int cb = grp * TC_EPI_CW; cb < ((flags & TC_DBG_NOEPI) ? 0 : bn); cb += TC_EPI_CW * EG) {   // the groups interleave the chunks
        if (n0 + cb >= N) break;          // warp-uniform: nothing of this chunk is inside the matrix
        // the saved activations this thread's 4 output float4s are masked with: issued first, so that their L2 latency
        // hides behind the TMEM reads and the transpose (a load -> use -> store chain per row serialises the round trips)
        const int col = n0 + cb + c4;
        const bool full = col + 3 < N;
        float4 mk4[4];
        if (flags & (GF_MASK_RELU | GF_MASK_TANH)) {
#pragma unroll
          for (int r4 = 0; r4 < 4; ++r4) {
            const int row = m0 + q * 32 + r4 * 8 + sub;
            mk4[r4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < M && col < N) {
              const float* mp = mask + (size_t)row * ldmask + col;
              if (m_vec && full) mk4[r4] = *reinterpret_cast<const float4*>(mp);
              else {
                mk4[r4].x = mp[0];
                if (col + 1 < N) mk4[r4].y = mp[1];
                if (col + 2 < N) mk4[r4].z = mp[2];
                if (col + 3 < N) mk4[r4].w = mp[3];
              }
            }
          }
        }
        float bz[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias && col < N) {
          if (b_vec && full) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(bias + col)); bz[0] = t4.x; bz[1] = t4.y; bz[2] = t4.z; bz[3] = t4.w; }
          else { for (int j = 0; j < 4; ++j) if (col + j < N) bz[j] = __ldg(bias + col + j); }
        }
        float v[16], u[16];
        tc_tmem_ld16(lane_addr + (uint32_t)(n_hh_max * bn + cb), v);   // correction chains first (smallest terms)
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int a = n_hh - 1; a >= 0; --a) {
          tc_tmem_ld16(lane_addr + (uint32_t)(a * bn + cb), u);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += u[j];
        }
        // thread = one row of the chunk -> scratch -> thread = 4 consecutive columns of 4 different rows
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(scr + lane * TC_EPI_LD + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        float4 xf[4];   // final values (kept for the transposed copy)
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          const int r = r4 * 8 + sub;
          const int row = m0 + q * 32 + r;
          const float4 x4 = *reinterpret_cast<const float4*>(scr + r * TC_EPI_LD + c4);
          xf[r4] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row >= M || col >= N) continue;
          float x[4] = {x4.x + bz[0], x4.y + bz[1], x4.z + bz[2], x4.w + bz[3]};
          if (flags & GF_RELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (flags & (GF_MASK_RELU | GF_MASK_TANH)) {
            const float mk[4] = {mk4[r4].x, mk4[r4].y, mk4[r4].z, mk4[r4].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (flags & GF_MASK_RELU) x[j] = (mk[j] > 0.f) ? x[j] : 0.f;
              if (flags & GF_MASK_TANH) x[j] *= (1.f - mk[j] * mk[j]);
            }
          }
          float* cp = C + (size_t)row * ldc + col;
          if (split) {
            if (c_vec && full) atomicAdd(reinterpret_cast<float4*>(cp), make_float4(x[0], x[1], x[2], x[3]));
            else { for (int j = 0; j < 4; ++j) if (col + j < N) atomicAdd(cp + j, x[j]); }
          } else if (c_vec && full) *reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
          else { for (int j = 0; j < 4; ++j) if (col + j < N) cp[j] = x[j]; }
          xf[r4] = make_float4(x[0], x[1], x[2], x[3]);
        }
        if (CT) {
          // transposed copy: final values back into the scratch, column-major with an odd pitch (conflict-free both ways), then
          // lane = (column j of 4, row group g of 8) writes 4 consecutive rows of one column = 16 contiguous bytes of C^T,
          // 8 lanes = one 128-byte run
          __syncwarp();
#pragma unroll
          for (int r4 = 0; r4 < 4; ++r4) {
            const int r = r4 * 8 + sub;
            scr[(c4 + 0) * 33 + r] = xf[r4].x; scr[(c4 + 1) * 33 + r] = xf[r4].y;
            scr[(c4 + 2) * 33 + r] = xf[r4].z; scr[(c4 + 3) * 33 + r] = xf[r4].w;
          }
          __syncwarp();
          const int g = lane & 7, jj = lane >> 3;
          const int rowT = m0 + q * 32 + 4 * g;
#pragma unroll
          for (int c4i = 0; c4i < 4; ++c4i) {
            const int cl = c4i * 4 + jj, colT = n0 + cb + cl;
            if (colT < N && rowT < M) {
              const float4 t4 = make_float4(scr[cl * 33 + 4 * g], scr[cl * 33 + 4 * g + 1], scr[cl * 33 + 4 * g + 2], scr[cl * 33 + 4 * g + 3]);
              *reinterpret_cast<float4*>(CT + (size_t)colT * ldct + rowT) = t4;
              if (CT_lo) *reinterpret_cast<float4*>(CT_lo + (size_t)colT * ldct + rowT) = tc_lo4(t4);
            }
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) { if constexpr (NCTA == 2) tc_mbar_arrive_remote(&bar_tmem_empty, 0); else tc_mbar_arrive(&bar_tmem_empty); }
    }
  }
}

// ring_bn: B-tile rows the ring geometry is laid out for (>= every problem's bn); total: tiles of the whole group
template <bool TS, int NCTA>
__global__ void __launch_bounds__(TC_THREADS, 1) k_gemm_tc(const TcGemmDesc* __restrict__ descs, const __grid_constant__ TcLaunch L) {
  static_assert(!TS || TC_BK == 32, "the TS form reads SWIZZLE_128B rows");
  fb_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  __shared__ __align__(16) float epi_scratch[4 * TC_EPI_GROUPS * TC_EPI_SCR];   // one region per epilogue warp
  __shared__ __align__(8) TcShared sh;

  const int warp = threadIdx.x >> 5;
  const uint32_t smem_base = (tc_smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = tc_smem_raw + (smem_base - tc_smem_u32(tc_smem_raw));
  const uint32_t rank = NCTA == 2 ? tc_cluster_rank() : 0u;

  if (threadIdx.x == 0) tc_init_barriers(&sh, false, NCTA, TC_EPI_GROUPS);
  if (warp == 1) {   // (a pair allocates the same columns in both CTAs: the same warp of each issues the cta_group::2 form)
    if constexpr (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&sh.tmem_base)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&sh.tmem_base)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (warp == 0 && (threadIdx.x & 31) == 0 && (int)(blockIdx.x / NCTA) < L.total) {
    // the tensor maps of this CTA's first work item: written at bind time, not by the previous kernel, so their fetch (~1 us on the
    // first TMA of a launch) may overlap that kernel's tail
    int m0, n0, kb0, kb1;
    const int pi = tc_locate<NCTA>(L, tc_snake(0, (int)(blockIdx.x / NCTA), (int)(gridDim.x / NCTA)), &m0, &n0, &kb0, &kb1);
    tc_prefetch_map(&descs[pi].mapA); tc_prefetch_map(&descs[pi].mapB);
    if (L.p[pi].flags & TC_B_PRE) tc_prefetch_map(&descs[pi].mapBlo);
    if (!TS && (L.p[pi].flags & TC_A_PRE)) tc_prefetch_map(&descs[pi].mapAlo);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (NCTA == 2) tc_cluster_sync();   // the partner's barriers are initialised before anything arrives on them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  fb_pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

  tc_gemm_roles<TS, NCTA, TC_EPI_GROUPS>(descs, L, (int)(blockIdx.x / NCTA), (int)(gridDim.x / NCTA), &sh, epi_scratch, smem_base, smem_gen, rank);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (NCTA == 2) tc_cluster_sync();   // neither CTA leaves (or frees tensor memory) while the pair's MMAs / commits can still touch it
  else __syncthreads();
  if (warp == 1) {
    if constexpr (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(sh.tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh.tmem_base), "r"(512) : "memory");
  }
}

// operand staging for the tensor-core GEMM: out = in^T (transpose = 1) or an aligned copy of in (transpose = 0), with a
// 16-byte aligned base and leading dimension so that TMA can address it; out_lo (optional) receives the lo plane
// x - trunc_tf32(x) of the same elements (out may then be null: the source itself is TMA-addressable and only its lo plane
// is wanted).  transpose = 2: zero-fill out (rows x cols).  32 x 32 tiles through shared memory.
struct TransposeDesc { const float* in; float* out; float* out_lo; int rows, cols, ld_in, ld_out, transpose, cta_begin, ctas_x, relu; };   // relu: the source is a lazily-ReLU'd pre-activation

__device__ __forceinline__ float tc_lo1(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// tile_smem: 32 x 33 floats of shared memory
__device__ __forceinline__ void transpose_grouped_body(const TransposeDesc* descs, int nprob, float* tile_smem, const int bid);
__global__ void __launch_bounds__(256) k_transpose_grouped(const TransposeDesc* __restrict__ descs, int nprob) {
  fb_pdl_trigger();
  fb_pdl_wait();
  __shared__ float tile_smem[32 * 33];
  transpose_grouped_body(descs, nprob, tile_smem, blockIdx.x);
}
// the same with the descriptors by value (the late staging launches on the main lane: no dependent descriptor loads)
__global__ void __launch_bounds__(256) k_transpose_grouped_tab(const __grid_constant__ DescTable<TransposeDesc, 8> T) {
  fb_pdl_trigger();
  fb_pdl_wait();
  __shared__ float tile_smem[32 * 33];
  transpose_grouped_body(T.d, T.n, tile_smem, blockIdx.x);
}
__device__ __forceinline__ void transpose_grouped_body(const TransposeDesc* descs, int nprob, float* tile_smem, const int bid) {
  float (*tile)[33] = reinterpret_cast<float (*)[33]>(tile_smem);
  int p = 0;
  while (p + 1 < nprob && descs[p + 1].cta_begin <= bid) ++p;
  const TransposeDesc d = descs[p];
  const int local = bid - d.cta_begin;
  const int bx = local % d.ctas_x, by = local / d.ctas_x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (d.transpose == 2) {   // zero fill (outputs of split-K GEMMs)
    for (int j = ty; j < 32; j += 8) {
      const int r = by * 32 + j, c = bx * 32 + tx;
      if (r < d.rows && c < d.cols) d.out[(size_t)r * d.ld_out + c] = 0.f;
    }
    return;
  }
  if (!d.transpose) {
    for (int j = ty; j < 32; j += 8) {
      const int r = by * 32 + j, c = bx * 32 + tx;
      if (r < d.rows && c < d.cols) {
        float x = d.in[(size_t)r * d.ld_in + c];
        if (d.relu) x = fmaxf(x, 0.f);
        if (d.out) d.out[(size_t)r * d.ld_out + c] = x;
        if (d.out_lo) d.out_lo[(size_t)r * d.ld_out + c] = tc_lo1(x);
      }
    }
    return;
  }
  for (int j = ty; j < 32; j += 8) {
    const int r = by * 32 + j, c = bx * 32 + tx;
    float x = (r < d.rows && c < d.cols) ? d.in[(size_t)r * d.ld_in + c] : 0.f;
    if (d.relu) x = fmaxf(x, 0.f);
    tile[j][tx] = x;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = bx * 32 + j, r = by * 32 + tx;
    if (c < d.cols && r < d.rows) {
      const float x = tile[tx][j];
      d.out[(size_t)c * d.ld_out + r] = x;
      if (d.out_lo) d.out_lo[(size_t)c * d.ld_out + r] = tc_lo1(x);
    }
  }
}
