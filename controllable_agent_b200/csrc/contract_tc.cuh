// The batch x batch successor-measure contraction of update_fb on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   M_k = F_k . B^T,  tM = min_k(tF_k . tB^T),  Cov = B . B^T          (fb_ddpg.py:313-315, 320-321, 344)
//   loss sums + dL/dM_k, dL/dCov                                        (fb_ddpg.py:322-326, 345-347)
//
// are computed tile by tile without ever materialising the einsum outputs or the boolean off-diagonal mask: one CTA owns
// a 128 x 64 tile of the [rows x cols] pair space, accumulates up to five products for it in TMEM (5 x 64 fp32 columns),
// and its epilogue turns them straight into the gradient tiles G1, G2 (= dL/dM_k), Gc (= ortho_coef * dL/dCov) and the
// fp64 loss sums.
//
// Precision: kind::tf32 reads fp32 operands and drops the low 13 mantissa bits, which alone would put a ~1e-3 relative
// error on every product (the parity gate is 1e-3 on gradients).  Each operand is therefore split once per step
// (k_contract_split) into  x = hi + lo,  hi = x with the low 13 bits cleared (what the tensor core sees when given x),
// lo = x - hi (exact in fp32), and each product is issued as three MMA chains  hi.hi + lo.hi + hi.lo  ("3xTF32"),
// which restores ~2^-21 relative accuracy at 3x the tensor work of a contraction that is <2% of the step.
//
// Operand storage: X2[m] = [n_rows][2 * KP] fp32 with columns [0, KP) = x (zero padded from Z to KP = 32 * ceil(Z / 32)),
// [KP, 2 KP) = lo.  TMA (SWIZZLE_128B, boxes of 64 rows x 32 floats) stages them K-major in shared memory exactly in
// the canonical UMMA K-major SW128 layout (8-row x 128-byte atoms, SBO = 1024 B).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "kernels.cuh"

#define CT_TILE_M 128
#define CT_TILE_N 64
#define CT_MAX_PRODUCTS 5
#define CT_TMEM_COLS 512
#define CT_EPI_GROUPS 2
#define CT_THREADS (64 + 128 * CT_EPI_GROUPS)   // warp 0: TMA producer, warp 1: MMA issuer, then CT_EPI_GROUPS x 4 epilogue warps

enum { CT_MODE_ROW = 0, CT_MODE_COL = 1 };
enum { CT_F1 = 0, CT_F2, CT_TF1, CT_TF2, CT_B, CT_TB, CT_NUM_OPERANDS };

struct __align__(64) ContractParams {
  CUtensorMap maps[CT_NUM_OPERANDS];  // one per split operand, box = 32 floats x 64 rows
  int prod_a[CT_MAX_PRODUCTS], prod_b[CT_MAX_PRODUCTS];  // operand ids of each product; A rows = tile rows, B rows = tile cols
  int n_products, mode;
  int nr, nc;              // rows / cols of the pair space handled by this launch
  int a_row0;              // first row of the A-side operands inside their arrays (row_offset for ROW mode)
  int diag0;               // column index of the diagonal for row 0 (row_offset)
  int nbox;                // KP / 32
  int ksteps;              // ceil(Z / 8) MMA k-steps per chain
  int ld;                  // leading dimension of the G outputs
  float* G1; float* G2; float* Gc;   // ROW: [nr, ld]; COL: G1, G2 only
  float* Gt1; float* Gt2;            // ROW, optional (single GPU): transposed copies [nc, ld] so the COL launch is not needed
  const float* disc; int disc_stride;  // ROW: discount of local row i; COL: discount of global column j
  float inv_noff, inv_n, c4;           // 1/(n(n-1)), 1/n, 4 * ortho_coef / (n(n-1))
  double* acc;
};

__device__ __forceinline__ uint32_t ct_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ct_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ct_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ct_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ct_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(ct_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void ct_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   ct_smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(ct_smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start address, LBO = 0, SBO = 1024 B, version 1, layout type 2
__device__ __forceinline__ uint64_t ct_umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = 64
__device__ __forceinline__ uint32_t ct_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CT_TILE_N >> 3) << 17) | ((uint32_t)(CT_TILE_M >> 4) << 24);
}
__device__ __forceinline__ void ct_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ct_mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ct_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ct_tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// split the exchange block into tensor-core operands: out[m][row] = [x (Z, zero padded to KP) | lo (KP)]
__global__ void __launch_bounds__(256) k_contract_split(const float* __restrict__ blk, int blk_pitch, int ldz, int rows, int Z, int KP,
                                                        float* __restrict__ out) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = CT_NUM_OPERANDS * rows * KP;
  if (idx >= total) return;
  const int c = idx % KP;
  const int r = (idx / KP) % rows;
  const int m = idx / (KP * rows);
  float x = 0.f, lo = 0.f;
  if (c < Z) {
    x = blk[(size_t)r * blk_pitch + m * ldz + c];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
  }
  float* o = out + ((size_t)m * rows + r) * (2 * KP);
  o[c] = x;
  o[KP + c] = lo;
}

// Persistent, warp-specialised: min(tiles, SMs) CTAs walk the 128 x 64 tiles of the pair space column by column (consecutive CTAs
// share one column block of the B-side operands in L2).
//   warp 0      TMA producer   one K-box (32 floats) of one product per ring stage: [A raw | A lo | B raw | B lo] = 48 KB, 4 stages;
//                              runs ahead into the next tile while the current one is in its epilogue
//   warp 1      MMA issuer     3 chains x <= 4 tcgen05.mma (M128 x N64 x K8) per stage into the product's 64 TMEM columns
//   warps 2..9  epilogue       two groups of four warps (one per TMEM lane quadrant) that take alternate 16-column chunks (one warp per
//                              scheduler is instruction-latency-bound): tcgen05.ld of the five accumulators, loss math, G tiles out
//                              through a shared-memory transpose
// (The first version ran one tile per CTA with ONE thread as producer and issuer and whole products as stages: at z = 100 a single
// 192 KB stage, i.e. no overlap of loads and MMAs at all; at batch 4096 that was 15 % of the step.)
// dynamic smem: [4 stages x 48 KB | epilogue scratch, 32 x 20 floats per epilogue warp], 1024-byte aligned
#define CT_STAGES 4
#define CT_STAGE_BYTES (2 * CT_TILE_M * 128 + 2 * CT_TILE_N * 128)
#define CT_SCR_LD 20
#define CT_SCRATCH_BYTES (4 * CT_EPI_GROUPS * 32 * CT_SCR_LD * 4)   // one 32 x 16 chunk per epilogue warp (G1, G2, Gc pass through it in turn)
#define CT_SMEM_BYTES (CT_STAGES * CT_STAGE_BYTES + CT_SCRATCH_BYTES + 1024)

__global__ void __launch_bounds__(CT_THREADS, 1) k_contract_tc(const __grid_constant__ ContractParams P) {
  fb_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t ct_smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[CT_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[CT_STAGES];
  __shared__ __align__(8) uint64_t bar_accum;        // all MMAs of a tile retired -> epilogue
  __shared__ __align__(8) uint64_t bar_tmem_empty;   // epilogue has read the accumulators -> MMA issuer (next tile)
  __shared__ uint32_t tmem_base_smem;
  __shared__ double red[6][4 * CT_EPI_GROUPS];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ct_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (P.nr + CT_TILE_M - 1) / CT_TILE_M, tiles_n = (P.nc + CT_TILE_N - 1) / CT_TILE_N;
  const int ntiles = tiles_m * tiles_n;
  const int nbox = P.nbox, np = P.n_products;
  constexpr uint32_t A_BOX = CT_TILE_M * 128u, B_BOX = CT_TILE_N * 128u;   // bytes of one K-box of the A / B side (raw or lo)

  if (threadIdx.x == 0) {
    for (int s = 0; s < CT_STAGES; ++s) { ct_mbar_init(&bar_full[s], 1); ct_mbar_init(&bar_empty[s], 1); }
    ct_mbar_init(&bar_accum, 1);
    ct_mbar_init(&bar_tmem_empty, 4 * CT_EPI_GROUPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem_u32(&tmem_base_smem)), "r"(CT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  fb_pdl_wait();   // the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t kbg = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = (t % tiles_m) * CT_TILE_M, col0 = (t / tiles_m) * CT_TILE_N;
        const int KP = nbox * 32;
        for (int p = 0; p < np; ++p) {
          const CUtensorMap* ma = &P.maps[P.prod_a[p]];
          const CUtensorMap* mb = &P.maps[P.prod_b[p]];
          for (int b = 0; b < nbox; ++b, ++kbg) {
            const uint32_t s = kbg % CT_STAGES;
            if (kbg >= CT_STAGES) ct_mbar_wait(&bar_empty[s], ((kbg / CT_STAGES) - 1u) & 1u);
            uint8_t* st = smem + (size_t)s * CT_STAGE_BYTES;
            ct_mbar_expect_tx(&bar_full[s], CT_STAGE_BYTES);
            for (int part = 0; part < 2; ++part) {       // 0: raw, 1: lo
              const int kcol = part * KP + b * 32;
              uint8_t* da = st + part * A_BOX;
              ct_tma_load_2d(da, ma, &bar_full[s], kcol, P.a_row0 + row0);
              ct_tma_load_2d(da + 64 * 128, ma, &bar_full[s], kcol, P.a_row0 + row0 + 64);
              ct_tma_load_2d(st + 2 * A_BOX + part * B_BOX, mb, &bar_full[s], kcol, col0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = ct_idesc();
      uint32_t kbg = 0, it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        if (it > 0) {   // the previous tile's accumulators have been read out
          ct_mbar_wait(&bar_tmem_empty, (it - 1u) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int p = 0; p < np; ++p) {
          const uint32_t d_tmem = tmem_base + (uint32_t)p * CT_TILE_N;
          for (int b = 0; b < nbox; ++b, ++kbg) {
            const uint32_t s = kbg % CT_STAGES;
            ct_mbar_wait(&bar_full[s], (kbg / CT_STAGES) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sbase = ct_smem_u32(smem + (size_t)s * CT_STAGE_BYTES);
            const int nks = min(4, P.ksteps - 4 * b);   // 8-float k-steps of this box that hold columns < Z (the rest is zero padding)
            // chains: (A raw, B raw), (A lo, B raw), (A raw, B lo)
            for (int chain = 0; chain < 3; ++chain) {
              const uint32_t abase = sbase + (chain == 1 ? A_BOX : 0u);
              const uint32_t bbase = sbase + 2u * A_BOX + (chain == 2 ? B_BOX : 0u);
              for (int ks = 0; ks < nks; ++ks)
                ct_mma_tf32(d_tmem, ct_umma_desc(abase + (uint32_t)ks * 32u), ct_umma_desc(bbase + (uint32_t)ks * 32u), idesc,
                            (b | chain | ks) != 0 ? 1u : 0u);
            }
            ct_mma_commit(&bar_empty[s]);   // arrives when the MMAs above have finished reading this stage
          }
        }
        ct_mma_commit(&bar_accum);
      }
    }
  } else {
    // ===== epilogue: thread = one row of the tile (TMEM lane), 16 columns at a time =====
    const int q = warp & 3;   // TMEM lane quadrant this warp may read
    const int grp = (warp - 2) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* scr = reinterpret_cast<float*>(smem + (size_t)CT_STAGES * CT_STAGE_BYTES) + (grp * 4 + q) * (32 * CT_SCR_LD);
    double a_off = 0.0, a_diag = 0.0, a_cov = 0.0, a_covd = 0.0, a_tm = 0.0, a_m1 = 0.0;
    const float inv_noff = P.inv_noff, inv_n = P.inv_n, c4 = P.c4;
    const bool g_vec = (P.ld % 4 == 0) && (((reinterpret_cast<uintptr_t>(P.G1) | reinterpret_cast<uintptr_t>(P.G2) |
                                              reinterpret_cast<uintptr_t>(P.Gc ? P.Gc : P.G1)) & 15u) == 0);
    const int n_arr = P.mode == CT_MODE_ROW ? 3 : 2;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int row0 = (t % tiles_m) * CT_TILE_M, col0 = (t / tiles_m) * CT_TILE_N;
      const int row = row0 + q * 32 + lane;
      const bool row_ok = row < P.nr;
      const int diag_col = P.diag0 + row;
      float g_row = 0.f;
      if (P.mode == CT_MODE_ROW && row_ok) g_row = P.disc[(size_t)row * P.disc_stride];
      ct_mbar_wait(&bar_accum, it & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // G tiles leave through shared memory: a thread owns one ROW of the tile (its TMEM lane), so direct stores would put the 32 lanes
      // of every store instruction on 32 different rows (one 32-byte sector each for 4 useful bytes).  Each warp stages its 32 x 16
      // chunk of G1 / G2 / Gc and writes it back as 128-bit stores, 4 lanes per row: full sectors.  The transposed copies Gt1 / Gt2 are
      // coalesced as they are (for a fixed column the lanes are consecutive rows).
#pragma unroll 1
      for (int cb = grp * 16; cb < CT_TILE_N; cb += 16 * CT_EPI_GROUPS) {
        float m1[16], m2[16], t1[16], t2[16], cv[16];
        ct_tmem_ld16(lane_addr + 0 * CT_TILE_N + cb, m1);
        ct_tmem_ld16(lane_addr + 1 * CT_TILE_N + cb, m2);
        ct_tmem_ld16(lane_addr + 2 * CT_TILE_N + cb, t1);
        ct_tmem_ld16(lane_addr + 3 * CT_TILE_N + cb, t2);
        if (P.mode == CT_MODE_ROW) ct_tmem_ld16(lane_addr + 4 * CT_TILE_N + cb, cv);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cb + 16 * CT_EPI_GROUPS >= CT_TILE_N) {   // this warp's last chunk is read: the issuer may overwrite the accumulators with the next tile
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ct_smem_u32(&bar_tmem_empty)) : "memory");
        }
        float g1v[16], g2v[16], gcv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = col0 + cb + j;
          float g1 = 0.f, g2 = 0.f, gc = 0.f;
          if (row_ok && col < P.nc) {
            const float tm = fminf(t1[j], t2[j]);
            if (P.mode == CT_MODE_ROW) {
              a_tm += tm; a_m1 += m1[j];
              if (col != diag_col) {
                const float d1 = m1[j] - g_row * tm, d2 = m2[j] - g_row * tm;
                a_off += (double)d1 * d1 + (double)d2 * d2;
                a_cov += (double)cv[j] * cv[j];
                g1 = d1 * inv_noff; g2 = d2 * inv_noff; gc = c4 * cv[j];
              } else {
                a_diag += (double)m1[j] + (double)m2[j];
                a_covd += cv[j];
                g1 = -inv_n; g2 = -inv_n; gc = 0.f;
              }
              if (P.Gt1) {  // transposed copies: for a fixed column the 32 lanes of a warp write 32 consecutive floats
                const size_t ot = (size_t)col * P.ld + row;
                P.Gt1[ot] = g1; P.Gt2[ot] = g2;
              }
            } else {
              // COL mode: rows are the LOCAL columns t of the loss matrices, cols run over all global rows s
              if (col != diag_col) {
                const float g = __ldg(P.disc + (size_t)col * P.disc_stride);
                g1 = (m1[j] - g * tm) * inv_noff; g2 = (m2[j] - g * tm) * inv_noff;
              } else {
                g1 = -inv_n; g2 = -inv_n;
              }
            }
          }
          g1v[j] = g1; g2v[j] = g2; gcv[j] = gc;
        }
        const int rr = lane >> 2, c4o = (lane & 3) * 4;
        const int colo = col0 + cb + c4o;
        auto flush = [&](const float (&gv)[16], float* G) {   // one G array through the warp's 32 x 16 scratch
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(scr + lane * CT_SCR_LD + j) = make_float4(gv[j], gv[j + 1], gv[j + 2], gv[j + 3]);
          __syncwarp();
#pragma unroll
          for (int r8 = 0; r8 < 4; ++r8) {
            const int r = r8 * 8 + rr;
            const int grow = row0 + q * 32 + r;
            if (grow >= P.nr || colo >= P.nc) continue;
            const float4 v = *reinterpret_cast<const float4*>(scr + r * CT_SCR_LD + c4o);
            float* gp = G + (size_t)grow * P.ld + colo;
            if (g_vec && colo + 3 < P.nc) *reinterpret_cast<float4*>(gp) = v;
            else {
              gp[0] = v.x;
              if (colo + 1 < P.nc) gp[1] = v.y;
              if (colo + 2 < P.nc) gp[2] = v.z;
              if (colo + 3 < P.nc) gp[3] = v.w;
            }
          }
          __syncwarp();
        };
        flush(g1v, P.G1);
        flush(g2v, P.G2);
        if (n_arr == 3) flush(gcv, P.Gc);
      }
    }
    if (P.mode == CT_MODE_ROW) {   // the loss sums of every tile this CTA walked: one reduction, six atomics per CTA
      double vals[6] = {a_off, a_diag, a_cov, a_covd, a_tm, a_m1};
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double v = warp_sum_d(vals[i]);
        if (lane == 0) red[i][grp * 4 + q] = v;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (P.mode == CT_MODE_ROW && threadIdx.x < 6) {
    const int slot[6] = {ACC_OFFDIAG_SQ, ACC_DIAG, ACC_COV_OFF_SQ, ACC_COV_DIAG, ACC_TARGET_M, ACC_M1};
    double tot = 0.0;
    for (int w = 0; w < 4 * CT_EPI_GROUPS; ++w) tot += red[threadIdx.x][w];
    atomicAdd(P.acc + slot[threadIdx.x], tot);
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CT_TMEM_COLS) : "memory");
  }
}
