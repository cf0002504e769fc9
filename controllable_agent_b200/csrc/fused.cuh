// Fused MLP-stack kernel: the launch plan of a run of step phases executed by ONE persistent kernel.
//
// The per-layer plan (plan.cuh) is a dependency chain of ~60 launches per gradient step on the caller's stream plus side / staging
// lanes; at batch 1024 a third of the step, and at the per-rank batch of 8 GPUs (128 rows) most of it, is launch latency, tail
// drain and pipeline fill between dependent kernels (profiles/r1d_step_table_cuda_events.txt).  k_fused_stack
// removes the kernel boundaries: one CTA per SM stays resident for a whole segment of the plan (the forward stacks of F, B and the
// actor; the backward stacks; ...) and walks a PROGRAM of stages.  A stage is the set of plan launches that may run concurrently
// (one main-lane launch plus whatever the side and staging lanes had in flight next to it); its work items — 128 x BN GEMM tiles
// of the tcgen05 grouped GEMM, LayerNorm rows, transpose tiles, column sums, the small elementwise kernels — are dealt round-robin
// to the CTAs, and a device-side grid barrier (one atomic per CTA, ~1-2 us) separates it from the next stage where a kernel
// boundary (launch + drain + prologue + pipeline fill, ~8-10 us) used to be.  The device code of every item is the body of the
// stand-alone kernel of the same name, so the fused and the per-launch paths compute the same values.
//
// Memory ordering between stages: writers' generic stores -> fence.proxy.async (TMA reads them through the async proxy) ->
// __syncthreads -> thread 0: __threadfence + atomicAdd (release) ... spin on an acquire load -> __threadfence -> __syncthreads ->
// fence.proxy.async.  No body reads in-kernel-produced data through the non-coherent path (__ldg is kept for parameters only).
#pragma once
#include "gemm_tc.cuh"
#include "kernels.cuh"

#define FS_THREADS 256
#define FS_MAX_STAGES 96
#define FS_MAX_ITEMS 224
#define FS_ARG_WORDS 256              // 1 KB of per-item arguments staged in shared memory
#define FS_NUM_BARRIERS 64
#define FS_SPIN_TIMEOUT_CYCLES (6000000000ll)   // ~3 s at 1.9 GHz: a lost CTA traps instead of hanging the device

enum { FS_NONE = 0, FS_GEMM_TC, FS_LN_FWD, FS_LN_BWD, FS_TRANSPOSE, FS_COLSUM, FS_L2_FWD, FS_L2_BWD, FS_STAGE_INPUTS, FS_Z_FINAL,
       FS_ACTOR_OUT, FS_ACTOR_Q };

struct FsItem { int type, count, arg_off, arg_bytes; };   // count: virtual blocks (GEMM: work items); arg_off: bytes from the program base
struct FsHeader { int n_stages, n_items, items_off, pad; int first_item[FS_MAX_STAGES + 1]; };

struct FsGemmArgs { const TcGemmDesc* descs; long long pad; TcLaunch hdr; };
struct FsLnFwdArgs { DescTable<LnDesc, 8> tab; int rows, vec, pad0, pad1; };
struct FsLnBwdArgs { DescTable<LnBwdDesc, 4> tab; int vec, pad0, pad1, pad2; };
struct FsPtrArgs { const void* descs; int n, extra; };
struct FsL2BwdArgs { const float* dy0; const float* dy1; const float* dy2; float* dsum; const float* y; const float* nrm; float* dx;
                     int lddy, ldsum, ldy, lddx, rows, Z, normalize; float coef; };
struct FsStageInArgs { StageParams P; const float* packed; };
struct FsZFinalArgs { ZFinalParams P; };
struct FsActorOutArgs { ActorOutParams P; const DevScalars* sc; };
struct FsActorQArgs { const float* F1; const float* F2; const float* z; float* dF1; float* dF2; double* acc; int ldf, ldz, lddf, rows, Z; float inv_n; };

static_assert(sizeof(FsGemmArgs) <= FS_ARG_WORDS * 4 && sizeof(FsLnFwdArgs) <= FS_ARG_WORDS * 4 && sizeof(FsLnBwdArgs) <= FS_ARG_WORDS * 4 &&
              sizeof(FsStageInArgs) <= FS_ARG_WORDS * 4 && sizeof(FsZFinalArgs) <= FS_ARG_WORDS * 4 && sizeof(FsActorOutArgs) <= FS_ARG_WORDS * 4,
              "item arguments must fit the shared staging block");

__device__ __forceinline__ unsigned long long fs_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the launch arrive; the counter only ever grows (never reset: a launch whose grid is G CTAs moves it by a multiple of G),
// so it needs no initialisation between launches as long as every launch on this counter uses the same grid.
__device__ __forceinline__ void fs_grid_barrier(unsigned long long* counter, unsigned int G, unsigned int* err) {
  asm volatile("fence.proxy.async.global;" ::: "memory");   // this thread's global stores -> later TMA (async proxy) reads on other SMs
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long old = atomicAdd(counter, 1ull);
    const unsigned long long target = (old / G + 1ull) * G;
    const long long t0 = clock64();
    while (fs_ld_acquire(counter) < target) {
      if (clock64() - t0 > FS_SPIN_TIMEOUT_CYCLES) { *err = 0xDEAD0001u; __threadfence_system(); __trap(); }
    }
    __threadfence();
  }
  __syncthreads();
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

__device__ __forceinline__ unsigned long long fs_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// times (nullable): [FS_MAX_STAGES + 1] nanosecond stamps written by CTA 0 at the start of every stage and at the end (fb_fused_profile)
__global__ void __launch_bounds__(FS_THREADS, 1) k_fused_stack(const char* __restrict__ prog, unsigned long long* barrier, unsigned int* err,
                                                               unsigned long long* times) {
  fb_pdl_trigger();
  extern __shared__ __align__(1024) uint8_t fs_smem_raw[];
  __shared__ __align__(16) float epi_scratch[4 * TC_EPI_SCR];
  __shared__ __align__(8) TcShared sh;
  __shared__ __align__(16) FsHeader hdr;
  __shared__ __align__(16) FsItem items[FS_MAX_ITEMS];
  __shared__ __align__(16) uint32_t args[FS_ARG_WORDS];

  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned int G = gridDim.x;
  const uint32_t smem_base = (tc_smem_u32(fs_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = fs_smem_raw + (smem_base - tc_smem_u32(fs_smem_raw));
  float* scratch = reinterpret_cast<float*>(smem_gen);   // block-style items borrow the (idle) operand ring

  // the program is constant data: fetched while the previous kernel of the stream still drains (before fb_pdl_wait)
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(prog);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&hdr);
    for (int i = tid; i < (int)(sizeof(FsHeader) / 4); i += FS_THREADS) dst[i] = src[i];
  }
  if (tid == 0) tc_init_barriers(&sh, false);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&sh.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(prog + hdr.items_off);
    uint32_t* dst = reinterpret_cast<uint32_t*>(items);
    const int n = hdr.n_items * (int)(sizeof(FsItem) / 4);
    for (int i = tid; i < n; i += FS_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  fb_pdl_wait();   // results of the kernels before this one are visible from here on

  const int n_stages = hdr.n_stages;
  for (int st = 0; st < n_stages; ++st) {
    if (times && blockIdx.x == 0 && tid == 0) times[st] = fs_globaltimer();
    unsigned int rot = 0;   // items of a stage start on consecutive CTAs, so that small items spread over the SMs the GEMM tiles leave idle
    for (int ii = hdr.first_item[st]; ii < hdr.first_item[st + 1]; ++ii) {
      const FsItem item = items[ii];
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes to the ring (scratch) before TMA refills it
      __syncthreads();          // the previous item is done with `args`, the ring and the GEMM barriers
      {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(prog + item.arg_off);
        for (int i = tid; i < item.arg_bytes / 4; i += FS_THREADS) args[i] = src[i];
      }
      if (item.type == FS_GEMM_TC && tid == 0) tc_init_barriers(&sh, true);
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int vcta = (int)((blockIdx.x + G - rot % G) % G);
      const int count = item.count, ng = (int)G;
      switch (item.type) {
        case FS_GEMM_TC: {
          const FsGemmArgs* a = reinterpret_cast<const FsGemmArgs*>(args);
          tc_gemm_roles<false, 1, 1>(a->descs, a->hdr, vcta, ng, &sh, epi_scratch, smem_base, smem_gen);
          break;
        }
        case FS_LN_FWD: {
          const FsLnFwdArgs* a = reinterpret_cast<const FsLnFwdArgs*>(args);
          if (a->vec) { for (int vb = vcta; vb < count; vb += ng) ln_tanh_fwd_v4_body(a->tab.d, a->tab.n, a->rows, vb); }
          else { for (int vb = vcta; vb < count; vb += ng) ln_tanh_fwd_body(a->tab.d, a->tab.n, a->rows, vb); }
          break;
        }
        case FS_LN_BWD: {
          const FsLnBwdArgs* a = reinterpret_cast<const FsLnBwdArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng) {
            if (a->vec) ln_tanh_bwd_v4_body(a->tab.d, a->tab.n, scratch, scratch + FB_MAX_LN_DIM, vb);
            else ln_tanh_bwd_body(a->tab.d, a->tab.n, scratch, scratch + FB_MAX_LN_DIM, vb);
            __syncthreads();   // the shared accumulators are reused by the next virtual block
          }
          break;
        }
        case FS_TRANSPOSE: {
          const FsPtrArgs* a = reinterpret_cast<const FsPtrArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng) {
            transpose_grouped_body(reinterpret_cast<const TransposeDesc*>(a->descs), a->n, scratch, vb);
            __syncthreads();
          }
          break;
        }
        case FS_COLSUM: {
          const FsPtrArgs* a = reinterpret_cast<const FsPtrArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng) {
            colsum_body(reinterpret_cast<const ColsumDesc*>(a->descs), a->n, scratch, vb);
            __syncthreads();
          }
          break;
        }
        case FS_L2_FWD: {
          const FsPtrArgs* a = reinterpret_cast<const FsPtrArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng) l2norm_fwd_body(reinterpret_cast<const L2Desc*>(a->descs), a->n, a->extra, vb);
          break;
        }
        case FS_L2_BWD: {
          const FsL2BwdArgs* a = reinterpret_cast<const FsL2BwdArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng)
            l2norm_bwd_body(a->dy0, a->dy1, a->dy2, a->lddy, a->coef, a->dsum, a->ldsum, a->y, a->ldy, a->nrm, a->dx, a->lddx, a->rows, a->Z,
                            a->normalize, vb);
          break;
        }
        case FS_STAGE_INPUTS: {
          const FsStageInArgs* a = reinterpret_cast<const FsStageInArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng) stage_inputs_body(a->P, a->packed, vb);
          break;
        }
        case FS_Z_FINAL: {
          const FsZFinalArgs* a = reinterpret_cast<const FsZFinalArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng) z_final_body(a->P, vb);
          break;
        }
        case FS_ACTOR_OUT: {
          const FsActorOutArgs* a = reinterpret_cast<const FsActorOutArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng) actor_out_body(a->P, a->sc, vb);
          break;
        }
        case FS_ACTOR_Q: {
          const FsActorQArgs* a = reinterpret_cast<const FsActorQArgs*>(args);
          for (int vb = vcta; vb < count; vb += ng)
            actor_q_body(a->F1, a->F2, a->ldf, a->z, a->ldz, a->dF1, a->dF2, a->lddf, a->rows, a->Z, a->inv_n, a->acc, vb);
          break;
        }
        default: break;
      }
      rot += (unsigned int)count;
    }
    if (st + 1 < n_stages) fs_grid_barrier(barrier, G, err);
  }
  if (times) {   // end stamp: after every CTA is done (one more barrier episode, profiling runs only)
    fs_grid_barrier(barrier, G, err);
    if (blockIdx.x == 0 && tid == 0) times[n_stages] = fs_globaltimer();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh.tmem_base), "r"(512) : "memory");
  }
}
