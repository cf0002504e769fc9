// libfb_b200.so — the FB-DDPG gradient step (url_benchmark/agent/fb_ddpg.py:427-520) on one B200, C ABI of
// include/fb_b200.h.  sm_100a only; no CPU route.
#include "plan.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)
#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

static int phase_index(uint32_t bit) { int i = 0; while ((1u << i) != bit) ++i; return i; }

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen: the library torch already loaded, or the path the host passes): no link-time dependency.
// Minimal declarations of the stable NCCL 2 ABI (nccl.h): ncclUniqueId is 128 opaque bytes passed by value, ncclFloat32 = 7,
// ncclSum = 0, results are 0 on success.
// ------------------------------------------------------------------------------------------------
#include <dlfcn.h>
#include <array>
struct FbNcclId { char internal[128]; };
struct FbNccl {
  void* lib = nullptr;
  int (*GetUniqueId)(FbNcclId*) = nullptr;
  int (*CommInitRank)(void**, int, FbNcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
};
static FbNccl g_nccl;
static int nccl_load(const char* path) {
  if (g_nccl.lib) return FB_OK;
  void* lib = nullptr;
  if (path && *path) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return FB_E_STATE;
  g_nccl.GetUniqueId = (int (*)(FbNcclId*))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, FbNcclId, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.AllReduce) return FB_E_STATE;
  g_nccl.lib = lib;
  return FB_OK;
}

// ------------------------------------------------------------------------------------------------
// TMA tensor maps for the tcgen05 contraction (driver entry point fetched through the runtime: no -lcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_tiled_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows, int box_cols = TC_BK) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return FB_E_STATE;
    fn = (EncodeTiledFn)p;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};   // inner extent = one swizzle atom row (64 or 128 bytes)
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FB_OK : FB_E_STATE;
}

static int encode_operand_map(CUtensorMap* map, float* base, int rows, int cols) { return encode_tiled_map(map, base, rows, cols, cols, 64, 32); }   // contraction: 32-float boxes, SWIZZLE_128B

// Every kernel of the plan is launched with the programmatic-serialization attribute: it may become resident while its
// predecessor in the stream still runs and waits in fb_pdl_wait() (first statement of every kernel; after the prologue in
// k_gemm_tc), which hides the launch latency of the ~60 dependent launches of a step.  FB_NO_PDL=1 launches them plainly.
template <typename... KArgs, typename... Args>
static void fb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  static const bool pdl = getenv("FB_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // errors surface through cudaGetLastError() at the call site
}

static int make_tc_launch(const std::vector<TcGemmDesc>& v, int total, int ring_bn, TcLaunch* out) {
  if (v.size() > TC_MAX_PROBS) return FB_E_UNSUPPORTED;
  memset(out, 0, sizeof(*out));
  out->nprob = (int)v.size(); out->total = total; out->ring_bn = ring_bn;
  for (size_t i = 0; i < v.size(); ++i) {
    const TcGemmDesc& d = v[i];
    out->p[i] = TcProb{d.M, d.N, d.K, d.K2, d.bn, d.flags, d.tiles_n, d.work_begin, d.splitk, d.kb_per_split};
  }
  return FB_OK;
}

// ------------------------------------------------------------------------------------------------
// the tensor-core GEMM kernels: CTA pairs (cta_group::2, clusters of two) or single CTAs, operands in shared memory unless
// FB_TC_TS=1 asks for the A-through-tensor-memory variant (A/B measurements, profiles/r2_gemm_tc_ts.txt)
typedef void (*TcKernel)(const TcGemmDesc*, const TcLaunch);
static TcKernel tc_kernel(int ncta) {
  static const bool ts = getenv("FB_TC_TS") != nullptr;
  if (ncta == 2) return ts ? (TcKernel)k_gemm_tc<true, 2> : (TcKernel)k_gemm_tc<false, 2>;
  return ts ? (TcKernel)k_gemm_tc<true, 1> : (TcKernel)k_gemm_tc<false, 1>;
}
static cudaError_t tc_set_smem_attr() {
  for (int ncta = 1; ncta <= 2; ++ncta) {
    cudaError_t e = cudaFuncSetAttribute(tc_kernel(ncta), cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
// CTA pairs that can be resident at once (persistent grid of a pair launch): 74 on a full B200 unless a GPC holds an odd number of SMs
static int tc_max_pairs() {
  static int pairs = 0;
  if (pairs) return pairs;
  int n = 0;
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(FB_SM_COUNT); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (tc_set_smem_attr() == cudaSuccess && cudaOccupancyMaxActiveClusters(&n, tc_kernel(2), &cfg) == cudaSuccess && n > 0) pairs = std::min(n, FB_SM_COUNT / 2);
  else { cudaGetLastError(); return FB_SM_COUNT / 2; }   // (no device yet: the sizing pass of fb_create)
  return pairs;
}
// one grouped launch: `work` items (tiles x k-ranges; pair tiles when ncta == 2) on a persistent grid
static cudaError_t tc_launch(const TcGemmDesc* dd, const TcLaunch& hdr, int work, int ncta, cudaStream_t s, bool pdl) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  const int units = ncta == 2 ? tc_max_pairs() : FB_SM_COUNT;
  cfg.gridDim = dim3((work < units ? work : units) * ncta); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {   // programmatic dependent launch: the kernel's prologue may overlap the tail of the launch before it (fb_pdl_wait inside)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na;
  }
  if (ncta == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, tc_kernel(ncta), dd, hdr);
}

// CTAs of the fused reduce-scatter + Adam + all-gather pass (one per SM; FB_P2P_ADAM_CTAS overrides).  Measured at 8 ranks
// (tools/gpu_p2p_adam_sweep.sh): 61.4 / 59.4 / 59.6 us for the fb segment with 296 / 148 / 74 CTAs, i.e. 12.9 MB in + 12.9 MB out per
// rank at ~220 GB/s each: the pass is bound by the rate of its peer LOADS (ld.volatile over NVLink), not by how the inbound and outbound
// transfers overlap.  What would change that is a push design (every rank stores its gradient slices into the owners' arenas as the
// backward produces them) or the switch's multimem.ld_reduce.
static int p2p_adam_ctas() {
  static const int n = getenv("FB_P2P_ADAM_CTAS") ? std::max(1, atoi(getenv("FB_P2P_ADAM_CTAS"))) : FB_SM_COUNT;
  return n;
}

// op recorders
// ------------------------------------------------------------------------------------------------
struct Builder {
  fb_handle* h;
  char* d_arena;
  int phase = 0;
  void set_phase(uint32_t bit) { phase = phase_index(bit); produced.clear(); }
  // outputs of tensor-core GEMMs of the current phase (direct stores, no split-K): a later dW product that needs such an output
  // transposed (K = batch contiguous) gets it from the producer's epilogue (TcGemmDesc::CT) instead of a staging launch
  struct Produced { const float* C; int M, N, ldc; size_t arena_off; };
  std::vector<Produced> produced;
  void invalidate_produced(const float* p) {   // another kernel rewrites the buffer in place
    for (size_t i = 0; i < produced.size();)
      if (p >= produced[i].C && p < produced[i].C + (size_t)produced[i].M * produced[i].ldc) produced.erase(produced.begin() + i); else ++i;
  }
  int cur_lane = 0;        // lane of the ops recorded next (1: a side-lane chain, e.g. the z-mixing forward)
  bool fork_next = false;  // the next side-lane op recorded starts a chain: it forks from the main lane
  void push(OpFn fn, int kind = FB_OPK_ELEMENTWISE, double flops = 0.0, double bytes = 0.0, int lane = -1) {
    int fork = 0;
    if (lane < 0) { lane = cur_lane; if (lane == 1 && fork_next) { fork = 1; fork_next = false; } }
    else if (lane == 1) fork = 1;   // a lone side-lane leaf (bias column sums) forks where it is issued
    h->ops[phase].push_back(Op{std::move(fn), kind, flops, bytes, lane, 0, fork});
  }

  int rc = FB_OK;  // first error met while building (tensor-map encoding)

  // can this problem run on the tensor cores (gemm_tc.cuh)?  Operands TMA cannot address directly are staged (see stage_operand)
  // Lazily-ReLU'd activations: a GF_RELU | GF_RELU_LAZY_OK product that the cost model wants to split along K stores its
  // PRE-activation (atomic partial sums cannot be clamped); every consumer applies the ReLU itself: a tensor-core A operand in
  // the builder warps (TC_A_RELU), a staged copy in the staging launch (TransposeDesc::relu), backward masks only test the
  // sign.  Whole-plan registry (a buffer produced in FB_FWD is consumed in FB_BWD).
  struct LazyRange { const float* p; size_t n; };
  std::vector<LazyRange> lazy;
  bool is_lazy(const float* p) const {
    for (auto& r : lazy) if (p >= r.p && p < r.p + r.n) return true;
    return false;
  }
  bool force_simt = false;   // inference plans: a handful of rows, the fp32 CUDA-core kernel
  bool tc_ok(const GemmDesc& g) const {
    if (force_simt) return false;
    if (h->cfg.mlp_mode != FB_MLP_TCGEN05 || (g.flags & GF_SHARED_C) || g.K < 4) return false;
    // a first-layer product (K = obs + action / z, tens of columns): its weight is staged anyway (pre-split lo plane, early on
    // the side lane); only an activation operand that would need a late staging launch keeps it on the SIMT kernel
    const bool a_direct = g.a_kmajor && aligned16(g.A) && g.lda % 4 == 0;
    if (g.K < 128 && g.a_kmajor && g.b_kmajor && !(a_direct || is_early(g.A))) return false;
    if (getenv("FB_TC_STRICT")) {   // debugging: the narrower rule of the first tensor-core plan
      const bool b_direct = g.b_kmajor && aligned16(g.B) && g.ldb % 4 == 0;
      if (g.M < 8 || (g.K < 128 && g.a_kmajor && g.b_kmajor && !(a_direct && b_direct))) return false;
    }
    return true;
  }
  // K-major, TMA-addressable view of an operand: the operand itself when it already is one, otherwise a staged copy
  // (transposed for mn-major operands) produced by the transpose launch that precedes the GEMM launch
  // is the source of this operand final when the current phase starts?  (weights and targets: always; workspace buffers
  // written by the forward phases: for the loss / backward phases)
  bool is_early(const float* p) const {
    const fb_buffers& bf = h->bufs;
    auto in = [&](const float* base, size_t n) { return base && p >= base && p < base + n; };
    if (in(bf.d_param_fb, h->seg_fb.size) || in(bf.d_target_fb, h->seg_fb.size) || in(bf.d_param_actor, h->seg_actor.size)) return true;
    const bool backward_phase = phase == phase_index(FB_PHASE_FB_LOSS) || phase == phase_index(FB_PHASE_FB_BWD) ||
                                phase == phase_index(FB_PHASE_ACTOR_BWD);
    const char* c = reinterpret_cast<const char*>(p);
    return backward_phase && c >= h->ws_base && c < h->ws_base + h->ws_fwd_end;
  }
  // Returns the K-major, TMA-addressable view of an operand (the operand itself when it already is one, otherwise a staged
  // copy).  *lo_out receives the operand's pre-split lo plane when one is produced: always for staged copies (the staging
  // launch writes it for free), and for directly addressable operands only if `want_lo` and the source is final when the
  // phase starts (weights: a lo-only staging entry on the side lane).
  // first phase (of a full step) at whose start `p` holds its final value for a consumer in the current phase
  int avail_of(const float* p) const {
    const fb_buffers& bf = h->bufs;
    auto in = [&](const float* base, size_t n) { return base && p >= base && p < base + n; };
    if (in(bf.d_param_actor, h->seg_actor.size)) return 0;                       // changes only in ACTOR_ADAM (last)
    if (in(bf.d_param_fb, h->seg_fb.size) || in(bf.d_target_fb, h->seg_fb.size))   // changes in FB_ADAM
      return phase <= phase_index(FB_PHASE_FB_ADAM) ? 0 : phase_index(FB_PHASE_ACTOR_FWD);
    // forward activations: MIX / FB_FWD outputs are final when FB_LOSS starts; ACTOR_FWD outputs when ACTOR_BWD starts
    if (phase == phase_index(FB_PHASE_FB_LOSS) || phase == phase_index(FB_PHASE_FB_BWD)) return phase_index(FB_PHASE_FB_LOSS);
    return phase;
  }
  void add_early(const TransposeDesc& t, int avail) {
    h->early_stage[phase].push_back(t);
    h->early_avail[phase].push_back(avail);
  }
  bool in_grad(const float* p) const {
    const fb_buffers& bf = h->bufs;
    return (bf.d_grad_fb && p >= bf.d_grad_fb && p < bf.d_grad_fb + h->seg_fb.size) ||
           (bf.d_grad_actor && p >= bf.d_grad_actor && p < bf.d_grad_actor + h->seg_actor.size);
  }
  const float* stage_operand(std::vector<TransposeDesc>& pending, bool* used_early, const float* p, int kmajor, int rows, int K, int ld,
                             bool want_lo, int* ld_out, const float** lo_out, int* ld_lo) {
    const bool direct = kmajor && aligned16(p) && ld % 4 == 0;
    const bool early = is_early(p);
    *lo_out = nullptr; *ld_lo = 0;
    if (direct && !(want_lo && early)) { *ld_out = ld; return p; }
    const int lds = fb_round_up(K, 4);
    if (!kmajor && !early && !getenv("FB_NO_CT")) {
      for (auto& pr : produced) {
        if (K != pr.M || ld != pr.ldc || p < pr.C || p + rows > pr.C + pr.N) continue;
        TcGemmDesc* hd = reinterpret_cast<TcGemmDesc*>(h->arena.data() + pr.arena_off);
        if (!hd->CT) {
          hd->ldct = fb_round_up(pr.M, 4);
          hd->CT = (float*)ws_alloc(h, (size_t)pr.N * hd->ldct * sizeof(float));
          hd->CT_lo = (float*)ws_alloc(h, (size_t)pr.N * hd->ldct * sizeof(float));
        }
        const size_t c0 = (size_t)(p - pr.C);
        *ld_out = hd->ldct; *ld_lo = hd->ldct; *lo_out = hd->CT_lo + c0 * hd->ldct;
        return hd->CT + c0 * hd->ldct;
      }
    }
    TransposeDesc t; memset(&t, 0, sizeof(t));
    t.in = p; t.ld_in = ld; t.ld_out = lds; t.transpose = kmajor ? 0 : 1;
    t.relu = is_lazy(p) ? 1 : 0;
    if (kmajor) { t.rows = rows; t.cols = K; } else { t.rows = K; t.cols = rows; }   // mn-major storage is [K][rows]
    std::vector<TransposeDesc>& list = early ? h->early_stage[phase] : pending;
    if (early) *used_early = true;
    for (auto& q : list)
      if (q.in == t.in && q.rows == t.rows && q.cols == t.cols && q.ld_in == t.ld_in && q.transpose == t.transpose) {
        *lo_out = q.out_lo; *ld_lo = q.ld_out;
        if (q.out) { *ld_out = q.ld_out; return q.out; }
        *ld_out = ld; return p;
      }
    const size_t bytes = (size_t)rows * lds * sizeof(float);
    t.out = direct ? nullptr : (float*)ws_alloc(h, bytes);
    t.out_lo = (float*)ws_alloc(h, bytes);
    if (getenv("FB_DEBUG_PLAN"))
      fprintf(stderr, "[fb plan %s] phase %d stage(%s%s) in=%p rows=%d cols=%d ld_in=%d T=%d bytes=%zu\n", h->ws_base ? "real" : "dry",
              phase, early ? "early" : "late", direct ? ", lo only" : "", (const void*)p, t.rows, t.cols, t.ld_in, t.transpose, bytes);
    if (early) add_early(t, avail_of(p)); else pending.push_back(t);
    *lo_out = t.out_lo; *ld_lo = lds;
    if (t.out) { *ld_out = lds; return t.out; }
    *ld_out = ld; return p;
  }
  void gemm_tc(const std::vector<GemmDesc>& g) {
    std::vector<TcGemmDesc> v;
    std::vector<TransposeDesc> pending;
    bool used_early = false;
    int work = 0;
    double flops = 0.0, bytes = 0.0;
    // Tile width and split-K of the launch, from a cost model of the persistent kernel (profiles/r1c_gemm_tc_breakdown.txt):
    // a k-block of a 128 x bn tile costs f(bn) (shared-memory bound: 1.0 / 0.8 / 0.72 for bn = 128 / 64 / 32, in units of
    // ~0.7 us), a work item additionally its epilogue; CTA c of min(items, 148) executes items c, c + grid, ... so the
    // launch lasts as long as its most loaded CTA (items dealt in rounds of alternating direction, see tc_snake).  Candidates: bn in {128, 64, 32} x split-K in {1, 2, 3, 4, 6, 8} (only
    // problems whose epilogue is linear may be split; their partial sums are added into a zeroed C).  Problems are ordered
    // longest k-chain first, which makes the round-robin an LPT schedule.
    std::vector<GemmDesc> gs = g;
    int bn_group = 128, sk_group = 1, pair_group = 0;
    // CTA pairs (cta_group::2): 256 x bn work items on 74 pairs.  Measured (profiles/r2_gemm_tc_pairs_ts.txt): correct, and a pair
    // needs no pre-split lo planes to match the single-CTA kernel fed with them, but it is not faster (the M = 256 tf32 instruction
    // issues at ~107 clk against 64 for M = 128, which cancels the halved operand traffic), so the plan uses pairs only on request
    // (FB_TC_PAIRS=1).  Never for plans that run through k_fused_stack (its GEMM stages are single-CTA) nor on the side lane (two
    // concurrent persistent launches could not both be resident).
    const bool pairs_ok = getenv("FB_TC_PAIRS") && !h->cfg.fused_stacks && cur_lane == 0;
    {
      auto natural_bn = [](int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : 128); };
      auto nkb_of = [](const GemmDesc& s) { return fb_ceil_div(s.K, 32) + fb_ceil_div(s.K2, 32); };   // the model counts 32-float k-steps
      // linear epilogues may be split along K (atomic partial sums); a ReLU epilogue only if its consumers take the pre-activation
      auto can_split = [](const GemmDesc& s, int nkb) {
        if (nkb < 8) return false;
        if (!(s.flags & GF_RELU)) return true;
        return (s.flags & GF_RELU_LAZY_OK) != 0 && !getenv("FB_NO_LAZY_RELU");
      };
      double best = 1e30;
      for (int pair : {0, 1}) {
      if (pair && !pairs_ok) continue;
      for (int bn : {128, 64, 32}) {
        for (int sk : {1, 2, 3, 4, 6, 8}) {
          if (sk > 1 && getenv("FB_NO_SPLITK")) continue;
          struct Item { double cost; int count; };
          std::vector<Item> items;
          for (const GemmDesc& s : gs) {
            const int bnp = std::min(bn, natural_bn(s.N)), nkb = nkb_of(s);
            const bool splittable = can_split(s, nkb);
            int len = nkb, parts = 1;
            if (splittable && sk > 1) { len = fb_ceil_div(nkb, std::min(sk, nkb / 4)); parts = fb_ceil_div(nkb, len); }
            const double f = (bnp == 128 ? 1.0 : (bnp == 64 ? 0.8 : 0.72)) * (pair ? 0.97 : 1.0);
            static const double epi_scale = getenv("FB_TC_EPI_COST") ? atof(getenv("FB_TC_EPI_COST")) : 1.0;
            const double epi = epi_scale * (2.0 + bnp / 64.0) * (parts > 1 ? 2.0 : 1.0);
            items.push_back(Item{len * f + epi, fb_ceil_div(s.M, TC_BM * (pair + 1)) * fb_ceil_div(s.N, bnp) * parts});
          }
          std::sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.cost > y.cost; });
          int total = 0;
          for (auto& it : items) total += it.count;
          const int grid = std::min(total, pair ? FB_SM_COUNT / 2 : FB_SM_COUNT);
          std::vector<double> load(grid, 0.0);
          int w = 0;
          for (auto& it : items)   // the kernel's deal: rounds of `grid` items in alternating direction (tc_snake)
            for (int i = 0; i < it.count; ++i, ++w) load[((w / grid) & 1) ? grid - 1 - w % grid : w % grid] += it.cost;
          double cost = 0.0;
          for (double l : load) cost = std::max(cost, l);
          cost += 0.02 * sk + (bn == 128 ? 0.0 : 0.01) + (pair ? 0.005 : 0.0);   // ties: fewer splits, wider tiles, single CTAs
          if (cost < best) { best = cost; bn_group = bn; sk_group = sk; pair_group = pair; }
        }
      }
      }
      if (getenv("FB_DEBUG_PLAN")) {
        fprintf(stderr, "[fb plan %s] phase %d gemm_tc group: bn=%d splitk=%d %s model cost %.1f :", h->ws_base ? "real" : "dry", phase, bn_group, sk_group, pair_group ? "pairs" : "single", best);
        for (const GemmDesc& s : gs) fprintf(stderr, " [%dx%dx%d%s%s]", s.M, s.N, s.K + s.K2, s.K2 ? "(K2)" : "", (s.flags & GF_RELU) ? " relu" : "");
        fprintf(stderr, "\n");
      }
      std::stable_sort(gs.begin(), gs.end(), [&](const GemmDesc& x, const GemmDesc& y) {
        auto len = [&](const GemmDesc& s) {
          const int nkb = nkb_of(s);
          const bool splittable = can_split(s, nkb);
          return (splittable && sk_group > 1) ? fb_ceil_div(nkb, std::min(sk_group, nkb / 4)) : nkb;
        };
        return len(x) > len(y);
      });
    }
    int ring_bn = 32;
    const int ncta = pair_group ? 2 : 1;
    if (ncta == 2) h->uses_pairs = true;
    for (const GemmDesc& s : gs) {
      TcGemmDesc d; memset(&d, 0, sizeof(d));
      d.C = s.C; d.bias = s.bias; d.mask = s.mask; d.M = s.M; d.N = s.N; d.K = s.K; d.K2 = s.K2;
      d.ldc = s.ldc; d.ldmask = s.ldmask; d.flags = s.flags & (GF_RELU | GF_MASK_RELU | GF_MASK_TANH);
      d.bn = s.N <= 32 ? 32 : (s.N <= 64 ? 64 : 128);
      if (d.bn > bn_group) d.bn = bn_group;
      if (d.bn > ring_bn) ring_bn = d.bn;
      d.tiles_m = fb_ceil_div(s.M, TC_BM * ncta); d.tiles_n = fb_ceil_div(s.N, d.bn);   // (pairs: 256-row work items)
      d.splitk = 1; d.kb_per_split = fb_ceil_div(s.K, TC_BK) + fb_ceil_div(s.K2, TC_BK);
      const bool may_split = !(s.flags & GF_RELU) || ((s.flags & GF_RELU_LAZY_OK) && !getenv("FB_NO_LAZY_RELU"));
      const int nkb32 = fb_ceil_div(s.K, 32) + fb_ceil_div(s.K2, 32);
      if (may_split && sk_group > 1 && nkb32 >= 8) {   // a k-range may straddle the two products
        const int per32 = fb_ceil_div(nkb32, std::min(sk_group, nkb32 / 4));
        d.splitk = fb_ceil_div(nkb32, per32);          // every k-range non-empty (also in TC_BK-sized k-blocks: see DESIGN.md)
        d.kb_per_split = per32 * TC_KB_PER_32;
      }
      if (d.splitk > 1 && (d.flags & GF_RELU)) {   // lazy ReLU: C receives the pre-activation, its consumers clamp it
        d.flags &= ~GF_RELU;
        lazy.push_back(LazyRange{s.C, (size_t)(s.M - 1) * s.ldc + s.N});
      }
      if (d.splitk > 1 && !in_grad(s.C)) {   // gradients are cleared by k_adam; any other split-K output is zeroed at phase start
        TransposeDesc z; memset(&z, 0, sizeof(z));
        z.out = s.C; z.rows = s.M; z.cols = s.N; z.ld_out = s.ldc; z.transpose = 2;
        if (cur_lane == 1 && phase == phase_index(FB_PHASE_MIX)) pending.push_back(z);   // side-lane chain: right before the launch
        else { add_early(z, phase); used_early = true; }
      }
      d.work_begin = work; d.work_count = d.tiles_m * d.tiles_n * d.splitk; work += d.work_count;
      int lda = 0, ldb = 0, lda2 = 0, ldb2 = 0, ldal = 0, ldbl = 0, ldal2 = 0, ldbl2 = 0;
      const float *Alo = nullptr, *Blo = nullptr, *A2lo = nullptr, *B2lo = nullptr;
      const float* A1 = stage_operand(pending, &used_early, s.A, s.a_kmajor, s.M, s.K, s.lda, false, &lda, &Alo, &ldal);
      const float* B1 = stage_operand(pending, &used_early, s.B, s.b_kmajor, s.N, s.K, s.ldb, true, &ldb, &Blo, &ldbl);
      const float* A2 = s.K2 ? stage_operand(pending, &used_early, s.A2, s.a_kmajor, s.M, s.K2, s.lda, false, &lda2, &A2lo, &ldal2) : nullptr;
      const float* B2 = s.K2 ? stage_operand(pending, &used_early, s.B2, s.b_kmajor, s.N, s.K2, s.ldb, true, &ldb2, &B2lo, &ldbl2) : nullptr;
      const bool a_pre = Alo && (!s.K2 || A2lo), b_pre = Blo && (!s.K2 || B2lo);
      if (a_pre) d.flags |= TC_A_PRE;
      if (b_pre) d.flags |= TC_B_PRE;
      // lazily-ReLU'd operands used WITHOUT a staged copy: A is clamped by the builder warps; anything else is a plan error
      if (A1 == s.A && is_lazy(s.A)) { if (a_pre || s.K2) { if (rc == FB_OK) rc = FB_E_STATE; } else d.flags |= TC_A_RELU; }
      if ((B1 == s.B && is_lazy(s.B)) || (s.K2 && (is_lazy(s.A2) || is_lazy(s.B2)))) { if (rc == FB_OK) rc = FB_E_STATE; }
      if (h->ws_base && rc == FB_OK) {
        rc = encode_tiled_map(&d.mapA, A1, s.M, s.K, lda, TC_BM);
        const int bbox = d.bn / ncta;   // B-tile rows one CTA loads
        if (rc == FB_OK) rc = encode_tiled_map(&d.mapB, B1, s.N, s.K, ldb, bbox);
        if (rc == FB_OK && s.K2) rc = encode_tiled_map(&d.mapA2, A2, s.M, s.K2, lda2, TC_BM);
        if (rc == FB_OK && s.K2) rc = encode_tiled_map(&d.mapB2, B2, s.N, s.K2, ldb2, bbox);
        if (rc == FB_OK && a_pre) rc = encode_tiled_map(&d.mapAlo, Alo, s.M, s.K, ldal, TC_BM);
        if (rc == FB_OK && b_pre) rc = encode_tiled_map(&d.mapBlo, Blo, s.N, s.K, ldbl, bbox);
        if (rc == FB_OK && a_pre && s.K2) rc = encode_tiled_map(&d.mapA2lo, A2lo, s.M, s.K2, ldal2, TC_BM);
        if (rc == FB_OK && b_pre && s.K2) rc = encode_tiled_map(&d.mapB2lo, B2lo, s.N, s.K2, ldbl2, bbox);
      }
      const double k = (double)s.K + s.K2;
      flops += 2.0 * s.M * (double)s.N * k;
      bytes += 4.0 * ((double)s.M * k + (double)s.N * k + (double)s.M * s.N);
      v.push_back(d);
    }
    if (!pending.empty()) {
      int ctas = 0;
      double tbytes = 0.0;
      for (auto& t : pending) {
        t.cta_begin = ctas; t.ctas_x = fb_ceil_div(t.cols, 32); ctas += t.ctas_x * fb_ceil_div(t.rows, 32);
        tbytes += (t.transpose == 2 ? 4.0 : (t.out ? 12.0 : 8.0)) * t.rows * (double)t.cols;
      }
      const TransposeDesc* td = arena_put(h, pending, d_arena);
      const int nt = (int)pending.size();
      if (nt <= 8) {
        DescTable<TransposeDesc, 8> tab; memset(&tab, 0, sizeof(tab));
        tab.n = nt;
        for (int i = 0; i < nt; ++i) tab.d[i] = pending[i];
        push([tab, ctas](cudaStream_t s) {
          fb_launch_pdl(k_transpose_grouped_tab, dim3(ctas), dim3(256), 0, s, tab);
          return cudaGetLastError();
        }, FB_OPK_TRANSPOSE, 0.0, tbytes);
      } else
      push([td, nt, ctas](cudaStream_t s) {
        fb_launch_pdl(k_transpose_grouped, dim3(ctas), dim3(256), 0, s, td, nt);
        return cudaGetLastError();
      }, FB_OPK_TRANSPOSE, 0.0, tbytes);
      h->ops[phase].back().dev.set(FS_TRANSPOSE, ctas, FsPtrArgs{td, nt, 0});
    }
    const TcGemmDesc* dd = arena_put(h, v, d_arena);
    const int n = (int)v.size();
    h->uses_gemm_tc = true;
    for (int i = 0; i < n; ++i)
      if (v[i].splitk == 1 && v[i].M % 4 == 0 && cur_lane == 0)
        produced.push_back(Produced{v[i].C, v[i].M, v[i].N, v[i].ldc, (size_t)((const char*)(dd + i) - d_arena)});
    TcLaunch hdr;   // persistent grid: one CTA per SM (or one pair per two SMs) walks the group's work items
    if (make_tc_launch(v, work, ring_bn, &hdr) != FB_OK && rc == FB_OK) rc = FB_E_UNSUPPORTED;
    push([dd, hdr, work, ncta](cudaStream_t s) {
      return tc_launch(dd, hdr, work, ncta, s, getenv("FB_NO_PDL") == nullptr);
    }, FB_OPK_GEMM_TC, flops, bytes);
    { FsGemmArgs fa; memset(&fa, 0, sizeof(fa)); fa.descs = dd; fa.hdr = hdr; h->ops[phase].back().dev.set(FS_GEMM_TC, work, fa); }
    if (used_early) h->ops[phase].back().wait_stage = 1;   // operands staged on the staging lane: wait for this phase's staging event
  }

  // Problems with an empty dimension are dropped: a configuration without one of the embeds (preprocess = False) builds the same
  // grouped launches with zero-row activations standing in for the missing network.
  void gemm(std::vector<GemmDesc> g) {
    g.erase(std::remove_if(g.begin(), g.end(), [](const GemmDesc& d) { return d.M <= 0 || d.N <= 0 || d.K <= 0; }), g.end());
    if (g.empty()) return;
    std::vector<GemmDesc> tc, simt;
    for (auto& d : g) (tc_ok(d) ? tc : simt).push_back(d);
    if (!tc.empty()) gemm_tc(tc);
    if (simt.empty()) return;
    g = std::move(simt);
    for (auto& d : g)   // the fp32 SIMT kernel has no lazy-ReLU operand path
      if (is_lazy(d.A) || is_lazy(d.B) || (d.K2 && (is_lazy(d.A2) || is_lazy(d.B2)))) { if (rc == FB_OK) rc = FB_E_STATE; }
    GroupLaunch gl = finalize_group(h, std::move(g), d_arena);
    push([gl](cudaStream_t s) {
      fb_launch_pdl(k_gemm_grouped, dim3(gl.ctas), dim3(GEMM_THREADS), GEMM_SMEM_BYTES, s, gl.d_descs, gl.nprob);
      return cudaGetLastError();
    }, FB_OPK_GEMM, gl.flops, gl.bytes);
  }
  void ln_fwd(std::vector<LnDesc> v) {
    v.erase(std::remove_if(v.begin(), v.end(), [](const LnDesc& d) { return d.rows <= 0; }), v.end());
    if (v.empty()) return;
    int rows = 0;
    double bytes = 0.0;
    for (auto& d : v) { d.row_begin = rows; rows += d.rows; bytes += 8.0 * d.rows * (double)d.D; }
    bool vec = true;
    for (auto& d : v) vec = vec && d.D <= 1024 && d.ld % 4 == 0 && aligned16(d.x) && aligned16(d.y);
    DescTable<LnDesc, 8> tab; memset(&tab, 0, sizeof(tab));
    if (v.size() > 8) { if (rc == FB_OK) rc = FB_E_UNSUPPORTED; return; }
    tab.n = (int)v.size();
    for (size_t i = 0; i < v.size(); ++i) tab.d[i] = v[i];
    push([tab, rows, vec](cudaStream_t s) {
      if (vec) fb_launch_pdl(k_ln_tanh_fwd_v4, dim3(fb_ceil_div(rows, 8)), dim3(256), 0, s, tab, rows);
      else fb_launch_pdl(k_ln_tanh_fwd, dim3(fb_ceil_div(rows, 8)), dim3(256), 0, s, tab, rows);
      return cudaGetLastError();
    }, FB_OPK_LAYERNORM, 0.0, bytes);
    { FsLnFwdArgs fa; memset(&fa, 0, sizeof(fa)); fa.tab = tab; fa.rows = rows; fa.vec = vec ? 1 : 0; h->ops[phase].back().dev.set(FS_LN_FWD, fb_ceil_div(rows, 8), fa); }
  }
  void ln_bwd(std::vector<LnBwdDesc> v) {
    v.erase(std::remove_if(v.begin(), v.end(), [](const LnBwdDesc& d) { return d.rows <= 0; }), v.end());
    if (v.empty()) return;
    int ctas = 0;
    double bytes = 0.0;
    for (auto& d : v) {
      d.cta_begin = ctas; d.cta_count = fb_ceil_div(d.rows, FB_LN_BWD_ROWS_PER_CTA); ctas += d.cta_count;
      bytes += 16.0 * d.rows * (double)d.D;
    }
    for (auto& d : v) invalidate_produced(d.dx);
    bool vec = true;  // every problem narrow enough and 16-byte aligned for the register-resident variant?
    for (auto& d : v)
      vec = vec && d.D <= 1024 && d.ld % 4 == 0 && d.ld_dy % 4 == 0 && aligned16(d.dy) && aligned16(d.y) && aligned16(d.x) && aligned16(d.dx);
    DescTable<LnBwdDesc, 4> tab; memset(&tab, 0, sizeof(tab));
    if (v.size() > 4) { if (rc == FB_OK) rc = FB_E_UNSUPPORTED; return; }
    tab.n = (int)v.size();
    for (size_t i = 0; i < v.size(); ++i) tab.d[i] = v[i];
    bool full = vec;   // every problem exactly 1024 wide: the kernel without column-bound tests (128 registers: two CTAs per SM)
    for (auto& d : v) full = full && d.D == 1024;
    push([tab, ctas, vec, full](cudaStream_t s) {
      if (full) fb_launch_pdl(k_ln_tanh_bwd_v4_full, dim3(ctas), dim3(256), 0, s, tab);
      else if (vec) fb_launch_pdl(k_ln_tanh_bwd_v4, dim3(ctas), dim3(256), 0, s, tab);
      else fb_launch_pdl(k_ln_tanh_bwd, dim3(ctas), dim3(256), 0, s, tab);
      return cudaGetLastError();
    }, FB_OPK_LAYERNORM, 0.0, bytes);
    { FsLnBwdArgs fa; memset(&fa, 0, sizeof(fa)); fa.tab = tab; fa.vec = vec ? 1 : 0; h->ops[phase].back().dev.set(FS_LN_BWD, ctas, fa); }
  }
  void l2_fwd(std::vector<L2Desc> v) {
    int rows = 0;
    for (auto& d : v) { d.row_begin = rows; rows += d.rows; }
    const L2Desc* dd = arena_put(h, v, d_arena);
    const int n = (int)v.size();
    push([dd, n, rows](cudaStream_t s) {
      fb_launch_pdl(k_l2norm_fwd, dim3(fb_ceil_div(rows, 8)), dim3(256), 0, s, dd, n, rows);
      return cudaGetLastError();
    });
    h->ops[phase].back().dev.set(FS_L2_FWD, fb_ceil_div(rows, 8), FsPtrArgs{dd, n, rows});
  }
  void colsum(std::vector<ColsumDesc> v) {
    v.erase(std::remove_if(v.begin(), v.end(), [](const ColsumDesc& d) { return d.rows <= 0; }), v.end());
    if (v.empty()) return;
    int ctas = 0;
    double bytes = 0.0;
    for (auto& d : v) {
      bytes += 4.0 * d.rows * (double)d.N;
      d.cta_begin = ctas; d.ctas_n = fb_ceil_div(d.N, 32); d.ctas_r = fb_ceil_div(d.rows, FB_COLSUM_ROWS_PER_CTA);
      ctas += d.ctas_n * d.ctas_r;
    }
    const ColsumDesc* dd = arena_put(h, v, d_arena);
    const int n = (int)v.size();
    push([dd, n, ctas](cudaStream_t s) {
      fb_launch_pdl(k_colsum, dim3(ctas), dim3(256), 0, s, dd, n);
      return cudaGetLastError();
    }, FB_OPK_COLSUM, 0.0, bytes, 1);  // side lane: bias gradients are leaves of the dependency graph (joined at phase end)
    h->ops[phase].back().dev.set(FS_COLSUM, ctas, FsPtrArgs{dd, n, 0});
  }
  void memset0(void* p, size_t bytes) {
    push([p, bytes](cudaStream_t s) { return cudaMemsetAsync(p, 0, bytes, s); }, FB_OPK_MEMSET, 0.0, (double)bytes);
  }
};

static ColsumDesc mk_colsum(const Mat& src, float* dst) {
  ColsumDesc d; memset(&d, 0, sizeof(d));
  d.src = src.p; d.dst = dst; d.rows = src.rows; d.N = src.cols; d.ld = src.ld;
  return d;
}

// activations of one "embed" block: Linear -> LayerNorm -> Tanh -> Linear -> ReLU  (fb_modules.py:60-78)
struct EmbedAct { Mat x, pre, y, out; float* mean; float* rstd; };
static EmbedAct embed_alloc(fb_handle* h, const Mat& x, const Mat& out, int H, const std::string& name) {
  EmbedAct e; e.x = x; e.out = out;
  e.pre = ws_mat(h, x.rows, H, (name + ".pre").c_str());
  e.y = ws_mat(h, x.rows, H, (name + ".y").c_str());
  e.mean = (float*)ws_alloc(h, x.rows * sizeof(float));
  e.rstd = (float*)ws_alloc(h, x.rows * sizeof(float));
  return e;
}
static LnDesc embed_ln(const EmbedAct& e, const PSet& p) {
  LnDesc d; memset(&d, 0, sizeof(d));
  d.x = e.pre.p; d.y = e.y.p; d.gamma = p.v(2); d.beta = p.v(3); d.mean = e.mean; d.rstd = e.rstd;
  d.rows = e.pre.rows; d.D = e.pre.cols; d.ld = e.pre.ld;
  return d;
}
// backward of LN+tanh for rows [r0, r0+n) of an embed block: dy (grad wrt tanh output) -> dx in place
static LnBwdDesc embed_ln_bwd(const EmbedAct& e, const PSet& p, const Mat& dy, int r0, bool affine) {
  LnBwdDesc d; memset(&d, 0, sizeof(d));
  d.dy = dy.p; d.dx = dy.p; d.y = e.y.p + (size_t)r0 * e.y.ld; d.x = e.pre.p + (size_t)r0 * e.pre.ld;
  d.gamma = p.v(2); d.mean = e.mean + r0; d.rstd = e.rstd + r0;
  d.dgamma = affine ? p.gv(2) : nullptr; d.dbeta = affine ? p.gv(3) : nullptr;
  d.rows = dy.rows; d.D = dy.cols; d.ld = e.y.ld; d.ld_dy = dy.ld;
  return d;
}

// activations of one BackwardMap instance: Linear -> LN -> Tanh -> Linear -> ReLU -> Linear -> sqrt(Z) normalize
struct BAct { Mat x, pre, y, h2, raw, out; float* mean; float* rstd; float* nrm; };
// width / out_dim: hidden and output widths (defaults: the BackwardMap's); the DiagGaussianActor of cfg.boltzmann has the same shape
// identity: cfg.debug (fb_ddpg.py:128-130, fb_modules.py:202-208): the backward map is nn.Identity — no layers, no activations; "raw" is
// the input itself (goal_dim == z_dim) and the L2 launch that follows degenerates to the copy input -> output
static BAct b_alloc(fb_handle* h, const Mat& x, const Mat& out, const std::string& name, int width = 0, int out_dim = 0, bool identity = false) {
  const fb_config& c = h->cfg;
  if (!width) width = c.backward_hidden_dim;
  if (!out_dim) out_dim = c.z_dim;
  BAct b; b.x = x; b.out = out;
  const int rows = identity ? 0 : x.rows;
  b.pre = ws_mat(h, rows, width, (name + ".pre").c_str());
  b.y = ws_mat(h, rows, width, (name + ".y").c_str());
  b.h2 = ws_mat(h, rows, width, (name + ".h2").c_str());
  if (identity) { b.raw = x; b.raw.cols = out_dim; }
  else b.raw = ws_mat(h, x.rows, out_dim, (name + ".raw").c_str());
  b.mean = (float*)ws_alloc(h, x.rows * sizeof(float));
  b.rstd = (float*)ws_alloc(h, x.rows * sizeof(float));
  b.nrm = (float*)ws_alloc(h, x.rows * sizeof(float));
  return b;
}
static LnDesc b_ln(const BAct& b, const PSet& p) {
  LnDesc d; memset(&d, 0, sizeof(d));
  d.x = b.pre.p; d.y = b.y.p; d.gamma = p.v(2); d.beta = p.v(3); d.mean = b.mean; d.rstd = b.rstd;
  d.rows = b.pre.rows; d.D = b.pre.cols; d.ld = b.pre.ld;
  return d;
}
static L2Desc b_l2(const BAct& b, int Z, int normalize) {
  L2Desc d; memset(&d, 0, sizeof(d));
  d.normalize = normalize;
  d.x = b.raw.p; d.y = b.out.p; d.nrm = b.nrm; d.rows = b.raw.rows; d.Z = Z; d.ldx = b.raw.ld; d.ldy = b.out.ld;
  return d;
}

// head: Linear(2Fd -> H) -> ReLU -> Linear(H -> out)
struct HeadAct { Mat h1, out; };

// ------------------------------------------------------------------------------------------------
// the plan
// ------------------------------------------------------------------------------------------------
static int build_gather_params(const fb_replay_view& v, GatherParams& gp, const BatchLayout& L, int out_ld);
static void make_batch_layout(BatchLayout& L, int O, int A, int G, int X, int with_future);

static int build_plan(fb_handle* h) {
  const fb_config& c = h->cfg;
  const int B = c.batch, n = c.global_batch, O = c.obs_dim, A = c.action_dim, Z = c.z_dim, H = c.hidden_dim, Fd = c.feature_dim;
  const int G = c.goal_dim;
  const bool use_goal = c.use_goal != 0;
  // preprocess = False: ONE embed-shaped block (trunk.0 / .1 / .3, output width H) on [obs | z (| action)] followed by trunk.5 = the
  // add_trunk layer with an H-wide input.  The plan below is the add_trunk plan in which the second embed of every pair has zero rows
  // (its launches drop out in the Builder): for forward_net the obs_action slot is the one that stays, for the actor the obs_z slot.
  const bool deep = c.no_preprocess != 0;
  const int Fe = deep ? H : Fd;          // output width of an embed
  const int Hc = deep ? H : 2 * Fd;      // width of the (concatenated) embed output the trunk / heads read
  const int act_col = deep ? O + Z : O;  // column of the action inside the forward-net inputs
  const int nz = c.no_norm_z ? 0 : 1;   // cfg.norm_z: sqrt(Z)-sphere projection of backward_net outputs and of the mixed z
  // cfg.boltzmann: the actor is ONE BackwardMap-shaped stack (Linear -> LN -> tanh -> Linear -> ReLU -> Linear) on [obs | z] with a
  // [mu | raw log-std] output; every launch of the default actor (two embeds, trunk, policy head) is built with zero rows and
  // drops out in the Builder, the stack's layers join the grouped launches of the same depth
  const bool bz = c.boltzmann != 0;
  // cfg.debug: backward_net / backward_target_net are identity maps (no parameters, no sqrt(Z) projection of their output)
  const bool dbg = c.debug_identity_b != 0;
  const int nzB = dbg ? 0 : nz;
  for (auto& v : h->ops) v.clear();
  for (auto& v : h->early_stage) v.clear();
  for (auto& v : h->early_avail) v.clear();
  for (auto& v : h->stage_batches) v.clear();
  h->views.clear();
  h->arena.clear();
  h->ws_off = 0;

  char* d_arena = (char*)ws_alloc(h, FB_DESC_ARENA_BYTES);
  h->d_sc = (DevScalars*)ws_alloc(h, sizeof(DevScalars));
  h->d_acc = (double*)ws_alloc(h, ACC_COUNT * sizeof(double));
  h->d_linf = (unsigned int*)ws_alloc(h, (size_t)c.z_dim * c.z_dim * sizeof(float));   // B^T B scratch of the metrics phase (k_metric_cov)
  h->d_metrics = (float*)ws_alloc(h, FB_METRIC_COUNT * sizeof(float));
  h->d_n_episodes = (int*)ws_alloc(h, 16);
  h->d_ep_idx = (int*)ws_alloc(h, B * sizeof(int));
  h->d_step_idx = (int*)ws_alloc(h, B * sizeof(int));
  h->d_future_idx = (int*)ws_alloc(h, B * sizeof(int));
  h->d_perm = (int*)ws_alloc(h, B * sizeof(int));
  h->d_identity_perm = (int*)ws_alloc(h, B * sizeof(int));
  h->d_mix_mask = (int*)ws_alloc(h, B * sizeof(int));
  h->d_future_mask = (int*)ws_alloc(h, B * sizeof(int));
  h->d_perm_keys = (unsigned int*)ws_alloc(h, B * sizeof(int));
  h->d_prog = (char*)ws_alloc(h, FB_PROG_ARENA_BYTES);
  h->d_fs_barrier = (unsigned long long*)ws_alloc(h, FS_NUM_BARRIERS * sizeof(unsigned long long));
  h->d_fs_err = (unsigned int*)ws_alloc(h, 16);
  h->d_fs_times = (unsigned long long*)ws_alloc(h, (FS_MAX_STAGES + 1) * sizeof(unsigned long long));
  h->uses_pairs = false;
  h->fused_plans.clear(); h->prog_host.clear(); h->prog_uploaded = 0; h->n_fs_barriers = 0;

  // ---- packed batch rows ---------------------------------------------------------------------
  BatchLayout& L = h->bl;
  const bool with_future = c.future_ratio > 0.f;   // hindsight z: the batch rows also carry future_obs / future_goal
  make_batch_layout(L, O, A, use_goal ? G : 0, 0, with_future ? 1 : 0);  // the step does not read meta fields
  h->packed = ws_mat(h, B, L.pitch, "packed");

  // ---- step inputs ---------------------------------------------------------------------------
  Mat actor_in_o = ws_mat(h, (deep || bz) ? 0 : 2 * B, O, "actor_in_o");
  Mat actor_in_oz = ws_mat(h, 2 * B, O + Z, "actor_in_oz");
  Mat in_oa = ws_mat(h, B, act_col + A, "in_oa");      // [obs | action], or [obs | z | action]
  Mat in_noa = ws_mat(h, B, act_col + A, "in_noa");
  Mat in_oa2 = ws_mat(h, B, act_col + A, "in_oa2");
  Mat goal_next = ws_mat(h, B, G, "next_goal");
  Mat mix_in = ws_mat(h, with_future ? 2 * B : B, G, "mix_input");   // [backward_input[perm] ; future goal]
  h->z_rand = ws_mat(h, B, Z, "z_rand");
  Mat z = ws_mat(h, B, Z, "z");
  h->noise_fb = ws_mat(h, B, A, "noise_fb");
  h->noise_actor = ws_mat(h, B, A, "noise_actor");
  Mat mu = ws_mat(h, 2 * B, A, "mu_all");
  h->views["mu"] = mu.rs(B, B);
  const bool rand_w = c.rand_weight != 0 && c.mix_ratio > 0.f;
  Mat b_mixw;
  if (rand_w) {
    h->mix_w = ws_mat(h, B, B, "mix_w");
    h->mix_u = (float*)ws_alloc(h, B * sizeof(float));
    { Mat u; u.p = h->mix_u; u.rows = B; u.cols = 1; u.ld = 1; h->views["mix_u"] = u; }
    b_mixw = ws_mat(h, B, Z, "B_mixw");
  }
  Mat next_action = ws_mat(h, B, A, "next_action");
  Mat action_new = ws_mat(h, B, A, "action_new");

  // gather block [F1|F2|tF1|tF2|B|tB|discount ...] : what a rank contributes to the all-gather
  const int ldZ = fb_round_up(Z, 4);
  const int blk_pitch = 6 * ldZ + 4;
  h->blk_local = ws_mat(h, B, blk_pitch, "blk_local");
  if (n != B) h->blk_global = ws_mat(h, n, blk_pitch, "blk_global");
  else h->blk_global = h->blk_local;
  if (n != B && h->p2p_on && h->ws_base) {   // peers store their rows straight into it: it lives in this rank's arena
    h->blk_global.p = reinterpret_cast<float*>(h->p2p_arena + h->p2p_off_blk);
    h->views["blk_global"] = h->blk_global;
  }
  const Mat& bl = h->blk_local; const Mat& bg = h->blk_global;
  auto blkcol = [&](const Mat& m, int i) { Mat r = m; r.p = m.p + i * ldZ; r.cols = Z; return r; };
  Mat F1 = blkcol(bl, 0), F2 = blkcol(bl, 1), tF1 = blkcol(bl, 2), tF2 = blkcol(bl, 3), Bm = blkcol(bl, 4), tB = blkcol(bl, 5);
  Mat F1g = blkcol(bg, 0), F2g = blkcol(bg, 1), tF1g = blkcol(bg, 2), tF2g = blkcol(bg, 3), Bg = blkcol(bg, 4), tBg = blkcol(bg, 5);
  const int disc_col = 6 * ldZ;
  h->views["F1"] = F1; h->views["F2"] = F2; h->views["tF1"] = tF1; h->views["tF2"] = tF2; h->views["B"] = Bm; h->views["tB"] = tB;
  { Mat d = bl; d.p = bl.p + disc_col; d.cols = 1; h->views["discount"] = d; }

  // ---- parameter sets ------------------------------------------------------------------------
  const fb_buffers& bf = h->bufs;
  PSet pF{&h->seg_fb, h->fwd_first, bf.d_param_fb, bf.d_grad_fb};
  PSet pFt{&h->seg_fb, h->fwd_first, bf.d_target_fb, nullptr};
  PSet pB{&h->seg_fb, h->bwd_first, bf.d_param_fb, bf.d_grad_fb};
  PSet pBt{&h->seg_fb, h->bwd_first, bf.d_target_fb, nullptr};
  PSet pA{&h->seg_actor, 0, bf.d_param_actor, bf.d_grad_actor};
  // sub-blocks: embed = 6 tensors (W0 b0 gamma beta W3 b3), [trunk = 2 tensors (W b), cfg.add_trunk], head = 4 tensors (W1 b1 W2 b2)
  const bool trunk = c.add_trunk != 0 || deep;
  const int T_N = trunk ? 2 : 0, NE = deep ? 6 : 12;
  const int E_OA = 0, E_OZ = deep ? 0 : 6, F_TR = NE, HD_1 = NE + T_N, HD_2 = NE + 4 + T_N;   // forward net (deep: E_OZ is never launched)
  const int A_O = 0, A_OZ = deep ? 0 : 6, A_TR = NE, A_POL = NE + T_N;                        // actor (deep: A_O is never launched)

  // ---- activations ---------------------------------------------------------------------------
  const int RA = bz ? 0 : 2 * B, RA1 = bz ? 0 : B;   // rows of the default actor's activations / gradients (none with cfg.boltzmann)
  Mat hA = ws_mat(h, RA, Hc, "hA");
  Mat hFt = ws_mat(h, B, Hc, "hFt");
  Mat hF = ws_mat(h, B, Hc, "hF");
  Mat hF2 = ws_mat(h, B, Hc, "hF2");
  const int oz0 = deep ? 0 : Fd;   // first column of the obs_z embed inside the concatenated output
  auto oz_rows = [&](int r0) { Mat m = actor_in_oz.rs(r0, B); if (deep) m.rows = 0; return m; };   // forward_net's obs_z embeds: absent when deep
  EmbedAct eAo = embed_alloc(h, actor_in_o, hA.cs(0, Fe), H, "actor.obs_net");
  Mat actor_oz_x = actor_in_oz; actor_oz_x.rows = RA;
  EmbedAct eAoz = embed_alloc(h, actor_oz_x, hA.cs(oz0, Fe), H, "actor.obs_z_net");
  EmbedAct eFtoa = embed_alloc(h, in_noa, hFt.cs(0, Fe), H, "Ft.obs_action_net");
  EmbedAct eFtoz = embed_alloc(h, oz_rows(0), hFt.cs(oz0, Fe), H, "Ft.obs_z_net");
  EmbedAct eFoa = embed_alloc(h, in_oa, hF.cs(0, Fe), H, "F.obs_action_net");
  EmbedAct eFoz = embed_alloc(h, oz_rows(B), hF.cs(oz0, Fe), H, "F.obs_z_net");
  EmbedAct eF2oa = embed_alloc(h, in_oa2, hF2.cs(0, Fe), H, "F2.obs_action_net");
  EmbedAct eF2oz = embed_alloc(h, oz_rows(B), hF2.cs(oz0, Fe), H, "F2.obs_z_net");
  // add_trunk: ReLU(Linear(2 Fd -> H)) of each concatenated embed pair; the heads then read these instead of hA / hFt / hF / hF2
  Mat trA, trFt, trF, trF2;
  if (trunk) { trA = ws_mat(h, RA, H, "actor.trunk"); trFt = ws_mat(h, B, H, "Ft.trunk"); trF = ws_mat(h, B, H, "F.trunk"); trF2 = ws_mat(h, B, H, "F2.trunk"); }
  const Mat& inFt = trunk ? trFt : hFt; const Mat& inF = trunk ? trF : hF; const Mat& inF2 = trunk ? trF2 : hF2;
  Mat h1A = ws_mat(h, RA, H, "actor.policy.h1");
  Mat preA = ws_mat(h, RA, A, "actor.policy.out");
  // cfg.boltzmann: the DiagGaussianActor stack on [next_obs | z ; obs | z]; raw = [mu | raw log-std]
  Mat pol_x = actor_in_oz; pol_x.rows = bz ? 2 * B : 0;
  BAct pol = b_alloc(h, pol_x, Mat(), "actor.policy", H, 2 * A);
  Mat bz_x = ws_mat(h, bz ? B : 0, A, "bz_x"), bz_sd = ws_mat(h, bz ? B : 0, A, "bz_std"), bz_t = ws_mat(h, bz ? B : 0, A, "bz_t");
  Mat h1Ft1 = ws_mat(h, B, H, "Ft.F1.h1"), h1Ft2 = ws_mat(h, B, H, "Ft.F2.h1");
  Mat h1F1 = ws_mat(h, B, H, "F.F1.h1"), h1F2 = ws_mat(h, B, H, "F.F2.h1");
  Mat h1Fa1 = ws_mat(h, B, H, "F2.F1.h1"), h1Fa2 = ws_mat(h, B, H, "F2.F2.h1");
  Mat Fa = ws_mat(h, B, 2 * ldZ, "Fa");
  Mat Fa1 = Fa.cs(0, Z), Fa2 = Fa.cs(ldZ, Z);
  h->views["F1a"] = Fa1; h->views["F2a"] = Fa2;
  Mat b_mix_out = ws_mat(h, with_future ? 2 * B : B, Z, "B_mix");
  BAct bMix = b_alloc(h, mix_in, b_mix_out, "Bmix", 0, 0, dbg);
  BAct bT = b_alloc(h, goal_next, tB, "Bt", 0, 0, dbg);
  BAct bO = b_alloc(h, goal_next, Bm, "Bo", 0, 0, dbg);

  h->ws_fwd_end = h->ws_off;  // everything allocated so far is written by MIX / FB_FWD / ACTOR_FWD (or is an input)

  // loss matrices (row block and column block); Mat cols = n
  Mat M1 = ws_mat(h, B, n, "M1"), M2 = ws_mat(h, B, n, "M2"), T1 = ws_mat(h, B, n, "T1"), T2 = ws_mat(h, B, n, "T2");
  Mat Cov = ws_mat(h, B, n, "Cov");
  Mat Mt1 = ws_mat(h, B, n, "Mt1"), Mt2 = ws_mat(h, B, n, "Mt2"), Tt1 = ws_mat(h, B, n, "Tt1"), Tt2 = ws_mat(h, B, n, "Tt2");
  Mat dblk = ws_mat(h, B, 3 * ldZ, "dblk");
  Mat dF1 = dblk.cs(0, Z), dF2 = dblk.cs(ldZ, Z), dB = dblk.cs(2 * ldZ, Z);
  h->views["dF1"] = dF1; h->views["dF2"] = dF2; h->views["dB"] = dB;
  Mat draw = ws_mat(h, B, Z, "dBraw");
  // backward scratch
  Mat dh1 = ws_mat(h, B, 2 * H, "dh1");          // [dh1_F1 | dh1_F2]
  Mat dh1_1 = dh1.cs(0, H), dh1_2 = dh1.cs(H, H);
  Mat dhF = ws_mat(h, B, Hc, "dhF");
  Mat dy_oa = ws_mat(h, B, H, "dy_oa"), dy_oz = ws_mat(h, deep ? 0 : B, H, "dy_oz");
  Mat dh2 = ws_mat(h, dbg ? 0 : B, c.backward_hidden_dim, "dh2"), dy1 = ws_mat(h, dbg ? 0 : B, c.backward_hidden_dim, "dy1");   // (cfg.debug: no backward_net to differentiate)
  Mat dFa = ws_mat(h, B, 2 * ldZ, "dFa");
  Mat dFa1 = dFa.cs(0, Z), dFa2 = dFa.cs(ldZ, Z);
  Mat dhoa = ws_mat(h, B, Fe, "dhoa");
  Mat dpreA = ws_mat(h, RA1, A, "dpreA");
  Mat dh1A = ws_mat(h, RA1, H, "dh1A");
  Mat dhA = ws_mat(h, RA1, Hc, "dhA");
  Mat dy_o = ws_mat(h, (deep || bz) ? 0 : B, H, "dy_o"), dy_aoz = ws_mat(h, RA1, H, "dy_aoz");
  const int RB = bz ? B : 0;   // cfg.boltzmann: gradients of the DiagGaussianActor stack
  Mat bz_da = ws_mat(h, RB, A, "bz_da"), bz_dpre = ws_mat(h, RB, 2 * A, "bz_dpre"), bz_dh2 = ws_mat(h, RB, H, "bz_dh2"), bz_dy = ws_mat(h, RB, H, "bz_dy");
  Mat dtF, dtA;   // add_trunk: gradients w.r.t. the trunk outputs
  if (trunk) { dtF = ws_mat(h, B, H, "dtF"); dtA = ws_mat(h, RA1, H, "dtA"); }

  Builder b{h, d_arena};
  DevScalars* sc = h->d_sc;
  double* acc = h->d_acc;

  // =========================== FB_PHASE_SAMPLE ===================================================
  b.set_phase(FB_PHASE_SAMPLE);
  if (c.rng_device) {
    RngParams rp; memset(&rp, 0, sizeof(rp));
    rp.seed = c.seed; rp.batch = B; rp.Z = Z; rp.A = A; rp.ldZ = h->z_rand.ld; rp.ldA = h->noise_fb.ld; rp.mix_ratio = c.mix_ratio; rp.norm_z = nz;
    rp.future_ratio = c.future_ratio; rp.future_mask = with_future ? h->d_future_mask : nullptr;
    rp.n_episodes = h->d_n_episodes; rp.ep_idx = h->d_ep_idx; rp.step_idx = h->d_step_idx; rp.future_idx = h->d_future_idx;
    rp.mix_mask = h->d_mix_mask; rp.perm_keys = h->d_perm_keys; rp.z_rand = h->z_rand.p; rp.noise_fb = h->noise_fb.p;
    rp.noise_actor = h->noise_actor.p;
    fb_handle* hh = h;
    b.push([rp, sc, hh](cudaStream_t s) mutable {
      rp.ep_len = hh->replay_bound ? hh->replay.d_episode_len : nullptr;   // no replay (host batches): no index draws
      rp.rows_per_episode = hh->replay_bound ? hh->replay.rows_per_episode : 2;
      fb_launch_pdl(k_rng_draw, dim3(fb_ceil_div(rp.batch, 8)), dim3(256), 0, s, rp, sc);
      return cudaGetLastError();
    });
    if (rand_w) {   // before k_randperm: it closes the step's draws by bumping the counter
      MixWeightRngParams wp; memset(&wp, 0, sizeof(wp));
      wp.seed = c.seed; wp.batch = B; wp.W = h->mix_w.p; wp.ldw = h->mix_w.ld; wp.u = h->mix_u;
      b.push([wp, sc](cudaStream_t s) {
        fb_launch_pdl(k_rng_mix_weights, dim3(FB_SM_COUNT * 4), dim3(256), 0, s, wp, sc);
        return cudaGetLastError();
      }, FB_OPK_ELEMENTWISE, 0.0, 4.0 * B * (double)B);
    }
    unsigned int* keys = h->d_perm_keys; int* perm = h->d_perm;
    b.push([keys, perm, B, sc](cudaStream_t s) {
      fb_launch_pdl(k_randperm, dim3(fb_ceil_div(8 * B, 256)), dim3(256), (size_t)B * sizeof(unsigned int), s, keys, B, perm, sc);
      return cudaGetLastError();
    });
  }
  {
    fb_handle* hh = h;
    b.push([hh, sc](cudaStream_t s) {
      if (!hh->replay_bound) return cudaErrorInvalidValue;
      GatherParams gp;
      if (build_gather_params(hh->replay, gp, hh->bl, hh->packed.ld) != FB_OK) return cudaErrorInvalidValue;
      fb_launch_pdl(k_gather_rows, dim3(fb_ceil_div(hh->cfg.batch, 8)), dim3(256), 0, s, gp, hh->d_ep_idx, hh->d_step_idx,
                    hh->bl.with_future ? hh->d_future_idx : nullptr, hh->cfg.batch,
                                                                 &sc->replay_discount, 0.f, hh->packed.p);
      return cudaGetLastError();
    }, FB_OPK_GATHER, 0.0, 2.0 * 4.0 * (double)B * (double)h->bl.pitch);
    h->ops[b.phase].back().replay_only = 1;
  }

  // =========================== FB_PHASE_MIX =====================================================
  b.set_phase(FB_PHASE_MIX);
  {
    StageParams sp; memset(&sp, 0, sizeof(sp));
    sp.L = L; sp.batch = B; sp.use_goal = use_goal ? 1 : 0;
    sp.actor_in_o = (deep || bz) ? nullptr : actor_in_o.p; sp.ldO = actor_in_o.ld; sp.act_col = act_col; sp.actor_in_oz = actor_in_oz.p; sp.ldOZ = actor_in_oz.ld;
    sp.in_oa = in_oa.p; sp.in_noa = in_noa.p; sp.in_oa2 = in_oa2.p; sp.ldOA = in_oa.ld;
    sp.goal_next = goal_next.p; sp.mix_in = mix_in.p; sp.ldG = goal_next.ld;
    sp.blk = bl.p; sp.blk_pitch = bl.ld; sp.disc_col = disc_col;
    sp.perm = h->d_perm; sp.mix_override = nullptr; sp.with_future = with_future ? 1 : 0;
    const float* packed = h->packed.p;
    b.push([sp, packed](cudaStream_t s) { fb_launch_pdl(k_stage_inputs, dim3(sp.batch), dim3(128), 0, s, sp, packed); return cudaGetLastError(); });
    { FsStageInArgs fa; memset(&fa, 0, sizeof(fa)); fa.P = sp; fa.packed = packed; h->ops[b.phase].back().dev.set(FS_STAGE_INPUTS, sp.batch, fa); }
  }
  const bool do_mix = c.mix_ratio > 0.f || with_future;   // one backward_net forward serves the mixing rows and the hindsight rows
  // The z-mixing forward (a chain of five small launches) runs on the side lane while the main lane already computes the
  // first layers that do not depend on z (actor.obs_net, F.obs_action_net, both backward nets); the phase end joins them.
  b.cur_lane = 1; b.fork_next = true;
  if (do_mix) {  // mix_z = backward_net(backward_input[perm]) on every row; rows outside the mask are ignored
    b.gemm({lin_fwd(bMix.x, pB.w(0), pB.v(1), bMix.pre, 0)});
    b.ln_fwd({b_ln(bMix, pB)});
    b.gemm({lin_fwd(bMix.y, pB.w(4), pB.v(5), bMix.h2, GF_RELU | GF_RELU_LAZY_OK)});
    b.gemm({lin_fwd(bMix.h2, pB.w(6), pB.v(7), bMix.raw, 0)});
    b.l2_fwd({b_l2(bMix, Z, nzB)});
  }
  if (rand_w) {   // mixed rows = random weighted sums of all B rows of backward_net(backward_input[perm])
    MixWeightParams mp; memset(&mp, 0, sizeof(mp));
    mp.batch = B; mp.Z = Z; mp.b = b_mix_out.p; mp.ldb = b_mix_out.ld; mp.W = h->mix_w.p; mp.ldw = h->mix_w.ld; mp.u = h->mix_u;
    mp.mix_mask = h->d_mix_mask; mp.out = b_mixw.p; mp.ldo = b_mixw.ld;
    const size_t smem = (size_t)FB_MIXW_TILE * Z * sizeof(float);
    b.push([mp, smem](cudaStream_t s) {
      fb_launch_pdl(k_mix_rand_weight, dim3(fb_ceil_div(mp.batch, 8)), dim3(256), smem, s, mp);
      return cudaGetLastError();
    }, FB_OPK_ELEMENTWISE, 2.0 * B * (double)B * Z, 4.0 * (B * (double)B + 2.0 * B * Z));
  }
  {
    ZFinalParams zp; memset(&zp, 0, sizeof(zp));
    zp.batch = B; zp.Z = Z; zp.O = O; zp.z_rand = h->z_rand.p; zp.ldZ = z.ld; zp.b_mix = b_mix_out.p; zp.ld_bmix = b_mix_out.ld;
    zp.mix_src = rand_w ? b_mixw.p : b_mix_out.p; zp.ld_mix_src = rand_w ? b_mixw.ld : b_mix_out.ld;
    zp.mix_mask = c.mix_ratio > 0.f ? h->d_mix_mask : nullptr; zp.future_mask = with_future ? h->d_future_mask : nullptr; zp.z = z.p; zp.actor_in_oz = actor_in_oz.p; zp.ldOZ = actor_in_oz.ld; zp.renorm = nz;
    if (deep) { zp.f_in[0] = in_oa.p; zp.f_in[1] = in_noa.p; zp.f_in[2] = in_oa2.p; zp.ldF = in_oa.ld; }
    b.push([zp](cudaStream_t s) { fb_launch_pdl(k_z_final, dim3(fb_ceil_div(zp.batch, 8)), dim3(256), 0, s, zp); return cudaGetLastError(); });
    { FsZFinalArgs fa; memset(&fa, 0, sizeof(fa)); fa.P = zp; h->ops[b.phase].back().dev.set(FS_Z_FINAL, fb_ceil_div(zp.batch, 8), fa); }
  }
  b.cur_lane = 0;
  // F's obs_action first layer does not read z in the default layout and runs here, next to the z-mixing chain; with
  // preprocess = False its input is [obs | z | action], so it waits for z and joins the first group of FB_FWD instead
  EmbedAct eFoa_early = eFoa, eFoa_late = eFoa;
  { EmbedAct& off = deep ? eFoa_early : eFoa_late; off.x.rows = 0; off.pre.rows = 0; off.y.rows = 0; }
  b.gemm({lin_fwd(eAo.x, pA.w(A_O + 0), pA.v(A_O + 1), eAo.pre, 0), lin_fwd(eFoa_early.x, pF.w(E_OA + 0), pF.v(E_OA + 1), eFoa_early.pre, 0),
          lin_fwd(bO.x, pB.w(0), pB.v(1), bO.pre, 0), lin_fwd(bT.x, pBt.w(0), pBt.v(1), bT.pre, 0)});
  b.ln_fwd({embed_ln(eAo, pA.sub(A_O)), embed_ln(eFoa_early, pF.sub(E_OA)), b_ln(bO, pB), b_ln(bT, pBt)});

  // =========================== FB_PHASE_FB_FWD ==================================================
  b.set_phase(FB_PHASE_FB_FWD);
  {   // loss / log-prob accumulators of update_fb: zeroed on the staging lane (consumers: k_actor_out here, the contraction later)
    TransposeDesc zd; memset(&zd, 0, sizeof(zd));
    zd.out = reinterpret_cast<float*>(acc); zd.rows = 1; zd.cols = 16; zd.ld_out = 16; zd.transpose = 2;
    b.add_early(zd, b.phase);
  }
  b.gemm({lin_fwd(eAoz.x, pA.w(A_OZ + 0), pA.v(A_OZ + 1), eAoz.pre, 0), lin_fwd(eFoz.x, pF.w(E_OZ + 0), pF.v(E_OZ + 1), eFoz.pre, 0),
          lin_fwd(eFtoz.x, pFt.w(E_OZ + 0), pFt.v(E_OZ + 1), eFtoz.pre, 0),
          lin_fwd(eFoa_late.x, pF.w(E_OA + 0), pF.v(E_OA + 1), eFoa_late.pre, 0),
          lin_fwd(pol.x, pA.w(0), pA.v(1), pol.pre, 0)});   // (cfg.boltzmann: policy.0)
  b.ln_fwd({embed_ln(eAoz, pA.sub(A_OZ)), embed_ln(eFoz, pF.sub(E_OZ)), embed_ln(eFtoz, pFt.sub(E_OZ)), embed_ln(eFoa_late, pF.sub(E_OA)),
            b_ln(pol, pA)});
  b.gemm({lin_fwd(eAo.y, pA.w(A_O + 4), pA.v(A_O + 5), eAo.out, GF_RELU | GF_RELU_LAZY_OK), lin_fwd(eAoz.y, pA.w(A_OZ + 4), pA.v(A_OZ + 5), eAoz.out, GF_RELU | GF_RELU_LAZY_OK),
          lin_fwd(eFoa.y, pF.w(E_OA + 4), pF.v(E_OA + 5), eFoa.out, GF_RELU | GF_RELU_LAZY_OK), lin_fwd(eFoz.y, pF.w(E_OZ + 4), pF.v(E_OZ + 5), eFoz.out, GF_RELU | GF_RELU_LAZY_OK),
          lin_fwd(eFtoz.y, pFt.w(E_OZ + 4), pFt.v(E_OZ + 5), eFtoz.out, GF_RELU | GF_RELU_LAZY_OK),
          lin_fwd(bO.y, pB.w(4), pB.v(5), bO.h2, GF_RELU | GF_RELU_LAZY_OK), lin_fwd(bT.y, pBt.w(4), pBt.v(5), bT.h2, GF_RELU | GF_RELU_LAZY_OK),
          lin_fwd(pol.y, pA.w(4), pA.v(5), pol.h2, GF_RELU | GF_RELU_LAZY_OK)});   // (cfg.boltzmann: policy.3)
  if (trunk) {
    b.gemm({lin_fwd(hA, pA.w(A_TR + 0), pA.v(A_TR + 1), trA, GF_RELU | GF_RELU_LAZY_OK), lin_fwd(hF, pF.w(F_TR + 0), pF.v(F_TR + 1), trF, GF_RELU | GF_RELU_LAZY_OK),
            lin_fwd(bO.h2, pB.w(6), pB.v(7), bO.raw, 0), lin_fwd(bT.h2, pBt.w(6), pBt.v(7), bT.raw, 0)});
    b.gemm({lin_fwd(trA, pA.w(A_POL + 0), pA.v(A_POL + 1), h1A, GF_RELU | GF_RELU_LAZY_OK)});
  } else {
    b.gemm({lin_fwd(hA, pA.w(A_POL + 0), pA.v(A_POL + 1), h1A, GF_RELU | GF_RELU_LAZY_OK),
            lin_fwd(bO.h2, pB.w(6), pB.v(7), bO.raw, 0), lin_fwd(bT.h2, pBt.w(6), pBt.v(7), bT.raw, 0)});
  }
  b.gemm({lin_fwd(h1A, pA.w(A_POL + 2), pA.v(A_POL + 3), preA, 0), lin_fwd(pol.h2, pA.w(6), pA.v(7), pol.raw, 0)});   // (policy.5)
  b.l2_fwd({b_l2(bO, Z, nzB), b_l2(bT, Z, nzB)});
  if (bz) {   // SquashedNormal samples of both sides + log pi of the obs side (fb_ddpg.py:304-306,391-396)
    ActorOutBzParams ap; memset(&ap, 0, sizeof(ap));
    ap.batch = B; ap.A = A; ap.act_col = act_col; ap.pre = pol.raw.p; ap.ldP = pol.raw.ld;
    ap.noise_fb = h->noise_fb.p; ap.noise_actor = h->noise_actor.p; ap.ldN = h->noise_fb.ld;
    ap.in_noa = in_noa.p; ap.in_oa2 = in_oa2.p; ap.ldOA = in_noa.ld; ap.next_action = next_action.p; ap.action_new = action_new.p;
    ap.ldA = next_action.ld; ap.xs = bz_x.p; ap.sds = bz_sd.p; ap.ts = bz_t.p;
    ap.log_std_min = c.log_std_min; ap.log_std_max = c.log_std_max; ap.acc = acc;
    b.push([ap](cudaStream_t s) {
      fb_launch_pdl(k_actor_out_bz, dim3(fb_ceil_div(2 * ap.batch * ap.A, 256)), dim3(256), 0, s, ap);
      return cudaGetLastError();
    });
    h->ops[b.phase].back().wait_stage = 1;   // the zeroed accumulators
  } else {
    ActorOutParams ap; memset(&ap, 0, sizeof(ap));
    ap.batch = B; ap.A = A; ap.O = O; ap.act_col = act_col; ap.pre = preA.p; ap.mu = mu.p; ap.ldA = preA.ld;
    ap.noise_fb = h->noise_fb.p; ap.noise_actor = h->noise_actor.p; ap.ldN = h->noise_fb.ld;
    ap.in_noa = in_noa.p; ap.in_oa2 = in_oa2.p; ap.ldOA = in_noa.ld; ap.next_action = next_action.p; ap.action_new = action_new.p;
    ap.acc = acc;
    b.push([ap, sc](cudaStream_t s) {
      fb_launch_pdl(k_actor_out, dim3(fb_ceil_div(2 * ap.batch * ap.A, 256)), dim3(256), 0, s, ap, sc);
      return cudaGetLastError();
    });
    { FsActorOutArgs fa; memset(&fa, 0, sizeof(fa)); fa.P = ap; fa.sc = sc; h->ops[b.phase].back().dev.set(FS_ACTOR_OUT, fb_ceil_div(2 * ap.batch * ap.A, 256), fa); }
    h->ops[b.phase].back().wait_stage = 1;   // the zeroed accumulators
  }
  b.gemm({lin_fwd(eFtoa.x, pFt.w(E_OA + 0), pFt.v(E_OA + 1), eFtoa.pre, 0)});
  b.ln_fwd({embed_ln(eFtoa, pFt.sub(E_OA))});
  b.gemm({lin_fwd(eFtoa.y, pFt.w(E_OA + 4), pFt.v(E_OA + 5), eFtoa.out, GF_RELU | GF_RELU_LAZY_OK)});
  if (trunk) b.gemm({lin_fwd(hFt, pFt.w(F_TR + 0), pFt.v(F_TR + 1), trFt, GF_RELU | GF_RELU_LAZY_OK)});
  b.gemm({lin_fwd(inFt, pFt.w(HD_1 + 0), pFt.v(HD_1 + 1), h1Ft1, GF_RELU | GF_RELU_LAZY_OK), lin_fwd(inFt, pFt.w(HD_2 + 0), pFt.v(HD_2 + 1), h1Ft2, GF_RELU | GF_RELU_LAZY_OK),
          lin_fwd(inF, pF.w(HD_1 + 0), pF.v(HD_1 + 1), h1F1, GF_RELU | GF_RELU_LAZY_OK), lin_fwd(inF, pF.w(HD_2 + 0), pF.v(HD_2 + 1), h1F2, GF_RELU | GF_RELU_LAZY_OK)});
  b.gemm({lin_fwd(h1Ft1, pFt.w(HD_1 + 2), pFt.v(HD_1 + 3), tF1, 0), lin_fwd(h1Ft2, pFt.w(HD_2 + 2), pFt.v(HD_2 + 3), tF2, 0),
          lin_fwd(h1F1, pF.w(HD_1 + 2), pF.v(HD_1 + 3), F1, 0), lin_fwd(h1F2, pF.w(HD_2 + 2), pF.v(HD_2 + 3), F2, 0)});

  // multi-GPU exchange 1: every rank's [F1|F2|tF1|tF2|B|tB|discount] rows -> blk_global, in rank order.  Inside the step (and its
  // CUDA graph) when the library owns a communicator (fb_nccl_init); otherwise the caller all-gathers between FB_FWD and FB_LOSS.
  if (h->nccl_comm && n != B) {
    fb_handle* hh = h;
    const float* src = bl.p; float* dst = bg.p; const size_t count = (size_t)B * bl.ld;
    b.push([hh, src, dst, count](cudaStream_t s) {
      return g_nccl.AllGather(src, dst, count, 7 /* ncclFloat32 */, hh->nccl_comm, s) == 0 ? cudaSuccess : cudaErrorUnknown;
    }, FB_OPK_COLLECTIVE, 0.0, 4.0 * (double)n * bl.ld);
  } else if (h->p2p_on && n != B) {
    // peer-memory form (p2p.cuh): every rank is past the previous step (nobody still reads its global block), then each rank
    // stores its rows into every arena, then waits until all R row blocks have landed in its own
    const P2pPeers pp = h->p2p;
    const float4* src = reinterpret_cast<const float4*>(bl.p);
    const size_t off_blk = h->p2p_off_blk, n4 = (size_t)B * bl.ld / 4;
    b.push([pp](cudaStream_t s) { fb_launch_pdl(k_p2p_barrier, dim3(1), dim3(32), 0, s, pp, (int)P2P_BAR_STEP, 1, 1); return cudaGetLastError(); },
           FB_OPK_COLLECTIVE);
    b.push([pp, src, off_blk, n4](cudaStream_t s) {
      fb_launch_pdl(k_p2p_scatter_rows, dim3((unsigned)std::min<size_t>(FB_SM_COUNT, (n4 + 255) / 256)), dim3(256), 0, s, pp, src, off_blk, n4);
      return cudaGetLastError();
    }, FB_OPK_COLLECTIVE, 0.0, 4.0 * (double)n * bl.ld);
    b.push([pp](cudaStream_t s) { fb_launch_pdl(k_p2p_barrier, dim3(1), dim3(32), 0, s, pp, (int)P2P_BAR_ROWS, 0, 1); return cudaGetLastError(); },
           FB_OPK_COLLECTIVE);
  }

  // =========================== FB_PHASE_FB_LOSS =================================================
  b.set_phase(FB_PHASE_FB_LOSS);
  const bool use_tc = c.contract_mode == FB_CONTRACT_TCGEN05 && Z <= 128;
  const float inv_noff = 1.0f / ((float)n * (float)(n - 1)), inv_n = 1.0f / (float)n;
  double* q_inv = nullptr;
  if (c.q_loss) {
    // Q loss (fb_ddpg.py:330-341): cov = B^T B / n over the GLOBAL batch and its inverse depend on B only -> side lane, next to
    // the contraction; the row pass that adds the term to dF_k opens FB_BWD (the phase end joins the lanes)
    double* q_cov = (double*)ws_alloc(h, (size_t)Z * Z * sizeof(double));
    q_inv = (double*)ws_alloc(h, (size_t)Z * Z * sizeof(double));
    const float* Ball = Bg.p; const int ldb = Bg.ld;
    b.push([Ball, ldb, n, Z, q_cov](cudaStream_t s) {
      fb_launch_pdl(k_qloss_cov, dim3(Z), dim3(256), 0, s, Ball, ldb, n, Z, q_cov);
      return cudaGetLastError();
    }, FB_OPK_ELEMENTWISE, 2.0 * n * (double)Z * Z, 4.0 * n * (double)Z, 1);
    const size_t inv_smem = (size_t)Z * 2 * Z * sizeof(double);
    h->qloss_smem = inv_smem;
    b.push([q_cov, q_inv, Z, inv_smem](cudaStream_t s) {
      fb_launch_pdl(k_qloss_inverse, dim3(1), dim3(FB_QLOSS_INV_THREADS), inv_smem, s, (const double*)q_cov, Z, q_inv);
      return cudaGetLastError();
    }, FB_OPK_ELEMENTWISE, 2.0 * (double)Z * Z * Z, 16.0 * Z * (double)Z, 1);
  }
  if (use_tc) {
    // tcgen05 path (contract_tc.cuh): split operands -> fused contraction + loss + dL/dM tiles
    const int nbox = fb_ceil_div(Z, 32), KP = nbox * 32;
    float* x2 = (float*)ws_alloc(h, (size_t)CT_NUM_OPERANDS * n * 2 * KP * sizeof(float));
    {
      const float* blk = bg.p; const int pitch = bg.ld;
      b.push([blk, pitch, ldZ, n, Z, KP, x2](cudaStream_t s) {
        fb_launch_pdl(k_contract_split, dim3(fb_ceil_div(CT_NUM_OPERANDS * n * KP, 256)), dim3(256), 0, s, blk, pitch, ldZ, n, Z, KP, x2);
        return cudaGetLastError();
      }, FB_OPK_ELEMENTWISE, 0.0, 4.0 * CT_NUM_OPERANDS * n * 3.0 * KP);
    }
    ContractParams cp; memset(&cp, 0, sizeof(cp));
    if (h->ws_base) {
      for (int m = 0; m < CT_NUM_OPERANDS; ++m) {
        int rc = encode_operand_map(&cp.maps[m], x2 + (size_t)m * n * 2 * KP, n, 2 * KP);
        if (rc != FB_OK) return rc;
      }
    }
    cp.nbox = nbox; cp.ksteps = fb_ceil_div(Z, 8); cp.ld = M1.ld; cp.inv_noff = inv_noff; cp.inv_n = inv_n;
    cp.c4 = 4.0f * c.ortho_coef * inv_noff; cp.acc = acc; cp.diag0 = c.row_offset;
    const size_t smem_bytes = CT_SMEM_BYTES;   // 4 ring stages of one K-box each + the epilogue scratch
    h->contract_smem = smem_bytes;
    const int ct_sms = h->sm_count > 0 ? h->sm_count : FB_SM_COUNT;
    const double tile_flops = 2.0 * 3.0 * cp.ksteps * 8.0;  // per pair, per product (3 chains)
    {
      ContractParams r = cp;
      const int pa[5] = {CT_F1, CT_F2, CT_TF1, CT_TF2, CT_B}, pb[5] = {CT_B, CT_B, CT_TB, CT_TB, CT_B};
      for (int i = 0; i < 5; ++i) { r.prod_a[i] = pa[i]; r.prod_b[i] = pb[i]; }
      r.n_products = 5; r.mode = CT_MODE_ROW; r.nr = B; r.nc = n; r.a_row0 = c.row_offset;
      r.G1 = M1.p; r.G2 = M2.p; r.Gc = Cov.p;
      r.Gt1 = (n == B) ? Mt1.p : nullptr; r.Gt2 = (n == B) ? Mt2.p : nullptr;
      r.disc = bl.p + disc_col; r.disc_stride = bl.ld;
      b.push([r, smem_bytes, ct_sms](cudaStream_t s) {
        const int tiles = fb_ceil_div(r.nc, CT_TILE_N) * fb_ceil_div(r.nr, CT_TILE_M);   // persistent: one CTA per SM walks them
        fb_launch_pdl(k_contract_tc, dim3(std::min(tiles, ct_sms)), dim3(CT_THREADS), smem_bytes, s, r);
        return cudaGetLastError();
      }, FB_OPK_CONTRACT, 5.0 * tile_flops * B * (double)n, 4.0 * (5.0 * B * (double)n + 6.0 * n * 2.0 * KP));
    }
    if (n != B) {  // multi-GPU: the column block of dL/dM (rows = local t, columns = all global s) for dB
      ContractParams r = cp;
      const int pa[4] = {CT_B, CT_B, CT_TB, CT_TB}, pb[4] = {CT_F1, CT_F2, CT_TF1, CT_TF2};
      for (int i = 0; i < 4; ++i) { r.prod_a[i] = pa[i]; r.prod_b[i] = pb[i]; }
      r.n_products = 4; r.mode = CT_MODE_COL; r.nr = B; r.nc = n; r.a_row0 = c.row_offset;
      r.G1 = Mt1.p; r.G2 = Mt2.p;
      r.disc = bg.p + disc_col; r.disc_stride = bg.ld;
      b.push([r, smem_bytes, ct_sms](cudaStream_t s) {
        const int tiles = fb_ceil_div(r.nc, CT_TILE_N) * fb_ceil_div(r.nr, CT_TILE_M);   // persistent: one CTA per SM walks them
        fb_launch_pdl(k_contract_tc, dim3(std::min(tiles, ct_sms)), dim3(CT_THREADS), smem_bytes, s, r);
        return cudaGetLastError();
      }, FB_OPK_CONTRACT, 4.0 * tile_flops * B * (double)n, 4.0 * (2.0 * B * (double)n + 6.0 * n * 2.0 * KP));
    }
  } else {
    auto outer = [&](const Mat& X, const Mat& Yall, const Mat& C) {  // C[rows(X), n] = X . Yall^T  (K = Z)
      return gemm_raw(X.p, X.ld, 1, Yall.p, Yall.ld, 1, C.p, C.ld, X.rows, Yall.rows, Z, nullptr, 0, nullptr, 0);
    };
    b.gemm({outer(F1, Bg, M1), outer(F2, Bg, M2), outer(tF1, tBg, T1), outer(tF2, tBg, T2), outer(Bm, Bg, Cov),
            outer(Bm, F1g, Mt1), outer(Bm, F2g, Mt2), outer(tB, tF1g, Tt1), outer(tB, tF2g, Tt2)});
    LossElemParams lp; memset(&lp, 0, sizeof(lp));
    lp.M1 = M1.p; lp.M2 = M2.p; lp.T1 = T1.p; lp.T2 = T2.p; lp.Cov = Cov.p; lp.nr = B; lp.nc = n; lp.ld = M1.ld; lp.row0 = c.row_offset;
    lp.disc = bl.p + disc_col; lp.disc_stride = bl.ld; lp.inv_noff = inv_noff; lp.inv_n = inv_n; lp.ortho_coef = c.ortho_coef; lp.acc = acc;
    b.push([lp](cudaStream_t s) {
      dim3 grid(fb_ceil_div(lp.nc, 1024) > 0 ? fb_ceil_div(lp.nc, 1024) : 1, lp.nr < 592 ? lp.nr : 592);
      fb_launch_pdl(k_fb_loss_elem, dim3(grid), dim3(256), 0, s, lp);
      return cudaGetLastError();
    }, FB_OPK_LOSS, 0.0, 4.0 * 8.0 * (double)B * (double)n);
    LossElemTParams lt; memset(&lt, 0, sizeof(lt));
    lt.M1 = Mt1.p; lt.M2 = Mt2.p; lt.T1 = Tt1.p; lt.T2 = Tt2.p; lt.nr = B; lt.nc = n; lt.ld = Mt1.ld; lt.row0 = c.row_offset;
    lt.disc = bg.p + disc_col; lt.disc_stride = bg.ld; lt.inv_noff = inv_noff; lt.inv_n = inv_n;
    b.push([lt](cudaStream_t s) {
      dim3 grid(fb_ceil_div(lt.nc, 1024) > 0 ? fb_ceil_div(lt.nc, 1024) : 1, lt.nr < 592 ? lt.nr : 592);
      fb_launch_pdl(k_fb_loss_elem_t, dim3(grid), dim3(256), 0, s, lt);
      return cudaGetLastError();
    }, FB_OPK_LOSS, 0.0, 4.0 * 6.0 * (double)B * (double)n);
  }
  // second stage: dF_k = G_k . B_all,  dB = Gt_1 . F1_all + Gt_2 . F2_all + Gc . B_all  (+ the diagonal term, added by k_l2norm_bwd)
  const float db_coef = -4.0f * c.ortho_coef * inv_n;   // d(-2 c mean_s Cov_ss)/dB_s = -(4c/n) B_s
  Mat dBparts = ws_mat(h, B, 3 * ldZ, "dBparts");
  const bool tc_inner = c.mlp_mode == FB_MLP_TCGEN05;
  if (tc_inner) {
    auto inner = [&](const Mat& Gm, const Mat& Yall, const Mat& C) {  // C[B, Z] = Gm[B, n] . Yall[n, Z]
      return gemm_raw(Gm.p, Gm.ld, 1, Yall.p, Yall.ld, 0, C.p, C.ld, Gm.rows, Z, n, nullptr, 0, nullptr, 0);
    };
    b.gemm({inner(M1, Bg, dF1), inner(M2, Bg, dF2), inner(Mt1, F1g, dBparts.cs(0, Z)), inner(Mt2, F2g, dBparts.cs(ldZ, Z)),
            inner(Cov, Bg, dBparts.cs(2 * ldZ, Z))});
  } else {
    b.memset0(dblk.p, (size_t)dblk.rows * dblk.ld * sizeof(float));
    b.push([dB, Bm, B, Z, db_coef](cudaStream_t s) {   // dB starts from the diagonal term, the products accumulate onto it
      fb_launch_pdl(k_loss_init_db, dim3(fb_ceil_div(B * Z, 256)), dim3(256), 0, s, dB.p, dB.ld, Bm.p, Bm.ld, B, Z, db_coef);
      return cudaGetLastError();
    });
    auto inner = [&](const Mat& Gm, const Mat& Yall, const Mat& C) {  // C[B, Z] += Gm[B, n] . Yall[n, Z]
      return gemm_raw(Gm.p, Gm.ld, 1, Yall.p, Yall.ld, 0, C.p, C.ld, Gm.rows, Z, n, nullptr, GF_ATOMIC | GF_SHARED_C, nullptr, 0);
    };
    b.gemm({inner(M1, Bg, dF1), inner(M2, Bg, dF2), inner(Mt1, F1g, dB), inner(Mt2, F2g, dB), inner(Cov, Bg, dB)});
  }

  // =========================== FB_PHASE_FB_BWD ==================================================
  b.set_phase(FB_PHASE_FB_BWD);
  if (c.q_loss) {
    QLossParams qp; memset(&qp, 0, sizeof(qp));
    qp.F1 = F1.p; qp.F2 = F2.p; qp.tF1 = tF1.p; qp.tF2 = tF2.p; qp.Bm = Bm.p; qp.ldblk = bl.ld;
    qp.disc = bl.p + disc_col; qp.disc_stride = bl.ld; qp.z = z.p; qp.ldz = z.ld;
    qp.dF1 = dF1.p; qp.dF2 = dF2.p; qp.lddf = dF1.ld; qp.inv = q_inv; qp.rows = B; qp.Z = Z;
    qp.gcoef = c.q_loss_coef * 2.0f * inv_n; qp.acc = acc;
    b.push([qp](cudaStream_t s) {
      fb_launch_pdl(k_qloss_rows, dim3(fb_ceil_div(qp.rows, 8)), dim3(256), 0, s, qp);
      return cudaGetLastError();
    }, FB_OPK_LOSS, 2.0 * B * (double)Z * Z, 4.0 * 8.0 * B * (double)Z);
  }
  {
    const float* p0 = tc_inner ? dBparts.p : dB.p;
    const float* p1 = tc_inner ? dBparts.p + ldZ : nullptr;
    const float* p2 = tc_inner ? dBparts.p + 2 * ldZ : nullptr;
    const int ldp = tc_inner ? dBparts.ld : dB.ld;
    float* dsum = tc_inner ? dB.p : nullptr;   // keep the "dB" view complete on both paths
    const float coef = tc_inner ? db_coef : 0.f;
    b.push([p0, p1, p2, ldp, coef, dsum, dB, Bm, bO, draw, B, Z, nzB](cudaStream_t s) {
      fb_launch_pdl(k_l2norm_bwd, dim3(fb_ceil_div(B, 8)), dim3(256), 0, s, p0, p1, p2, ldp, coef, dsum, dB.ld, Bm.p, Bm.ld, bO.nrm, draw.p, draw.ld, B, Z, nzB);
      return cudaGetLastError();
    });
    {
      FsL2BwdArgs fa; memset(&fa, 0, sizeof(fa));
      fa.dy0 = p0; fa.dy1 = p1; fa.dy2 = p2; fa.dsum = dsum; fa.y = Bm.p; fa.nrm = bO.nrm; fa.dx = draw.p;
      fa.lddy = ldp; fa.ldsum = dB.ld; fa.ldy = Bm.ld; fa.lddx = draw.ld; fa.rows = B; fa.Z = Z; fa.normalize = nzB; fa.coef = coef;
      h->ops[b.phase].back().dev.set(FS_L2_BWD, fb_ceil_div(B, 8), fa);
    }
  }
  Mat draw_c = draw;   // the gradient w.r.t. backward_net's last Linear output, as its backward launches see it (cfg.debug: nobody)
  if (dbg) draw_c.rows = 0;
  b.colsum({mk_colsum(dF1, pF.gv(HD_1 + 3)), mk_colsum(dF2, pF.gv(HD_2 + 3)), mk_colsum(draw_c, pB.gv(7))});
  b.gemm({lin_dw(dF1, h1F1, pF.gw(HD_1 + 2)), lin_dw(dF2, h1F2, pF.gw(HD_2 + 2)), lin_dw(draw_c, bO.h2, pB.gw(6)),
          lin_dx(dF1, pF.w(HD_1 + 2), dh1_1, GF_MASK_RELU, &h1F1), lin_dx(dF2, pF.w(HD_2 + 2), dh1_2, GF_MASK_RELU, &h1F2),
          lin_dx(draw_c, pB.w(6), dh2, GF_MASK_RELU, &bO.h2)});
  b.colsum({mk_colsum(dh1_1, pF.gv(HD_1 + 1)), mk_colsum(dh1_2, pF.gv(HD_2 + 1)), mk_colsum(dh2, pB.gv(5))});
  {
    // both heads' dX products land in one buffer (K2 = the second head): the gradient of the heads' common input, masked by its ReLU
    GemmDesc d = lin_dx(dh1_1, pF.w(HD_1 + 0), trunk ? dtF : dhF, GF_MASK_RELU, &inF);
    Mat w2 = pF.w(HD_2 + 0);
    d.A2 = dh1_2.p; d.B2 = w2.p; d.K2 = w2.rows;
    b.gemm({lin_dw(dh1_1, inF, pF.gw(HD_1 + 0)), lin_dw(dh1_2, inF, pF.gw(HD_2 + 0)), d, lin_dw(dh2, bO.y, pB.gw(4)),
            lin_dx(dh2, pB.w(4), dy1, 0, nullptr)});
  }
  if (trunk) {   // through the trunk: bias / weight gradients, then the gradient of the concatenated embeds
    b.colsum({mk_colsum(dtF, pF.gv(F_TR + 1))});
    b.gemm({lin_dw(dtF, hF, pF.gw(F_TR + 0)), lin_dx(dtF, pF.w(F_TR + 0), dhF, GF_MASK_RELU, &hF)});
  }
  Mat dhF_oa = dhF.cs(0, Fe), dhF_oz = dhF.cs(oz0, Fe);
  if (deep) dhF_oz.rows = 0;
  b.colsum({mk_colsum(dhF_oa, pF.gv(E_OA + 5)), mk_colsum(dhF_oz, pF.gv(E_OZ + 5))});
  {
    LnBwdDesc d; memset(&d, 0, sizeof(d));
    d.dy = dy1.p; d.dx = dy1.p; d.y = bO.y.p; d.x = bO.pre.p; d.gamma = pB.v(2); d.mean = bO.mean; d.rstd = bO.rstd;
    d.dgamma = pB.gv(2); d.dbeta = pB.gv(3); d.rows = dy1.rows; d.D = dy1.cols; d.ld = bO.y.ld; d.ld_dy = dy1.ld;
    b.ln_bwd({d});
  }
  b.colsum({mk_colsum(dy1, pB.gv(1))});
  b.gemm({lin_dw(dhF_oa, eFoa.y, pF.gw(E_OA + 4)), lin_dw(dhF_oz, eFoz.y, pF.gw(E_OZ + 4)),
          lin_dx(dhF_oa, pF.w(E_OA + 4), dy_oa, 0, nullptr), lin_dx(dhF_oz, pF.w(E_OZ + 4), dy_oz, 0, nullptr),
          lin_dw(dy1, bO.x, pB.gw(0))});
  b.ln_bwd({embed_ln_bwd(eFoa, pF.sub(E_OA), dy_oa, 0, true), embed_ln_bwd(eFoz, pF.sub(E_OZ), dy_oz, 0, true)});
  b.colsum({mk_colsum(dy_oa, pF.gv(E_OA + 1)), mk_colsum(dy_oz, pF.gv(E_OZ + 1))});
  b.gemm({lin_dw(dy_oa, eFoa.x, pF.gw(E_OA + 0)), lin_dw(dy_oz, eFoz.x, pF.gw(E_OZ + 0))});

  // multi-GPU exchange 2: the flat forward_net | backward_net gradient, summed over ranks (gradients only cross NVLink)
  const bool p2p_step = h->p2p_on && n != B;
  if (h->nccl_comm && n != B) {
    fb_handle* hh = h;
    float* gbuf = bf.d_grad_fb; const size_t count = h->seg_fb.size;
    b.push([hh, gbuf, count](cudaStream_t s) {
      return g_nccl.AllReduce(gbuf, gbuf, count, 7, 0 /* ncclSum */, hh->nccl_comm, s) == 0 ? cudaSuccess : cudaErrorUnknown;
    }, FB_OPK_COLLECTIVE, 0.0, 8.0 * (double)count);
  } else if (p2p_step) {   // this rank's fb gradient is final: tell the peers, wait for theirs
    const P2pPeers pp = h->p2p;
    b.push([pp](cudaStream_t s) { fb_launch_pdl(k_p2p_barrier, dim3(1), dim3(32), 0, s, pp, (int)P2P_BAR_GRAD_FB, 1, 1); return cudaGetLastError(); },
           FB_OPK_COLLECTIVE);
  }

  // =========================== FB_PHASE_FB_ADAM =================================================
  b.set_phase(FB_PHASE_FB_ADAM);
  {
    const float b1 = c.beta1, b2 = c.beta2, eps = c.adam_eps;
    float4 *p = (float4*)bf.d_param_fb, *g = (float4*)bf.d_grad_fb, *m = (float4*)bf.d_m_fb, *v = (float4*)bf.d_v_fb, *t = (float4*)bf.d_target_fb;
    const size_t n4 = h->seg_fb.size / 4, split4 = h->bwd_offset / 4;
    if (p2p_step) {
      // reduce-scatter + Adam + all-gather over the arenas in one kernel, then (all slices landed) the local target lerp + gradient clear
      const P2pPeers pp = h->p2p;
      P2pAdamParams ap; memset(&ap, 0, sizeof(ap));
      ap.off_grad = h->p2p_off_grad_fb; ap.off_param = h->p2p_off_param_fb; ap.m = m; ap.v = v; ap.n4 = n4; ap.split4 = split4;
      ap.slice4 = (n4 + pp.world - 1) / pp.world; ap.sc = sc; ap.which = 0; ap.bar_param = P2P_BAR_PARAM_FB; ap.beta1 = b1; ap.beta2 = b2; ap.eps = eps;
      b.push([pp, ap](cudaStream_t s) {
        fb_launch_pdl(k_p2p_adam, dim3(p2p_adam_ctas()), dim3(256), 0, s, pp, ap);
        return cudaGetLastError();
      }, FB_OPK_ADAM, 0.0, 16.0 * (double)ap.slice4 * (pp.world + 5.0 + pp.world));   // r(g x R, p, m, v) + w(m, v, p x R) of one slice
      b.push([pp](cudaStream_t s) { fb_launch_pdl(k_p2p_barrier, dim3(1), dim3(32), 0, s, pp, (int)P2P_BAR_PARAM_FB, 0, 1); return cudaGetLastError(); },
             FB_OPK_COLLECTIVE);
      b.push([=](cudaStream_t s) {
        fb_launch_pdl(k_p2p_finish, dim3(FB_SM_COUNT * 8), dim3(256), 0, s, (const float4*)p, g, t, n4, (const DevScalars*)sc);
        return cudaGetLastError();
      }, FB_OPK_ADAM, 0.0, 16.0 * (double)n4 * 4.0);
    } else {
    b.push([=](cudaStream_t s) {
      fb_launch_pdl(k_adam, dim3(FB_SM_COUNT * 8), dim3(256), 0, s, p, g, m, v, t, n4, split4, sc, 0, b1, b2, eps);
      return cudaGetLastError();
    }, FB_OPK_ADAM, 0.0, 4.0 * 4.0 * (double)n4 * 10.0);  // r(p,g,m,v,target) + w(p,g,m,v,target)
    // Adam step count + the next step's bias corrections: one thread on the side lane, behind k_adam (its CTAs read the scalars)
    b.push([=](cudaStream_t s) { fb_launch_pdl(k_tick, dim3(1), dim3(32), 0, s, sc, 0, b1, b2); return cudaGetLastError(); },
           FB_OPK_ELEMENTWISE, 0.0, 0.0, 1);
    }
  }

  // =========================== FB_PHASE_ACTOR_FWD ===============================================
  b.set_phase(FB_PHASE_ACTOR_FWD);
  {   // Q accumulator of update_actor: zeroed on the staging lane (consumer: k_actor_q)
    TransposeDesc zd; memset(&zd, 0, sizeof(zd));
    zd.out = reinterpret_cast<float*>(acc + ACC_Q); zd.rows = 1; zd.cols = 4; zd.ld_out = 4; zd.transpose = 2;
    b.add_early(zd, b.phase);
  }
  b.gemm({lin_fwd(eF2oa.x, pF.w(E_OA + 0), pF.v(E_OA + 1), eF2oa.pre, 0), lin_fwd(eF2oz.x, pF.w(E_OZ + 0), pF.v(E_OZ + 1), eF2oz.pre, 0)});
  b.ln_fwd({embed_ln(eF2oa, pF.sub(E_OA)), embed_ln(eF2oz, pF.sub(E_OZ))});
  b.gemm({lin_fwd(eF2oa.y, pF.w(E_OA + 4), pF.v(E_OA + 5), eF2oa.out, GF_RELU | GF_RELU_LAZY_OK), lin_fwd(eF2oz.y, pF.w(E_OZ + 4), pF.v(E_OZ + 5), eF2oz.out, GF_RELU | GF_RELU_LAZY_OK)});
  if (trunk) b.gemm({lin_fwd(hF2, pF.w(F_TR + 0), pF.v(F_TR + 1), trF2, GF_RELU | GF_RELU_LAZY_OK)});
  b.gemm({lin_fwd(inF2, pF.w(HD_1 + 0), pF.v(HD_1 + 1), h1Fa1, GF_RELU | GF_RELU_LAZY_OK), lin_fwd(inF2, pF.w(HD_2 + 0), pF.v(HD_2 + 1), h1Fa2, GF_RELU | GF_RELU_LAZY_OK)});
  b.gemm({lin_fwd(h1Fa1, pF.w(HD_1 + 2), pF.v(HD_1 + 3), Fa1, 0), lin_fwd(h1Fa2, pF.w(HD_2 + 2), pF.v(HD_2 + 3), Fa2, 0)});
  {
    const float inv_n = 1.0f / (float)n;
    b.push([Fa1, Fa2, z, dFa1, dFa2, B, Z, inv_n, acc](cudaStream_t s) {
      fb_launch_pdl(k_actor_q, dim3(fb_ceil_div(B, 8)), dim3(256), 0, s, Fa1.p, Fa2.p, Fa1.ld, z.p, z.ld, dFa1.p, dFa2.p, dFa1.ld, B, Z, inv_n, acc);
      return cudaGetLastError();
    });
    {
      FsActorQArgs fa; memset(&fa, 0, sizeof(fa));
      fa.F1 = Fa1.p; fa.F2 = Fa2.p; fa.z = z.p; fa.dF1 = dFa1.p; fa.dF2 = dFa2.p; fa.acc = acc;
      fa.ldf = Fa1.ld; fa.ldz = z.ld; fa.lddf = dFa1.ld; fa.rows = B; fa.Z = Z; fa.inv_n = inv_n;
      h->ops[b.phase].back().dev.set(FS_ACTOR_Q, fb_ceil_div(B, 8), fa);
    }
    h->ops[b.phase].back().wait_stage = 1;   // the zeroed accumulator
  }

  // =========================== FB_PHASE_ACTOR_BWD ===============================================
  b.set_phase(FB_PHASE_ACTOR_BWD);
  b.gemm({lin_dx(dFa1, pF.w(HD_1 + 2), dh1_1, GF_MASK_RELU, &h1Fa1), lin_dx(dFa2, pF.w(HD_2 + 2), dh1_2, GF_MASK_RELU, &h1Fa2)});
  {
    Mat hF2oa = hF2.cs(0, Fe);
    if (trunk) {   // heads -> trunk output (full width), then the obs_action half of the trunk's input
      GemmDesc d = lin_dx(dh1_1, pF.w(HD_1 + 0), dtF, GF_MASK_RELU, &trF2);
      Mat w2 = pF.w(HD_2 + 0);
      d.A2 = dh1_2.p; d.B2 = w2.p; d.K2 = w2.rows;
      b.gemm({d});
      b.gemm({lin_dx(dtF, pF.w(F_TR + 0).cs(0, Fe), dhoa, GF_MASK_RELU, &hF2oa)});
    } else {
      GemmDesc d = lin_dx(dh1_1, pF.w(HD_1 + 0).cs(0, Fd), dhoa, GF_MASK_RELU, &hF2oa);
      Mat w2 = pF.w(HD_2 + 0).cs(0, Fd);
      d.A2 = dh1_2.p; d.B2 = w2.p; d.K2 = w2.rows;
      b.gemm({d});
    }
  }
  b.gemm({lin_dx(dhoa, pF.w(E_OA + 4), dy_oa, 0, nullptr)});
  b.ln_bwd({embed_ln_bwd(eF2oa, pF.sub(E_OA), dy_oa, 0, false)});
  if (bz) {
    // d(-mean Q)/d action, then the SquashedNormal / log-std chain rule and the entropy term (k_boltz_bwd), then the
    // DiagGaussianActor stack backwards: policy.5 -> ReLU -> policy.3 -> LayerNorm + tanh -> policy.0
    b.gemm({lin_dx(dy_oa, pF.w(E_OA + 0).cs(act_col, A), bz_da, 0, nullptr)});
    BoltzBwdParams bp; memset(&bp, 0, sizeof(bp));
    bp.batch = B; bp.A = A; bp.da = bz_da.p; bp.ldda = bz_da.ld; bp.xs = bz_x.p; bp.sds = bz_sd.p; bp.ts = bz_t.p; bp.ldA = bz_x.ld;
    bp.noise_actor = h->noise_actor.p; bp.ldN = h->noise_actor.ld; bp.dpre = bz_dpre.p; bp.ldP = bz_dpre.ld;
    bp.temp_over_n = c.temp / (float)n; bp.half_range = 0.5f * (c.log_std_max - c.log_std_min);
    b.push([bp](cudaStream_t s) {
      fb_launch_pdl(k_boltz_bwd, dim3(fb_ceil_div(bp.batch * bp.A, 256)), dim3(256), 0, s, bp);
      return cudaGetLastError();
    });
    Mat h2_o = pol.h2.rs(B, B), y_o = pol.y.rs(B, B), x_o = pol.x.rs(B, B);
    b.colsum({mk_colsum(bz_dpre, pA.gv(7))});
    b.gemm({lin_dw(bz_dpre, h2_o, pA.gw(6)), lin_dx(bz_dpre, pA.w(6), bz_dh2, GF_MASK_RELU, &h2_o)});
    b.colsum({mk_colsum(bz_dh2, pA.gv(5))});
    b.gemm({lin_dw(bz_dh2, y_o, pA.gw(4)), lin_dx(bz_dh2, pA.w(4), bz_dy, 0, nullptr)});
    {
      LnBwdDesc d; memset(&d, 0, sizeof(d));
      d.dy = bz_dy.p; d.dx = bz_dy.p; d.y = y_o.p; d.x = pol.pre.p + (size_t)B * pol.pre.ld; d.gamma = pA.v(2); d.mean = pol.mean + B; d.rstd = pol.rstd + B;
      d.dgamma = pA.gv(2); d.dbeta = pA.gv(3); d.rows = B; d.D = bz_dy.cols; d.ld = pol.y.ld; d.ld_dy = bz_dy.ld;
      b.ln_bwd({d});
    }
    b.colsum({mk_colsum(bz_dy, pA.gv(1))});
    b.gemm({lin_dw(bz_dy, x_o, pA.gw(0))});
  } else {
    Mat muB = mu.rs(B, B);
    b.gemm({lin_dx(dy_oa, pF.w(E_OA + 0).cs(act_col, A), dpreA, GF_MASK_TANH, &muB)});
  }
  Mat h1A_o = h1A.rs(B, B), hA_o = hA.rs(B, B);
  b.colsum({mk_colsum(dpreA, pA.gv(A_POL + 3))});
  b.gemm({lin_dw(dpreA, h1A_o, pA.gw(A_POL + 2)), lin_dx(dpreA, pA.w(A_POL + 2), dh1A, GF_MASK_RELU, &h1A_o)});
  b.colsum({mk_colsum(dh1A, pA.gv(A_POL + 1))});
  if (trunk) {
    Mat tA_o = trA.rs(B, B);
    b.gemm({lin_dw(dh1A, tA_o, pA.gw(A_POL + 0)), lin_dx(dh1A, pA.w(A_POL + 0), dtA, GF_MASK_RELU, &tA_o)});
    b.colsum({mk_colsum(dtA, pA.gv(A_TR + 1))});
    b.gemm({lin_dw(dtA, hA_o, pA.gw(A_TR + 0)), lin_dx(dtA, pA.w(A_TR + 0), dhA, GF_MASK_RELU, &hA_o)});
  } else {
    b.gemm({lin_dw(dh1A, hA_o, pA.gw(A_POL + 0)), lin_dx(dh1A, pA.w(A_POL + 0), dhA, GF_MASK_RELU, &hA_o)});
  }
  Mat dhA_o = dhA.cs(0, Fe), dhA_oz = dhA.cs(oz0, Fe);
  if (deep) dhA_o.rows = 0;
  b.colsum({mk_colsum(dhA_o, pA.gv(A_O + 5)), mk_colsum(dhA_oz, pA.gv(A_OZ + 5))});
  b.gemm({lin_dw(dhA_o, eAo.y.rs(B, B), pA.gw(A_O + 4)), lin_dw(dhA_oz, eAoz.y.rs(B, B), pA.gw(A_OZ + 4)),
          lin_dx(dhA_o, pA.w(A_O + 4), dy_o, 0, nullptr), lin_dx(dhA_oz, pA.w(A_OZ + 4), dy_aoz, 0, nullptr)});
  b.ln_bwd({embed_ln_bwd(eAo, pA.sub(A_O), dy_o, B, true), embed_ln_bwd(eAoz, pA.sub(A_OZ), dy_aoz, B, true)});
  b.colsum({mk_colsum(dy_o, pA.gv(A_O + 1)), mk_colsum(dy_aoz, pA.gv(A_OZ + 1))});
  b.gemm({lin_dw(dy_o, eAo.x.rs(B, B), pA.gw(A_O + 0)), lin_dw(dy_aoz, eAoz.x.rs(B, B), pA.gw(A_OZ + 0))});

  if (h->nccl_comm && n != B) {   // multi-GPU exchange 3: the flat actor gradient
    fb_handle* hh = h;
    float* gbuf = bf.d_grad_actor; const size_t count = h->seg_actor.size;
    b.push([hh, gbuf, count](cudaStream_t s) {
      return g_nccl.AllReduce(gbuf, gbuf, count, 7, 0, hh->nccl_comm, s) == 0 ? cudaSuccess : cudaErrorUnknown;
    }, FB_OPK_COLLECTIVE, 0.0, 8.0 * (double)count);
  } else if (p2p_step) {
    const P2pPeers pp = h->p2p;
    b.push([pp](cudaStream_t s) { fb_launch_pdl(k_p2p_barrier, dim3(1), dim3(32), 0, s, pp, (int)P2P_BAR_GRAD_ACTOR, 1, 1); return cudaGetLastError(); },
           FB_OPK_COLLECTIVE);
  }

  // =========================== FB_PHASE_ACTOR_ADAM ==============================================
  b.set_phase(FB_PHASE_ACTOR_ADAM);
  {
    const float b1 = c.beta1, b2 = c.beta2, eps = c.adam_eps;
    float4 *p = (float4*)bf.d_param_actor, *g = (float4*)bf.d_grad_actor, *m = (float4*)bf.d_m_actor, *v = (float4*)bf.d_v_actor;
    const size_t n4 = h->seg_actor.size / 4;
    if (p2p_step) {
      const P2pPeers pp = h->p2p;
      P2pAdamParams ap; memset(&ap, 0, sizeof(ap));
      ap.off_grad = h->p2p_off_grad_actor; ap.off_param = h->p2p_off_param_actor; ap.m = m; ap.v = v; ap.n4 = n4; ap.split4 = n4;
      ap.slice4 = (n4 + pp.world - 1) / pp.world; ap.sc = sc; ap.which = 1; ap.bar_param = P2P_BAR_PARAM_ACTOR; ap.beta1 = b1; ap.beta2 = b2; ap.eps = eps;
      b.push([pp, ap](cudaStream_t s) {
        fb_launch_pdl(k_p2p_adam, dim3(p2p_adam_ctas()), dim3(256), 0, s, pp, ap);
        return cudaGetLastError();
      }, FB_OPK_ADAM, 0.0, 16.0 * (double)ap.slice4 * (pp.world + 5.0 + pp.world));
      b.push([pp](cudaStream_t s) { fb_launch_pdl(k_p2p_barrier, dim3(1), dim3(32), 0, s, pp, (int)P2P_BAR_PARAM_ACTOR, 0, 1); return cudaGetLastError(); },
             FB_OPK_COLLECTIVE);
      b.push([=](cudaStream_t s) {
        fb_launch_pdl(k_p2p_finish, dim3(FB_SM_COUNT * 8), dim3(256), 0, s, (const float4*)p, g, (float4*)nullptr, n4, (const DevScalars*)sc);
        return cudaGetLastError();
      }, FB_OPK_ADAM, 0.0, 16.0 * (double)n4);
    } else {
    b.push([=](cudaStream_t s) {
      fb_launch_pdl(k_adam, dim3(FB_SM_COUNT * 8), dim3(256), 0, s, p, g, m, v, nullptr, n4, n4, sc, 1, b1, b2, eps);
      return cudaGetLastError();
    }, FB_OPK_ADAM, 0.0, 4.0 * 4.0 * (double)n4 * 8.0);
    b.push([=](cudaStream_t s) { fb_launch_pdl(k_tick, dim3(1), dim3(32), 0, s, sc, 1, b1, b2); return cudaGetLastError(); },
           FB_OPK_ELEMENTWISE, 0.0, 0.0, 1);
    }
  }

  // =========================== FB_PHASE_METRICS =================================================
  b.set_phase(FB_PHASE_METRICS);
  b.memset0(acc + ACC_F1, 6 * sizeof(double));
  b.memset0(h->d_linf, (size_t)Z * Z * sizeof(float));
  b.push([F1, Bm, z, B, Z, acc](cudaStream_t s) {
    fb_launch_pdl(k_metric_rows, dim3(fb_ceil_div(B, 8)), dim3(256), 0, s, F1.p, F1.ld, Bm.p, Bm.ld, z.p, z.ld, B, Z, acc);
    return cudaGetLastError();
  });
  {
    float* cov = reinterpret_cast<float*>(h->d_linf);
    const int chunk = std::max(1, std::min(FB_COV_ROWS, (int)(48 * 1024 / ((Z + 1) * sizeof(float)))));
    b.push([Bg, n, Z, cov, chunk](cudaStream_t s) {
      fb_launch_pdl(k_metric_cov, dim3(fb_ceil_div(n, chunk)), dim3(256), (size_t)chunk * (Z + 1) * sizeof(float), s, Bg.p, Bg.ld, n, Z, cov, chunk);
      return cudaGetLastError();
    });
    MetricFinalParams mp; memset(&mp, 0, sizeof(mp));
    mp.acc = acc; mp.cov = cov; mp.out = h->d_metrics; mp.n_local = B; mp.n_global = n; mp.Z = Z; mp.ortho_coef = c.ortho_coef;
    mp.q_loss_coef = c.q_loss ? c.q_loss_coef : 0.f;
    mp.temp = bz ? c.temp : 0.f;
    b.push([mp](cudaStream_t s) { fb_launch_pdl(k_metric_final, dim3(1), dim3(256), 0, s, mp); return cudaGetLastError(); });
  }

  // =========================== inference plans ==================================================
  // act / get_goal_meta / compute_z_correl / infer_meta_from_obs_and_rewards (fb_ddpg.py:177-222,258-289): forward passes of
  // the online actor / backward_net on caller-filled blocks, no gradients; a few rows -> the fp32 SIMT kernel
  b.force_simt = true;
  {
    const int R = FB_INFER_ROWS;
    Mat io = ws_mat(h, R, O, "infer_obs"), iz = ws_mat(h, R, Z, "infer_z"), ioz = ws_mat(h, R, O + Z, "infer_oz");
    const int RI = bz ? 0 : R;   // the default actor's inference activations (cfg.boltzmann: none; infer_mu then holds [mu | std])
    Mat ihA = ws_mat(h, RI, Hc, "infer_hA"), ih1 = ws_mat(h, RI, H, "infer_h1"), ipre = ws_mat(h, RI, A, "infer_pre");
    Mat imu = ws_mat(h, R, bz ? 2 * A : A, "infer_mu");
    Mat io_embed = io;
    if (deep || bz) io_embed.rows = 0;   // no obs-only embed
    Mat ioz_embed = ioz; ioz_embed.rows = RI;
    EmbedAct eo = embed_alloc(h, io_embed, ihA.cs(0, Fe), H, "infer.obs_net"), eoz = embed_alloc(h, ioz_embed, ihA.cs(oz0, Fe), H, "infer.obs_z_net");
    Mat ipol_x = ioz; ipol_x.rows = bz ? R : 0;
    BAct ipol = b_alloc(h, ipol_x, Mat(), "infer.policy", H, 2 * A);
    b.set_phase(FB_PHASE_INFER_ACTOR);
    b.push([=](cudaStream_t s) {
      fb_launch_pdl(k_infer_concat, dim3(R), dim3(128), 0, s, io.p, io.ld, iz.p, iz.ld, ioz.p, ioz.ld, R, O, Z);
      return cudaGetLastError();
    });
    b.gemm({lin_fwd(eo.x, pA.w(A_O + 0), pA.v(A_O + 1), eo.pre, 0), lin_fwd(eoz.x, pA.w(A_OZ + 0), pA.v(A_OZ + 1), eoz.pre, 0),
            lin_fwd(ipol.x, pA.w(0), pA.v(1), ipol.pre, 0)});
    b.ln_fwd({embed_ln(eo, pA.sub(A_O)), embed_ln(eoz, pA.sub(A_OZ)), b_ln(ipol, pA)});
    b.gemm({lin_fwd(eo.y, pA.w(A_O + 4), pA.v(A_O + 5), eo.out, GF_RELU), lin_fwd(eoz.y, pA.w(A_OZ + 4), pA.v(A_OZ + 5), eoz.out, GF_RELU),
            lin_fwd(ipol.y, pA.w(4), pA.v(5), ipol.h2, GF_RELU)});
    if (trunk) {
      Mat ihT = ws_mat(h, R, H, "infer_trunk");
      b.gemm({lin_fwd(ihA, pA.w(A_TR + 0), pA.v(A_TR + 1), ihT, GF_RELU)});
      b.gemm({lin_fwd(ihT, pA.w(A_POL + 0), pA.v(A_POL + 1), ih1, GF_RELU)});
    } else {
      b.gemm({lin_fwd(ihA, pA.w(A_POL + 0), pA.v(A_POL + 1), ih1, GF_RELU)});
    }
    b.gemm({lin_fwd(ih1, pA.w(A_POL + 2), pA.v(A_POL + 3), ipre, 0), lin_fwd(ipol.h2, pA.w(6), pA.v(7), ipol.raw, 0)});
    if (bz) {
      const float lo = c.log_std_min, hi = c.log_std_max;
      b.push([=](cudaStream_t s) {
        fb_launch_pdl(k_infer_gauss, dim3(fb_ceil_div(R * A, 128)), dim3(128), 0, s, ipol.raw.p, imu.p, ipol.raw.ld, R, A, lo, hi);
        return cudaGetLastError();
      });
    } else
    b.push([=](cudaStream_t s) {
      fb_launch_pdl(k_infer_tanh, dim3(fb_ceil_div(R * A, 128)), dim3(128), 0, s, ipre.p, imu.p, ipre.ld, R, A);
      return cudaGetLastError();
    });

    Mat ig = ws_mat(h, R, G, "infer_goal"), ib = ws_mat(h, R, Z, "infer_b");
    BAct b1 = b_alloc(h, ig, ib, "infer.B", 0, 0, dbg);
    b.set_phase(FB_PHASE_INFER_B);
    b.gemm({lin_fwd(b1.x, pB.w(0), pB.v(1), b1.pre, 0)});
    b.ln_fwd({b_ln(b1, pB)});
    b.gemm({lin_fwd(b1.y, pB.w(4), pB.v(5), b1.h2, GF_RELU)});
    b.gemm({lin_fwd(b1.h2, pB.w(6), pB.v(7), b1.raw, 0)});
    b.l2_fwd({b_l2(b1, Z, nzB)});

    Mat igN = ws_mat(h, B, G, "infer_goal_batch"), irN = ws_mat(h, B, 1, "infer_reward"), ibN = ws_mat(h, B, Z, "infer_b_batch");
    Mat izs = ws_mat(h, 1, Z, "infer_zsum");
    BAct bN = b_alloc(h, igN, ibN, "infer.BN", 0, 0, dbg);
    b.set_phase(FB_PHASE_INFER_BN);
    b.gemm({lin_fwd(bN.x, pB.w(0), pB.v(1), bN.pre, 0)});
    b.ln_fwd({b_ln(bN, pB)});
    b.gemm({lin_fwd(bN.y, pB.w(4), pB.v(5), bN.h2, GF_RELU)});
    b.gemm({lin_fwd(bN.h2, pB.w(6), pB.v(7), bN.raw, 0)});
    b.l2_fwd({b_l2(bN, Z, nzB)});
    b.push([=](cudaStream_t s) {
      fb_launch_pdl(k_infer_weighted_colsum, dim3(fb_ceil_div(Z, 32)), dim3(256), 0, s, ibN.p, ibN.ld, irN.p, irN.ld, B, Z, izs.p);
      return cudaGetLastError();
    });
  }
  b.force_simt = false;

  // staged operands: one grouped launch per (consumer phase, availability) on the staging lane
  for (int ph = 0; ph < FB_NUM_PHASES; ++ph) {
    std::vector<TransposeDesc>& tv = h->early_stage[ph];
    std::vector<int> avails = h->early_avail[ph];
    std::sort(avails.begin(), avails.end());
    avails.erase(std::unique(avails.begin(), avails.end()), avails.end());
    for (int av : avails) {
      std::vector<TransposeDesc> sub;
      for (size_t i = 0; i < tv.size(); ++i)
        if (h->early_avail[ph][i] == av) sub.push_back(tv[i]);
      int ctas = 0;
      double tbytes = 0.0;
      for (auto& t : sub) {
        t.cta_begin = ctas; t.ctas_x = fb_ceil_div(t.cols, 32); ctas += t.ctas_x * fb_ceil_div(t.rows, 32);
        tbytes += (t.transpose == 2 ? 4.0 : (t.out ? 12.0 : 8.0)) * t.rows * (double)t.cols;
      }
      fb_handle::StageBatch sb;
      sb.avail = av; sb.d_descs = arena_put(h, sub, d_arena); sb.n = (int)sub.size(); sb.ctas = ctas; sb.bytes = tbytes;
      sb.dev.set(FS_TRANSPOSE, ctas, FsPtrArgs{sb.d_descs, sb.n, 0});
      h->stage_batches[ph].push_back(sb);
    }
  }
  if (b.rc != FB_OK) return b.rc;
  if (h->arena.size() > FB_DESC_ARENA_BYTES) return FB_E_STATE;
  h->ws_off = (h->ws_off + 255) / 256 * 256;
  if (getenv("FB_DEBUG_PLAN")) fprintf(stderr, "[fb plan %s] workspace %zu bytes, arena %zu bytes\n", h->ws_base ? "real" : "dry", h->ws_off, h->arena.size());
  if (h->ws_base && h->ws_off > h->ws_bytes) return FB_E_STATE;  // the sizing pass and the binding pass must agree
  return FB_OK;
}

// gather slots: which 16-byte chunk of which storage row feeds each 16-byte chunk of the output row
static int build_gather_params(const fb_replay_view& v, GatherParams& gp, const BatchLayout& L, int out_ld) {
  memset(&gp, 0, sizeof(gp));
  if (v.row_stride % 4 || out_ld % 4 || v.off_obs % 4 || v.off_action % 4 || v.off_reward % 4 || v.off_discount != v.off_reward + 1)
    return FB_E_ARG;
  if (L.G > 0 && (v.off_goal < 0 || v.off_goal % 4)) return FB_E_ARG;
  if (L.X > 0 && (v.off_extra < 0 || v.off_extra % 4)) return FB_E_ARG;
  gp.rows = v.d_rows; gp.ep_len = v.d_episode_len; gp.rows_per_episode = v.rows_per_episode; gp.row_stride = v.row_stride;
  gp.out_ld = out_ld;
  int n = 0;
  auto field = [&](int dst_off, int src_row, int src_off, int dim, int special) {
    for (int c = 0; c < fb_round_up(dim, 4); c += 4) {
      if (n >= FB_MAX_GATHER_SLOTS) return;
      GatherSlot& s = gp.slots[n];
      s.dst_f4 = (dst_off + c) / 4; s.src_row = (short)src_row; s.special = (short)special; s.src_f4 = (src_off + c) / 4;
      ++n;
    }
  };
  field(L.off_obs, 0, v.off_obs, L.O, 0);
  field(L.off_action, 1, v.off_action, L.A, 0);
  field(L.off_rd, 1, v.off_reward, 4, 1);
  field(L.off_next_obs, 1, v.off_obs, L.O, 0);
  if (L.G > 0) { field(L.off_goal, 0, v.off_goal, L.G, 0); field(L.off_next_goal, 1, v.off_goal, L.G, 0); }
  if (L.X > 0) field(L.off_extra, 0, v.off_extra, L.X, 0);
  if (L.with_future) {
    field(L.off_future_obs, 2, v.off_obs, L.O, 0);
    if (L.G > 0) field(L.off_future_goal, 2, v.off_goal, L.G, 0);
  }
  if (n >= FB_MAX_GATHER_SLOTS) return FB_E_UNSUPPORTED;
  gp.n_slots = n;
  return FB_OK;
}

static void make_batch_layout(BatchLayout& L, int O, int A, int G, int X, int with_future) {
  memset(&L, 0, sizeof(L));
  L.O = O; L.A = A; L.G = G; L.X = X; L.with_future = with_future;
  int off = 0;
  L.off_obs = off; off += fb_round_up(O, 4);
  L.off_action = off; off += fb_round_up(A, 4);
  L.off_rd = off; off += 4;
  L.off_next_obs = off; off += fb_round_up(O, 4);
  L.off_goal = off; off += fb_round_up(G, 4);
  L.off_next_goal = off; off += fb_round_up(G, 4);
  L.off_extra = off; off += fb_round_up(X, 4);
  L.off_future_obs = off; if (with_future) off += fb_round_up(O, 4);
  L.off_future_goal = off; if (with_future) off += fb_round_up(G, 4);
  L.pitch = off;
}

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int fb_abi_version(void) { return FB_ABI_VERSION; }

const char* fb_error_string(int code) {
  switch (code) {
    case FB_OK: return "ok";
    case FB_E_ARG: return "bad argument";
    case FB_E_STATE: return "bad state (fb_bind / fb_bind_replay not called, or descriptor arena overflow)";
    case FB_E_UNSUPPORTED: return "unsupported configuration";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

int fb_create(const fb_config* cfg, fb_handle** out) {
  if (!cfg || !out) return FB_E_ARG;
  if (cfg->abi_version != FB_ABI_VERSION) return FB_E_ARG;
  if (cfg->batch < 2 || cfg->global_batch < cfg->batch || cfg->row_offset < 0 || cfg->row_offset + cfg->batch > cfg->global_batch)
    return FB_E_ARG;
  if (cfg->obs_dim < 1 || cfg->action_dim < 1 || cfg->z_dim < 1 || cfg->goal_dim < 1 || cfg->hidden_dim < 1 || cfg->feature_dim < 1 ||
      cfg->backward_hidden_dim < 1)
    return FB_E_ARG;
  if (cfg->hidden_dim > FB_MAX_LN_DIM || cfg->backward_hidden_dim > FB_MAX_LN_DIM) return FB_E_UNSUPPORTED;
  if (!cfg->use_goal && cfg->goal_dim != cfg->obs_dim) return FB_E_ARG;
  if (!(cfg->future_ratio >= 0.f && cfg->future_ratio <= 1.f) || !(cfg->mix_ratio >= 0.f && cfg->mix_ratio <= 1.f)) return FB_E_ARG;
  if (cfg->q_loss && cfg->z_dim > FB_QLOSS_MAX_Z) return FB_E_UNSUPPORTED;
  if (cfg->boltzmann && !(cfg->log_std_max > cfg->log_std_min)) return FB_E_ARG;
  if (cfg->debug_identity_b && cfg->z_dim != cfg->goal_dim) return FB_E_ARG;   // an identity backward map: z lives in goal space
  if (cfg->rand_weight && cfg->z_dim > FB_MIXW_MAX_Z) return FB_E_UNSUPPORTED;
  fb_handle* h = new fb_handle();
  h->cfg = *cfg;
  memset(&h->bufs, 0, sizeof(h->bufs));
  memset(&h->replay, 0, sizeof(h->replay));
  build_layout(h);
  // sizing pass: build the plan against fake, never dereferenced, mutually distinct base addresses (the plan de-duplicates
  // staged operands by address, so the caller's segments must not alias each other here any more than they will later)
  h->ws_base = nullptr;
  {
    float** seg[] = {&h->bufs.d_param_fb, &h->bufs.d_grad_fb, &h->bufs.d_m_fb, &h->bufs.d_v_fb, &h->bufs.d_target_fb,
                     &h->bufs.d_param_actor, &h->bufs.d_grad_actor, &h->bufs.d_m_actor, &h->bufs.d_v_actor};
    for (int i = 0; i < 9; ++i) *seg[i] = reinterpret_cast<float*>((uintptr_t)(i + 1) << 40);
  }
  int rc = build_plan(h);
  memset(&h->bufs, 0, sizeof(h->bufs));
  if (rc != FB_OK) { delete h; return rc; }
  h->ws_bytes = h->ws_off;
  for (auto& v : h->ops) v.clear();
  *out = h;
  return FB_OK;
}

void fb_destroy(fb_handle* h) {
  if (!h) return;
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->stage_stream) cudaStreamDestroy(h->stage_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_stage_fork) cudaEventDestroy(h->ev_stage_fork);
  for (auto& e : h->ev_stage) if (e) cudaEventDestroy(e);
  if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl_comm);
  for (int q = 0; q < P2P_MAX_WORLD; ++q)
    if (h->p2p_ipc[q] && h->p2p.base[q]) cudaIpcCloseMemHandle(h->p2p.base[q]);
  if (h->p2p_arena) cudaFree(h->p2p_arena);
  delete h;
}

size_t fb_flat_size(const fb_handle* h, int actor) { return actor ? h->seg_actor.size : h->seg_fb.size; }

static const SegmentLayout* net_tensors(const fb_handle* h, int net, int* first, int* count) {
  switch (net) {
    case FB_NET_FORWARD: *first = h->fwd_first; *count = h->bwd_first - h->fwd_first; return &h->seg_fb;
    case FB_NET_BACKWARD: *first = h->bwd_first; *count = (int)h->seg_fb.t.size() - h->bwd_first; return &h->seg_fb;
    case FB_NET_ACTOR: *first = 0; *count = (int)h->seg_actor.t.size(); return &h->seg_actor;
    default: return nullptr;
  }
}

int fb_num_tensors(const fb_handle* h, int net) {
  int first, count;
  return net_tensors(h, net, &first, &count) ? count : FB_E_ARG;
}

int fb_tensor_info(const fb_handle* h, int net, int index, size_t* offset, int* rows, int* cols, char* name, size_t name_cap) {
  int first, count;
  const SegmentLayout* L = net_tensors(h, net, &first, &count);
  if (!L || index < 0 || index >= count) return FB_E_ARG;
  const TensorInfo& t = L->t[first + index];
  if (offset) *offset = t.off;
  if (rows) *rows = t.rows;
  if (cols) *cols = t.cols;
  if (name && name_cap) { strncpy(name, t.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  return FB_OK;
}

size_t fb_workspace_bytes(const fb_handle* h) { return h->ws_bytes; }

__global__ void k_iota(int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

int fb_bind(fb_handle* h, const fb_buffers* bufs, void* stream) {
  if (!h || !bufs) return FB_E_ARG;
  if (!bufs->d_param_fb || !bufs->d_grad_fb || !bufs->d_m_fb || !bufs->d_v_fb || !bufs->d_target_fb || !bufs->d_param_actor ||
      !bufs->d_grad_actor || !bufs->d_m_actor || !bufs->d_v_actor || !bufs->d_workspace)
    return FB_E_ARG;
  if (bufs->workspace_bytes < h->ws_bytes) return FB_E_ARG;
  if (h->p2p_on) {
    if (!h->p2p_attached) return FB_E_STATE;
    if ((char*)bufs->d_grad_fb != h->p2p_arena + h->p2p_off_grad_fb || (char*)bufs->d_grad_actor != h->p2p_arena + h->p2p_off_grad_actor ||
        (char*)bufs->d_param_fb != h->p2p_arena + h->p2p_off_param_fb || (char*)bufs->d_param_actor != h->p2p_arena + h->p2p_off_param_actor)
      return FB_E_ARG;   // gradients and parameters must be the arena's segments (fb_p2p_create filled them in)
  }
  if (((uintptr_t)bufs->d_workspace & 255u) || ((uintptr_t)bufs->d_param_fb & 15u) || ((uintptr_t)bufs->d_param_actor & 15u) ||
      ((uintptr_t)bufs->d_grad_fb & 15u) || ((uintptr_t)bufs->d_grad_actor & 15u) || ((uintptr_t)bufs->d_target_fb & 15u))
    return FB_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  h->graphs.clear();
  h->bufs = *bufs;
  h->ws_base = (char*)bufs->d_workspace;
  int rc = build_plan(h);
  if (rc != FB_OK) return rc;
  CK(cudaFuncSetAttribute(k_gemm_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  if (h->uses_gemm_tc) CK(tc_set_smem_attr());
  if (h->uses_gemm_tc) {
    CK(cudaFuncSetAttribute(k_fused_stack, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    int dev = 0, sms = 0, per_sm = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_stack, FS_THREADS, TC_SMEM_BYTES));
    if (per_sm < 1) return FB_E_UNSUPPORTED;
    h->sm_count = sms;   // one resident CTA per SM: the grid barrier of k_fused_stack needs every CTA of the launch co-resident
  }
  if (h->contract_smem) CK(cudaFuncSetAttribute(k_contract_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->contract_smem));
  if (h->qloss_smem > 48u * 1024u) CK(cudaFuncSetAttribute(k_qloss_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->qloss_smem));
  CK(cudaMemsetAsync(h->ws_base, 0, h->ws_bytes, s));
  CK(cudaMemcpyAsync(h->ws_base, h->arena.data(), h->arena.size(), cudaMemcpyHostToDevice, s));
  k_iota<<<fb_ceil_div(h->cfg.batch, 256), 256, 0, s>>>(h->d_perm, h->cfg.batch);
  k_set_adam_steps<<<1, 32, 0, s>>>(h->d_sc, 0ll, 0ll, h->cfg.beta1, h->cfg.beta2);   // step counts 0, bias corrections of step 1
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  h->bound = true;
  return FB_OK;
}

int fb_bind_replay(fb_handle* h, const fb_replay_view* view, void* stream) {
  if (!h || !view || !h->bound) return h && view ? FB_E_STATE : FB_E_ARG;
  GatherParams gp;
  int rc = build_gather_params(*view, gp, h->bl, h->packed.ld);
  if (rc != FB_OK) return rc;
  const bool same = h->replay_bound && h->replay.d_rows == view->d_rows && h->replay.d_episode_len == view->d_episode_len &&
                    h->replay.rows_per_episode == view->rows_per_episode && h->replay.row_stride == view->row_stride;
  if (!same) {  // graphs bake the storage pointers
    for (auto it = h->graphs.begin(); it != h->graphs.end();) {
      if (it->first & FB_PHASE_SAMPLE) { cudaGraphExecDestroy(it->second); it = h->graphs.erase(it); } else ++it;
    }
  }
  h->replay = *view;
  h->replay_bound = true;
  CK(cudaMemcpyAsync(h->d_n_episodes, &h->replay.n_episodes, sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return FB_OK;
}

int fb_set_step_scalars(fb_handle* h, const fb_step_scalars* sv, void* stream) {
  if (!h || !sv || !h->bound) return FB_E_STATE;
  HostScalars hs{sv->stddev, sv->stddev_clip, sv->lr_forward, sv->lr_backward, sv->lr_actor, sv->tau, sv->replay_discount,
                 sv->replay_future, sv->grad_scale};
  k_set_scalars<<<1, 32, 0, (cudaStream_t)stream>>>(h->d_sc, hs);
  CK(cudaGetLastError());
  return FB_OK;
}

int fb_set_indices(fb_handle* h, const int32_t* d_ep_idx, const int32_t* d_step_idx, const int32_t* d_future_idx, const int32_t* d_perm,
                   const int32_t* d_mix_mask, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nb = (size_t)h->cfg.batch * sizeof(int32_t);
  if (d_ep_idx) CK(cudaMemcpyAsync(h->d_ep_idx, d_ep_idx, nb, cudaMemcpyDeviceToDevice, s));
  if (d_step_idx) CK(cudaMemcpyAsync(h->d_step_idx, d_step_idx, nb, cudaMemcpyDeviceToDevice, s));
  if (d_future_idx) CK(cudaMemcpyAsync(h->d_future_idx, d_future_idx, nb, cudaMemcpyDeviceToDevice, s));
  if (d_perm) CK(cudaMemcpyAsync(h->d_perm, d_perm, nb, cudaMemcpyDeviceToDevice, s));
  if (d_mix_mask) CK(cudaMemcpyAsync(h->d_mix_mask, d_mix_mask, nb, cudaMemcpyDeviceToDevice, s));
  return FB_OK;
}

int fb_set_batch(fb_handle* h, const float* d_obs, const float* d_action, const float* d_discount, const float* d_next_obs,
                 const float* d_goal, const float* d_next_goal, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  if (!d_obs || !d_action || !d_discount || !d_next_obs) return FB_E_ARG;
  if (h->cfg.use_goal && (!d_goal || !d_next_goal)) return FB_E_ARG;
  k_pack_batch<<<h->cfg.batch, 64, 0, (cudaStream_t)stream>>>(h->bl, h->cfg.batch, d_obs, d_action, d_discount, d_next_obs, d_goal,
                                                             d_next_goal, h->packed.p);
  CK(cudaGetLastError());
  return FB_OK;
}

int fb_nccl_unique_id(const char* libnccl_path, void* id128) {
  if (!id128) return FB_E_ARG;
  int rc = nccl_load(libnccl_path);
  if (rc != FB_OK) return rc;
  return g_nccl.GetUniqueId(reinterpret_cast<FbNcclId*>(id128)) == 0 ? FB_OK : FB_E_STATE;
}

int fb_nccl_init(fb_handle* h, const char* libnccl_path, const void* id128, int world, int rank) {
  if (!h || !id128 || world < 2 || rank < 0 || rank >= world) return FB_E_ARG;
  if (h->bound || h->nccl_comm) return FB_E_STATE;   // before fb_bind: the plan places the collectives
  if (h->cfg.global_batch != h->cfg.batch * world || h->cfg.row_offset != rank * h->cfg.batch) return FB_E_ARG;
  int rc = nccl_load(libnccl_path);
  if (rc != FB_OK) return rc;
  FbNcclId id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  if (g_nccl.CommInitRank(&comm, world, id, rank) != 0 || !comm) return FB_E_STATE;
  h->nccl_comm = comm; h->nccl_world = world; h->nccl_rank = rank;
  return FB_OK;
}

// ---- peer-memory exchange (p2p.cuh) ----------------------------------------------------------------------------------------
static size_t p2p_round(size_t x) { return (x + 255) / 256 * 256; }

int fb_p2p_create(fb_handle* h, int world, int rank, void* ipc_handle64, fb_buffers* bufs) {
  if (!h || !bufs || world < 2 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) return FB_E_ARG;
  if (h->bound || h->p2p_on || h->nccl_comm) return FB_E_STATE;   // before fb_bind: the plan places the exchange kernels
  if (h->cfg.global_batch != h->cfg.batch * world || h->cfg.row_offset != rank * h->cfg.batch) return FB_E_ARG;
  const int ldZ = fb_round_up(h->cfg.z_dim, 4), blk_pitch = 6 * ldZ + 4;
  size_t off = P2P_FLAG_BYTES;
  h->p2p_off_grad_fb = off; off = p2p_round(off + h->seg_fb.size * sizeof(float));
  h->p2p_off_grad_actor = off; off = p2p_round(off + h->seg_actor.size * sizeof(float));
  h->p2p_off_param_fb = off; off = p2p_round(off + h->seg_fb.size * sizeof(float));
  h->p2p_off_param_actor = off; off = p2p_round(off + h->seg_actor.size * sizeof(float));
  h->p2p_off_blk = off; off = p2p_round(off + (size_t)h->cfg.global_batch * blk_pitch * sizeof(float));
  h->p2p_bytes = off;
  {   // load the exchange kernels now: a lazy module load at first launch may synchronise the context, which deadlocks when a
      // barrier kernel of this context is already spinning on a peer that lives in the same context (several ranks on one device)
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_p2p_barrier)); CK(cudaFuncGetAttributes(&fa, k_p2p_scatter_rows));
    CK(cudaFuncGetAttributes(&fa, k_p2p_adam)); CK(cudaFuncGetAttributes(&fa, k_p2p_finish));
  }
  CK(cudaMalloc((void**)&h->p2p_arena, off));
  CK(cudaMemset(h->p2p_arena, 0, off));
  CK(cudaDeviceSynchronize());
  if (ipc_handle64) {
    cudaIpcMemHandle_t ih;
    static_assert(sizeof(ih) == 64, "CUDA IPC handles are 64 bytes");
    CK(cudaIpcGetMemHandle(&ih, h->p2p_arena));
    memcpy(ipc_handle64, &ih, sizeof(ih));
  }
  memset(&h->p2p, 0, sizeof(h->p2p));
  h->p2p.world = world; h->p2p.rank = rank; h->p2p.base[rank] = h->p2p_arena;
  h->p2p_on = true;
  bufs->d_grad_fb = reinterpret_cast<float*>(h->p2p_arena + h->p2p_off_grad_fb);
  bufs->d_grad_actor = reinterpret_cast<float*>(h->p2p_arena + h->p2p_off_grad_actor);
  bufs->d_param_fb = reinterpret_cast<float*>(h->p2p_arena + h->p2p_off_param_fb);
  bufs->d_param_actor = reinterpret_cast<float*>(h->p2p_arena + h->p2p_off_param_actor);
  return FB_OK;
}

int fb_p2p_attach(fb_handle* h, const void* ipc_handles, void* const* local_arenas) {
  if (!h || !h->p2p_on || h->bound || h->p2p_attached) return FB_E_STATE;
  if (!ipc_handles && !local_arenas) return FB_E_ARG;
  for (int q = 0; q < h->p2p.world; ++q) {
    if (q == h->p2p.rank) continue;
    if (local_arenas) {   // arenas of engines of this process (one device, or peer-enabled devices): plain pointers
      if (!local_arenas[q]) return FB_E_ARG;
      h->p2p.base[q] = (char*)local_arenas[q];
    } else {
      cudaIpcMemHandle_t ih;
      memcpy(&ih, (const char*)ipc_handles + (size_t)q * sizeof(ih), sizeof(ih));
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
      h->p2p.base[q] = (char*)p; h->p2p_ipc[q] = true;
    }
  }
  h->p2p_attached = true;
  return FB_OK;
}

void* fb_p2p_arena(fb_handle* h) { return h && h->p2p_on ? h->p2p_arena : nullptr; }

int fb_p2p_status(fb_handle* h, uint32_t* code, uint64_t* epochs8, void* stream) {
  if (!h || !h->p2p_on || !code) return FB_E_STATE;
  P2pFlags f;
  CK(cudaMemcpyAsync(&f, h->p2p_arena, sizeof(f), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  *code = f.error;
  if (epochs8) for (int i = 0; i < P2P_NUM_BARRIERS; ++i) epochs8[i] = f.epoch[i];
  return FB_OK;
}

int fb_p2p_slice(const fb_handle* h, int actor, size_t* first, size_t* count) {
  if (!h || !h->p2p_on || !first || !count) return FB_E_STATE;
  const size_t n4 = (actor ? h->seg_actor.size : h->seg_fb.size) / 4, slice4 = (n4 + h->p2p.world - 1) / h->p2p.world;
  const size_t lo = std::min(n4, (size_t)h->p2p.rank * slice4), hi = std::min(n4, lo + slice4);
  *first = 4 * lo; *count = 4 * (hi - lo);
  return FB_OK;
}

int fb_set_future_mask(fb_handle* h, const int32_t* d_future_mask, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  if (!d_future_mask || !(h->cfg.future_ratio > 0.f)) return FB_E_ARG;
  CK(cudaMemcpyAsync(h->d_future_mask, d_future_mask, (size_t)h->cfg.batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return FB_OK;
}

int fb_upload_batch(fb_handle* h, const float* h_rows, int pitch, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  if (!h_rows || pitch != h->bl.pitch) return FB_E_ARG;
  if (h->packed.ld == pitch) {   // one contiguous block: a 1-D copy (the 2-D form hands the copy engine one descriptor per 240-byte row)
    CK(cudaMemcpyAsync(h->packed.p, h_rows, (size_t)pitch * sizeof(float) * (size_t)h->cfg.batch, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return FB_OK;
  }
  CK(cudaMemcpy2DAsync(h->packed.p, (size_t)h->packed.ld * sizeof(float), h_rows, (size_t)pitch * sizeof(float), (size_t)pitch * sizeof(float),
                       (size_t)h->cfg.batch, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return FB_OK;
}

static int copy_rows(const Mat& dst, const float* src, cudaStream_t s) {
  CK(cudaMemcpy2DAsync(dst.p, dst.ld * sizeof(float), src, dst.cols * sizeof(float), dst.cols * sizeof(float), dst.rows,
                       cudaMemcpyDeviceToDevice, s));
  return FB_OK;
}

int fb_set_mix_weights(fb_handle* h, const float* d_weight, const float* d_row_scale, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  if (!d_weight || !d_row_scale || !h->mix_u) return FB_E_ARG;   // mix_u: the plan was built with rand_weight (and mix_ratio > 0)
  CK(cudaMemcpyAsync(h->mix_u, d_row_scale, (size_t)h->cfg.batch * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return copy_rows(h->mix_w, d_weight, (cudaStream_t)stream);
}

int fb_set_z(fb_handle* h, const float* d_z, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  if (!d_z) return FB_E_ARG;
  return copy_rows(h->z_rand, d_z, (cudaStream_t)stream);
}

int fb_set_noise(fb_handle* h, const float* d_noise_fb, const float* d_noise_actor, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  int rc = FB_OK;
  if (d_noise_fb) rc = copy_rows(h->noise_fb, d_noise_fb, (cudaStream_t)stream);
  if (rc == FB_OK && d_noise_actor) rc = copy_rows(h->noise_actor, d_noise_actor, (cudaStream_t)stream);
  return rc;
}

static cudaError_t ensure_streams(fb_handle* h) {
  if (h->side_stream) return cudaSuccess;
  CKE(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
  CKE(cudaStreamCreateWithFlags(&h->stage_stream, cudaStreamNonBlocking));
  CKE(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CKE(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CKE(cudaEventCreateWithFlags(&h->ev_stage_fork, cudaEventDisableTiming));
  for (auto& e : h->ev_stage) CKE(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return cudaSuccess;
}

static cudaError_t launch_stage_batch(const fb_handle::StageBatch& sb, cudaStream_t s) {
  fb_launch_pdl(k_transpose_grouped, dim3(sb.ctas), dim3(256), 0, s, sb.d_descs, sb.n);
  return cudaGetLastError();
}

// Three lanes: the caller's stream carries the dependency chain; the side stream carries chains whose inputs are final early
// (z mixing, bias column sums); the staging stream carries operand staging, launched as soon as its sources are final.
static cudaError_t run_eager(fb_handle* h, uint32_t mask, cudaStream_t s) {
  bool side_pending = false;
  CKE(ensure_streams(h));
  size_t issued[FB_NUM_PHASES] = {};   // staging batches of each consumer phase launched so far (batches are sorted by availability)
  for (int ph = 0; ph < FB_NUM_PHASES; ++ph) {
    if (!(mask & (1u << ph))) continue;
    // staging whose sources are final by now, for this and every later phase of the mask
    bool forked = false;
    for (int P = ph; P < FB_NUM_PHASES; ++P) {
      if (!(mask & (1u << P))) continue;
      auto& sbs = h->stage_batches[P];
      const bool no_hoist = getenv("FB_NO_HOIST") != nullptr;
      while (issued[P] < sbs.size() && (sbs[issued[P]].avail <= ph && (!no_hoist || P == ph))) {
        if (!forked) {   // the staging lane sees everything issued so far on the main lane
          CKE(cudaEventRecord(h->ev_stage_fork, s));
          CKE(cudaStreamWaitEvent(h->stage_stream, h->ev_stage_fork, 0));
          forked = true;
        }
        CKE(launch_stage_batch(sbs[issued[P]], h->stage_stream));
        if (++issued[P] == sbs.size()) CKE(cudaEventRecord(h->ev_stage[P], h->stage_stream));
      }
    }
    for (auto& op : h->ops[ph]) {
      if (op.replay_only && (mask & FB_RUN_HOST_BATCH)) continue;
      const bool staged = op.wait_stage && !h->stage_batches[ph].empty();
      if (op.lane == 1) {
        if (op.fork || !side_pending) {                    // fork: the side lane sees everything issued so far on the main lane
          CKE(cudaEventRecord(h->ev_fork, s));
          CKE(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        }
        if (staged) CKE(cudaStreamWaitEvent(h->side_stream, h->ev_stage[ph], 0));
        CKE(op(h->side_stream));
        side_pending = true;
      } else {
        if (op.join && side_pending) {
          CKE(cudaEventRecord(h->ev_join, h->side_stream));
          CKE(cudaStreamWaitEvent(s, h->ev_join, 0));
          side_pending = false;
        }
        if (staged) CKE(cudaStreamWaitEvent(s, h->ev_stage[ph], 0));
        CKE(op(s));
      }
    }
    if (side_pending) {  // a phase never leaves side work dangling (graph capture needs every fork joined)
      CKE(cudaEventRecord(h->ev_join, h->side_stream));
      CKE(cudaStreamWaitEvent(s, h->ev_join, 0));
      side_pending = false;
    }
    if (!h->stage_batches[ph].empty()) CKE(cudaStreamWaitEvent(s, h->ev_stage[ph], 0));   // the staging lane is joined at the latest here
  }
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// fused execution (fused.cuh): the plan of a phase mask cut into units — fused segments and stand-alone launches
// ------------------------------------------------------------------------------------------------
// The walk below issues the launches in exactly the order run_eager does and assigns every fusable launch a STAGE of the current
// segment: a main-lane launch opens the next stage; a side-lane launch (z-mixing chain, bias column sums) goes to the stage that
// runs next to the following main-lane launch (fork), or behind the previous launch of its chain; a staging batch runs next to the
// first launch of the phase that makes its sources final, and its consumers are held behind it.  A launch without a fused form
// (RNG, replay gather, the contraction, q_loss, SIMT GEMMs, Adam, collectives, metrics) closes the segment and runs as a kernel of
// its own in stream order — list order is a valid serialisation of the lanes, so any cut is correct.
struct FusedSegment { std::vector<std::vector<const DevItem*>> stages; };

static int fused_emit_program(fb_handle* h, const FusedSegment& seg) {
  FsHeader hd; memset(&hd, 0, sizeof(hd));
  std::vector<FsItem> items;
  std::vector<char> blob;
  if (seg.stages.size() > FS_MAX_STAGES) return -1;
  hd.n_stages = (int)seg.stages.size();
  for (size_t st = 0; st < seg.stages.size(); ++st) {
    hd.first_item[st] = (int)items.size();
    // GEMM items first: their tiles are the long poles of a stage, the block-style items fill the CTAs behind them
    std::vector<const DevItem*> order;
    for (auto* it : seg.stages[st]) if (it->type == FS_GEMM_TC) order.push_back(it);
    for (auto* it : seg.stages[st]) if (it->type != FS_GEMM_TC) order.push_back(it);
    for (auto* it : order) {
      FsItem fi; fi.type = it->type; fi.count = it->count; fi.arg_off = (int)blob.size(); fi.arg_bytes = (int)((it->args.size() + 3) / 4 * 4);
      if (fi.arg_bytes > FS_ARG_WORDS * 4) return -1;
      blob.insert(blob.end(), it->args.begin(), it->args.end());
      blob.resize((blob.size() + 15) / 16 * 16);
      items.push_back(fi);
    }
  }
  hd.first_item[seg.stages.size()] = (int)items.size();
  hd.n_items = (int)items.size();
  if (items.size() > FS_MAX_ITEMS) return -1;
  const size_t base = (h->prog_host.size() + 127) / 128 * 128;
  const size_t items_off = (sizeof(FsHeader) + 15) / 16 * 16;
  const size_t args_off = items_off + (items.size() * sizeof(FsItem) + 15) / 16 * 16;
  hd.items_off = (int)items_off;
  for (auto& fi : items) fi.arg_off += (int)args_off;
  if (base + args_off + blob.size() > FB_PROG_ARENA_BYTES) return -1;
  h->prog_host.resize(base + args_off + blob.size());
  memcpy(h->prog_host.data() + base, &hd, sizeof(hd));
  memcpy(h->prog_host.data() + base + items_off, items.data(), items.size() * sizeof(FsItem));
  memcpy(h->prog_host.data() + base + args_off, blob.data(), blob.size());
  return (int)base;
}

static int build_fused_plan(fb_handle* h, uint32_t mask, fb_handle::FusedPlan& plan) {
  FusedSegment seg;
  int main_next = 0, side_next = 0;
  bool side_pending = false;
  int ready[FB_NUM_PHASES] = {};
  size_t issued[FB_NUM_PHASES] = {};
  int rc = FB_OK;
  auto put = [&](int st, const DevItem* it) {
    if ((int)seg.stages.size() <= st) seg.stages.resize(st + 1);
    seg.stages[st].push_back(it);
  };
  auto close = [&]() {
    if (!seg.stages.empty()) {
      const int off = fused_emit_program(h, seg);
      if (off < 0 || h->n_fs_barriers >= FS_NUM_BARRIERS) { rc = FB_E_STATE; }
      else { fb_handle::Unit u; u.program = off; plan.units.push_back(u); plan.barrier_of.push_back(h->n_fs_barriers++); }
    }
    seg.stages.clear();
    main_next = side_next = 0; side_pending = false;
    for (auto& r : ready) r = 0;
  };
  for (int ph = 0; ph < FB_NUM_PHASES; ++ph) {
    if (!(mask & (1u << ph))) continue;
    for (int P = ph; P < FB_NUM_PHASES; ++P) {
      if (!(mask & (1u << P))) continue;
      auto& sbs = h->stage_batches[P];
      while (issued[P] < sbs.size() && sbs[issued[P]].avail <= ph) {
        if (getenv("FB_FUSE_STAGING_ALONE")) {   // experiment: staging batches as kernels of their own (full occupancy), between segments
          close();
          fb_handle::Unit u; u.sb = &sbs[issued[P]]; plan.units.push_back(u); plan.barrier_of.push_back(-1);
          ++issued[P];
          continue;
        }
        put(main_next, &sbs[issued[P]].dev);
        ready[P] = std::max(ready[P], main_next + 1);
        ++issued[P];
      }
    }
    for (auto& op : h->ops[ph]) {
      if (op.replay_only && (mask & FB_RUN_HOST_BATCH)) continue;
      const bool staged = op.wait_stage && !h->stage_batches[ph].empty();
      if (op.dev.type == FS_NONE) {
        close();
        fb_handle::Unit u; u.op = &op; plan.units.push_back(u); plan.barrier_of.push_back(-1);
        continue;
      }
      if (op.lane == 1) {
        int st = (op.fork || !side_pending) ? std::max(main_next, side_next) : side_next;
        if (staged) st = std::max(st, ready[ph]);
        put(st, &op.dev);
        side_next = st + 1; side_pending = true;
      } else {
        int st = main_next;
        if (op.join && side_pending) { st = std::max(st, side_next); side_pending = false; }
        if (staged) st = std::max(st, ready[ph]);
        put(st, &op.dev);
        main_next = st + 1;
      }
    }
    if (side_pending) { main_next = std::max(main_next, side_next); side_pending = false; }
    if (!h->stage_batches[ph].empty()) main_next = std::max(main_next, ready[ph]);
  }
  close();
  return rc;
}

static cudaError_t launch_fused(fb_handle* h, int program, int barrier, cudaStream_t s, unsigned long long* times = nullptr) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(h->sm_count); cfg.blockDim = dim3(FS_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = getenv("FB_NO_PDL") ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, k_fused_stack, (const char*)(h->d_prog + program), h->d_fs_barrier + barrier, h->d_fs_err, times);
}

// the fused plan of `mask`, built (and its programs uploaded) on first use
static int get_fused_plan(fb_handle* h, uint32_t mask, cudaStream_t s, fb_handle::FusedPlan** out) {
  auto it = h->fused_plans.find(mask);
  if (it == h->fused_plans.end()) {
    fb_handle::FusedPlan plan;
    int rc = build_fused_plan(h, mask, plan);
    if (rc != FB_OK) return rc;
    if (h->prog_host.size() > h->prog_uploaded) {
      CK(cudaMemcpyAsync(h->d_prog + h->prog_uploaded, h->prog_host.data() + h->prog_uploaded, h->prog_host.size() - h->prog_uploaded,
                         cudaMemcpyHostToDevice, s));
      CK(cudaStreamSynchronize(s));   // prog_host may reallocate when the next plan is built
      h->prog_uploaded = h->prog_host.size();
    }
    it = h->fused_plans.emplace(mask, std::move(plan)).first;
  }
  *out = &it->second;
  return FB_OK;
}

static cudaError_t run_fused(fb_handle* h, const fb_handle::FusedPlan& plan, cudaStream_t s) {
  for (size_t i = 0; i < plan.units.size(); ++i) {
    const fb_handle::Unit& u = plan.units[i];
    if (u.program >= 0) CKE(launch_fused(h, u.program, plan.barrier_of[i], s));
    else if (u.sb) CKE(launch_stage_batch(*u.sb, s));
    else CKE((*u.op)(s));
  }
  return cudaSuccess;
}

static bool fused_wanted(const fb_handle* h, uint32_t mask) {
  static const bool off = getenv("FB_NO_FUSE") != nullptr;
  if (off || (mask & FB_RUN_UNFUSED) || !h->uses_gemm_tc) return false;
  return (mask & (FB_PHASE_INFER_ACTOR | FB_PHASE_INFER_B | FB_PHASE_INFER_BN)) == 0;   // a few rows on the SIMT kernel: nothing to fuse
}

int fb_run(fb_handle* h, uint32_t phase_mask, int use_graph, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  if ((phase_mask & FB_PHASE_SAMPLE) && !(phase_mask & FB_RUN_HOST_BATCH) && !h->replay_bound) return FB_E_STATE;
  cudaStream_t s = (cudaStream_t)stream;
  const bool fused = fused_wanted(h, phase_mask);
  if (fused && h->uses_pairs) return FB_E_STATE;   // CTA-pair GEMM launches have no fused form: create the handle with cfg.fused_stacks
  fb_handle::FusedPlan* fplan = nullptr;
  if (fused) {   // before any capture: building the plan uploads its programs
    int rc = get_fused_plan(h, phase_mask, s, &fplan);
    if (rc != FB_OK) return rc;
  }
  if (!use_graph) return (int)(fused ? run_fused(h, *fplan, s) : run_eager(h, phase_mask, s));
  auto it = h->graphs.find(phase_mask);
  if (it == h->graphs.end()) {
    // capture on a private stream (the caller's may be the legacy default stream, which cannot capture);
    // nothing executes during capture, the instantiated graph is launched on the caller's stream
    cudaGraph_t graph = nullptr;
    if (!h->capture_stream) CK(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
    cudaStream_t cs = h->capture_stream;
    CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = fused ? run_fused(h, *fplan, cs) : run_eager(h, phase_mask, cs);
    cudaError_t e2 = cudaStreamEndCapture(cs, &graph);
    if (e != cudaSuccess || e2 != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      return (int)(e != cudaSuccess ? e : e2);
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return (int)e;
    it = h->graphs.emplace(phase_mask, exec).first;
  }
  if (use_graph == 2) return FB_OK;   // capture + instantiate only (ranks sharing a context instantiate before any of them spins)
  CK(cudaGraphLaunch(it->second, s));
  return FB_OK;
}

int fb_launch_count(fb_handle* h, uint32_t phase_mask) {
  if (!h || !h->bound) return FB_E_STATE;
  if (fused_wanted(h, phase_mask) && h->uses_pairs) return FB_E_STATE;
  if (fused_wanted(h, phase_mask)) {   // fused execution: one launch per unit (fused segment or stand-alone kernel)
    auto it = h->fused_plans.find(phase_mask);
    if (it != h->fused_plans.end()) return (int)it->second.units.size();
    fb_handle::FusedPlan plan;
    const size_t keep_prog = h->prog_host.size(); const int keep_bar = h->n_fs_barriers;
    int rc = build_fused_plan(h, phase_mask, plan);
    h->prog_host.resize(keep_prog); h->n_fs_barriers = keep_bar;   // a dry run: the real plan is built by fb_run
    return rc == FB_OK ? (int)plan.units.size() : rc;
  }
  phase_mask &= ~(uint32_t)FB_RUN_UNFUSED;
  int n = 0;
  for (int ph = 0; ph < FB_NUM_PHASES; ++ph)
    if (phase_mask & (1u << ph)) {
      n += (int)h->stage_batches[ph].size();
      for (auto& op : h->ops[ph]) n += (op.replay_only && (phase_mask & FB_RUN_HOST_BATCH)) ? 0 : 1;
    }
  return n;
}

int fb_profile_ops(fb_handle* h, uint32_t phase_mask, int reps, void* stream, float* ms_out, int32_t* kind_out, double* flops_out,
                   double* bytes_out, int cap) {
  if (!h || !h->bound) return FB_E_STATE;
  if (reps < 1 || cap < 1 || !ms_out) return FB_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  // every launch of the phases, the staging batches of a phase listed before its ops (all serialised on `s` here)
  std::vector<Op> stage_ops;
  for (int ph = 0; ph < FB_NUM_PHASES; ++ph)
    if (phase_mask & (1u << ph))
      for (auto& sb : h->stage_batches[ph]) {
        fb_handle::StageBatch c = sb;
        stage_ops.push_back(Op{[c](cudaStream_t st) { return launch_stage_batch(c, st); }, FB_OPK_TRANSPOSE, 0.0, sb.bytes, 2, 0, 0});
      }
  std::vector<const Op*> ops;
  size_t si = 0;
  for (int ph = 0; ph < FB_NUM_PHASES; ++ph)
    if (phase_mask & (1u << ph)) {
      for (size_t k = 0; k < h->stage_batches[ph].size(); ++k) ops.push_back(&stage_ops[si++]);
      for (auto& op : h->ops[ph]) ops.push_back(&op);
    }
  const int n = (int)ops.size();
  if (n > cap) return FB_E_ARG;
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CK(cudaEventCreate(&e));
  std::vector<double> acc(n, 0.0);
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(ev[0], s));
    for (int i = 0; i < n; ++i) {
      CK((*ops[i])(s));
      CK(cudaEventRecord(ev[i + 1], s));
    }
    CK(cudaEventSynchronize(ev[n]));
    for (int i = 0; i < n; ++i) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      acc[i] += ms;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  for (int i = 0; i < n; ++i) {
    ms_out[i] = (float)(acc[i] / reps);
    if (kind_out) kind_out[i] = ops[i]->kind;
    if (flops_out) flops_out[i] = ops[i]->flops;
    if (bytes_out) bytes_out[i] = ops[i]->bytes;
  }
  return n;
}

int fb_fused_profile(fb_handle* h, uint32_t phase_mask, int reps, void* stream, float* us_out, int32_t* info_out, int cap) {
  if (!h || !h->bound) return FB_E_STATE;
  if (reps < 1 || cap < 1 || !us_out || !info_out || !fused_wanted(h, phase_mask)) return FB_E_ARG;
  if (h->uses_pairs) return FB_E_STATE;
  cudaStream_t s = (cudaStream_t)stream;
  fb_handle::FusedPlan* plan = nullptr;
  int rc = get_fused_plan(h, phase_mask, s, &plan);
  if (rc != FB_OK) return rc;
  // rows: one per stage of every fused unit, one per stand-alone unit (CUDA events); info = [unit, stage (-1: stand-alone), items,
  // type of the first item (FS_*), its count]
  std::vector<double> acc;
  std::vector<std::array<int32_t, 5>> info;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int r = 0; r < reps; ++r) {
    size_t row = 0;
    for (size_t i = 0; i < plan->units.size(); ++i) {
      const fb_handle::Unit& u = plan->units[i];
      if (u.program >= 0) {
        const FsHeader* hd = reinterpret_cast<const FsHeader*>(h->prog_host.data() + u.program);
        const FsItem* items = reinterpret_cast<const FsItem*>(h->prog_host.data() + u.program + hd->items_off);
        CK(launch_fused(h, u.program, plan->barrier_of[i], s, h->d_fs_times));
        std::vector<unsigned long long> t(hd->n_stages + 1);
        CK(cudaMemcpyAsync(t.data(), h->d_fs_times, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int st = 0; st < hd->n_stages; ++st, ++row) {
          if (r == 0) {
            const int a = hd->first_item[st], b = hd->first_item[st + 1];
            acc.push_back(0.0);
            info.push_back({(int32_t)i, st, b - a, b > a ? items[a].type : 0, b > a ? items[a].count : 0});
          }
          acc[row] += (double)(t[st + 1] - t[st]) * 1e-3;
        }
      } else {
        CK(cudaEventRecord(e0, s));
        if (u.sb) CK(launch_stage_batch(*u.sb, s)); else CK((*u.op)(s));
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r == 0) { acc.push_back(0.0); info.push_back({(int32_t)i, -1, 1, u.sb ? -2 : -1 - u.op->kind, 0}); }
        acc[row++] += (double)ms * 1e3;
      }
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const int n = (int)acc.size();
  if (n > cap) return FB_E_ARG;
  for (int i = 0; i < n; ++i) {
    us_out[i] = (float)(acc[i] / reps);
    for (int j = 0; j < 5; ++j) info_out[5 * i + j] = info[i][j];
  }
  return n;
}

const float* fb_metrics_ptr(const fb_handle* h) { return h && h->bound ? h->d_metrics : nullptr; }

int fb_set_adam_steps(fb_handle* h, int64_t fb_step, int64_t actor_step, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  k_set_adam_steps<<<1, 32, 0, (cudaStream_t)stream>>>(h->d_sc, (long long)fb_step, (long long)actor_step, h->cfg.beta1, h->cfg.beta2);
  CK(cudaGetLastError());
  return FB_OK;
}

int fb_get_adam_steps(fb_handle* h, int64_t* fb_step, int64_t* actor_step, void* stream) {
  if (!h || !h->bound) return FB_E_STATE;
  DevScalars hs;
  CK(cudaMemcpyAsync(&hs, h->d_sc, sizeof(hs), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  if (fb_step) *fb_step = hs.step_fb;
  if (actor_step) *actor_step = hs.step_actor;
  return FB_OK;
}

int fb_gather_block(fb_handle* h, int* floats_per_row, float** d_local, float** d_global) {
  if (!h || !h->bound) return FB_E_STATE;
  if (floats_per_row) *floats_per_row = h->blk_local.ld;
  if (d_local) *d_local = h->blk_local.p;
  if (d_global) *d_global = h->blk_global.p;
  return FB_OK;
}

int fb_workspace_view(fb_handle* h, const char* name, float** d_ptr, int* rows, int* cols, int* ld) {
  if (!h || !h->bound || !name) return FB_E_STATE;
  auto it = h->views.find(name);
  if (it == h->views.end()) return FB_E_ARG;
  if (d_ptr) *d_ptr = it->second.p;
  if (rows) *rows = it->second.rows;
  if (cols) *cols = it->second.cols;
  if (ld) *ld = it->second.ld;
  return FB_OK;
}

// ---- stand-alone operators -----------------------------------------------------------------------
int fb_batch_row_layout(int obs_dim, int action_dim, int goal_dim, int extra_dim, int with_future, int32_t* offsets9, int32_t* pitch) {
  BatchLayout L;
  make_batch_layout(L, obs_dim, action_dim, goal_dim, extra_dim, with_future);
  if (offsets9) {
    offsets9[0] = L.off_obs; offsets9[1] = L.off_action; offsets9[2] = L.off_rd; offsets9[3] = L.off_next_obs; offsets9[4] = L.off_goal;
    offsets9[5] = L.off_next_goal; offsets9[6] = L.off_extra; offsets9[7] = L.off_future_obs; offsets9[8] = L.off_future_goal;
  }
  if (pitch) *pitch = L.pitch;
  return FB_OK;
}

int fb_host_gather_rows(const fb_host_storage* st, const int32_t* ep_idx, const int32_t* step_idx, const int32_t* future_idx,
                        int batch, float replay_discount, float* rows, int pitch) {
  if (!st || !st->observation || !st->action || !st->discount || !ep_idx || !step_idx || !rows || batch < 1) return FB_E_ARG;
  const int O = st->obs_dim, A = st->action_dim, G = st->goal ? st->goal_dim : 0, R = st->rows_per_episode;
  BatchLayout L;
  make_batch_layout(L, O, A, G, 0, future_idx != nullptr);
  if (pitch < L.pitch || O < 1 || A < 1 || R < 2 || st->max_episodes < 1) return FB_E_ARG;
  auto row = [R](const float* base, int dim, int ep, int t) { return base + ((size_t)ep * R + t) * dim; };
  for (int i = 0; i < batch; ++i) {   // indices first: nothing is read through a bad one
    if (ep_idx[i] < 0 || ep_idx[i] >= st->max_episodes || step_idx[i] < 1 || step_idx[i] >= R) return FB_E_ARG;
    if (future_idx && (future_idx[i] < 1 || future_idx[i] > R)) return FB_E_ARG;
  }
  // The rows are random ~100-byte reads of a buffer far larger than the caches: the copy is bound by how many misses are in flight.
  // Each worker first touches every source line of its share (the misses overlap), then copies; a few workers multiply the lines
  // in flight (one core sustains ~12).
  static const int max_workers = getenv("FB_HOST_GATHER_WORKERS") ? std::max(1, atoi(getenv("FB_HOST_GATHER_WORKERS"))) : 4;
  const int workers = batch >= 256 ? max_workers : 1;
#pragma omp parallel for num_threads(workers) schedule(static)
  for (int w = 0; w < workers; ++w) {
    const int i0 = (int)((long long)batch * w / workers), i1 = (int)((long long)batch * (w + 1) / workers);
    for (int i = i0; i < i1; ++i) {
      const int ep = ep_idx[i], t = step_idx[i];
      const float* o0 = row(st->observation, O, ep, t - 1);
      for (int b = 0; b < 2 * O * 4; b += 64) __builtin_prefetch((const char*)o0 + b);   // rows t-1 and t are adjacent
      __builtin_prefetch(row(st->action, A, ep, t));
      __builtin_prefetch(row(st->discount, 1, ep, t));
      if (st->reward) __builtin_prefetch(row(st->reward, 1, ep, t));
      if (G) { const float* g0 = row(st->goal, G, ep, t - 1); for (int b = 0; b < 2 * G * 4; b += 64) __builtin_prefetch((const char*)g0 + b); }
      if (future_idx) {
        const int f = future_idx[i];
        __builtin_prefetch(row(st->observation, O, ep, f - 1));
        if (G) __builtin_prefetch(row(st->goal, G, ep, f - 1));
      }
    }
    for (int i = i0; i < i1; ++i) {
      const int ep = ep_idx[i], t = step_idx[i];
      float* out = rows + (size_t)i * pitch;
      memcpy(out + L.off_obs, row(st->observation, O, ep, t - 1), O * sizeof(float));
      memcpy(out + L.off_action, row(st->action, A, ep, t), A * sizeof(float));
      out[L.off_rd + 0] = st->reward ? *row(st->reward, 1, ep, t) : 0.f;
      out[L.off_rd + 1] = replay_discount * *row(st->discount, 1, ep, t);
      out[L.off_rd + 2] = 0.f; out[L.off_rd + 3] = 0.f;
      memcpy(out + L.off_next_obs, row(st->observation, O, ep, t), O * sizeof(float));
      if (G) {
        memcpy(out + L.off_goal, row(st->goal, G, ep, t - 1), G * sizeof(float));
        memcpy(out + L.off_next_goal, row(st->goal, G, ep, t), G * sizeof(float));
      }
      if (future_idx) {
        const int f = future_idx[i];
        memcpy(out + L.off_future_obs, row(st->observation, O, ep, f - 1), O * sizeof(float));
        if (G) memcpy(out + L.off_future_goal, row(st->goal, G, ep, f - 1), G * sizeof(float));
      }
    }
  }
  return FB_OK;
}

int fb_replay_gather(const fb_replay_view* view, int obs_dim, int action_dim, const int32_t* d_ep_idx, const int32_t* d_step_idx,
                     const int32_t* d_future_idx, int batch, float replay_discount, float* d_out, int out_ld, void* stream) {
  if (!view || !d_ep_idx || !d_step_idx || !d_out || batch < 1) return FB_E_ARG;
  BatchLayout L;
  make_batch_layout(L, obs_dim, action_dim, view->off_goal >= 0 ? view->goal_dim : 0, view->off_extra >= 0 ? view->extra_dim : 0,
                    d_future_idx != nullptr);
  if (out_ld < L.pitch) return FB_E_ARG;
  GatherParams gp;
  int rc = build_gather_params(*view, gp, L, out_ld);
  if (rc != FB_OK) return rc;
  k_gather_rows<<<fb_ceil_div(batch, 8), 256, 0, (cudaStream_t)stream>>>(gp, d_ep_idx, d_step_idx, d_future_idx, batch, nullptr,
                                                                        replay_discount, d_out);
  CK(cudaGetLastError());
  return FB_OK;
}

int fb_replay_pack_episode(const fb_replay_view* view, float* d_rows_mut, int slot, int rows, int obs_dim, int action_dim,
                           const float* d_obs, const float* d_action, const float* d_reward, const float* d_discount, const float* d_goal,
                           const float* d_extra, void* stream) {
  if (!view || !d_rows_mut || slot < 0 || slot >= view->max_episodes || rows < 1 || rows > view->rows_per_episode) return FB_E_ARG;
  if (!d_obs || !d_action || !d_reward || !d_discount) return FB_E_ARG;
  PackEpisodeParams P; memset(&P, 0, sizeof(P));
  P.row_stride = view->row_stride; P.rows = rows; P.O = obs_dim; P.A = action_dim;
  P.G = (view->off_goal >= 0 && d_goal) ? view->goal_dim : 0; P.X = (view->off_extra >= 0 && d_extra) ? view->extra_dim : 0;
  P.off_obs = view->off_obs; P.off_action = view->off_action; P.off_reward = view->off_reward; P.off_goal = view->off_goal;
  P.off_extra = view->off_extra;
  float* dst = d_rows_mut + (size_t)slot * view->rows_per_episode * view->row_stride;
  k_pack_episode<<<rows, 64, 0, (cudaStream_t)stream>>>(dst, P, d_obs, d_action, d_reward, d_discount, d_goal, d_extra);
  CK(cudaGetLastError());
  return FB_OK;
}

int fb_sgemm(const float* dA, const float* dB, float* dC, const float* d_bias, int M, int N, int K, int lda, int ldb, int ldc,
             int a_kmajor, int b_kmajor, int relu, int splitk, int tile_cfg, void* stream) {
  if (!dA || !dB || !dC || M < 1 || N < 1 || K < 1) return FB_E_ARG;
  if (tile_cfg > 4) return FB_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (tile_cfg >= 3) {  // tcgen05 3xTF32 kernel (4: CTA pairs, cta_group::2)
    const int ncta = tile_cfg == 4 ? 2 : 1;
    if (!a_kmajor || !aligned16(dA) || lda % 4) return FB_E_UNSUPPORTED;
    const float* Bp = dB;
    int ldbp = ldb;
    float *tmp = nullptr, *tmp_lo = nullptr;
    if (!b_kmajor) {   // staged transposed copy + its pre-split lo plane (the TC_B_PRE path of the plan)
      ldbp = fb_round_up(K, 4);
      CK(cudaMallocAsync(&tmp, (size_t)N * ldbp * sizeof(float), s));
      CK(cudaMallocAsync(&tmp_lo, (size_t)N * ldbp * sizeof(float), s));
      TransposeDesc t; memset(&t, 0, sizeof(t));
      t.in = dB; t.out = tmp; t.out_lo = tmp_lo; t.rows = K; t.cols = N; t.ld_in = ldb; t.ld_out = ldbp; t.transpose = 1; t.cta_begin = 0; t.ctas_x = fb_ceil_div(N, 32);
      TransposeDesc* dt = nullptr;
      CK(cudaMallocAsync(&dt, sizeof(t), s));
      CK(cudaMemcpyAsync(dt, &t, sizeof(t), cudaMemcpyHostToDevice, s));
      CK(cudaStreamSynchronize(s));
      k_transpose_grouped<<<t.ctas_x * fb_ceil_div(K, 32), 256, 0, s>>>(dt, 1);
      CK(cudaGetLastError());
      CK(cudaFreeAsync(dt, s));
      Bp = tmp;
    } else if (!aligned16(dB) || ldb % 4) {
      return FB_E_UNSUPPORTED;
    }
    TcGemmDesc d; memset(&d, 0, sizeof(d));
    d.C = dC; d.bias = d_bias; d.M = M; d.N = N; d.K = K; d.ldc = ldc; d.flags = relu ? GF_RELU : 0; d.bn = N > 64 ? 128 : 64;
    d.tiles_m = fb_ceil_div(M, TC_BM * ncta); d.tiles_n = fb_ceil_div(N, d.bn); d.work_begin = 0;
    d.splitk = 1; d.kb_per_split = fb_ceil_div(K, TC_BK);
    if (splitk > 1) {   // C must be zeroed by the caller
      if (relu) return FB_E_ARG;
      d.kb_per_split = fb_ceil_div(d.kb_per_split, splitk);
      d.splitk = fb_ceil_div(fb_ceil_div(K, TC_BK), d.kb_per_split);
    }
    d.work_count = d.tiles_m * d.tiles_n * d.splitk;
    int rc = encode_tiled_map(&d.mapA, dA, M, K, lda, TC_BM);
    if (rc == FB_OK) rc = encode_tiled_map(&d.mapB, Bp, N, K, ldbp, d.bn / ncta);
    if (rc == FB_OK && tmp_lo) { rc = encode_tiled_map(&d.mapBlo, tmp_lo, N, K, ldbp, d.bn / ncta); d.flags |= TC_B_PRE; }
    if (rc != FB_OK) return rc;
    TcGemmDesc* dd = nullptr;
    CK(cudaMallocAsync(&dd, sizeof(d), s));
    CK(cudaMemcpyAsync(dd, &d, sizeof(d), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    CK(tc_set_smem_attr());
    TcLaunch hdr;
    make_tc_launch(std::vector<TcGemmDesc>(1, d), d.work_count, d.bn, &hdr);
    CK(tc_launch(dd, hdr, d.work_count, ncta, s, false));
    CK(cudaGetLastError());
    CK(cudaFreeAsync(dd, s));
    if (tmp) CK(cudaFreeAsync(tmp, s));
    if (tmp_lo) CK(cudaFreeAsync(tmp_lo, s));
    CK(cudaStreamSynchronize(s));
    return FB_OK;
  }
  GemmDesc d = gemm_raw(dA, lda, a_kmajor, dB, ldb, b_kmajor, dC, ldc, M, N, K, d_bias, relu ? GF_RELU : 0, nullptr, 0);
  gemm_set_tile(d, tile_cfg >= 0 ? tile_cfg : ((M > 64 && N > 64) ? GEMM_CFG_BIG : GEMM_CFG_SMALL));
  if (splitk < 1) splitk = 1;
  if (splitk > 1) { if (relu) return FB_E_ARG; d.flags |= GF_ATOMIC; }
  d.k_per_split = fb_round_up(fb_ceil_div(K, splitk), GEMM_BK);
  d.splitk = fb_ceil_div(K, d.k_per_split);
  d.a_vec = aligned16(dA) && lda % 4 == 0; d.b_vec = aligned16(dB) && ldb % 4 == 0; d.c_vec = aligned16(dC) && ldc % 4 == 0;
  d.work_begin = 0; d.work_count = d.tiles_m * d.tiles_n * d.splitk;
  GemmDesc* dd = nullptr;
  CK(cudaMallocAsync(&dd, sizeof(GemmDesc), s));
  CK(cudaMemcpyAsync(dd, &d, sizeof(GemmDesc), cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));  // d is a stack object
  CK(cudaFuncSetAttribute(k_gemm_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  k_gemm_grouped<<<d.work_count, GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(dd, 1);
  CK(cudaGetLastError());
  CK(cudaFreeAsync(dd, s));
  return FB_OK;
}

// Timing harness of the tcgen05 grouped GEMM on one synthetic problem replicated `nprob` times (a "group"): `reps` back-to-back
// launches between two CUDA events.  dbg = TC_DBG_* knobs (0: the product kernel).  Synchronises; allocates its own operands.
__global__ void k_fill_hash(float* p, size_t n, unsigned int seed) {   // pseudo-random values in [-1, 1) (bench operands)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned int x = (unsigned int)i * 2654435761u + seed;
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    p[i] = (float)(int)x * (1.0f / 2147483648.0f);
  }
}

int fb_gemm_tc_bench(int M, int N, int K, int bn, int nprob, int splitk, int dbg, int reps, float* ms_per_launch, void* stream) {
  if (M < 1 || N < 1 || K < 8 || K % 4 || nprob < 1 || nprob > 16 || reps < 1 || !ms_per_launch) return FB_E_ARG;
  if (bn != 32 && bn != 64 && bn != 128) return FB_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  float *A = nullptr, *B = nullptr, *Cm = nullptr;
  TcGemmDesc* dd = nullptr;
  const bool pre_b = dbg & (1 << 19), pre_a = dbg & (1 << 20);   // lo planes: a second (zero) copy behind the raw one
  const bool random_fill = dbg & (1 << 21);                       // operands: pseudo-random instead of a constant
  const int ncta = (dbg & (1 << 22)) ? 2 : 1;                     // CTA pairs
  dbg &= ~((1 << 19) | (1 << 20) | (1 << 21) | (1 << 22));
  CK(cudaMallocAsync(&A, (size_t)nprob * M * K * 4 * 2, s));
  CK(cudaMallocAsync(&B, (size_t)nprob * N * K * 4 * 2, s));
  CK(cudaMemsetAsync(A, 0, (size_t)nprob * M * K * 4 * 2, s));
  CK(cudaMemsetAsync(B, 0, (size_t)nprob * N * K * 4 * 2, s));
  CK(cudaMallocAsync(&Cm, (size_t)nprob * M * N * 4, s));
  CK(cudaMallocAsync(&dd, sizeof(TcGemmDesc) * nprob, s));
  CK(cudaMemsetAsync(A, 0x3c, (size_t)nprob * M * K * 4, s));
  CK(cudaMemsetAsync(B, 0x3c, (size_t)nprob * N * K * 4, s));
  if (random_fill) {
    k_fill_hash<<<1184, 256, 0, s>>>(A, (size_t)nprob * M * K * 2, 1u);
    k_fill_hash<<<1184, 256, 0, s>>>(B, (size_t)nprob * N * K * 2, 2u);
  }
  std::vector<TcGemmDesc> v(nprob);
  int work = 0, rc = FB_OK;
  for (int i = 0; i < nprob && rc == FB_OK; ++i) {
    TcGemmDesc& d = v[i]; memset(&d, 0, sizeof(d));
    d.C = Cm + (size_t)i * M * N; d.M = M; d.N = N; d.K = K; d.ldc = N; d.flags = dbg; d.bn = bn;
    d.tiles_m = fb_ceil_div(M, TC_BM * ncta); d.tiles_n = fb_ceil_div(N, bn); d.work_begin = work;
    d.splitk = 1; d.kb_per_split = fb_ceil_div(K, TC_BK);
    if (splitk > 1) { d.kb_per_split = fb_ceil_div(d.kb_per_split, splitk); d.splitk = fb_ceil_div(fb_ceil_div(K, TC_BK), d.kb_per_split); }
    d.work_count = d.tiles_m * d.tiles_n * d.splitk;
    work += d.work_count;
    rc = encode_tiled_map(&d.mapA, A + (size_t)i * M * K, M, K, K, TC_BM);
    if (rc == FB_OK) rc = encode_tiled_map(&d.mapB, B + (size_t)i * N * K, N, K, K, bn / ncta);
    if (rc == FB_OK && pre_a) { rc = encode_tiled_map(&d.mapAlo, A + (size_t)(nprob + i) * M * K, M, K, K, TC_BM); d.flags |= TC_A_PRE; }
    if (rc == FB_OK && pre_b) { rc = encode_tiled_map(&d.mapBlo, B + (size_t)(nprob + i) * N * K, N, K, K, bn / ncta); d.flags |= TC_B_PRE; }
  }
  if (rc == FB_OK) {
    cudaEvent_t e0, e1;
    CK(cudaMemcpyAsync(dd, v.data(), sizeof(TcGemmDesc) * nprob, cudaMemcpyHostToDevice, s));
    CK(tc_set_smem_attr());
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    TcLaunch hdr;
    make_tc_launch(v, work, bn, &hdr);
    for (int r = 0; r < 3; ++r) CK(tc_launch(dd, hdr, work, ncta, s, false));
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r) CK(tc_launch(dd, hdr, work, ncta, s, false));
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  CK(cudaFreeAsync(A, s)); CK(cudaFreeAsync(B, s)); CK(cudaFreeAsync(Cm, s)); CK(cudaFreeAsync(dd, s));
  CK(cudaStreamSynchronize(s));
  return rc;
}

int fb_fp32_peak_tflops(double* out_tflops, void* stream) {
  if (!out_tflops) return FB_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  float* d = nullptr;
  CK(cudaMallocAsync(&d, 16, s));
  const int iters = 1 << 15, blocks = FB_SM_COUNT * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_fma_peak<<<blocks, 256, 0, s>>>(d, 1 << 10);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0, s));
    k_fma_peak<<<blocks, 256, 0, s>>>(d, iters);
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  CK(cudaFreeAsync(d, s));
  *out_tflops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks / ((double)best * 1e-3) / 1e12;
  return FB_OK;
}

}  // extern "C"
