// Host-side launch plan of one FB-DDPG gradient step (fb_ddpg.py:427-520) on one B200.
//
// The plan is built once per handle (fb_bind): every activation lives at a fixed address inside the
// caller's workspace, every kernel argument is a constant, so a phase is a fixed launch sequence that can
// be replayed eagerly or as a CUDA graph.  Layers of different networks that sit at the same depth of the
// step's dependency graph are issued as ONE grouped SGEMM launch (gemm_simt.cuh) so a launch fills 148 SMs.
#pragma once
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/fb_b200.h"
#include "gemm_simt.cuh"
#include "kernels.cuh"
#include "contract_tc.cuh"
#include "gemm_tc.cuh"
#include "fused.cuh"
#include "p2p.cuh"

#define FB_NUM_PHASES 13
#define FB_INFER_ROWS 8   // rows of the per-environment-step inference plans (act / get_goal_meta / compute_z_correl)
#define FB_DESC_ARENA_BYTES (1u << 20)
#define FB_PROG_ARENA_BYTES (1u << 20)
#define FB_SM_COUNT 148

struct Mat {
  float* p = nullptr;
  int rows = 0, cols = 0, ld = 0;
  Mat rs(int r0, int n) const { Mat m = *this; m.p = p + (size_t)r0 * ld; m.rows = rows ? n : 0; return m; }   // a row block of an absent (zero-row) matrix is absent
  Mat cs(int c0, int n) const { Mat m = *this; m.p = p + c0; m.cols = n; return m; }
};

struct TensorInfo { std::string name; size_t off; int rows, cols; size_t numel() const { return (size_t)rows * (cols ? cols : 1); } };

struct SegmentLayout {
  std::vector<TensorInfo> t;
  size_t size = 0;
  int add(const std::string& name, int rows, int cols) {
    TensorInfo ti{name, size, rows, cols};
    t.push_back(ti);
    size += (ti.numel() + 31) / 32 * 32;  // every tensor starts on a 128-byte boundary
    return (int)t.size() - 1;
  }
};

typedef std::function<cudaError_t(cudaStream_t)> OpFn;
// what a launch looks like as an item of the fused stack kernel (fused.cuh): body type, virtual blocks, argument block.
// type FS_NONE: the launch has no fused form and runs as a kernel of its own (it ends the current fused segment).
struct DevItem {
  int type = FS_NONE, count = 0;
  std::vector<char> args;
  template <typename T> void set(int t, int n, const T& a) {
    type = t; count = n;
    args.assign(reinterpret_cast<const char*>(&a), reinterpret_cast<const char*>(&a) + sizeof(T));
  }
};
// kinds of launch, reported by fb_profile_ops (FB_OPK_* in fb_b200.h)
struct Op {
  OpFn fn;
  int kind;
  double flops;  // algorithmic FLOPs of this launch (2 x MACs for GEMM groups)
  double bytes;  // algorithmic bytes read + written
  int lane;      // 0: main stream; 1: side stream
  int join;      // main-lane op that must wait for the side-lane work issued before it
  int fork;      // side-lane op that must first wait for the main-lane work issued before it (otherwise it only follows the
                 // side-lane ops issued before it)
  int replay_only = 0;  // skipped under FB_RUN_HOST_BATCH (the replay gather)
  int wait_stage = 0;   // consumes operands staged for this phase on the staging lane: waits for the phase's staging event
  DevItem dev;          // fused form (FS_NONE: stand-alone only)
  cudaError_t operator()(cudaStream_t s) const { return fn(s); }
};

struct fb_handle {
  fb_config cfg;
  // parameter layout: fb segment = forward_net tensors [0,20) then backward_net tensors [20,28); actor segment
  SegmentLayout seg_fb, seg_actor;
  int fwd_first = 0, bwd_first = 0;
  size_t bwd_offset = 0;  // float offset of the first backward_net tensor inside the fb segment
  fb_buffers bufs;
  bool bound = false;
  size_t ws_bytes = 0;
  // workspace bump allocator
  char* ws_base = nullptr;
  size_t ws_off = 0;
  // descriptor arena (host mirror, uploaded once at bind)
  std::vector<char> arena;
  // plan
  std::vector<Op> ops[FB_NUM_PHASES];
  std::map<std::string, Mat> views;
  std::map<uint32_t, cudaGraphExec_t> graphs;
  cudaStream_t capture_stream = nullptr;
  size_t contract_smem = 0;  // dynamic shared memory of k_contract_tc (0: SIMT contraction)
  size_t qloss_smem = 0;     // dynamic shared memory of k_qloss_inverse (0: cfg.q_loss off)
  bool uses_gemm_tc = false;
  bool uses_pairs = false;   // some GEMM launch of the plan is a CTA-pair (cta_group::2) launch: no fused (k_fused_stack) execution
  // Operands whose source is final before the consuming phase starts (weights, activations of earlier phases) are staged
  // (aligned / transposed copies, pre-split lo planes, zero fills) on a staging lane, as early as their source allows: an
  // entry of phase P with availability a is launched at the start of the first phase >= a of the running mask, so that the
  // staging of FB_FWD / FB_BWD / ACTOR_BWD operands hides behind the phases before them.
  std::vector<TransposeDesc> early_stage[FB_NUM_PHASES];
  std::vector<int> early_avail[FB_NUM_PHASES];   // per entry: first phase at whose start the source is final
  struct StageBatch { int avail; const TransposeDesc* d_descs; int n, ctas; double bytes; DevItem dev; };
  // fused execution (fused.cuh): per phase mask, the plan cut into units = fused segments (one k_fused_stack launch walking a
  // program of stages) and the launches that stay kernels of their own
  struct Unit { int program = -1; const Op* op = nullptr; const StageBatch* sb = nullptr; };   // program >= 0: offset of its header in the program arena
  struct FusedPlan { std::vector<Unit> units; std::vector<int> barrier_of; };
  std::map<uint32_t, FusedPlan> fused_plans;
  std::vector<char> prog_host;        // host mirror of the program arena
  size_t prog_uploaded = 0;           // bytes of prog_host already on the device
  char* d_prog = nullptr;             // program arena (workspace)
  unsigned long long* d_fs_barrier = nullptr;   // FS_NUM_BARRIERS monotonic grid-barrier counters
  unsigned int* d_fs_err = nullptr;
  unsigned long long* d_fs_times = nullptr;     // [FS_MAX_STAGES + 1] stage stamps of a profiled fused launch
  int n_fs_barriers = 0;
  int sm_count = FB_SM_COUNT;
  std::vector<StageBatch> stage_batches[FB_NUM_PHASES];
  size_t ws_fwd_end = 0;             // workspace offset below which every buffer is written by a forward phase
  cudaStream_t side_stream = nullptr, stage_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_stage_fork = nullptr;
  cudaEvent_t ev_stage[FB_NUM_PHASES] = {};
  // fixed workspace objects
  DevScalars* d_sc = nullptr;
  double* d_acc = nullptr;
  unsigned int* d_linf = nullptr;
  float* d_metrics = nullptr;
  int* d_n_episodes = nullptr;
  int *d_ep_idx = nullptr, *d_step_idx = nullptr, *d_future_idx = nullptr, *d_perm = nullptr, *d_mix_mask = nullptr, *d_future_mask = nullptr;
  unsigned int* d_perm_keys = nullptr;
  int* d_identity_perm = nullptr;
  Mat packed, z_rand, noise_fb, noise_actor, blk_local, blk_global;
  Mat mix_w;                 // cfg.rand_weight: [batch, batch] weight rows
  float* mix_u = nullptr;    //                  [batch] row scales
  BatchLayout bl;
  fb_replay_view replay;
  bool replay_bound = false;
  bool have_perm = false, have_mix_mask = false;
  // data-parallel exchange inside the step (fb_nccl_init before fb_bind): NCCL communicator of this rank, loaded with dlopen
  void* nccl_comm = nullptr;
  int nccl_world = 1, nccl_rank = 0;
  // data-parallel exchange by the library's own kernels over NVLink peer memory (p2p.cuh; fb_p2p_create / fb_p2p_attach before fb_bind)
  bool p2p_on = false, p2p_attached = false;
  P2pPeers p2p;
  char* p2p_arena = nullptr;          // this rank's arena (owned: cudaMalloc)
  bool p2p_ipc[P2P_MAX_WORLD] = {};   // peer arenas opened through CUDA IPC (to be closed)
  size_t p2p_bytes = 0, p2p_off_grad_fb = 0, p2p_off_grad_actor = 0, p2p_off_param_fb = 0, p2p_off_param_actor = 0, p2p_off_blk = 0;
};

// ------------------------------------------------------------------------------------------------
// layout (registration order of fb_modules.py: Actor :81-108, ForwardMap :154-185, BackwardMap :211-221)
// ------------------------------------------------------------------------------------------------
static void add_embed(SegmentLayout& s, const std::string& p, int in_dim, int H, int Fd) {
  s.add(p + ".0.weight", H, in_dim); s.add(p + ".0.bias", H, 0);
  s.add(p + ".1.weight", H, 0); s.add(p + ".1.bias", H, 0);
  s.add(p + ".3.weight", Fd, H); s.add(p + ".3.bias", Fd, 0);
}
static void add_head(SegmentLayout& s, const std::string& p, int in_dim, int H, int out) {
  s.add(p + ".0.weight", H, in_dim); s.add(p + ".0.bias", H, 0);
  s.add(p + ".2.weight", out, H); s.add(p + ".2.bias", out, 0);
}
static void build_layout(fb_handle* h) {
  const fb_config& c = h->cfg;
  const int O = c.obs_dim, A = c.action_dim, Z = c.z_dim, G = c.goal_dim, H = c.hidden_dim, Fd = c.feature_dim, Hb = c.backward_hidden_dim;
  SegmentLayout& f = h->seg_fb;
  h->fwd_first = 0;
  const bool deep = c.no_preprocess != 0;         // preprocess = False: trunk.{0,1,3} = one embed-shaped block of width H, trunk.5 = Linear(H -> H)
  const int head_in = (c.add_trunk || deep) ? H : 2 * Fd;   // add_trunk: trunk = Linear(2 Fd -> H) + ReLU in front of the heads
  if (deep) {
    add_embed(f, "trunk", O + Z + A, H, H);
    f.add("trunk.5.weight", H, H); f.add("trunk.5.bias", H, 0);
  } else {
    add_embed(f, "obs_action_net", O + A, H, Fd);
    add_embed(f, "obs_z_net", O + Z, H, Fd);
    if (c.add_trunk) { f.add("trunk.0.weight", H, 2 * Fd); f.add("trunk.0.bias", H, 0); }
  }
  add_head(f, "F1", head_in, H, Z);
  add_head(f, "F2", head_in, H, Z);
  h->bwd_first = (int)f.t.size();
  h->bwd_offset = f.size;
  if (!c.debug_identity_b) {   // (cfg.debug: the backward map is nn.Identity, no tensors)
    f.add("B.0.weight", Hb, G); f.add("B.0.bias", Hb, 0); f.add("B.1.weight", Hb, 0); f.add("B.1.bias", Hb, 0);
    f.add("B.3.weight", Hb, Hb); f.add("B.3.bias", Hb, 0); f.add("B.5.weight", Z, Hb); f.add("B.5.bias", Z, 0);
  }
  SegmentLayout& a = h->seg_actor;
  if (c.boltzmann) {   // DiagGaussianActor: policy = mlp(obs + z, hidden, "ntanh", hidden, "relu", 2 * action)  (fb_modules.py:137)
    a.add("policy.0.weight", H, O + Z); a.add("policy.0.bias", H, 0); a.add("policy.1.weight", H, 0); a.add("policy.1.bias", H, 0);
    a.add("policy.3.weight", H, H); a.add("policy.3.bias", H, 0); a.add("policy.5.weight", 2 * A, H); a.add("policy.5.bias", 2 * A, 0);
    return;
  }
  if (deep) {
    add_embed(a, "trunk", O + Z, H, H);
    a.add("trunk.5.weight", H, H); a.add("trunk.5.bias", H, 0);
  } else {
    add_embed(a, "obs_net", O, H, Fd);
    add_embed(a, "obs_z_net", O + Z, H, Fd);
    if (c.add_trunk) { a.add("trunk.0.weight", H, 2 * Fd); a.add("trunk.0.bias", H, 0); }
  }
  add_head(a, "policy", head_in, H, A);
}

// a parameter set: values + (optional) gradient buffer sharing one layout
struct PSet {
  const SegmentLayout* L; int first; float* val; float* grad;
  // an index past the layout (a network variant without that tensor: the plan builds its launches with zero rows) reads as empty
  bool has(int i) const { return first + i >= 0 && first + i < (int)L->t.size(); }
  Mat w(int i) const { Mat m; if (!has(i)) return m; const TensorInfo& t = L->t[first + i]; m.p = val + t.off; m.rows = t.rows; m.cols = t.cols ? t.cols : 1; m.ld = m.cols; return m; }
  float* v(int i) const { return has(i) ? val + L->t[first + i].off : nullptr; }
  Mat gw(int i) const { Mat m = w(i); if (has(i)) m.p = grad + L->t[first + i].off; return m; }
  float* gv(int i) const { return (grad && has(i)) ? grad + L->t[first + i].off : nullptr; }
  PSet sub(int d) const { PSet p = *this; p.first += d; return p; }
};

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
static void* ws_alloc(fb_handle* h, size_t bytes) {
  h->ws_off = (h->ws_off + 255) / 256 * 256;
  void* p = h->ws_base + h->ws_off;
  h->ws_off += bytes;
  return p;
}
static Mat ws_mat(fb_handle* h, int rows, int cols, const char* name = nullptr) {
  Mat m; m.rows = rows; m.cols = cols; m.ld = fb_round_up(cols, 4);
  m.p = (float*)ws_alloc(h, (size_t)rows * m.ld * sizeof(float));
  if (name) h->views[name] = m;
  return m;
}
template <typename T>
static T* arena_put(fb_handle* h, const std::vector<T>& v, char* d_arena) {
  size_t off = (h->arena.size() + 127) / 128 * 128;  // tensor maps inside descriptors need 64-byte alignment
  h->arena.resize(off + v.size() * sizeof(T));
  memcpy(h->arena.data() + off, v.data(), v.size() * sizeof(T));
  return reinterpret_cast<T*>(d_arena + off);
}

// ------------------------------------------------------------------------------------------------
// grouped GEMM construction
// ------------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static GemmDesc gemm_raw(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor, float* C, int ldc, int M, int N,
                         int K, const float* bias, int flags, const float* mask, int ldmask) {
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  d.A = A; d.B = B; d.C = C; d.bias = bias; d.mask = mask;
  d.M = M; d.N = N; d.K = K; d.K2 = 0;
  d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.ldmask = ldmask;
  d.a_kmajor = a_kmajor; d.b_kmajor = b_kmajor;
  d.flags = flags; d.alpha = 1.f; d.splitk = 1;
  return d;
}
// Y[rows, out] = X . W^T + b            (nn.Linear forward)
static GemmDesc lin_fwd(const Mat& X, const Mat& W, const float* bias, const Mat& Y, int flags) {
  return gemm_raw(X.p, X.ld, 1, W.p, W.ld, 1, Y.p, Y.ld, X.rows, W.rows, W.cols, bias, flags, nullptr, 0);
}
// dX[rows, in] = dY . W   (optionally masked by a saved activation)
static GemmDesc lin_dx(const Mat& dY, const Mat& W, const Mat& dX, int flags, const Mat* mask) {
  return gemm_raw(dY.p, dY.ld, 1, W.p, W.ld, 0, dX.p, dX.ld, dY.rows, W.cols, W.rows, nullptr, flags, mask ? mask->p : nullptr,
                  mask ? mask->ld : 0);
}
// dW[out, in] += dY^T . X  (accumulates into the zeroed flat gradient; split-K over the batch)
static GemmDesc lin_dw(const Mat& dY, const Mat& X, const Mat& dW) {
  GemmDesc d = gemm_raw(dY.p, dY.ld, 0, X.p, X.ld, 0, dW.p, dW.ld, dW.rows, dW.cols, dY.rows, nullptr, GF_ATOMIC, nullptr, 0);
  d.splitk = -1;  // let finalize choose
  return d;
}

struct GroupLaunch { const GemmDesc* d_descs; int nprob, ctas; double flops, bytes; };

static void gemm_set_tile(GemmDesc& d, int cfg) {
  d.cfg = cfg;
  const int bm = cfg == GEMM_CFG_SMALL ? 64 : 128, bn = cfg == GEMM_CFG_BIG ? 128 : 64;
  d.tiles_m = fb_ceil_div(d.M, bm); d.tiles_n = fb_ceil_div(d.N, bn);
}

static GroupLaunch finalize_group(fb_handle* h, std::vector<GemmDesc> g, char* d_arena) {
  // Tile configuration: the largest CTA tile that still gives the launch >= 2 CTAs per SM (a lone 8-warp CTA cannot hide
  // the shared-memory latency of the FFMA loop); narrow outputs (N <= 64) never use the 128-wide tile.
  const int target = 2 * FB_SM_COUNT;
  const int order[3] = {GEMM_CFG_BIG, GEMM_CFG_WIDE, GEMM_CFG_SMALL};
  int total = 0;
  for (int c = 0; c < 3; ++c) {
    total = 0;
    for (auto& d : g) {
      int cfg = order[c];
      if (d.N <= 64 && cfg == GEMM_CFG_BIG) cfg = GEMM_CFG_WIDE;
      if (d.M <= 64) cfg = GEMM_CFG_SMALL;
      gemm_set_tile(d, cfg);
      int sk_possible = 1;
      if ((d.flags & GF_ATOMIC) && d.K2 == 0 && !(d.flags & GF_RELU)) sk_possible = d.K / 256 > 0 ? d.K / 256 : 1;
      total += d.tiles_m * d.tiles_n * (sk_possible > 8 ? 8 : sk_possible);
    }
    if (total >= target) break;
  }
  int tiles_total = 0;
  for (auto& d : g) tiles_total += d.tiles_m * d.tiles_n;
  int work = 0;
  for (auto& d : g) {
    const int tiles = d.tiles_m * d.tiles_n;
    int sk = 1;
    const bool can_split = (d.flags & GF_ATOMIC) && d.K2 == 0 && !(d.flags & GF_RELU);
    if (can_split && tiles_total < target) {
      sk = fb_ceil_div(target, tiles_total);
      const int max_sk = d.K / 256 > 0 ? d.K / 256 : 1;
      if (sk > max_sk) sk = max_sk;
      if (sk > 8) sk = 8;
    }
    int kps = fb_round_up(fb_ceil_div(d.K, sk), GEMM_BK);
    sk = fb_ceil_div(d.K, kps);
    d.splitk = sk; d.k_per_split = kps;
    d.a_vec = aligned16(d.A) && (d.lda % 4 == 0) && (d.K2 == 0 || aligned16(d.A2));
    d.b_vec = aligned16(d.B) && (d.ldb % 4 == 0) && (d.K2 == 0 || aligned16(d.B2));
    d.c_vec = aligned16(d.C) && (d.ldc % 4 == 0);
    d.work_begin = work; d.work_count = tiles * sk;
    work += d.work_count;
  }
  GroupLaunch gl;
  gl.flops = 0.0; gl.bytes = 0.0;
  for (auto& d : g) {
    const double k = (double)d.K + (double)d.K2;
    gl.flops += 2.0 * d.M * (double)d.N * k;
    gl.bytes += 4.0 * ((double)d.M * k + (double)d.N * k + (double)d.M * d.N * ((d.flags & GF_ATOMIC) ? 2.0 : 1.0));
  }
  gl.d_descs = arena_put(h, g, d_arena);
  gl.nprob = (int)g.size(); gl.ctas = work;
  return gl;
}
