// Non-GEMM kernels of the FB-DDPG step: replay gather, input staging, device RNG, LayerNorm+tanh,
// L2 projection, action sampling, batch x batch loss, Q loss, bias column sums, Adam + soft update.
#pragma once
#include "common.cuh"

#define FB_LN_EPS 1e-5f
#define FB_NORMALIZE_EPS 1e-12f
#define FB_CLAMP_EPS 1e-6f
#define FB_MAX_LN_DIM 2048

// ---- device-resident per-step scalars -------------------------------------------------------------
struct DevScalars {
  float stddev, stddev_clip, lr_forward, lr_backward, lr_actor, tau, replay_discount, replay_future, grad_scale;
  float bc1_fb, bc2s_fb, bc1_actor, bc2s_actor;  // Adam bias corrections 1-beta1^t, sqrt(1-beta2^t) of the NEXT step t = step + 1
  long long step_fb, step_actor;                 // 1-based Adam step counts
  unsigned long long rng_counter;                // bumped once per FB_PHASE_SAMPLE
  unsigned int adam_ticket[2];                   // CTAs of k_adam that have finished (the last one advances the step count)
};

struct HostScalars {  // mirrors fb_step_scalars (fb_b200.h)
  float stddev, stddev_clip, lr_forward, lr_backward, lr_actor, tau, replay_discount, replay_future, grad_scale;
};

__global__ void k_set_scalars(DevScalars* s, HostScalars h) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    s->stddev = h.stddev; s->stddev_clip = h.stddev_clip; s->lr_forward = h.lr_forward;
    s->lr_backward = h.lr_backward; s->lr_actor = h.lr_actor; s->tau = h.tau;
    s->replay_discount = h.replay_discount; s->replay_future = h.replay_future; s->grad_scale = h.grad_scale;
  }
}

__device__ __forceinline__ void adam_bias_corrections(float beta1, float beta2, long long t, float* bc1, float* bc2s) {
  *bc1 = (float)(1.0 - pow((double)beta1, (double)t));
  *bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)t));
}
__global__ void k_set_adam_steps(DevScalars* s, long long fb, long long actor, float beta1, float beta2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    s->step_fb = fb; s->step_actor = actor;
    adam_bias_corrections(beta1, beta2, fb + 1, &s->bc1_fb, &s->bc2s_fb);
    adam_bias_corrections(beta1, beta2, actor + 1, &s->bc1_actor, &s->bc2s_actor);
    s->adam_ticket[0] = 0u; s->adam_ticket[1] = 0u;
  }
}

// which: 0 = fb optimizer, 1 = actor optimizer, 2 = rng counter
__global__ void k_tick(DevScalars* s, int which, float beta1, float beta2) {
  fb_pdl_trigger();
  fb_pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (which == 2) { s->rng_counter += 1ull; return; }
  // runs BEHIND the k_adam launch of step t: publishes t and the bias corrections of step t + 1 (what the next k_adam reads)
  const long long t = (which == 0 ? s->step_fb : s->step_actor) + 1;
  float n1, n2;
  adam_bias_corrections(beta1, beta2, t + 1, &n1, &n2);
  if (which == 0) { s->step_fb = t; s->bc1_fb = n1; s->bc2s_fb = n2; }
  else { s->step_actor = t; s->bc1_actor = n1; s->bc2s_actor = n2; }
}

// ---- replay gather ----------------------------------------------------------------------------------
// Packed storage row (fp32, every field starts on a 16-byte boundary):
//   [observation(O) pad | action(A) pad | reward discount 0 0 | goal(G) pad]
// Packed batch row produced by the gather:
//   [obs | action | reward discount*gamma 0 0 | next_obs | goal | next_goal | future_obs | future_goal]
// (in_memory_replay_buffer.py:162-183).  One warp handles one sampled transition; lane l moves the l-th,
// (l+32)-th ... 128-bit chunk of the output row, so a warp reads the two adjacent storage rows (t-1, t)
// as contiguous coalesced 16-byte loads.
struct GatherSlot {   // one 16-byte chunk of the output row
  short src_row;      // 0: row t-1, 1: row t, 2: row future-1
  short special;      // 1: the [reward, discount, 0, 0] chunk (discount is scaled)
  int src_f4;         // float4 index inside the storage row
  int dst_f4;         // float4 index inside the output row
};
#define FB_MAX_GATHER_SLOTS 192

struct GatherParams {
  const float* rows; const int* ep_len;
  int rows_per_episode, row_stride;  // row_stride in floats (multiple of 4)
  int n_slots, out_ld;               // out_ld in floats (multiple of 4)
  GatherSlot slots[FB_MAX_GATHER_SLOTS];
};

__global__ void __launch_bounds__(256) k_gather_rows(const __grid_constant__ GatherParams gpv, const int* __restrict__ ep_idx,
                                                     const int* __restrict__ step_idx, const int* __restrict__ future_idx,
                                                     int batch, const float* __restrict__ discount_scale_dev,
                                                     float discount_scale_host, float* __restrict__ out) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const GatherParams* gp = &gpv;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= batch) return;
  const float gamma = discount_scale_dev ? *discount_scale_dev : discount_scale_host;
  const int ep = ep_idx[warp], t = step_idx[warp];
  const int fut = future_idx ? future_idx[warp] : t;
  const int stride4 = gp->row_stride >> 2;
  const float4* base = reinterpret_cast<const float4*>(gp->rows) + (size_t)ep * gp->rows_per_episode * stride4;
  const float4* r_prev = base + (size_t)(t - 1) * stride4;
  const float4* r_cur = base + (size_t)t * stride4;
  const float4* r_fut = base + (size_t)(fut - 1) * stride4;
  float4* o = reinterpret_cast<float4*>(out + (size_t)warp * gp->out_ld);
  const int n = gp->n_slots;
  for (int s = lane; s < n; s += 32) {
    const GatherSlot sl = gp->slots[s];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sl.src_row == 0) v = ld_stream_f4(r_prev + sl.src_f4);
    else if (sl.src_row == 1) v = ld_stream_f4(r_cur + sl.src_f4);
    else if (sl.src_row == 2) { if (future_idx) v = ld_stream_f4(r_fut + sl.src_f4); else continue; }
    if (sl.special) { v.y *= gamma; v.z = 0.f; v.w = 0.f; }
    o[sl.dst_f4] = v;
  }
}

// write one finished episode into packed storage (device-side repack of tight [rows, dim] field arrays)
struct PackEpisodeParams { int row_stride, rows, O, A, G, X, off_obs, off_action, off_reward, off_goal, off_extra; };

__global__ void k_pack_episode(float* __restrict__ dst_rows, PackEpisodeParams P, const float* __restrict__ obs,
                               const float* __restrict__ act, const float* __restrict__ rew, const float* __restrict__ disc,
                               const float* __restrict__ goal, const float* __restrict__ extra) {
  const int r = blockIdx.x;
  if (r >= P.rows) return;
  float* d = dst_rows + (size_t)r * P.row_stride;
  for (int c = threadIdx.x; c < P.row_stride; c += blockDim.x) {
    float v = 0.f;
    if (c >= P.off_obs && c < P.off_obs + P.O) v = obs[(size_t)r * P.O + (c - P.off_obs)];
    else if (c >= P.off_action && c < P.off_action + P.A) v = act[(size_t)r * P.A + (c - P.off_action)];
    else if (c == P.off_reward) v = rew[r];
    else if (c == P.off_reward + 1) v = disc[r];
    else if (P.G > 0 && c >= P.off_goal && c < P.off_goal + P.G) v = goal[(size_t)r * P.G + (c - P.off_goal)];
    else if (P.X > 0 && c >= P.off_extra && c < P.off_extra + P.X) v = extra[(size_t)r * P.X + (c - P.off_extra)];
    d[c] = v;
  }
}

// layout of the packed batch row (floats)
struct BatchLayout {
  int O, A, G, X;       // G == 0: no goal columns; X: extra (meta) columns, gathered at t-1
  int with_future;
  int off_obs, off_action, off_rd, off_next_obs, off_goal, off_next_goal, off_extra, off_future_obs, off_future_goal, pitch;
};

// tight user arrays -> packed batch rows (explicit-input path; fb_set_batch)
__global__ void k_pack_batch(BatchLayout L, int batch, const float* __restrict__ obs, const float* __restrict__ action,
                             const float* __restrict__ discount, const float* __restrict__ next_obs,
                             const float* __restrict__ goal, const float* __restrict__ next_goal, float* __restrict__ out) {
  const int r = blockIdx.x;
  if (r >= batch) return;
  float* o = out + (size_t)r * L.pitch;
  for (int c = threadIdx.x; c < L.O; c += blockDim.x) {
    o[L.off_obs + c] = obs[(size_t)r * L.O + c];
    o[L.off_next_obs + c] = next_obs[(size_t)r * L.O + c];
  }
  for (int c = threadIdx.x; c < L.A; c += blockDim.x) o[L.off_action + c] = action[(size_t)r * L.A + c];
  if (L.G > 0 && goal && next_goal)
    for (int c = threadIdx.x; c < L.G; c += blockDim.x) {
      o[L.off_goal + c] = goal[(size_t)r * L.G + c];
      o[L.off_next_goal + c] = next_goal[(size_t)r * L.G + c];
    }
  if (threadIdx.x == 0) { o[L.off_rd] = 0.f; o[L.off_rd + 1] = discount[r]; }
}

// packed batch rows -> the concatenated network inputs of the step.
//   actor_in_o  [2B, ldO ] = [next_obs ; obs]          actor_in_oz [2B, ldOZ] = [next_obs|. ; obs|.]  (z filled later)
//   in_oa [B, ldOA] = [obs|action]   in_noa/in_oa2 [B, ldOA] = [next_obs|.] / [obs|.] (actions filled later)
//   goal_next [B, ldG] = next_goal (or next_obs)        mix_in [B, ldG] = (goal or obs)[perm]   (fb_ddpg.py:460-468)
//   disc -> column `disc_col` of the gather block
struct StageParams {
  BatchLayout L;
  int batch, use_goal;
  float* actor_in_o; int ldO;
  float* actor_in_oz; int ldOZ;
  float* in_oa; float* in_noa; float* in_oa2; int ldOA;
  int act_col;                // column of the action inside in_oa / in_noa / in_oa2: O, or O + Z when z sits between (preprocess = False)
  float* goal_next; float* mix_in; int ldG;
  float* blk; int blk_pitch, disc_col;
  int with_future;            // rows [B, 2B) of mix_in = future_goal / future_obs (hindsight input, not permuted)
  const int* perm;            // null: identity (mix_in then holds un-permuted rows)
  const float* mix_override;  // tight [B, G]: explicit backward_input[perm] from fb_set_batch, or null
};

__device__ __forceinline__ void stage_inputs_body(StageParams P, const float* packed, const int bid) {

  const int r = bid;
  if (r >= P.batch) return;
  const BatchLayout& L = P.L;
  const float* row = packed + (size_t)r * L.pitch;
  const int B = P.batch;
  for (int c = threadIdx.x; c < L.O; c += blockDim.x) {
    const float o = row[L.off_obs + c], no = row[L.off_next_obs + c];
    if (P.actor_in_o) {   // null: no obs-only embed (preprocess = False)
      P.actor_in_o[(size_t)r * P.ldO + c] = no;
      P.actor_in_o[(size_t)(B + r) * P.ldO + c] = o;
    }
    P.actor_in_oz[(size_t)r * P.ldOZ + c] = no;
    P.actor_in_oz[(size_t)(B + r) * P.ldOZ + c] = o;
    P.in_oa[(size_t)r * P.ldOA + c] = o;
    P.in_oa2[(size_t)r * P.ldOA + c] = o;
    P.in_noa[(size_t)r * P.ldOA + c] = no;
  }
  for (int c = threadIdx.x; c < L.A; c += blockDim.x) P.in_oa[(size_t)r * P.ldOA + P.act_col + c] = row[L.off_action + c];
  const int G = P.use_goal ? L.G : L.O;
  const int src = P.perm ? P.perm[r] : r;
  const float* prow = packed + (size_t)src * L.pitch;
  for (int c = threadIdx.x; c < G; c += blockDim.x) {
    P.goal_next[(size_t)r * P.ldG + c] = P.use_goal ? row[L.off_next_goal + c] : row[L.off_next_obs + c];
    float m;
    if (P.mix_override) m = P.mix_override[(size_t)r * G + c];
    else m = P.use_goal ? prow[L.off_goal + c] : prow[L.off_obs + c];
    P.mix_in[(size_t)r * P.ldG + c] = m;
    if (P.with_future) P.mix_in[(size_t)(P.batch + r) * P.ldG + c] = P.use_goal ? row[L.off_future_goal + c] : row[L.off_future_obs + c];
  }
  if (threadIdx.x == 0) P.blk[(size_t)r * P.blk_pitch + P.disc_col] = row[L.off_rd + 1];
}
__global__ void __launch_bounds__(128) k_stage_inputs(StageParams P, const float* __restrict__ packed) {
  fb_pdl_trigger();
  fb_pdl_wait();
  stage_inputs_body(P, packed, blockIdx.x);
}

// ---- device RNG (rng_device = 1) ----------------------------------------------------------------------
// One warp per sampled transition: episode / step / future indices (in_memory_replay_buffer.py:147-161
// semantics: uniform episode, uniform step in [1, len], future = step + Geometric(1-future) clipped),
// the mix mask (fb_ddpg.py:471), z ~ sqrt(Z) * normalize(N(0,I)) (fb_ddpg.py:224-232) and both action-noise rows.
struct RngParams {
  unsigned long long seed;
  int batch, rows_per_episode, Z, A, ldZ, ldA;
  float mix_ratio, future_ratio;
  int norm_z;             // 0: z = sqrt(Z) * U(0,1) (x) normalize(N(0,I))  (cfg.norm_z == False, fb_ddpg.py:230-231)
  int* future_mask;       // null: no hindsight
  const int* n_episodes;  // device scalar: len(buffer)
  const int* ep_len;
  int* ep_idx; int* step_idx; int* future_idx; int* mix_mask; unsigned int* perm_keys;
  float* z_rand; float* noise_fb; float* noise_actor;
};

__global__ void __launch_bounds__(256) k_rng_draw(RngParams P, const DevScalars* __restrict__ sc) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.batch) return;
  Philox ph(P.seed);
  const unsigned long long ctr = sc->rng_counter;
  if (lane == 0) {
    // uniform over the valid transitions (ep, t), t in [1, len(ep)]: draw a padded slot and reject the
    // padding.  Equals "uniform episode, uniform step" for fixed-length buffers and the length-weighted
    // episode draw of in_memory_replay_buffer.py:149-151 for ragged ones.
    const int n_ep = max(*P.n_episodes, 1);
    const int tmax = max(P.rows_per_episode - 1, 1);
    int ep = 0, t = 1, len = 1;
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t attempt = 0; attempt < (P.ep_len ? 64u : 1u); ++attempt) {
      r = ph(ctr, (uint32_t)warp, 8u + attempt);
      if (!P.ep_len) break;   // host-supplied batches: only the mix mask / permutation key of this row are needed
      ep = (int)(((unsigned long long)r.x * (unsigned long long)n_ep) >> 32);
      t = 1 + (int)(((unsigned long long)r.y * (unsigned long long)tmax) >> 32);
      len = P.ep_len[ep];
      if (t <= len) break;
      t = max(len, 1);  // only reached if all 64 attempts hit padding
    }
    int fut = t;
    const float future = sc->replay_future;
    if (future < 1.f) {
      // Geometric(p = 1-future) on {1,2,...} by inversion
      const float u = u01(r.z);
      const float g = (future <= 0.f) ? 1.f : ceilf(logf(u) / logf(future));
      fut = t + (int)fminf(fmaxf(g, 1.f), 1.0e9f);
      fut = min(fut, len);
    }
    P.ep_idx[warp] = ep; P.step_idx[warp] = t; P.future_idx[warp] = fut;
    P.mix_mask[warp] = (u01(r.w) - (1.0f / 16777216.0f) < P.mix_ratio) ? 1 : 0;
    const uint4 q = ph(ctr, (uint32_t)warp, 1u);
    P.perm_keys[warp] = q.x;
    if (P.future_mask) P.future_mask[warp] = (u01(q.y) - (1.0f / 16777216.0f) < P.future_ratio) ? 1 : 0;
  }
  // z: Z normals, 4 per lane per pass
  float ss = 0.f;
  float zv[4];
  for (int base = 0; base < P.Z; base += 128) {
    const uint4 r = ph(ctr, (uint32_t)warp, 1024u + (uint32_t)(base / 128) * 32u + lane);
    const float2 n0 = box_muller(r.x, r.y), n1 = box_muller(r.z, r.w);
    zv[0] = n0.x; zv[1] = n0.y; zv[2] = n1.x; zv[3] = n1.y;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = base + lane * 4 + j;
      if (c < P.Z) { ss += zv[j] * zv[j]; P.z_rand[(size_t)warp * P.ldZ + c] = zv[j]; }
    }
  }
  ss = warp_sum(ss);
  const float scale = sqrtf((float)P.Z) / fmaxf(sqrtf(ss), FB_NORMALIZE_EPS);
  __syncwarp();
  for (int c = lane; c < P.Z; c += 32) {
    float v = P.z_rand[(size_t)warp * P.ldZ + c] * scale;
    if (!P.norm_z) {   // every coordinate scaled by its own U[0,1)
      const uint4 q = ph(ctr, (uint32_t)warp, 2048u + (uint32_t)(c >> 2));
      const uint32_t w = (c & 3) == 0 ? q.x : ((c & 3) == 1 ? q.y : ((c & 3) == 2 ? q.z : q.w));
      v *= u01(w) - (1.0f / 16777216.0f);
    }
    P.z_rand[(size_t)warp * P.ldZ + c] = v;
  }
  // action noise rows
  for (int base = 0; base < P.A; base += 64) {
    const uint4 r = ph(ctr, (uint32_t)warp, 4096u + (uint32_t)(base / 64) * 32u + lane);
    const float2 n0 = box_muller(r.x, r.y), n1 = box_muller(r.z, r.w);
    const int c = base + lane * 2;
    if (c < P.A) { P.noise_fb[(size_t)warp * P.ldA + c] = n0.x; P.noise_actor[(size_t)warp * P.ldA + c] = n1.x; }
    if (c + 1 < P.A) { P.noise_fb[(size_t)warp * P.ldA + c + 1] = n0.y; P.noise_actor[(size_t)warp * P.ldA + c + 1] = n1.y; }
  }
}

// random permutation of [0, n) = argsort of n random keys (torch.randperm, fb_ddpg.py:467): every (key, index) pair is ranked against
// all others held in shared memory (no dependent sort network: one pass, any number of CTAs).  EIGHT lanes share a pair, each scanning
// every eighth 16-byte group of the key table (one thread per pair was a 2 500-instruction serial loop at batch 1024: 13 us at the very
// start of every step, in front of the replay gather).  The lane group that owns pair 0 also bumps the RNG counter (the draws of this
// step are done: k_rng_draw ran before).  Launch with ceil(8 n / 256) CTAs.
__global__ void __launch_bounds__(256) k_randperm(const unsigned int* __restrict__ keys, int n, int* __restrict__ perm, DevScalars* sc) {
  fb_pdl_trigger();
  fb_pdl_wait();
  extern __shared__ __align__(16) unsigned int sk[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sk[i] = keys[i];
  __syncthreads();
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = gt >> 3, sub = gt & 7;
  const bool live = i < n;
  const unsigned int mine = live ? sk[i] : 0u;
  int rank = 0;
  const int n4 = n & ~3;
  if (live) {
    for (int j = sub * 4; j < n4; j += 32) {   // the 8 lanes of a pair read 128 contiguous bytes: conflict-free, broadcast across pairs
      const uint4 k = *reinterpret_cast<const uint4*>(sk + j);
      rank += (k.x < mine || (k.x == mine && j < i)) ? 1 : 0;
      rank += (k.y < mine || (k.y == mine && j + 1 < i)) ? 1 : 0;
      rank += (k.z < mine || (k.z == mine && j + 2 < i)) ? 1 : 0;
      rank += (k.w < mine || (k.w == mine && j + 3 < i)) ? 1 : 0;
    }
    if (sub == 0) {
      for (int j = n4; j < n; ++j) {
        const unsigned int k = sk[j];
        rank += (k < mine || (k == mine && j < i)) ? 1 : 0;
      }
    }
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  rank += __shfl_xor_sync(0xffffffffu, rank, 4);
  if (live && sub == 0) {
    perm[rank] = i;
    if (i == 0) sc->rng_counter += 1ull;
  }
}

// ---- LayerNorm + tanh ("ntanh", fb_modules.py:49-50) ------------------------------------------------
struct LnDesc {
  const float* x; float* y; const float* gamma; const float* beta; float* mean; float* rstd;
  int rows, D, ld, row_begin;  // row_begin: first global warp index of this problem
};

__device__ __forceinline__ void ln_tanh_fwd_body(const LnDesc* descs, const int nprob, int total_rows, const int bid) {

  const int gw = (bid * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= total_rows) return;
  int p = 0;
  while (p + 1 < nprob && descs[p + 1].row_begin <= gw) ++p;
  const LnDesc d = descs[p];
  const int r = gw - d.row_begin;
  const float* x = d.x + (size_t)r * d.ld;
  float s = 0.f;
  for (int c = lane; c < d.D; c += 32) s += x[c];
  const float mean = warp_sum(s) / (float)d.D;
  float v = 0.f;
  for (int c = lane; c < d.D; c += 32) { const float t = x[c] - mean; v += t * t; }
  const float var = warp_sum(v) / (float)d.D;
  const float rstd = 1.0f / sqrtf(var + FB_LN_EPS);
  float* y = d.y + (size_t)r * d.ld;
  for (int c = lane; c < d.D; c += 32) y[c] = tanhf((x[c] - mean) * rstd * __ldg(d.gamma + c) + __ldg(d.beta + c));
  if (lane == 0) { d.mean[r] = mean; d.rstd[r] = rstd; }
}
__global__ void __launch_bounds__(256) k_ln_tanh_fwd(const __grid_constant__ DescTable<LnDesc, 8> T, int total_rows) {
  fb_pdl_trigger();
  fb_pdl_wait();
  ln_tanh_fwd_body(T.d, T.n, total_rows, blockIdx.x);
}

// Vectorised forward for D <= 1024 and 16-byte aligned rows: the row lives in registers (one global read), two-pass
// mean / variance on the register copy like nn.LayerNorm, float4 stores.
// FULL: D == 1024 exactly: no column-bound tests (see ln_tanh_bwd_v4_rows)
template <bool FULL>
__device__ __forceinline__ void ln_tanh_fwd_v4_row(const LnDesc& d, const int r, const int lane) {
  constexpr int NV = 8;
  const int D = d.D;
  const float4* x = reinterpret_cast<const float4*>(d.x + (size_t)r * d.ld);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int q = lane + 32 * i, c = q * 4;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FULL || c < D) {
      v[i] = x[q];
      if (!FULL) {
        if (c + 1 >= D) v[i].y = 0.f;
        if (c + 2 >= D) v[i].z = 0.f;
        if (c + 3 >= D) v[i].w = 0.f;
      }
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / (float)D;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (FULL || c < D) {
      const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
      var += a * a;
      if (FULL || c + 1 < D) var += b * b;
      if (FULL || c + 2 < D) var += e * e;
      if (FULL || c + 3 < D) var += f * f;
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(var) / (float)D + FB_LN_EPS);
  float4* y = reinterpret_cast<float4*>(d.y + (size_t)r * d.ld);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int q = lane + 32 * i, c = q * 4;
    if (FULL || c < D) {
      float o[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      // gamma / beta: 128-bit loads (every parameter tensor starts on a 128-byte boundary and is padded to 32 floats)
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(d.gamma + c)), b4 = __ldg(reinterpret_cast<const float4*>(d.beta + c));
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (FULL || c + e < D) o[e] = tanhf((o[e] - mean) * rstd * gg[e] + bb[e]);
      if (FULL || c + 3 < D) {
        y[q] = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        float* ys = reinterpret_cast<float*>(y + q);
        for (int e = 0; e < 4; ++e)
          if (c + e < D) ys[e] = o[e];
      }
    }
  }
  if (lane == 0) { d.mean[r] = mean; d.rstd[r] = rstd; }
}
__device__ __forceinline__ void ln_tanh_fwd_v4_body(const LnDesc* descs, const int nprob, int total_rows, const int bid) {
  const int gw = (bid * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= total_rows) return;
  int p = 0;
  while (p + 1 < nprob && descs[p + 1].row_begin <= gw) ++p;
  const LnDesc d = descs[p];
  if (d.D == 1024) ln_tanh_fwd_v4_row<true>(d, gw - d.row_begin, lane);
  else ln_tanh_fwd_v4_row<false>(d, gw - d.row_begin, lane);
}
__global__ void __launch_bounds__(256) k_ln_tanh_fwd_v4(const __grid_constant__ DescTable<LnDesc, 8> T, int total_rows) {
  fb_pdl_trigger();
  fb_pdl_wait();
  ln_tanh_fwd_v4_body(T.d, T.n, total_rows, blockIdx.x);
}

struct LnBwdDesc {
  const float* dy;   // grad w.r.t. the tanh output, [rows, ld_dy]
  const float* y;    // saved tanh output
  const float* x;    // saved pre-LayerNorm activations
  const float* gamma; const float* mean; const float* rstd;
  float* dx;         // grad w.r.t. x (may alias dy), [rows, ld_dy]
  float* dgamma; float* dbeta;  // accumulated (atomics); both null when the affine grads are not needed
  int rows, D, ld, ld_dy, cta_begin, cta_count;
};
#define FB_LN_BWD_ROWS_PER_CTA 8

// s_dg, s_db: FB_MAX_LN_DIM floats of shared memory each
__device__ __forceinline__ void ln_tanh_bwd_body(const LnBwdDesc* descs, const int nprob, float* s_dg, float* s_db, const int bid) {

  int p = 0;
  while (p + 1 < nprob && descs[p + 1].cta_begin <= (int)bid) ++p;
  const LnBwdDesc d = descs[p];
  const int cta = bid - d.cta_begin;
  const bool affine = d.dgamma != nullptr;
  if (affine) {
    for (int c = threadIdx.x; c < d.D; c += blockDim.x) { s_dg[c] = 0.f; s_db[c] = 0.f; }
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = cta * FB_LN_BWD_ROWS_PER_CTA;
  const float invD = 1.0f / (float)d.D;
  for (int rr = warp; rr < FB_LN_BWD_ROWS_PER_CTA; rr += 8) {
    const int r = r0 + rr;
    if (r >= d.rows) break;
    const float* dy = d.dy + (size_t)r * d.ld_dy;
    const float* y = d.y + (size_t)r * d.ld;
    const float* x = d.x + (size_t)r * d.ld;
    const float mean = d.mean[r], rstd = d.rstd[r];
    float a = 0.f, b = 0.f;
    for (int c = lane; c < d.D; c += 32) {
      const float yy = y[c];
      const float g = dy[c] * (1.f - yy * yy);
      const float gg = g * __ldg(d.gamma + c);
      const float xh = (x[c] - mean) * rstd;
      a += gg; b += gg * xh;
    }
    a = warp_sum(a) * invD; b = warp_sum(b) * invD;
    float* dx = d.dx + (size_t)r * d.ld_dy;
    for (int c = lane; c < d.D; c += 32) {
      const float yy = y[c];
      const float g = dy[c] * (1.f - yy * yy);
      const float gg = g * __ldg(d.gamma + c);
      const float xh = (x[c] - mean) * rstd;
      dx[c] = rstd * (gg - a - xh * b);
      if (affine) { atomicAdd(&s_dg[c], g * xh); atomicAdd(&s_db[c], g); }
    }
  }
  if (affine) {
    __syncthreads();
    for (int c = threadIdx.x; c < d.D; c += blockDim.x) { atomicAdd(d.dgamma + c, s_dg[c]); atomicAdd(d.dbeta + c, s_db[c]); }
  }
}
__global__ void __launch_bounds__(256) k_ln_tanh_bwd(const __grid_constant__ DescTable<LnBwdDesc, 4> T) {
  fb_pdl_trigger();
  fb_pdl_wait();
  __shared__ float s_dg[FB_MAX_LN_DIM];
  __shared__ float s_db[FB_MAX_LN_DIM];
  ln_tanh_bwd_body(T.d, T.n, s_dg, s_db, blockIdx.x);
}

// Vectorised variant for D <= 1024 with 16-byte aligned rows: each lane keeps its 8 float4 columns of the row in registers
// (one pass over dy / y / x), and its share of dgamma / dbeta in registers across the rows of the CTA (no per-element
// shared-memory atomics); per CTA one shared-memory combine across the 8 warps, then one global atomic per column.
// s_dg, s_db: 1024 floats of shared memory each
// FULL: D == 1024 exactly (the hidden width of the F / actor embeds: the large launches): every lane owns 8 complete float4s, so the
// column-bound tests of the general path (two thirds of its instructions; the kernel is instruction-latency-bound at one CTA of 8 warps
// per SM) compile away
template <bool FULL>
__device__ __forceinline__ void ln_tanh_bwd_v4_rows(const LnBwdDesc& d, float* s_dg, float* s_db, const int cta) {

  constexpr int NV = 8;  // float4s per lane: D <= 32 * 4 * 8 = 1024
  const bool affine = d.dgamma != nullptr;
  const int D = d.D;
  if (affine) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) { s_dg[c] = 0.f; s_db[c] = 0.f; }
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = cta * FB_LN_BWD_ROWS_PER_CTA;
  const float invD = 1.0f / (float)D;
  float4 gam[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    gam[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FULL || c < D) {   // 128-bit load: parameter tensors are 128-byte aligned and padded to 32 floats; columns >= D are masked below
      gam[i] = __ldg(reinterpret_cast<const float4*>(d.gamma + c));
      if (!FULL) {
        if (c + 1 >= D) gam[i].y = 0.f;
        if (c + 2 >= D) gam[i].z = 0.f;
        if (c + 3 >= D) gam[i].w = 0.f;
      }
    }
  }
  float4 adg[NV], adb[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { adg[i] = make_float4(0.f, 0.f, 0.f, 0.f); adb[i] = adg[i]; }
  for (int rr = warp; rr < FB_LN_BWD_ROWS_PER_CTA; rr += 8) {
    const int r = r0 + rr;
    if (r >= d.rows) break;
    const float4* dy = reinterpret_cast<const float4*>(d.dy + (size_t)r * d.ld_dy);
    const float4* y = reinterpret_cast<const float4*>(d.y + (size_t)r * d.ld);
    const float4* x = reinterpret_cast<const float4*>(d.x + (size_t)r * d.ld);
    const float mean = d.mean[r], rstd = d.rstd[r];
    float4 g[NV], xh[NV];
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int q = lane + 32 * i, c = q * 4;
      g[i] = make_float4(0.f, 0.f, 0.f, 0.f); xh[i] = g[i];
      if (FULL || c < D) {
        const float4 dv = dy[q], yv = y[q], xv = x[q];
        g[i].x = dv.x * (1.f - yv.x * yv.x); xh[i].x = (xv.x - mean) * rstd;
        if (FULL || c + 1 < D) { g[i].y = dv.y * (1.f - yv.y * yv.y); xh[i].y = (xv.y - mean) * rstd; }
        if (FULL || c + 2 < D) { g[i].z = dv.z * (1.f - yv.z * yv.z); xh[i].z = (xv.z - mean) * rstd; }
        if (FULL || c + 3 < D) { g[i].w = dv.w * (1.f - yv.w * yv.w); xh[i].w = (xv.w - mean) * rstd; }
        const float4 gg = make_float4(g[i].x * gam[i].x, g[i].y * gam[i].y, g[i].z * gam[i].z, g[i].w * gam[i].w);
        a += gg.x + gg.y + gg.z + gg.w;
        b += gg.x * xh[i].x + gg.y * xh[i].y + gg.z * xh[i].z + gg.w * xh[i].w;
      }
    }
    a = warp_sum(a) * invD; b = warp_sum(b) * invD;
    float4* dx = reinterpret_cast<float4*>(d.dx + (size_t)r * d.ld_dy);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int q = lane + 32 * i, c = q * 4;
      if (FULL || c < D) {
        float4 o;
        o.x = rstd * (g[i].x * gam[i].x - a - xh[i].x * b);
        o.y = rstd * (g[i].y * gam[i].y - a - xh[i].y * b);
        o.z = rstd * (g[i].z * gam[i].z - a - xh[i].z * b);
        o.w = rstd * (g[i].w * gam[i].w - a - xh[i].w * b);
        if (FULL || c + 3 < D) {
          dx[q] = o;
        } else {  // ragged tail (D not a multiple of 4): the row pitch still holds a full float4, only valid lanes are stored
          float* ds = reinterpret_cast<float*>(dx + q);
          ds[0] = o.x;
          if (c + 1 < D) ds[1] = o.y;
          if (c + 2 < D) ds[2] = o.z;
        }
        adg[i].x += g[i].x * xh[i].x; adg[i].y += g[i].y * xh[i].y; adg[i].z += g[i].z * xh[i].z; adg[i].w += g[i].w * xh[i].w;
        adb[i].x += g[i].x; adb[i].y += g[i].y; adb[i].z += g[i].z; adb[i].w += g[i].w;
      }
    }
  }
  if (affine) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (FULL || c < D) {
        atomicAdd(&s_dg[c], adg[i].x); atomicAdd(&s_db[c], adb[i].x);
        if (FULL || c + 1 < D) { atomicAdd(&s_dg[c + 1], adg[i].y); atomicAdd(&s_db[c + 1], adb[i].y); }
        if (FULL || c + 2 < D) { atomicAdd(&s_dg[c + 2], adg[i].z); atomicAdd(&s_db[c + 2], adb[i].z); }
        if (FULL || c + 3 < D) { atomicAdd(&s_dg[c + 3], adg[i].w); atomicAdd(&s_db[c + 3], adb[i].w); }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) { atomicAdd(d.dgamma + c, s_dg[c]); atomicAdd(d.dbeta + c, s_db[c]); }
  }
}
__device__ __forceinline__ void ln_tanh_bwd_v4_body(const LnBwdDesc* descs, const int nprob, float* s_dg, float* s_db, const int bid) {
  int p = 0;
  while (p + 1 < nprob && descs[p + 1].cta_begin <= (int)bid) ++p;
  const LnBwdDesc d = descs[p];
  if (d.D == 1024) ln_tanh_bwd_v4_rows<true>(d, s_dg, s_db, bid - d.cta_begin);
  else ln_tanh_bwd_v4_rows<false>(d, s_dg, s_db, bid - d.cta_begin);
}
__global__ void __launch_bounds__(256) k_ln_tanh_bwd_v4(const __grid_constant__ DescTable<LnBwdDesc, 4> T) {
  fb_pdl_trigger();
  fb_pdl_wait();
  __shared__ float s_dg[1024];
  __shared__ float s_db[1024];
  ln_tanh_bwd_v4_body(T.d, T.n, s_dg, s_db, blockIdx.x);
}
// every problem of the launch has D == 1024: the full-width path alone (a kernel of its own so that its register count, not the general
// path's, decides how many CTAs an SM holds)
__global__ void __launch_bounds__(256) k_ln_tanh_bwd_v4_full(const __grid_constant__ DescTable<LnBwdDesc, 4> T) {
  fb_pdl_trigger();
  fb_pdl_wait();
  __shared__ float s_dg[1024];
  __shared__ float s_db[1024];
  int p = 0;
  while (p + 1 < T.n && T.d[p + 1].cta_begin <= (int)blockIdx.x) ++p;
  ln_tanh_bwd_v4_rows<true>(T.d[p], s_dg, s_db, blockIdx.x - T.d[p].cta_begin);
}

// ---- sqrt(Z) * F.normalize (fb_modules.py:33-40, 227-229) ------------------------------------------
struct L2Desc { const float* x; float* y; float* nrm; int rows, Z, ldx, ldy, row_begin; int normalize; };   // normalize = 0: y = x (norm_z off)

__device__ __forceinline__ void l2norm_fwd_body(const L2Desc* descs, int nprob, int total_rows, const int bid) {

  const int gw = (bid * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= total_rows) return;
  int p = 0;
  while (p + 1 < nprob && descs[p + 1].row_begin <= gw) ++p;
  const L2Desc d = descs[p];
  const int r = gw - d.row_begin;
  const float* x = d.x + (size_t)r * d.ldx;
  float s = 0.f;
  for (int c = lane; c < d.Z; c += 32) s += x[c] * x[c];
  const float nrm = fmaxf(sqrtf(warp_sum(s)), FB_NORMALIZE_EPS);
  const float sq = sqrtf((float)d.Z);
  float* y = d.y + (size_t)r * d.ldy;
  for (int c = lane; c < d.Z; c += 32) y[c] = d.normalize ? sq * (x[c] / nrm) : x[c];
  if (lane == 0 && d.nrm) d.nrm[r] = nrm;
}
__global__ void __launch_bounds__(256) k_l2norm_fwd(const L2Desc* __restrict__ descs, int nprob, int total_rows) {
  fb_pdl_trigger();
  fb_pdl_wait();
  l2norm_fwd_body(descs, nprob, total_rows, blockIdx.x);
}

// dx = (sqrt(Z)/nrm) * (dy - xh * (xh . dy)),  xh = y / sqrt(Z).  dy is the sum of up to three partial gradients plus coef * y
// (the partial products of the tensor-core loss path and the diagonal orthonormality term); the total is also written to dsum.
__device__ __forceinline__ void l2norm_bwd_body(const float* dy0, const float* dy1,
                                                    const float* dy2, int lddy, float coef, float* dsum,
                                                    int ldsum, const float* y, int ldy, const float* nrm,
                                                    float* dx, int lddx, int rows, int Z, int normalize, const int bid) {

  const int r = (bid * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  if (!normalize) {   // norm_z off: the projection is the identity, dx = the summed gradient (the diagonal term survives)
    for (int c = lane; c < Z; c += 32) {
      float g = dy0[(size_t)r * lddy + c] + coef * y[(size_t)r * ldy + c];
      if (dy1) g += dy1[(size_t)r * lddy + c];
      if (dy2) g += dy2[(size_t)r * lddy + c];
      if (dsum) dsum[(size_t)r * ldsum + c] = g;
      dx[(size_t)r * lddx + c] = g;
    }
    return;
  }
  const float sq = sqrtf((float)Z), isq = 1.0f / sq;
  float dot = 0.f;
  for (int c = lane; c < Z; c += 32) {
    const float yv = y[(size_t)r * ldy + c];
    float g = dy0[(size_t)r * lddy + c] + coef * yv;
    if (dy1) g += dy1[(size_t)r * lddy + c];
    if (dy2) g += dy2[(size_t)r * lddy + c];
    if (dsum) dsum[(size_t)r * ldsum + c] = g;
    dot += yv * isq * g;
  }
  dot = warp_sum(dot);
  const float k = sq / nrm[r];
  for (int c = lane; c < Z; c += 32) {
    const float yv = y[(size_t)r * ldy + c];
    float g = dy0[(size_t)r * lddy + c] + coef * yv;
    if (dy1) g += dy1[(size_t)r * lddy + c];
    if (dy2) g += dy2[(size_t)r * lddy + c];
    dx[(size_t)r * lddx + c] = k * (g - yv * isq * dot);
  }
}
__global__ void __launch_bounds__(256) k_l2norm_bwd(const float* __restrict__ dy0, const float* __restrict__ dy1,
                                                    const float* __restrict__ dy2, int lddy, float coef, float* __restrict__ dsum,
                                                    int ldsum, const float* __restrict__ y, int ldy, const float* __restrict__ nrm,
                                                    float* __restrict__ dx, int lddx, int rows, int Z, int normalize) {
  fb_pdl_trigger();
  fb_pdl_wait();
  l2norm_bwd_body(dy0, dy1, dy2, lddy, coef, dsum, ldsum, y, ldy, nrm, dx, lddx, rows, Z, normalize, blockIdx.x);
}

// ---- rand_weight mixing (cfg.rand_weight, fb_ddpg.py:475-482) -----------------------------------------------------
//   weight = U(0,1)^[nmix, B], rows L2-normalised, each scaled by its own U(0,1);  mix_z = weight . backward_net(backward_input[perm])
// The weight rows live in a [B, ldw] block (row s belongs to batch row s, rows outside the mix mask are never read): drawn here
// with Philox (rng_device) or uploaded by the caller (fb_set_mix_weights, the reference's torch.rand draws).
struct MixWeightRngParams { unsigned long long seed; int batch; float* W; int ldw; float* u; };
__global__ void __launch_bounds__(256) k_rng_mix_weights(MixWeightRngParams P, const DevScalars* __restrict__ sc) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const unsigned long long ctr = sc->rng_counter;   // k_randperm (the next launch) bumps it
  Philox ph(P.seed);
  const int q4 = P.ldw / 4;
  const size_t total = (size_t)P.batch * q4;
  const float lo = 1.0f / 16777216.0f;   // u01 is (0, 1]; torch.rand is [0, 1)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(i / q4), c4 = (int)(i - (size_t)s * q4);
    const uint4 r = ph(ctr, (uint32_t)s, 8192u + (uint32_t)c4);
    *reinterpret_cast<float4*>(P.W + (size_t)s * P.ldw + 4 * c4) = make_float4(u01(r.x) - lo, u01(r.y) - lo, u01(r.z) - lo, u01(r.w) - lo);
    if (c4 == 0) P.u[s] = u01(ph(ctr, (uint32_t)s, 2u).x) - lo;
  }
}

struct MixWeightParams {
  int batch, Z;
  const float* b; int ldb;        // backward_net(backward_input[perm]), [batch, Z]
  const float* W; int ldw;        // [batch, ldw] weight rows
  const float* u;                 // [batch] row scales
  const int* mix_mask;
  float* out; int ldo;            // [batch, Z]: (u_s / |W_s|) sum_t W[s,t] b[t]  for the rows of the mix mask
};
#define FB_MIXW_TILE 64
#define FB_MIXW_MAX_Z 128
// one warp per output row, 8 rows per CTA; the CTA walks b in 64-row tiles staged in shared memory
__global__ void __launch_bounds__(256) k_mix_rand_weight(MixWeightParams P) {
  fb_pdl_trigger();
  fb_pdl_wait();
  extern __shared__ float mixw_sb[];   // [FB_MIXW_TILE][Z]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 8 + warp;
  const bool active = s < P.batch && P.mix_mask[s] != 0;   // warp-uniform
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ss = 0.f;
  for (int t0 = 0; t0 < P.batch; t0 += FB_MIXW_TILE) {
    const int nt = min(FB_MIXW_TILE, P.batch - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < nt * P.Z; i += blockDim.x) {
      const int tt = i / P.Z, c = i - tt * P.Z;
      mixw_sb[i] = P.b[(size_t)(t0 + tt) * P.ldb + c];
    }
    __syncthreads();
    if (active) {
      const float w0 = lane < nt ? P.W[(size_t)s * P.ldw + t0 + lane] : 0.f;
      const float w1 = lane + 32 < nt ? P.W[(size_t)s * P.ldw + t0 + 32 + lane] : 0.f;
      ss += w0 * w0 + w1 * w1;
      for (int tt = 0; tt < nt; ++tt) {
        const float wv = __shfl_sync(FB_FULL_MASK, tt < 32 ? w0 : w1, tt & 31);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = lane + 32 * j;
          if (c < P.Z) acc[j] = fmaf(wv, mixw_sb[tt * P.Z + c], acc[j]);
        }
      }
    }
  }
  if (!active) return;
  ss = warp_sum(ss);
  const float scale = P.u[s] / fmaxf(sqrtf(ss), FB_NORMALIZE_EPS);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane + 32 * j;
    if (c < P.Z) P.out[(size_t)s * P.ldo + c] = scale * acc[j];
  }
}

// ---- z finalisation (fb_ddpg.py:470-485): z[mix] = sqrt(Z) normalize(B(backward_input[perm])[mix]) ----
struct ZFinalParams {
  int batch, Z, O;
  const float* z_rand; int ldZ;
  const float* b_mix; int ld_bmix;      // already sqrt(Z)-normalised by BackwardMap; the reference renormalises
  const float* mix_src; int ld_mix_src; // where the mixing rows come from: b_mix, or the rand_weight combinations of its rows
  const int* mix_mask;                  // null: no mixing
  const int* future_mask;               // null: no hindsight; else rows [B, 2B) of b_mix hold backward_net(future goal)
  float* z; float* actor_in_oz; int ldOZ;
  float* f_in[3]; int ldF;              // preprocess = False: z also goes to column O of the three forward-net inputs [obs | z | action]
  int renorm;                           // cfg.norm_z: re-project the mixed rows (fb_ddpg.py:483-484); 0 leaves them raw
};

__device__ __forceinline__ void z_final_body(ZFinalParams P, const int bid) {

  const int r = (bid * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= P.batch) return;
  const bool fut = P.future_mask && P.future_mask[r] != 0;   // applied after the mixing: it wins (fb_ddpg.py:488-491)
  const bool mix = !fut && P.mix_mask && P.mix_mask[r] != 0;
  const bool renorm = mix && P.renorm != 0;
  const float sq = sqrtf((float)P.Z);
  const float* src = fut ? (P.b_mix + (size_t)(P.batch + r) * P.ld_bmix)
                         : (mix ? (P.mix_src + (size_t)r * P.ld_mix_src) : (P.z_rand + (size_t)r * P.ldZ));
  float nrm = 1.f;
  if (renorm) {
    float s = 0.f;
    for (int c = lane; c < P.Z; c += 32) s += src[c] * src[c];
    nrm = fmaxf(sqrtf(warp_sum(s)), FB_NORMALIZE_EPS);
  }
  for (int c = lane; c < P.Z; c += 32) {
    const float v = renorm ? sq * (src[c] / nrm) : src[c];
    P.z[(size_t)r * P.ldZ + c] = v;
    P.actor_in_oz[(size_t)r * P.ldOZ + P.O + c] = v;
    P.actor_in_oz[(size_t)(P.batch + r) * P.ldOZ + P.O + c] = v;
    if (P.f_in[0]) {
#pragma unroll
      for (int i = 0; i < 3; ++i) P.f_in[i][(size_t)r * P.ldF + P.O + c] = v;
    }
  }
}
__global__ void __launch_bounds__(256) k_z_final(ZFinalParams P) {
  fb_pdl_trigger();
  fb_pdl_wait();
  z_final_body(P, blockIdx.x);
}

// ---- action sampling: mu = tanh(pre); TruncatedNormal.sample(clip) (utils.py:164-185) ---------------
// rows [0,B): next_obs side with update_fb's noise -> next_action into in_noa[:, O:]
// rows [B,2B): obs side with update_actor's noise  -> action into in_oa2[:, O:], log-prob metric
struct ActorOutParams {
  int batch, A, O;
  int act_col;   // column of the action inside in_noa / in_oa2 (O, or O + Z for preprocess = False)
  const float* pre; float* mu; int ldA;
  const float* noise_fb; const float* noise_actor; int ldN;
  float* in_noa; float* in_oa2; int ldOA;
  float* next_action; float* action_new;  // tight-ish copies [B, ldA] for inspection
  double* acc;  // acc[ACC_LOGPROB]
};

// ---- accumulators (doubles) ------------------------------------------------------------------------
enum {
  // zeroed at the start of FB_PHASE_FB_FWD
  ACC_OFFDIAG_SQ = 0,  // sum_{s!=t} sum_k (M_k - g*tM)^2
  ACC_DIAG,            // sum_s sum_k M_k[s,s]
  ACC_COV_OFF_SQ,      // sum_{s!=t} Cov^2
  ACC_COV_DIAG,        // sum_s Cov[s,s]
  ACC_TARGET_M,        // sum_{s,t} tM
  ACC_M1,              // sum_{s,t} M1
  ACC_LOGPROB,         // sum_s log N(a; mu, std)
  ACC_QLOSS,           // q_loss = 1: sum_s sum_k (F_k[s].z_s - target_Q[s])^2
  // zeroed at the start of FB_PHASE_ACTOR_FWD
  ACC_Q,               // sum_s min_k Q_k
  ACC_Q1_SUCCESS,      // number of rows with Q1 > Q2 (additional_metric, fb_ddpg.py:403-404)
  // zeroed at the start of FB_PHASE_METRICS
  ACC_F1, ACC_B, ACC_B_NORM, ACC_Z_NORM, ACC_ORTH_SQ, ACC_LINF_BITS,
  ACC_COUNT = 16
};

__device__ __forceinline__ void actor_out_body(ActorOutParams P, const DevScalars* sc, const int bid) {

  const int idx = bid * blockDim.x + threadIdx.x;
  const int total = 2 * P.batch * P.A;
  double lp = 0.0;
  if (idx < total) {
    const int r = idx / P.A, a = idx - r * P.A;
    const float std = sc->stddev, clip = sc->stddev_clip;
    const float mu = tanhf(P.pre[(size_t)r * P.ldA + a]);
    P.mu[(size_t)r * P.ldA + a] = mu;
    const bool fb_side = r < P.batch;
    const int rb = fb_side ? r : r - P.batch;
    const float nz = fb_side ? P.noise_fb[(size_t)rb * P.ldN + a] : P.noise_actor[(size_t)rb * P.ldN + a];
    float eps = nz * std;
    eps = fminf(fmaxf(eps, -clip), clip);
    const float x = mu + eps;
    const float act = fminf(fmaxf(x, -1.0f + FB_CLAMP_EPS), 1.0f - FB_CLAMP_EPS);
    if (fb_side) {
      P.in_noa[(size_t)rb * P.ldOA + P.act_col + a] = act;
      P.next_action[(size_t)rb * P.ldA + a] = act;
    } else {
      P.in_oa2[(size_t)rb * P.ldOA + P.act_col + a] = act;
      P.action_new[(size_t)rb * P.ldA + a] = act;
      const float dlt = act - mu;
      lp = (double)(-(dlt * dlt) / (2.f * std * std) - logf(std) - 0.91893853320467274178f);
    }
  }
  lp = warp_sum_d(lp);
  if ((threadIdx.x & 31) == 0 && lp != 0.0) atomicAdd(P.acc + ACC_LOGPROB, lp);
}
__global__ void __launch_bounds__(256) k_actor_out(ActorOutParams P, const DevScalars* __restrict__ sc) {
  fb_pdl_trigger();
  fb_pdl_wait();
  actor_out_body(P, sc, blockIdx.x);
}

// ---- cfg.boltzmann: DiagGaussianActor + SquashedNormal (fb_modules.py:129-151, utils.py:188-233) -----------------
// pre [2B, ldP] = policy output [mu | raw log-std]; std = exp(lo + (hi - lo) (tanh(raw) + 1) / 2).
// rows [0,B): next_obs side, next_action = tanh(mu + std * noise_fb) (dist.sample(), no clipping, fb_ddpg.py:304-306)
// rows [B,2B): obs side, x = mu + std * noise_actor (rsample), action = tanh(x); log pi(a) = N(x; mu, std) - 2 (log 2 - x - softplus(-2x))
// summed into acc[ACC_LOGPROB]; x, std and tanh(raw) are kept for the backward.
struct ActorOutBzParams {
  int batch, A, act_col;
  const float* pre; int ldP;
  const float* noise_fb; const float* noise_actor; int ldN;
  float* in_noa; float* in_oa2; int ldOA;
  float* next_action; float* action_new; int ldA;
  float* xs; float* sds; float* ts;   // [B, ldA]: pre-tanh sample, std, tanh(raw log-std) of the obs side
  float log_std_min, log_std_max;
  double* acc;
};

__device__ __forceinline__ float fb_softplus(float v) { return v > 20.f ? v : log1pf(expf(v)); }   // F.softplus (beta 1, threshold 20)

__global__ void __launch_bounds__(256) k_actor_out_bz(ActorOutBzParams P) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 2 * P.batch * P.A;
  double lp = 0.0;
  if (idx < total) {
    const int r = idx / P.A, a = idx - r * P.A;
    const float mu = P.pre[(size_t)r * P.ldP + a];
    const float t = tanhf(P.pre[(size_t)r * P.ldP + P.A + a]);
    const float ls = P.log_std_min + 0.5f * (P.log_std_max - P.log_std_min) * (t + 1.f);
    const float sd = expf(ls);
    const bool fb_side = r < P.batch;
    const int rb = fb_side ? r : r - P.batch;
    const float nz = fb_side ? P.noise_fb[(size_t)rb * P.ldN + a] : P.noise_actor[(size_t)rb * P.ldN + a];
    const float x = mu + sd * nz;
    const float act = tanhf(x);
    if (fb_side) {
      P.in_noa[(size_t)rb * P.ldOA + P.act_col + a] = act;
      P.next_action[(size_t)rb * P.ldA + a] = act;
    } else {
      P.in_oa2[(size_t)rb * P.ldOA + P.act_col + a] = act;
      P.action_new[(size_t)rb * P.ldA + a] = act;
      P.xs[(size_t)rb * P.ldA + a] = x; P.sds[(size_t)rb * P.ldA + a] = sd; P.ts[(size_t)rb * P.ldA + a] = t;
      const float d = x - mu;
      const float base = -(d * d) / (2.f * sd * sd) - ls - 0.91893853320467274178f;
      lp = (double)(base - 2.f * (0.69314718055994530942f - x - fb_softplus(-2.f * x)));
    }
  }
  lp = warp_sum_d(lp);
  if ((threadIdx.x & 31) == 0 && lp != 0.0) atomicAdd(P.acc + ACC_LOGPROB, lp);
}

// gradient of mean_s(temp * log pi(a_s) - Q_s) w.r.t. the policy output [mu | raw], from da = d(-mean Q)/d action (the dX product
// through forward_net): with x = mu + std eps, a = tanh(x):  g_x = da (1 - a^2) + (temp / n) 2 a;  d/dmu = g_x;
// d/dstd = g_x eps - (temp / n) / std;  d/draw = d/dstd * std * (hi - lo) / 2 * (1 - tanh(raw)^2)
struct BoltzBwdParams {
  int batch, A;
  const float* da; int ldda;
  const float* xs; const float* sds; const float* ts; int ldA;
  const float* noise_actor; int ldN;
  float* dpre; int ldP;
  float temp_over_n, half_range;
};
__global__ void __launch_bounds__(256) k_boltz_bwd(BoltzBwdParams P) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P.batch * P.A) return;
  const int r = idx / P.A, a = idx - r * P.A;
  const float x = P.xs[(size_t)r * P.ldA + a], sd = P.sds[(size_t)r * P.ldA + a], t = P.ts[(size_t)r * P.ldA + a];
  const float eps = P.noise_actor[(size_t)r * P.ldN + a];
  const float act = tanhf(x);
  const float gx = P.da[(size_t)r * P.ldda + a] * (1.f - act * act) + P.temp_over_n * 2.f * act;
  const float dsd = gx * eps - P.temp_over_n / sd;
  P.dpre[(size_t)r * P.ldP + a] = gx;
  P.dpre[(size_t)r * P.ldP + P.A + a] = dsd * sd * P.half_range * (1.f - t * t);
}

// ---- batch x batch loss, elementwise stage ----------------------------------------------------------
// In:  M1,M2 = F_k . B^T; T1,T2 = tF_k . tB^T; Cov = B . B^T for a block of `nr` rows (global row index
//      row0 + i) by `nc` columns (global).  Out (in place): G1,G2 = dL/dM_k, Gc = ortho_coef * dL_orth/dCov
//      restricted to off-diagonal entries (the diagonal -2 mean(Cov_ss) term is added by k_loss_init_db).
// fb_ddpg.py:320-326,344-347; gradient identities in SURVEY.md section 8a.
struct LossElemParams {
  float* M1; float* M2; const float* T1; const float* T2; float* Cov;
  int nr, nc, ld, row0;
  const float* disc; int disc_stride;  // discount of local row i at disc[i*disc_stride]
  float inv_noff, inv_n, ortho_coef;
  double* acc;
};

__global__ void __launch_bounds__(256) k_fb_loss_elem(LossElemParams P) {
  fb_pdl_trigger();
  fb_pdl_wait();
  double a_off = 0.0, a_diag = 0.0, a_cov = 0.0, a_covd = 0.0, a_tm = 0.0, a_m1 = 0.0;
  for (int row = blockIdx.y; row < P.nr; row += gridDim.y) {
  const float g = P.disc[(size_t)row * P.disc_stride];
  const int diag_col = P.row0 + row;
  const size_t base = (size_t)row * P.ld;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < P.nc; c += gridDim.x * blockDim.x) {
    const float m1 = P.M1[base + c], m2 = P.M2[base + c];
    const float tm = fminf(P.T1[base + c], P.T2[base + c]);
    const float cv = P.Cov[base + c];
    a_tm += tm; a_m1 += m1;
    if (c != diag_col) {
      const float d1 = m1 - g * tm, d2 = m2 - g * tm;
      a_off += (double)d1 * d1 + (double)d2 * d2;
      a_cov += (double)cv * cv;
      P.M1[base + c] = d1 * P.inv_noff;
      P.M2[base + c] = d2 * P.inv_noff;
      P.Cov[base + c] = P.ortho_coef * 4.f * P.inv_noff * cv;
    } else {
      a_diag += (double)m1 + (double)m2;
      a_covd += cv;
      P.M1[base + c] = -P.inv_n;
      P.M2[base + c] = -P.inv_n;
      P.Cov[base + c] = 0.f;
    }
  }
  }
  __shared__ double red[6][8];
  double vals[6] = {a_off, a_diag, a_cov, a_covd, a_tm, a_m1};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double v = warp_sum_d(vals[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
    const int slot[6] = {ACC_OFFDIAG_SQ, ACC_DIAG, ACC_COV_OFF_SQ, ACC_COV_DIAG, ACC_TARGET_M, ACC_M1};
    atomicAdd(P.acc + slot[threadIdx.x], s);
  }
}

// Column-block twin of k_fb_loss_elem: rows are the LOCAL columns t of the loss matrices, columns run over
// all global rows s:  Mt_k[t,s] = B_t . F_k[s],  Tt_k[t,s] = tB_t . tF_k[s].  Out (in place): Gt_k = dL/dM_k[s,t].
// The discount now belongs to the column (s).  No accumulators (the row-block pass owns the loss sums).
struct LossElemTParams {
  float* M1; float* M2; const float* T1; const float* T2;
  int nr, nc, ld, row0;
  const float* disc; int disc_stride;  // discount of GLOBAL row s at disc[s*disc_stride]
  float inv_noff, inv_n;
};

__global__ void __launch_bounds__(256) k_fb_loss_elem_t(LossElemTParams P) {
  fb_pdl_trigger();
  fb_pdl_wait();
  for (int row = blockIdx.y; row < P.nr; row += gridDim.y) {
    const int diag_col = P.row0 + row;
    const size_t base = (size_t)row * P.ld;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < P.nc; c += gridDim.x * blockDim.x) {
      if (c != diag_col) {
        const float g = __ldg(P.disc + (size_t)c * P.disc_stride);
        const float tm = fminf(P.T1[base + c], P.T2[base + c]);
        P.M1[base + c] = (P.M1[base + c] - g * tm) * P.inv_noff;
        P.M2[base + c] = (P.M2[base + c] - g * tm) * P.inv_noff;
      } else {
        P.M1[base + c] = -P.inv_n;
        P.M2[base + c] = -P.inv_n;
      }
    }
  }
}

// dB starts from the diagonal term of the orthonormality loss: d(-2 c mean_s Cov_ss)/dB_s = -(4c/n) B_s
__global__ void k_loss_init_db(float* __restrict__ dB, int lddb, const float* __restrict__ Bm, int ldb, int rows, int Z, float coef) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Z) return;
  const int r = idx / Z, c = idx - r * Z;
  dB[(size_t)r * lddb + c] = coef * Bm[(size_t)r * ldb + c];
}

// ---- actor Q loss (fb_ddpg.py:400-406): Q = min_k F_k . z, loss = -mean Q ------------------------
__device__ __forceinline__ void actor_q_body(const float* F1, const float* F2, int ldf,
                                                 const float* z, int ldz, float* dF1,
                                                 float* dF2, int lddf, int rows, int Z, float inv_n, double* acc, const int bid) {

  const int r = (bid * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  float q1 = 0.f, q2 = 0.f;
  for (int c = lane; c < Z; c += 32) {
    const float zz = z[(size_t)r * ldz + c];
    q1 += F1[(size_t)r * ldf + c] * zz;
    q2 += F2[(size_t)r * ldf + c] * zz;
  }
  q1 = warp_sum(q1); q2 = warp_sum(q2);
  // torch.min(Q1, Q2) routes the gradient to the smaller entry (ties: split evenly)
  const float w1 = (q1 < q2) ? 1.f : ((q1 == q2) ? 0.5f : 0.f);
  const float w2 = 1.f - w1;
  for (int c = lane; c < Z; c += 32) {
    const float gz = -z[(size_t)r * ldz + c] * inv_n;
    dF1[(size_t)r * lddf + c] = w1 * gz;
    dF2[(size_t)r * lddf + c] = w2 * gz;
  }
  if (lane == 0) {
    atomicAdd(acc + ACC_Q, (double)fminf(q1, q2));
    if (q1 > q2) atomicAdd(acc + ACC_Q1_SUCCESS, 1.0);
  }
}
__global__ void __launch_bounds__(256) k_actor_q(const float* __restrict__ F1, const float* __restrict__ F2, int ldf,
                                                 const float* __restrict__ z, int ldz, float* __restrict__ dF1,
                                                 float* __restrict__ dF2, int lddf, int rows, int Z, float inv_n, double* acc) {
  fb_pdl_trigger();
  fb_pdl_wait();
  actor_q_body(F1, F2, ldf, z, ldz, dF1, dF2, lddf, rows, Z, inv_n, acc, blockIdx.x);
}

// ---- optional Q loss of update_fb (cfg.q_loss, fb_ddpg.py:330-341) -------------------------------------------
//   next_Q = min_k tF_k[s].z_s;  cov = B^T B / n;  implicit_reward[s] = (B_s cov^-1) . z_s;  target_Q = implicit_reward + g_s next_Q
//   q_loss = sum_k mean_s (F_k[s].z_s - target_Q[s])^2, all of target_Q under no_grad  =>  dF_k[s] += coef (2/n) (Q_k - target_Q) z_s
// Three launches: cov (fp64 accumulation, one CTA per row, the work shape of k_metric_cov), its inverse (one CTA, Gauss-Jordan
// with partial pivoting on [cov | I] in shared memory, fp64 — the matrix is z_dim x z_dim), and a row pass (one warp per sample).
__global__ void __launch_bounds__(256) k_qloss_cov(const float* __restrict__ Bm, int ldb, int rows, int Z, double* __restrict__ cov) {
  fb_pdl_trigger();
  fb_pdl_wait();
  __shared__ double part[8][128];
  const int a = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < Z; b0 += 128) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = warp; r < rows; r += 8) {
      const float* row = Bm + (size_t)r * ldb;
      const double va = (double)__ldg(row + a);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = b0 + j * 32 + lane;
        if (b < Z) s[j] = fma(va, (double)__ldg(row + b), s[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) part[warp][j * 32 + lane] = s[j];
    __syncthreads();
    if (threadIdx.x < 128) {
      const int b = b0 + threadIdx.x;
      if (b < Z) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
        cov[(size_t)a * Z + b] = t / (double)rows;
      }
    }
    __syncthreads();
  }
}

#define FB_QLOSS_MAX_Z 118   // [Z][2Z] doubles of k_qloss_inverse must fit the 227 KB of shared memory a CTA can opt into
#define FB_QLOSS_INV_THREADS 512
__global__ void __launch_bounds__(FB_QLOSS_INV_THREADS) k_qloss_inverse(const double* __restrict__ cov, int Z, double* __restrict__ inv) {
  fb_pdl_trigger();
  fb_pdl_wait();
  extern __shared__ double aug[];   // [Z][2Z] = [cov | I] -> [I | cov^-1]
  __shared__ double fcol[FB_QLOSS_MAX_Z];
  __shared__ int piv_row;
  const int W = 2 * Z, tid = threadIdx.x, lane = threadIdx.x & 31;
  for (int i = tid; i < Z * W; i += blockDim.x) {
    const int r = i / W, c = i - r * W;
    aug[i] = c < Z ? cov[(size_t)r * Z + c] : (c - Z == r ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int k = 0; k < Z; ++k) {
    if (tid < 32) {   // pivot: the largest |entry| of column k at or below the diagonal (ties: the lowest row index)
      double best = -1.0;
      int bi = k;
      for (int r = k + lane; r < Z; r += 32) {
        const double v = fabs(aug[r * W + k]);
        if (v > best) { best = v; bi = r; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FB_FULL_MASK, best, o);
        const int oi = __shfl_xor_sync(FB_FULL_MASK, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) piv_row = bi;
    }
    __syncthreads();
    const int pr = piv_row;
    if (pr != k)
      for (int c = tid; c < W; c += blockDim.x) {
        const double t = aug[k * W + c];
        aug[k * W + c] = aug[pr * W + c];
        aug[pr * W + c] = t;
      }
    __syncthreads();
    const double p = aug[k * W + k];
    __syncthreads();   // every thread holds the pivot before row k is scaled
    for (int c = tid; c < W; c += blockDim.x) aug[k * W + c] /= p;
    for (int r = tid; r < Z; r += blockDim.x) fcol[r] = (r == k) ? 0.0 : aug[r * W + k];   // column k of the OTHER rows (row k: only its
    __syncthreads();                                                                      // entry [k][k] is touched by the scaling)
    for (int i = tid; i < Z * W; i += blockDim.x) {
      const int r = i / W, c = i - r * W;
      const double f = fcol[r];
      if (f != 0.0) aug[i] -= f * aug[k * W + c];
    }
    __syncthreads();
  }
  for (int i = tid; i < Z * Z; i += blockDim.x) {
    const int r = i / Z, c = i - r * Z;
    inv[i] = aug[r * W + Z + c];
  }
}

struct QLossParams {
  const float *F1, *F2, *tF1, *tF2, *Bm; int ldblk;   // local rows of the [F1|F2|tF1|tF2|B|tB|discount] block
  const float* disc; int disc_stride;
  const float* z; int ldz;
  float* dF1; float* dF2; int lddf;                   // dL/dF_k of the batch x batch loss, the Q-loss term is added in place
  const double* inv;                                  // cov^-1, [Z, Z]
  int rows, Z;
  float gcoef;                                        // q_loss_coef * 2 / n
  double* acc;
};
__global__ void __launch_bounds__(256) k_qloss_rows(QLossParams P) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= P.rows) return;
  const size_t ob = (size_t)r * P.ldblk;
  const float* zr = P.z + (size_t)r * P.ldz;
  double imp = 0.0;
  float q1 = 0.f, q2 = 0.f, n1 = 0.f, n2 = 0.f;
  for (int j = lane; j < P.Z; j += 32) {
    double u = 0.0;   // (B_s cov^-1)_j : lane j reads column j of cov^-1 (coalesced across the warp)
    for (int i = 0; i < P.Z; ++i) u = fma((double)__ldg(P.Bm + ob + i), __ldg(P.inv + (size_t)i * P.Z + j), u);
    const float zz = zr[j];
    imp = fma(u, (double)zz, imp);
    q1 = fmaf(P.F1[ob + j], zz, q1); q2 = fmaf(P.F2[ob + j], zz, q2);
    n1 = fmaf(P.tF1[ob + j], zz, n1); n2 = fmaf(P.tF2[ob + j], zz, n2);
  }
  imp = warp_sum_d(imp);
  q1 = warp_sum(q1); q2 = warp_sum(q2); n1 = warp_sum(n1); n2 = warp_sum(n2);
  const float target = (float)imp + P.disc[(size_t)r * P.disc_stride] * fminf(n1, n2);
  const float e1 = q1 - target, e2 = q2 - target;
  for (int j = lane; j < P.Z; j += 32) {
    const float gz = P.gcoef * zr[j];
    P.dF1[(size_t)r * P.lddf + j] += e1 * gz;
    P.dF2[(size_t)r * P.lddf + j] += e2 * gz;
  }
  if (lane == 0) atomicAdd(P.acc + ACC_QLOSS, (double)e1 * e1 + (double)e2 * e2);
}

// ---- bias gradients: column sums over the batch ----------------------------------------------------
struct ColsumDesc { const float* src; float* dst; int rows, N, ld, cta_begin, ctas_n, ctas_r; };
#define FB_COLSUM_ROWS_PER_CTA 128

// red_smem: 8 x 33 floats of shared memory
__device__ __forceinline__ void colsum_body(const ColsumDesc* descs, int nprob, float* red_smem, const int bid) {

  int p = 0;
  while (p + 1 < nprob && descs[p + 1].cta_begin <= (int)bid) ++p;
  const ColsumDesc d = descs[p];
  const int local = bid - d.cta_begin;
  const int cn = local % d.ctas_n, cr = local / d.ctas_n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = cn * 32 + tx;
  const int r0 = cr * FB_COLSUM_ROWS_PER_CTA;
  const int r1 = min(d.rows, r0 + FB_COLSUM_ROWS_PER_CTA);
  float s = 0.f;
  if (col < d.N)
    for (int r = r0 + ty; r < r1; r += 8) s += d.src[(size_t)r * d.ld + col];
  float (*red)[33] = reinterpret_cast<float (*)[33]>(red_smem);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < d.N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(d.dst + col, t);
  }
}
__global__ void __launch_bounds__(256) k_colsum(const ColsumDesc* __restrict__ descs, int nprob) {
  fb_pdl_trigger();
  fb_pdl_wait();
  __shared__ float red_smem[8 * 33];
  colsum_body(descs, nprob, red_smem, blockIdx.x);
}

// ---- Adam (torch.optim.Adam defaults, fb_ddpg.py:146-151) + target soft update (utils.py:66-69) -------
// p,g,m,v are flat fp32 segments of n floats (multiple of 4).  Elements [0, split) use lr_a, the rest lr_b
// (the fb optimizer's two param groups).  `target` (nullable) is lerped towards the NEW parameters.
// Gradients are cleared after use so the next step's atomically-accumulated dW start from zero.
__global__ void __launch_bounds__(256) k_adam(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                              float4* __restrict__ v, float4* __restrict__ target, size_t n4, size_t split4,
                                              DevScalars* __restrict__ sc, int which, float beta1, float beta2, float eps) {
  fb_pdl_trigger();
  fb_pdl_wait();
  // this launch is Adam step t = (steps so far) + 1; its bias corrections were published by the previous step (or by
  // fb_bind / fb_set_adam_steps).  The step count and the corrections of step t + 1 are published by k_tick, a one-thread launch
  // on the side lane right behind this kernel: the streaming CTAs end without a ticket (a ticket at the end of every CTA cost this
  // pass a sixth of its bandwidth: 0.68 instead of 0.81 of the copy peak, profiles/r1d) and nothing of it sits on the critical path.
  const float bc1 = which == 0 ? sc->bc1_fb : sc->bc1_actor;
  const float bc2s = which == 0 ? sc->bc2s_fb : sc->bc2s_actor;
  const float lr_a = which == 0 ? sc->lr_forward : sc->lr_actor;
  const float lr_b = which == 0 ? sc->lr_backward : sc->lr_actor;
  const float tau = sc->tau, gs = sc->grad_scale;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float step_size = (i < split4 ? lr_a : lr_b) / bc1;
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* pa = reinterpret_cast<float*>(&pp); float* ga = reinterpret_cast<float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm); float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * gs;
      ma[j] = ma[j] * beta1 + gr * (1.f - beta1);
      va[j] = va[j] * beta2 + (gr * gr) * (1.f - beta2);
      const float denom = sqrtf(va[j]) / bc2s + eps;
      pa[j] = pa[j] - step_size * (ma[j] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (target) {
      float4 tt = target[i];
      tt.x = tau * pp.x + (1.f - tau) * tt.x; tt.y = tau * pp.y + (1.f - tau) * tt.y;
      tt.z = tau * pp.z + (1.f - tau) * tt.z; tt.w = tau * pp.w + (1.f - tau) * tt.w;
      target[i] = tt;
    }
  }
}

// ---- metrics (fb_ddpg.py:356-377, 413-418) -----------------------------------------------------------
// sums over the local rows of F1, B, |B_s|, |z_s|
__global__ void __launch_bounds__(256) k_metric_rows(const float* __restrict__ F1, int ldf, const float* __restrict__ Bm, int ldb,
                                                     const float* __restrict__ z, int ldz, int rows, int Z, double* acc) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  float f = 0.f, b = 0.f, bb = 0.f, zz = 0.f;
  for (int c = lane; c < Z; c += 32) {
    const float bv = Bm[(size_t)r * ldb + c], zv = z[(size_t)r * ldz + c];
    f += F1[(size_t)r * ldf + c]; b += bv; bb += bv * bv; zz += zv * zv;
  }
  f = warp_sum(f); b = warp_sum(b); bb = warp_sum(bb); zz = warp_sum(zz);
  if (lane == 0) {
    atomicAdd(acc + ACC_F1, (double)f); atomicAdd(acc + ACC_B, (double)b);
    atomicAdd(acc + ACC_B_NORM, (double)sqrtf(bb)); atomicAdd(acc + ACC_Z_NORM, (double)sqrtf(zz));
  }
}

// cov += B_chunk^T B_chunk for a chunk of FB_COV_ROWS batch rows (B^T B of the whole batch = the sum over the chunks; Z x Z fp32,
// zeroed by the phase's memset): the chunk is staged in shared memory, thread t forms outputs t, t + 256, ... (a, b) = (idx / Z,
// idx % Z) and adds them with one atomic each.  (The first version gave one CTA per ROW of the covariance and walked the whole batch
// in each: 50 CTAs, 115 us at batch 1024.)  eye_diff = cov / n - I and its norms are taken by k_metric_final.
#define FB_COV_ROWS 64
__global__ void __launch_bounds__(256) k_metric_cov(const float* __restrict__ Bm, int ldb, int rows, int Z, float* __restrict__ cov, int chunk) {
  fb_pdl_trigger();
  fb_pdl_wait();
  extern __shared__ float cov_tile[];   // [chunk][Z + 1], chunk <= FB_COV_ROWS (fewer for a very wide z: 48 KB of shared memory)
  const int r0 = blockIdx.x * chunk, nr = min(chunk, rows - r0), zp = Z + 1;
  for (int i = threadIdx.x; i < nr * Z; i += blockDim.x) {
    const int r = i / Z, c = i - r * Z;
    cov_tile[r * zp + c] = Bm[(size_t)(r0 + r) * ldb + c];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < Z * Z; idx += blockDim.x) {
    const int a = idx / Z, b = idx - a * Z;
    float s0 = 0.f, s1 = 0.f;
    int r = 0;
    for (; r + 1 < nr; r += 2) {
      s0 = fmaf(cov_tile[r * zp + a], cov_tile[r * zp + b], s0);
      s1 = fmaf(cov_tile[(r + 1) * zp + a], cov_tile[(r + 1) * zp + b], s1);
    }
    if (r < nr) s0 = fmaf(cov_tile[r * zp + a], cov_tile[r * zp + b], s0);
    atomicAdd(cov + idx, s0 + s1);
  }
}

struct MetricFinalParams {
  const double* acc; const float* cov; float* out;   // cov: B^T B of the (global) batch, Z x Z (k_metric_cov)
  int n_local, n_global, Z; float ortho_coef;
  float q_loss_coef;   // 0 when cfg.q_loss is off (the accumulator then stays 0)
  float temp;          // cfg.boltzmann: actor_loss = mean(temp * log pi - Q) (fb_ddpg.py:406); 0 otherwise
};
// indices must match FB_M_* in fb_b200.h
__global__ void k_metric_final(MetricFinalParams P) {
  fb_pdl_trigger();
  fb_pdl_wait();
  if (blockIdx.x != 0) return;
  // eye_diff = B^T B / n - I: its max-norm and Frobenius norm (fb_ddpg.py:366-370), 256 threads over the Z x Z entries
  __shared__ float s_mx[8];
  __shared__ double s_sq[8];
  float mx = 0.f;
  double sq = 0.0;
  for (int idx = threadIdx.x; idx < P.Z * P.Z; idx += blockDim.x) {
    const int a = idx / P.Z, b = idx - a * P.Z;
    const float e = P.cov[idx] / (float)P.n_global - (a == b ? 1.f : 0.f);
    mx = fmaxf(mx, fabsf(e)); sq += (double)(e * e);
  }
  mx = warp_max(mx); sq = warp_sum_d(sq);
  if ((threadIdx.x & 31) == 0) { s_mx[threadIdx.x >> 5] = mx; s_sq[threadIdx.x >> 5] = sq; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  mx = 0.f; sq = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { mx = fmaxf(mx, s_mx[w]); sq += s_sq[w]; }
  const double n = (double)P.n_global, nl = (double)P.n_local;
  const double noff = n * (n - 1.0);
  const double fb_off = 0.5 * P.acc[ACC_OFFDIAG_SQ] / noff;
  const double fb_diag = -P.acc[ACC_DIAG] / n;
  const double orth_off = P.acc[ACC_COV_OFF_SQ] / noff;
  const double orth_diag = -2.0 * P.acc[ACC_COV_DIAG] / n;
  const double orth = orth_off + orth_diag;
  float* o = P.out;
  o[0] = (float)(P.acc[ACC_TARGET_M] / (nl * n));
  o[1] = (float)(P.acc[ACC_M1] / (nl * n));
  o[2] = (float)(P.acc[ACC_F1] / (nl * P.Z));
  o[3] = (float)(P.acc[ACC_B] / (nl * P.Z));
  o[4] = (float)(P.acc[ACC_B_NORM] / nl);
  o[5] = (float)(P.acc[ACC_Z_NORM] / nl);
  const double q_loss = P.acc[ACC_QLOSS] / n;
  o[6] = (float)(fb_off + fb_diag + (double)P.q_loss_coef * q_loss + (double)P.ortho_coef * orth);
  o[7] = (float)fb_diag;
  o[8] = (float)fb_off;
  o[9] = (float)orth;
  o[10] = (float)orth_diag;
  o[11] = (float)orth_off;
  o[12] = mx;
  o[13] = (float)(sqrt(sq) / sqrt((double)P.Z));
  o[14] = (float)(((double)P.temp * P.acc[ACC_LOGPROB] - P.acc[ACC_Q]) / n);
  o[15] = (float)(P.acc[ACC_Q] / n);
  o[16] = (float)(P.acc[ACC_LOGPROB] / n);
  o[17] = (float)q_loss;
  o[18] = (float)(P.acc[ACC_Q1_SUCCESS] / n);
}

// ---- fp32 FMA-chain microbenchmark (roofline denominator for the CUDA-core GEMMs) ---------------------
// ---- inference plans (FB_PHASE_INFER_*) --------------------------------------------------------------
// [obs | z] rows for the actor's obs_z_net (fb_modules.py:114)
__global__ void k_infer_concat(const float* __restrict__ obs, int ldo, const float* __restrict__ z, int ldz, float* __restrict__ out, int ldout,
                               int rows, int O, int Z) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int r = blockIdx.x;
  if (r >= rows) return;
  for (int c = threadIdx.x; c < O + Z; c += blockDim.x) out[(size_t)r * ldout + c] = c < O ? obs[(size_t)r * ldo + c] : z[(size_t)r * ldz + c - O];
}
// mu = tanh(policy output)  (fb_modules.py:121)
__global__ void k_infer_tanh(const float* __restrict__ pre, float* __restrict__ mu, int ld, int rows, int A) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * A) { const int r = i / A, a = i - r * A; mu[(size_t)r * ld + a] = tanhf(pre[(size_t)r * ld + a]); }
}
// cfg.boltzmann: out [rows, 2A] = [mu (pre-tanh mean) | std] of the SquashedNormal (fb_modules.py:141-151)
__global__ void k_infer_gauss(const float* __restrict__ pre, float* __restrict__ out, int ld, int rows, int A, float lo, float hi) {
  fb_pdl_trigger();
  fb_pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * A) {
    const int r = i / A, a = i - r * A;
    out[(size_t)r * ld + a] = pre[(size_t)r * ld + a];
    out[(size_t)r * ld + A + a] = expf(lo + 0.5f * (hi - lo) * (tanhf(pre[(size_t)r * ld + A + a]) + 1.f));
  }
}
// zsum[c] += sum_r reward[r] * b[r, c]   (fb_ddpg.py:215: z = reward^T . B): block = 32 columns, 8 warps stride the rows
__global__ void __launch_bounds__(256) k_infer_weighted_colsum(const float* __restrict__ b, int ldb, const float* __restrict__ reward, int ldr,
                                                               int rows, int Z, float* __restrict__ zsum) {
  fb_pdl_trigger();
  fb_pdl_wait();
  __shared__ float part[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < Z)
    for (int r = warp; r < rows; r += 8) s = fmaf(__ldg(reward + (size_t)r * ldr), __ldg(b + (size_t)r * ldb + c), s);
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < Z) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    atomicAdd(zsum + c, t);
  }
}

__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 1.000001f, c = 1e-7f;
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) out[0] = a0;
}
