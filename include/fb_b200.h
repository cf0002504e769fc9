/*
 * fb_b200.h — C ABI of libfb_b200.so: the FB-DDPG gradient step on one B200 (sm_100a).
 *
 * The reference (facebookresearch/controllable_agent) is pure Python/PyTorch and has no FFI; its
 * seam for this path is the Python method surface of FBDDPGAgent / ReplayBuffer.  This header is
 * the boundary a maintainer would bind (ctypes stub in INTEGRATION.md) to replace, per entry point:
 *
 *   fb_replay_*           url_benchmark/in_memory_replay_buffer.py:104-133 (add), :139-190 (sample)
 *                         url_benchmark/replay_buffer.py:50-63 (EpisodeBatch.to)
 *   fb_set_batch/z/noise  explicit-input form of fb_ddpg.py:433-468 (used by parity tests and by a
 *                         host-resident replay buffer feeding the device step)
 *   FB_PHASE_SAMPLE       fb_ddpg.py:433-434,451,467,471 + utils.py:178 (index/z/noise draws + gather)
 *   FB_PHASE_MIX          fb_ddpg.py:470-485 (z mixing through backward_net)
 *   FB_PHASE_FB_FWD/LOSS/BWD   fb_ddpg.py:303-348,380-383 (update_fb forward, loss incl. the optional Q loss :330-341, backward)
 *   FB_PHASE_FB_ADAM      fb_ddpg.py:384 (fb_opt.step) fused with utils.py:66-69 soft_update_params
 *                         (fb_ddpg.py:500-503; legal because update_actor never writes F/B)
 *   FB_PHASE_ACTOR_FWD/BWD     fb_ddpg.py:389-410 (update_actor forward, Q loss, backward)
 *   FB_PHASE_ACTOR_ADAM   fb_ddpg.py:411 (actor_opt.step)
 *
 * Conventions: every pointer argument named d_* is a DEVICE pointer into caller-owned fp32 (or
 * int32 where stated) contiguous memory; h_* is a HOST pointer.  The library never allocates
 * parameter/optimizer/replay memory and never takes ownership.  Calls enqueue work on `stream`
 * (a cudaStream_t passed as void*) and do not synchronise unless documented.  Return value:
 * 0 = ok, negative = argument/state error (FB_E_*), positive = cudaError_t.  A handle may be used
 * from one host thread at a time.  No mutable global state (one lazily dlopen'ed table of NCCL entry points is shared by the handles of a process).
 */
#ifndef FB_B200_H
#define FB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ABI_VERSION 6

enum {
  FB_OK = 0,
  FB_E_ARG = -1,        /* bad argument */
  FB_E_STATE = -2,      /* called before fb_bind / unsupported in this state */
  FB_E_UNSUPPORTED = -3 /* configuration branch the kernels do not implement */
};

enum { FB_CONTRACT_TCGEN05 = 0, FB_CONTRACT_SIMT = 1 };
enum { FB_MLP_TCGEN05 = 0, FB_MLP_SIMT = 1 };

/* nets, in the order their flat segments are described */
enum { FB_NET_FORWARD = 0, FB_NET_BACKWARD = 1, FB_NET_ACTOR = 2 };

/* phases of one gradient step; OR them into fb_run's mask */
enum {
  FB_PHASE_SAMPLE = 1 << 0,     /* device RNG draws (if rng_device) + replay gather into the step inputs */
  FB_PHASE_MIX = 1 << 1,        /* B(backward_input[perm]) on mix rows -> z */
  FB_PHASE_FB_FWD = 1 << 2,     /* actor(next_obs,z), actor(obs,z), F_tgt, B_tgt, F, B forwards */
  FB_PHASE_FB_LOSS = 1 << 3,    /* batch x batch contraction, fb/orth losses, dF1 dF2 dB */
  FB_PHASE_FB_BWD = 1 << 4,     /* backward through forward_net / backward_net -> grad_fb */
  FB_PHASE_FB_ADAM = 1 << 5,    /* Adam on F,B + target soft update + grad clear */
  FB_PHASE_ACTOR_FWD = 1 << 6,  /* F(obs,z,a) with updated F, Q = min_k F_k.z, actor loss */
  FB_PHASE_ACTOR_BWD = 1 << 7,  /* dQ -> dF -> da -> backward through actor -> grad_actor */
  FB_PHASE_ACTOR_ADAM = 1 << 8, /* Adam on actor + grad clear */
  FB_PHASE_METRICS = 1 << 9,    /* finalise the metrics block (means, orth_linf, orth_l2) */
  FB_PHASE_ALL = (1 << 10) - 1, /* the gradient step */
  /* Inference plans (no gradients, online networks; fb_ddpg.py:177-222,258-289), each run on its own with fb_run(mask, ...).
   * Inputs / outputs are the named workspace blocks below (fb_workspace_view), FB_INFER_ROWS = 8 rows for the per-step ones:
   *   FB_PHASE_INFER_ACTOR   "infer_obs" [8, obs], "infer_z" [8, z]  ->  "infer_mu" [8, action] = tanh(policy(obs, z))   (act)
   *   FB_PHASE_INFER_B       "infer_goal" [8, goal]  ->  "infer_b" [8, z] = sqrt(z_dim) normalize(backward_net(goal))
   *                          (get_goal_meta, compute_z_correl)
   *   FB_PHASE_INFER_BN      "infer_goal_batch" [batch, goal], "infer_reward" [batch, 1]  ->  "infer_zsum" [1, z] +=
   *                          sum_i reward_i * backward_net(goal_i)   (infer_meta_from_obs_and_rewards, one chunk of `batch` rows;
   *                          the caller zeroes infer_zsum before the first chunk and pads the last chunk with zero rewards) */
  FB_PHASE_INFER_ACTOR = 1 << 10,
  FB_PHASE_INFER_B = 1 << 11,
  FB_PHASE_INFER_BN = 1 << 12,
  /* modifier of FB_PHASE_SAMPLE: the caller supplied the batch rows (fb_upload_batch / fb_set_batch), e.g. sampled from a
   * host-resident replay buffer (in_memory_replay_buffer.py:139-190 run by the caller): the device RNG draws of the phase
   * still run (rng_device = 1), the replay gather is skipped and no replay needs to be bound */
  FB_RUN_HOST_BATCH = 1 << 15,
  /* modifier: run every launch of the plan as a kernel of its own (main / side / staging lanes).  Default (flag clear, tensor-core
   * plan): runs of consecutive launches execute inside ONE persistent kernel per segment (k_fused_stack: the fused forward /
   * backward MLP-stack kernels — stages separated by a device-side grid barrier instead of kernel boundaries); both paths run the
   * same device code per launch and produce the same values */
  FB_RUN_UNFUSED = 1 << 14
};

/* index of each scalar in the metrics block (float[FB_METRIC_COUNT]) — keys of the dict returned
 * by FBDDPGAgent.update, fb_ddpg.py:357-377,414-418 */
enum {
  FB_M_TARGET_M = 0, FB_M_M1, FB_M_F1, FB_M_B, FB_M_B_NORM, FB_M_Z_NORM, FB_M_FB_LOSS, FB_M_FB_DIAG,
  FB_M_FB_OFFDIAG, FB_M_ORTH_LOSS, FB_M_ORTH_LOSS_DIAG, FB_M_ORTH_LOSS_OFFDIAG, FB_M_ORTH_LINF, FB_M_ORTH_L2,
  FB_M_ACTOR_LOSS, FB_M_Q, FB_M_ACTOR_LOGPROB,
  FB_M_Q_LOSS,      /* cfg.q_loss (fb_ddpg.py:366-367); 0 otherwise */
  FB_M_Q1_SUCCESS,  /* mean(Q1 > Q2) of update_actor, reported when cfg.additional_metric (fb_ddpg.py:403-404,416-417) */
  FB_M_COUNT_USED,
  FB_METRIC_COUNT = 32
};

typedef struct fb_config {
  int32_t abi_version;         /* FB_ABI_VERSION */
  int32_t batch;               /* rows this process handles per step (B, or B/world for sharded runs) */
  int32_t global_batch;        /* n in the loss normalisation (1/n, 1/(n(n-1))); == batch on one GPU */
  int32_t row_offset;          /* global row index of local row 0 (diagonal position); 0 on one GPU */
  int32_t obs_dim, action_dim, z_dim, goal_dim; /* goal_dim == obs_dim when use_goal == 0 */
  int32_t hidden_dim, feature_dim, backward_hidden_dim;
  int32_t use_goal;            /* goal_space is not None: B reads goal / next_goal instead of obs */
  int32_t rng_device;          /* 1: indices, perm, mix mask, z and action noise are drawn on the device
                                  (Philox) inside FB_PHASE_SAMPLE; 0: the caller provides them */
  int32_t contract_mode;       /* batch x batch contraction + loss of FB_PHASE_FB_LOSS: FB_CONTRACT_TCGEN05 (tensor cores,
                                  3xTF32, z_dim <= 128) or FB_CONTRACT_SIMT (fp32 CUDA cores, materialised matrices) */
  int32_t mlp_mode;            /* wide nn.Linear forward / dX products: FB_MLP_TCGEN05 (tensor cores, 3xTF32, where TMA can
                                  address the operands; the rest and all dW products stay on the fp32 SIMT kernel) or
                                  FB_MLP_SIMT (every product on fp32 CUDA cores) */
  float ortho_coef, mix_ratio;
  float beta1, beta2, adam_eps; /* torch.optim.Adam defaults 0.9 / 0.999 / 1e-8 */
  float future_ratio;          /* hindsight z (fb_ddpg.py:488-491): rows drawn with this probability take
                                  z = backward_net(future_goal or future_obs); 0 disables it (the batch rows then carry no
                                  future fields) */
  uint64_t seed;               /* Philox seed for rng_device */
  int32_t q_loss;              /* cfg.q_loss (fb_ddpg.py:330-341): add q_loss_coef * sum_k mse(F_k.z, implicit reward + discount *
                                  min_k tF_k.z) to fb_loss, implicit reward = (B (B^T B / n)^-1) . z; z_dim <= 118 */
  float q_loss_coef;
  int32_t no_norm_z;           /* cfg.norm_z == False: backward_net outputs are not projected on the sqrt(z_dim)-sphere
                                  (fb_modules.py:227-229), mixed z is not re-projected (fb_ddpg.py:483-484) and the device z draw is
                                  sqrt(z_dim) * U(0,1) (x) direction (fb_ddpg.py:230-231) */
  int32_t rand_weight;         /* cfg.rand_weight (fb_ddpg.py:475-482): the mixed z rows are random weighted sums of ALL rows of
                                  backward_net(backward_input[perm]) (weights U(0,1), rows L2-normalised, scaled by one U(0,1) each)
                                  instead of single rows; z_dim <= 128.  rng_device = 0: the caller uploads the weights
                                  (fb_set_mix_weights) */
  int32_t add_trunk;           /* cfg.add_trunk (fb_modules.py:96-100,169-173): Actor and ForwardMap get a trunk Linear(2 feature_dim ->
                                  hidden_dim) + ReLU between the embeds and the policy / F1 / F2 heads (whose first layer then reads
                                  hidden_dim columns); tensors "trunk.0.weight", "trunk.0.bias" sit between the embeds and the heads */
  int32_t no_preprocess;       /* cfg.preprocess == False (fb_modules.py:102-104,175-177): Actor / ForwardMap are one deep trunk
                                  mlp(obs+z[+action], hidden, "ntanh", hidden, "irelu", hidden, "irelu") in front of the heads instead
                                  of the two embeds; tensors "trunk.{0,1,3,5}.*" then the heads */
  int32_t boltzmann;           /* cfg.boltzmann (fb_modules.py:129-151, fb_ddpg.py:118-120,304-306,391-393,406): the actor is the
                                  DiagGaussianActor mlp(obs+z, hidden, "ntanh", hidden, "relu", 2 action) (tensors "policy.{0,1,3,5}.*"),
                                  actions are tanh of a Normal(mu, std) sample and the actor loss is mean(temp * log pi - Q) */
  float temp;                  /* cfg.temp */
  float log_std_min, log_std_max; /* cfg.log_std_bounds */
  int32_t fused_stacks;        /* the handle will be run WITHOUT FB_RUN_UNFUSED (k_fused_stack segments): its GEMM launches are then planned
                                  for single CTAs.  0 (default): wide GEMM groups run on CTA pairs (tcgen05 cta_group::2) and fused
                                  execution of such a plan is refused with FB_E_STATE */
  int32_t debug_identity_b;    /* cfg.debug (fb_ddpg.py:128-130, fb_modules.py:202-208): backward_net and its target are nn.Identity: B(goal) =
                                  goal (needs z_dim == goal_dim), no "B.*" tensors, no projection of B's output, no backward_net gradients */
} fb_config;

/* per-step scalars (host values; copied to the device by fb_set_step_scalars) */
typedef struct fb_step_scalars {
  float stddev, stddev_clip;   /* utils.schedule(stddev_schedule, step), cfg.stddev_clip */
  float lr_forward, lr_backward, lr_actor; /* lr, lr_coef*lr, lr (fb_ddpg.py:146-151) */
  float tau;                   /* fb_target_tau */
  float replay_discount;       /* ReplayBuffer._discount */
  float replay_future;         /* ReplayBuffer._future (only read by rng_device draws) */
  float grad_scale;            /* multiplies gradients inside Adam (1 on one GPU) */
} fb_step_scalars;

/* caller-owned device memory the handle works on */
typedef struct fb_buffers {
  /* flat fp32 segments; sizes from fb_flat_size().  fb = [forward_net | backward_net] */
  float* d_param_fb; float* d_grad_fb; float* d_m_fb; float* d_v_fb; float* d_target_fb;
  float* d_param_actor; float* d_grad_actor; float* d_m_actor; float* d_v_actor;
  void* d_workspace; size_t workspace_bytes;  /* >= fb_workspace_bytes(), zero-initialised by fb_bind */
} fb_buffers;

/* replay storage resident in HBM: one packed fp32 row per (episode, time) of
 * [observation | action | reward | discount | goal], row stride `row_stride` floats (multiple of 4) */
typedef struct fb_replay_view {
  const float* d_rows;          /* [max_episodes, rows_per_episode, row_stride] */
  const int32_t* d_episode_len; /* [max_episodes] transitions per episode (rows-1) */
  int32_t max_episodes, rows_per_episode, row_stride;
  int32_t n_episodes;           /* len(buffer): episodes currently sampleable */
  /* column offsets (floats, each a multiple of 4; off_discount == off_reward + 1); off_goal / off_extra < 0 if absent.
   * `extra` = the non-TimeStep (meta) keys concatenated, in_memory_replay_buffer.py:162 */
  int32_t off_obs, off_action, off_reward, off_discount, off_goal, off_extra;
  int32_t goal_dim, extra_dim;
} fb_replay_view;

typedef struct fb_handle fb_handle;

/* ---- lifetime ------------------------------------------------------------------------------ */
int fb_abi_version(void);
const char* fb_error_string(int code);
int fb_create(const fb_config* cfg, fb_handle** out);
void fb_destroy(fb_handle* h);

/* ---- layout queries (valid right after fb_create) ------------------------------------------ */
/* total floats of the fb flat segment (forward_net then backward_net) / of the actor segment */
size_t fb_flat_size(const fb_handle* h, int actor /*0: fb, 1: actor*/);
/* number of parameter tensors of a net, in nn.Module registration order */
int fb_num_tensors(const fb_handle* h, int net);
/* tensor `index` of `net`: offset (floats) inside its flat segment (fb for FORWARD/BACKWARD, actor
 * for ACTOR), shape rows x cols (cols == 0 for 1-D tensors); name is written NUL-terminated */
int fb_tensor_info(const fb_handle* h, int net, int index, size_t* offset, int* rows, int* cols,
                   char* name, size_t name_cap);
size_t fb_workspace_bytes(const fb_handle* h);

/* ---- binding --------------------------------------------------------------------------------- */
/* record the caller's buffers, zero the workspace, build the launch plan (synchronises once) */
int fb_bind(fb_handle* h, const fb_buffers* bufs, void* stream);
/* (re)bind the HBM replay storage FB_PHASE_SAMPLE gathers from; copies n_episodes to the device.  Cheap when only
 * n_episodes changed (online training, pretrain.py:649). */
int fb_bind_replay(fb_handle* h, const fb_replay_view* view, void* stream);

/* ---- per-step inputs --------------------------------------------------------------------------- */
int fb_set_step_scalars(fb_handle* h, const fb_step_scalars* s, void* stream);
/* rng_device == 0: int32 draws made by the caller exactly as the reference makes them
 * (in_memory_replay_buffer.py:147-161, fb_ddpg.py:467,471): ep_idx, step_idx, future_idx (may be
 * NULL), perm, mix_mask (0/1).  Device pointers, length batch each. */
int fb_set_indices(fb_handle* h, const int32_t* d_ep_idx, const int32_t* d_step_idx, const int32_t* d_future_idx,
                   const int32_t* d_perm, const int32_t* d_mix_mask, void* stream);
/* explicit transitions instead of FB_PHASE_SAMPLE's gather (row-major [batch, dim], tight; what
 * EpisodeBatch.to(device) yields, replay_buffer.py:50-63).  d_goal / d_next_goal may be NULL when
 * use_goal == 0 (obs / next_obs are used).  d_goal is NOT permuted: the library applies `perm` from
 * fb_set_indices when it builds backward_input (fb_ddpg.py:460-468).  d_discount is [batch], already
 * multiplied by the replay discount. */
int fb_set_batch(fb_handle* h, const float* d_obs, const float* d_action, const float* d_discount,
                 const float* d_next_obs, const float* d_goal, const float* d_next_goal, void* stream);
/* Data-parallel exchange inside the step (the reference has no distributed path).  One process per GPU: rank 0
 * calls fb_nccl_unique_id (128 bytes) and ships the id to the other ranks; every rank then calls fb_nccl_init BEFORE fb_bind
 * (cfg.global_batch = world * cfg.batch, cfg.row_offset = rank * cfg.batch).  The plan then contains the all-gather of the
 * [F|B|targets|discount] row blocks at the end of FB_PHASE_FB_FWD and the all-reduce(sum) of the flat gradients at the end of
 * FB_PHASE_FB_BWD / FB_PHASE_ACTOR_BWD, captured into the step's CUDA graph.  libnccl_path: the NCCL shared object to dlopen
 * (NULL: "libnccl.so.2" as already loaded by the process). */
int fb_nccl_unique_id(const char* libnccl_path, void* id128);
int fb_nccl_init(fb_handle* h, const char* libnccl_path, const void* id128, int world, int rank);
/* The same exchange by the library's OWN kernels over NVLink peer memory (csrc/p2p.cuh), no NCCL on the step's path.  Every rank
 * calls fb_p2p_create BEFORE fb_bind: it allocates this rank's arena (flags | gradients | parameters | global exchange block; the
 * one allocation the library makes, because the peers must be able to open it), writes its CUDA IPC handle (64 bytes) to
 * ipc_handle64 (may be NULL) and fills bufs->d_grad_fb / d_grad_actor / d_param_fb / d_param_actor with the arena's segments — the
 * caller wraps those instead of allocating them and passes the SAME fb_buffers (moments, targets and workspace added) to fb_bind.
 * fb_p2p_attach then maps the peers: ipc_handles = world x 64 bytes in rank order (every rank's handle, shipped by the host, e.g.
 * torch.distributed.all_gather_object), or local_arenas = world pointers obtained from fb_p2p_arena when the engines share a
 * process (tests: several ranks on one device).  The plan then contains, per step: rows of the [F|B|targets|discount] block stored
 * into every arena (k_p2p_scatter_rows) between FB_FWD and FB_LOSS, and k_p2p_adam in FB_PHASE_FB_ADAM / ACTOR_ADAM: rank r sums
 * slice r of the flat gradient over the arenas, applies Adam to slice r (the Adam moments of slice r are live on rank r only:
 * fb_p2p_slice) and stores the new parameter slice into every arena; six epoch-flag barriers per step order the ranks. */
int fb_p2p_create(fb_handle* h, int world, int rank, void* ipc_handle64, fb_buffers* bufs);
int fb_p2p_attach(fb_handle* h, const void* ipc_handles, void* const* local_arenas);
void* fb_p2p_arena(fb_handle* h);
/* 0, or 0xDEAD0000 + (barrier id << 8) + peer rank of the first wait that gave up (~4 s without the peer's signal: a rank died
 * or the ranks issued different phase sequences); the results of that step are invalid.  epochs8 (nullable): this rank's epoch
 * counters of the 8 barriers.  Synchronises `stream`. */
int fb_p2p_status(fb_handle* h, uint32_t* code, uint64_t* epochs8, void* stream);
/* the float range [first, first + count) of the flat fb (actor = 0) / actor (actor = 1) segment whose Adam moments this rank owns */
int fb_p2p_slice(const fb_handle* h, int actor, size_t* first, size_t* count);
/* Host-buffer form of fb_set_batch = EpisodeBatch.to(device) (replay_buffer.py:50-63) as ONE copy: h_rows is [batch, pitch]
 * floats in HOST memory (pinned for an asynchronous copy) laid out by fb_batch_row_layout(obs, action, goal_dim or 0, 0,
 * future_ratio > 0): obs | action | reward, discount (already times the replay discount) | next_obs | goal | next_goal
 * [| future_obs | future_goal].  The copy is enqueued
 * on `stream`; the caller keeps h_rows alive until it has completed. */
int fb_upload_batch(fb_handle* h, const float* h_rows, int pitch, void* stream);
/* A replay buffer that lives in HOST memory in the reference's own layout (in_memory_replay_buffer.py:66-88,126: `_storage`
 * name -> C-contiguous fp32 [max_episodes, rows_per_episode, dim]); `goal` may be NULL. */
typedef struct fb_host_storage {
  const float* observation; const float* action; const float* reward; const float* discount; const float* goal;
  int32_t rows_per_episode, obs_dim, action_dim, goal_dim;
  int32_t max_episodes;   /* first dimension of every array: episode indices are checked against it */
} fb_host_storage;
/* ReplayBuffer.sample's gathers (in_memory_replay_buffer.py:162-183) for host storage, straight into the packed rows fb_upload_batch
 * takes (h_rows [batch, pitch], fb_batch_row_layout(obs, action, goal_dim or 0, 0, h_future_idx != NULL) order): row i reads
 * (episode h_ep_idx[i], steps h_step_idx[i] - 1 and h_step_idx[i], future step h_future_idx[i] - 1), discount times
 * replay_discount.  Runs on the calling thread's cores (software-prefetched row copies; no device work). */
int fb_host_gather_rows(const fb_host_storage* st, const int32_t* h_ep_idx, const int32_t* h_step_idx, const int32_t* h_future_idx,
                        int batch, float replay_discount, float* h_rows, int pitch);
/* rng_device == 0, future_ratio > 0: the hindsight row mask of fb_ddpg.py:490 ([batch] int32, non-zero = hindsight z) */
int fb_set_future_mask(fb_handle* h, const int32_t* d_future_mask, void* stream);
/* rng_device == 0, rand_weight = 1: the U(0,1) draws of fb_ddpg.py:477,479 as a [batch, batch] block (row s = the weight row of
 * batch row s; rows outside the mix mask are ignored) and [batch] row scales */
int fb_set_mix_weights(fb_handle* h, const float* d_weight, const float* d_row_scale, void* stream);
/* rng_device == 0: the random z of fb_ddpg.py:451 ([batch, z_dim], rows of norm sqrt(z_dim)) */
int fb_set_z(fb_handle* h, const float* d_z, void* stream);
/* rng_device == 0: the two N(0,1) draws of utils.py:178 ([batch, action_dim] each): update_fb's
 * next-action noise and update_actor's action noise */
int fb_set_noise(fb_handle* h, const float* d_noise_fb, const float* d_noise_actor, void* stream);

/* ---- the step -------------------------------------------------------------------------------- */
/* enqueue the phases in `phase_mask` in step order.  use_graph != 0 replays a CUDA graph captured
 * (once per mask) from the same launch sequence; use_graph == 2 only captures and instantiates that graph (nothing runs). */
int fb_run(fb_handle* h, uint32_t phase_mask, int use_graph, void* stream);
/* number of kernel launches (graph nodes) fb_run(mask) issues */
int fb_launch_count(fb_handle* h, uint32_t phase_mask);
/* kinds of launch reported by fb_profile_ops */
enum { FB_OPK_GEMM = 0, FB_OPK_LAYERNORM, FB_OPK_ELEMENTWISE, FB_OPK_COLSUM, FB_OPK_ADAM, FB_OPK_GATHER, FB_OPK_LOSS,
       FB_OPK_MEMSET, FB_OPK_CONTRACT, FB_OPK_GEMM_TC, FB_OPK_TRANSPOSE, FB_OPK_COLLECTIVE };
/* run the launches of `phase_mask` eagerly `reps` times with a CUDA event between consecutive launches (on `stream`) and
 * report, per launch: mean duration (ms), kind (FB_OPK_*), algorithmic FLOPs and algorithmic bytes.  Returns the number
 * of launches (<= cap) or a negative error.  Synchronises.  Executes the step for real (parameters move). */
int fb_profile_ops(fb_handle* h, uint32_t phase_mask, int reps, void* stream, float* ms_out, int32_t* kind_out,
                   double* flops_out, double* bytes_out, int cap);
/* the same for fused execution (FB_RUN_UNFUSED clear): runs the units of `phase_mask` `reps` times and reports one row per STAGE
 * of every fused segment (device-side %globaltimer stamps of CTA 0 between the grid barriers) and one row per stand-alone kernel
 * (CUDA events): mean duration (us) and info[5] = {unit index, stage (-1: stand-alone kernel), items in the stage, body type of its
 * first item (0 none, 1 grouped tcgen05 GEMM, 2 LayerNorm fwd, 3 LayerNorm bwd, 4 staging, 5 column sums, 6 / 7 L2 projection
 * fwd / bwd, 8 input staging, 9 z mixing, 10 action sampling, 11 actor Q; stand-alone: -1 - FB_OPK_*, staging batch: -2), its
 * work items}.  Returns the number of rows (<= cap) or a negative error.  Synchronises; executes the step for real. */
int fb_fused_profile(fb_handle* h, uint32_t phase_mask, int reps, void* stream, float* us_out, int32_t* info_out, int cap);
/* device pointer to the metrics block, float[FB_METRIC_COUNT], indices FB_M_* */
const float* fb_metrics_ptr(const fb_handle* h);
/* 1-based Adam step counters live on the device; these set them (checkpoint restore) */
int fb_set_adam_steps(fb_handle* h, int64_t fb_step, int64_t actor_step, void* stream);
int fb_get_adam_steps(fb_handle* h, int64_t* fb_step, int64_t* actor_step, void* stream); /* synchronises */

/* ---- multi-GPU hooks (exact global-batch loss; see DESIGN.md "Multi-GPU") --------------------- */
/* packed per-row block [F1|F2|tF1|tF2|B|tB|discount] each rank contributes to the all-gather that
 * sits between FB_PHASE_FB_FWD and FB_PHASE_FB_LOSS; floats per row / local + gathered pointers */
int fb_gather_block(fb_handle* h, int* floats_per_row, float** d_local, float** d_global);

/* ---- introspection for tests: named views into the workspace --------------------------------- */
/* names: "obs","next_obs","action","discount","z","next_action","action_new","F1","F2","tF1","tF2",
 * "B","tB","dF1","dF2","dB","mu","mix_input","next_goal", ... returns rows, cols, leading dim */
int fb_workspace_view(fb_handle* h, const char* name, float** d_ptr, int* rows, int* cols, int* ld);

/* ---- stand-alone operators (ReplayBuffer.sample / add, tests, microbenchmarks) ------------------- */
/* layout of a packed batch row: float offsets of [obs, action, (reward,discount,0,0), next_obs, goal, next_goal, extra,
 * future_obs, future_goal] (every field padded to a multiple of 4 floats) and the row pitch */
int fb_batch_row_layout(int obs_dim, int action_dim, int goal_dim, int extra_dim, int with_future, int32_t* offsets9,
                        int32_t* pitch);
/* the replay gather alone (in_memory_replay_buffer.py:162-183): out rows [batch, out_ld] in fb_batch_row_layout order
 * (goal / extra parts only if the view has them, future parts only if d_future_idx != NULL), discount multiplied by
 * replay_discount.  obs = row step-1, action/reward/discount/next_obs = row step, future_* = row future-1. */
int fb_replay_gather(const fb_replay_view* view, int obs_dim, int action_dim, const int32_t* d_ep_idx,
                     const int32_t* d_step_idx, const int32_t* d_future_idx, int batch, float replay_discount, float* d_out,
                     int out_ld, void* stream);
/* write one finished episode (tight [rows, dim] fp32 field arrays already on the device) into the packed storage rows of
 * episode `slot` (in_memory_replay_buffer.py:114-133) */
int fb_replay_pack_episode(const fb_replay_view* view, float* d_rows_mut, int slot, int rows, int obs_dim, int action_dim,
                           const float* d_obs, const float* d_action, const float* d_reward, const float* d_discount,
                           const float* d_goal, const float* d_extra, void* stream);
/* C[M,N] (+)= op(A)·op(B)^T with the grouped SIMT SGEMM used by the plan (single problem).  tile_cfg: -1 automatic,
 * 0 = 128x128, 1 = 64x64, 2 = 128x64 CTA tile, 3 = the tcgen05 3xTF32 kernel (a_kmajor = 1, 16-byte aligned operands;
 * synchronises).  splitk > 1 accumulates into C with atomics (C must be zeroed). */
int fb_sgemm(const float* dA, const float* dB, float* dC, const float* d_bias, int M, int N, int K, int lda, int ldb,
             int ldc, int a_kmajor, int b_kmajor, int relu, int splitk, int tile_cfg, void* stream);
/* Timing harness of the tcgen05 3xTF32 grouped GEMM (the kernel behind every wide nn.Linear product, fb_modules.py:76 and its
 * autograd): one synthetic [M,K]x[N,K]^T problem replicated nprob times in one launch, tile width bn (32/64/128), `reps` launches
 * between two CUDA events -> *ms_per_launch; splitk > 1 cuts every tile's K into that many work items.  dbg: 0 = the product kernel with both operands split in shared memory; 1<<19 / 1<<20 =
 * B / A operand with a pre-split lo plane (what the plan does for staged operands); 1<<21 = pseudo-random operands instead of a constant fill; 1<<16 no lo-split, 1<<17 one MMA chain,
 * 1<<18 no epilogue (profiling knobs, results invalid).  Synchronises; allocates its operands from the stream's pool. */
int fb_gemm_tc_bench(int M, int N, int K, int bn, int nprob, int splitk, int dbg, int reps, float* ms_per_launch, void* stream);
/* FMA-chain microbenchmark: returns measured fp32 TFLOP/s of the CUDA cores (synchronises) */
int fb_fp32_peak_tflops(double* out_tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FB_B200_H */
