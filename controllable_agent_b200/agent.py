"""FBDDPGAgent — the reference's agent API (url_benchmark/agent/fb_ddpg.py:37-520) over the B200 step engine.

Same Hydra config surface (`agent=fb_ddpg`, every field of FBDDPGAgentConfig with the same default), same
constructor (`FBDDPGAgent(**kwargs)`), same public methods and attributes the workspaces touch (SURVEY.md
section 8b).  `update()` is one CUDA-graph launch of libfb_b200's step; the nn.Modules hanging off the agent
(`actor`, `forward_net`, ...) are parameter *views* on the flat segments that step trains in place and are only
evaluated by the per-environment-step helpers (`act`, `get_goal_meta`, `infer_meta*`, `compute_z_correl`).
"""
from __future__ import annotations

import copy
import dataclasses
import logging
import math
import os
import typing as tp
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.distributions.utils import _standard_normal

from . import _lib as L
from . import modules as M
from .dist_utils import reduce_metrics, shard_layout
from .engine import EngineConfig, FBStepEngine
from .replay import HostStorageView, ReplayBuffer, draw_sample_indices

logger = logging.getLogger(__name__)
MetaDict = tp.Mapping[str, np.ndarray]

try:  # the Hydra surface, when hydra / omegaconf are installed (they are not in the build image)
    import omegaconf
    MISSING: tp.Any = omegaconf.MISSING
    II = omegaconf.II
except ImportError:  # pragma: no cover
    MISSING = "???"

    def II(key: str) -> tp.Any:  # noqa: N802
        return "${" + key + "}"

# goal-space widths (url_benchmark/goals.py:44-112 evaluated once; get_goal_space_dim builds a MuJoCo env, goals.py:218-221)
GOAL_SPACE_DIMS = {"simplified_walker": 3, "walker_pos_speed": 4, "walker_pos_speed_z": 6, "simplified_quadruped": 2,
                   "quad_pos_speed": 7, "simplified_jaco": 3, "simplified_point_mass_maze": 2}


def get_goal_space_dim(name: str) -> int:
    if name in GOAL_SPACE_DIMS:
        return GOAL_SPACE_DIMS[name]
    from url_benchmark import goals as _goals  # needs dm_control
    return int(_goals.get_goal_space_dim(name))


@dataclasses.dataclass
class FBDDPGAgentConfig:
    # @package agent  — field for field the reference's dataclass (fb_ddpg.py:37-82)
    _target_: str = "controllable_agent_b200.agent.FBDDPGAgent"
    name: str = "fb_ddpg"
    obs_type: str = MISSING
    obs_shape: tp.Tuple[int, ...] = MISSING
    action_shape: tp.Tuple[int, ...] = MISSING
    device: str = II("device")
    lr: float = 1e-4
    lr_coef: float = 1
    fb_target_tau: float = 0.01
    update_every_steps: int = 2
    use_tb: bool = II("use_tb")
    use_wandb: bool = II("use_wandb")
    use_hiplog: bool = II("use_hiplog")
    num_expl_steps: int = MISSING
    num_inference_steps: int = 5120
    hidden_dim: int = 1024
    backward_hidden_dim: int = 526
    feature_dim: int = 512
    z_dim: int = 50
    stddev_schedule: str = "0.2"
    stddev_clip: float = 0.3
    update_z_every_step: int = 300
    update_z_proba: float = 1.0
    nstep: int = 1
    batch_size: int = 1024
    init_fb: bool = True
    update_encoder: bool = II("update_encoder")
    goal_space: tp.Optional[str] = II("goal_space")
    ortho_coef: float = 1.0
    log_std_bounds: tp.Tuple[float, float] = (-5, 2)
    temp: float = 1
    boltzmann: bool = False
    debug: bool = False
    future_ratio: float = 0.0
    mix_ratio: float = 0.5
    rand_weight: bool = False
    preprocess: bool = True
    norm_z: bool = True
    q_loss: bool = False
    q_loss_coef: float = 0.01
    additional_metric: bool = False
    add_trunk: bool = False
    # --- additions of this implementation (defaults keep the reference's semantics) ---
    rng_mode: str = "device"     # "device": Philox draws inside the step graph; "reference": numpy/torch draws in the
    #                              reference's order (Appendix B of SURVEY.md), uploaded per step
    use_cuda_graph: bool = True
    mlp_mode: str = "tcgen05"   # wide Linear products: "tcgen05" (3xTF32 tensor cores) or "simt" (fp32 CUDA cores)
    contract_mode: str = "tcgen05"  # batch x batch contraction: "tcgen05" or "simt"
    collectives: str = "p2p"     # multi-GPU exchange: "p2p" = the library's own kernels over NVLink peer memory (row scatter + fused
    #                              reduce-scatter / Adam / all-gather, csrc/p2p.cuh); "graph" = the library's own NCCL communicator,
    #                              all-gather / all-reduce captured inside the step graph; "torch" = torch.distributed calls between
    #                              graph segments
    fuse_stacks: bool = False   # True: the MLP stacks as fused persistent kernels (one launch per forward / backward segment of the
    #                             plan, k_fused_stack: 80 -> 13 launches per step); False (default, measured faster: the chain is bound by
    #                             the latency inside each GEMM, not by launches): one kernel per layer-level launch on three lanes
    prefetch_host_batch: bool = False   # host replay only: sample + upload the NEXT update's batch while this update's step runs
    #                                     on the GPU (same numpy draw order as the reference as long as nothing else samples the
    #                                     replay between updates; for a static replay, e.g. train_offline)


def register_hydra() -> None:
    """cs.store(group="agent", name="fb_ddpg", node=FBDDPGAgentConfig) as fb_ddpg.py:85-86 does (needs hydra)."""
    from hydra.core.config_store import ConfigStore
    ConfigStore.instance().store(group="agent", name="fb_ddpg", node=FBDDPGAgentConfig)


_UNSUPPORTED: tp.Dict[str, tp.Any] = {}   # every non-default branch of fb_ddpg.py runs on the device (pixels: see __init__)


class _EngineAdam(torch.optim.Adam):
    """A real torch.optim.Adam (init_from / checkpoint code looks for Optimizer instances, fb_ddpg.py:173-175) whose moments are views
    of the flat segments the CUDA step updates.  The Adam step COUNT lives on the device; `state_dict()` refreshes the per-parameter
    `step` entries from it first, so whoever saves or copies the optimizer sees consistent bias corrections."""
    _refresh: tp.Optional[tp.Callable[[], None]] = None

    def state_dict(self) -> tp.Dict[str, tp.Any]:   # type: ignore[override]
        if self._refresh is not None:
            self._refresh()
        return super().state_dict()


class FBDDPGAgent:

    def __init__(self, **kwargs: tp.Any) -> None:
        cfg = FBDDPGAgentConfig(**kwargs)
        self.cfg = cfg
        assert len(cfg.action_shape) == 1
        self.action_dim = cfg.action_shape[0]
        self.solved_meta: tp.Any = None
        if cfg.obs_type == "pixels":
            raise NotImplementedError("controllable_agent_b200 implements the states-only FB-DDPG path (obs_type=pixels "
                                      "needs the DDPG conv encoder, out of scope)")
        for key, ok in _UNSUPPORTED.items():
            if getattr(cfg, key) != ok:
                raise NotImplementedError(f"agent.{key}={getattr(cfg, key)} is a non-default branch of fb_ddpg.py the CUDA step "
                                          f"does not implement (supported: {key}={ok})")
        if not 0.0 <= cfg.future_ratio <= 1.0:
            raise ValueError(f"agent.future_ratio must be in [0, 1] (got {cfg.future_ratio})")
        device = torch.device(cfg.device)
        if device.type != "cuda":
            raise RuntimeError(f"controllable_agent_b200.FBDDPGAgent needs device=cuda (got {cfg.device!r}); there is no "
                               "CPU fallback — use the reference agent on CPU")
        self.aug: nn.Module = nn.Identity()
        self.encoder: nn.Module = nn.Identity()
        self.obs_dim = cfg.obs_shape[0]
        if cfg.feature_dim < self.obs_dim:
            logger.warning(f"feature_dim {cfg.feature_dim} should not be smaller that obs_dim {self.obs_dim}")
        goal_dim = self.obs_dim
        if cfg.goal_space is not None:
            goal_dim = get_goal_space_dim(cfg.goal_space)
        if cfg.debug and cfg.z_dim != goal_dim:
            raise ValueError(f"agent.debug=True makes backward_net an identity map: z_dim ({cfg.z_dim}) must equal the goal dimension ({goal_dim})")
        if cfg.z_dim < goal_dim:
            logger.warning(f"z_dim {cfg.z_dim} should not be smaller that goal_dim {goal_dim}")
        self.goal_dim = goal_dim

        # data-parallel layout: cfg.batch_size is the GLOBAL batch; each rank steps batch_size / world rows
        self.world, self.rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world, self.rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
        local, row_offset = shard_layout(cfg.batch_size, self.world, self.rank)

        seed = int(torch.initial_seed() % (2 ** 63)) + 7919 * self.rank
        nccl, p2p = None, None
        if cfg.collectives not in ("p2p", "graph", "torch"):
            raise ValueError(f"agent.collectives must be p2p, graph or torch (got {cfg.collectives!r})")
        if self.world > 1 and cfg.collectives == "p2p":
            p2p = (self.world, self.rank)   # the library's own exchange kernels over NVLink peer memory (csrc/p2p.cuh)
        elif self.world > 1 and cfg.collectives == "graph":
            # the library's own NCCL communicator: rank 0 creates the 128-byte id, torch.distributed (plumbing) ships it
            import ctypes as C
            uid = C.create_string_buffer(128)
            ok = 1
            if self.rank == 0:
                ok = int(L.load().fb_nccl_unique_id(L.nccl_library_path(), uid) == L.FB_OK)
            box = [uid.raw, ok]
            torch.distributed.broadcast_object_list(box, src=0, device=device)
            if not box[1]:   # never a silent change of path: the caller asked for in-graph NCCL
                raise RuntimeError("libfb_b200 could not load NCCL for collectives='graph'; use collectives='p2p' (no NCCL needed) or 'torch'")
            nccl = (box[0], self.world, self.rank)
        self.collectives_mode = cfg.collectives if self.world > 1 else "none"
        self.engine = FBStepEngine(EngineConfig(nccl=nccl, p2p=p2p, fused=bool(cfg.fuse_stacks),
            batch=local, obs_dim=self.obs_dim, action_dim=self.action_dim, z_dim=cfg.z_dim, goal_dim=goal_dim,
            hidden_dim=cfg.hidden_dim, feature_dim=cfg.feature_dim, backward_hidden_dim=cfg.backward_hidden_dim,
            use_goal=cfg.goal_space is not None, rng_device=cfg.rng_mode == "device", ortho_coef=cfg.ortho_coef,
            mix_ratio=cfg.mix_ratio, future_ratio=cfg.future_ratio, q_loss=bool(cfg.q_loss), q_loss_coef=float(cfg.q_loss_coef), norm_z=bool(cfg.norm_z), rand_weight=bool(cfg.rand_weight), add_trunk=bool(cfg.add_trunk), preprocess=bool(cfg.preprocess),
            boltzmann=bool(cfg.boltzmann), temp=float(cfg.temp), log_std_bounds=(float(cfg.log_std_bounds[0]), float(cfg.log_std_bounds[1])),
            debug=bool(cfg.debug), seed=seed, global_batch=cfg.batch_size, row_offset=row_offset,
            mlp_mode=L.MLP_SIMT if cfg.mlp_mode == "simt" else L.MLP_TCGEN05,
            contract_mode=L.CONTRACT_SIMT if cfg.contract_mode == "simt" else L.CONTRACT_TCGEN05), device)

        # networks: constructed on the CPU in the reference's order (Actor, ForwardMap, BackwardMap, BackwardMap target,
        # ForwardMap target — fb_ddpg.py:117-139) so that the torch CPU generator is consumed identically, then moved
        # onto the flat device segments
        e = self.engine
        if cfg.boltzmann:   # fb_ddpg.py:118-120
            self.actor: nn.Module = M.DiagGaussianActor(self.obs_dim, cfg.z_dim, self.action_dim, cfg.hidden_dim, cfg.log_std_bounds)
        else:
            self.actor = M.Actor(self.obs_dim, cfg.z_dim, self.action_dim, cfg.feature_dim, cfg.hidden_dim, add_trunk=cfg.add_trunk, preprocess=cfg.preprocess)
        self.forward_net = M.ForwardMap(self.obs_dim, cfg.z_dim, self.action_dim, cfg.feature_dim, cfg.hidden_dim, add_trunk=cfg.add_trunk, preprocess=cfg.preprocess)
        if cfg.debug:   # fb_ddpg.py:128-130
            self.backward_net: nn.Module = M.IdentityMap()
            self.backward_target_net: nn.Module = M.IdentityMap()
        else:
            self.backward_net = M.BackwardMap(goal_dim, cfg.z_dim, cfg.backward_hidden_dim, norm_z=cfg.norm_z)
            self.backward_target_net = M.BackwardMap(goal_dim, cfg.z_dim, cfg.backward_hidden_dim, norm_z=cfg.norm_z)
        self.forward_target_net = M.ForwardMap(self.obs_dim, cfg.z_dim, self.action_dim, cfg.feature_dim, cfg.hidden_dim, add_trunk=cfg.add_trunk, preprocess=cfg.preprocess)
        M.adopt_flat(self.actor, e.tensors(L.NET_ACTOR, "param"))
        M.adopt_flat(self.forward_net, e.tensors(L.NET_FORWARD, "param"))
        M.adopt_flat(self.backward_net, e.tensors(L.NET_BACKWARD, "param"))
        M.adopt_flat(self.backward_target_net, e.tensors(L.NET_BACKWARD, "target"))
        M.adopt_flat(self.forward_target_net, e.tensors(L.NET_FORWARD, "target"))
        e.target_fb.copy_(e.param_fb)   # load_state_dict of the online nets into the targets (fb_ddpg.py:141-142)
        if self.world > 1:  # replicas start from rank 0's parameters
            for flat in (e.param_fb, e.target_fb, e.param_actor):
                torch.distributed.broadcast(flat, src=0)

        # real torch optimizers (init_from / checkpoints look for them in __dict__, fb_ddpg.py:173-175); their state
        # tensors are views of the flat Adam moments the CUDA step updates
        self.encoder_opt: tp.Optional[torch.optim.Adam] = None
        self.actor_opt = _EngineAdam(self.actor.parameters(), lr=cfg.lr)
        self.fb_opt = _EngineAdam([{"params": self.forward_net.parameters()},
                                   {"params": self.backward_net.parameters(), "lr": cfg.lr_coef * cfg.lr}], lr=cfg.lr)
        self._link_optimizer_state()
        self.actor_opt._refresh = self.fb_opt._refresh = self._sync_optimizer_steps
        self.train()
        self.forward_target_net.train()
        self.backward_target_net.train()
        self.actor_success: tp.List[float] = []
        self._replay_key: tp.Any = None
        self._prefetched: tp.Any = None   # id of the host replay whose next batch is already in the packed block
        self.native_host_sampling = True   # reference-layout host replays are sampled by the library (False: always their own sample())
        self._host_views: tp.Dict[int, HostStorageView] = {}
        self.last_update_launches = 0
        # where rng_mode="reference" makes its torch draws (z, action noise): the agent's device, like the reference run
        # with device=cuda; tests point it at "cpu" to replay the CPU-generated golden trajectories
        self.draw_device: tp.Union[str, torch.device] = device

    # ------------------------------------------------------------------------------------------------
    def _link_optimizer_state(self) -> None:
        e = self.engine
        fb_step, actor_step = e.get_adam_steps()
        for opt, nets, step in ((self.actor_opt, [(self.actor, L.NET_ACTOR)], actor_step),
                                (self.fb_opt, [(self.forward_net, L.NET_FORWARD), (self.backward_net, L.NET_BACKWARD)], fb_step)):
            for net, net_id in nets:
                m, v = e.tensors(net_id, "m"), e.tensors(net_id, "v")
                for (name, p) in net.named_parameters():
                    opt.state[p] = {"step": torch.tensor(float(step)), "exp_avg": m[name], "exp_avg_sq": v[name]}

    def _sync_optimizer_steps(self) -> None:
        fb_step, actor_step = self.engine.get_adam_steps()
        for opt, step in ((self.actor_opt, actor_step), (self.fb_opt, fb_step)):
            for st in opt.state.values():
                st["step"] = torch.tensor(float(step))

    def train(self, training: bool = True) -> None:
        self.training = training
        for net in [self.encoder, self.actor, self.forward_net, self.backward_net]:
            net.train(training)

    def init_from(self, other: tp.Any) -> None:
        """fb_ddpg.py:166-175: copy parameters net by net, then the optimizers' state."""
        names = ["encoder", "actor"]
        if self.cfg.init_fb:
            names += ["forward_net", "backward_net", "backward_target_net", "forward_target_net"]
        for name in names:
            M.hard_update_params(getattr(other, name), getattr(self, name))
        if isinstance(other, FBDDPGAgent):
            other._sync_optimizer_steps()
        steps = {}
        for key in ("actor_opt", "fb_opt"):
            src = getattr(other, key).state_dict()
            mine = getattr(self, key)
            step = 0.0
            for group_s, group_m in zip(src["param_groups"], mine.param_groups):
                group_m["lr"] = group_s["lr"]
                for idx_s, p in zip(group_s["params"], group_m["params"]):
                    st = src["state"].get(idx_s)
                    if st is None:
                        continue
                    mine.state[p]["exp_avg"].copy_(st["exp_avg"])
                    mine.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                    step = float(st["step"])
            steps[key] = int(step)
        self.engine.set_adam_steps(steps["fb_opt"], steps["actor_opt"])
        self._sync_optimizer_steps()

    # -- pickling: pretrain.py:437-449 saves the whole agent object ---------------------------------
    def __getstate__(self) -> tp.Dict[str, tp.Any]:
        e = self.engine
        fb_step, actor_step = e.get_adam_steps()
        flats = {k: getattr(e, k).detach().cpu() for k in ("param_fb", "m_fb", "v_fb", "target_fb", "param_actor", "m_actor", "v_actor")}
        if e.has_p2p:   # the Adam moments are sharded over the ranks: gather the slices (collective: every rank pickles together)
            flats.update({k: v.cpu() for k, v in e.full_moments().items()})
        return {"cfg": dataclasses.asdict(self.cfg), "flats": flats, "adam_steps": (fb_step, actor_step),
                "solved_meta": self.solved_meta, "training": self.training,
                "lrs": ([g["lr"] for g in self.fb_opt.param_groups], [g["lr"] for g in self.actor_opt.param_groups])}

    def __setstate__(self, state: tp.Dict[str, tp.Any]) -> None:
        cpu_rng = torch.get_rng_state()
        self.__init__(**state["cfg"])  # type: ignore[misc]
        torch.set_rng_state(cpu_rng)
        for k, v in state["flats"].items():
            getattr(self.engine, k).copy_(v)
        self.engine.set_adam_steps(*state["adam_steps"])
        self._sync_optimizer_steps()
        for g, lr in zip(self.fb_opt.param_groups, state["lrs"][0]):
            g["lr"] = lr
        for g, lr in zip(self.actor_opt.param_groups, state["lrs"][1]):
            g["lr"] = lr
        self.solved_meta = state["solved_meta"]
        self.train(state["training"])

    # -- inference helpers (per environment step): forward passes through the library's inference plans ----
    def get_goal_meta(self, goal_array: np.ndarray) -> MetaDict:
        """fb_ddpg.py:177-186."""
        z = self.engine.infer_backward(np.asarray(goal_array, dtype=np.float32).reshape(1, -1))[0]
        if self.cfg.norm_z:
            z = math.sqrt(self.cfg.z_dim) * z / max(float(np.linalg.norm(z)), 1e-12)
        meta = OrderedDict()
        meta["z"] = z.astype(np.float32)
        return meta

    def infer_meta(self, replay_loader: tp.Any) -> MetaDict:
        """fb_ddpg.py:188-199."""
        obs_list, reward_list = [], []
        batch_size = 0
        while batch_size < self.cfg.num_inference_steps:
            batch = replay_loader.sample(self.cfg.batch_size)
            batch = batch.to(self.cfg.device)
            obs_list.append(batch.next_goal if self.cfg.goal_space is not None else batch.next_obs)
            reward_list.append(batch.reward)
            batch_size += batch.next_obs.size(0)
        obs, reward = torch.cat(obs_list, 0), torch.cat(reward_list, 0)
        obs, reward = obs[:self.cfg.num_inference_steps], reward[:self.cfg.num_inference_steps]
        return self.infer_meta_from_obs_and_rewards(obs, reward)

    def infer_meta_from_obs_and_rewards(self, obs: torch.Tensor, reward: torch.Tensor) -> MetaDict:
        """fb_ddpg.py:201-222: z = reward^T . B(obs) / N, then the sqrt(z_dim) projection."""
        z = self.engine.infer_backward_weighted_sum(obs, reward) / float(reward.shape[0])
        if self.cfg.norm_z:
            z = math.sqrt(self.cfg.z_dim) * z / max(float(np.linalg.norm(z)), 1e-12)
        meta = OrderedDict()
        meta["z"] = z.astype(np.float32)
        return meta

    def sample_z(self, size: int, device: tp.Union[str, torch.device] = "cpu") -> torch.Tensor:
        gaussian_rdv = torch.randn((size, self.cfg.z_dim), dtype=torch.float32, device=device)
        gaussian_rdv = F.normalize(gaussian_rdv, dim=1)
        if self.cfg.norm_z:
            return math.sqrt(self.cfg.z_dim) * gaussian_rdv
        uniform_rdv = torch.rand((size, self.cfg.z_dim), dtype=torch.float32, device=device)   # fb_ddpg.py:230-231
        return np.sqrt(self.cfg.z_dim) * uniform_rdv * gaussian_rdv

    def init_meta(self) -> MetaDict:
        if self.solved_meta is not None:
            return self.solved_meta
        meta = OrderedDict()
        meta["z"] = self.sample_z(1).squeeze().numpy()
        return meta

    def update_meta(self, meta: MetaDict, global_step: int, time_step: tp.Any, finetune: bool = False,
                    replay_loader: tp.Optional[tp.Any] = None) -> MetaDict:
        if global_step % self.cfg.update_z_every_step == 0 and np.random.rand() < self.cfg.update_z_proba:
            return self.init_meta()
        return meta

    def act(self, obs: tp.Any, meta: MetaDict, step: int, eval_mode: bool) -> tp.Any:
        """fb_ddpg.py:258-281: one actor forward per environment step (the library's FB_PHASE_INFER_ACTOR graph)."""
        obs_np = np.asarray(obs, dtype=np.float32).reshape(1, -1)
        z_np = np.asarray(meta["z"], dtype=np.float32).reshape(1, -1)
        mu = self.engine.infer_actor(obs_np, z_np)
        if self.cfg.boltzmann:
            # the plan returns [mu | std] of the pre-tanh Normal (fb_modules.py:141-151); dist.mean = tanh(mu), dist.sample() = tanh of a
            # torch.normal draw on the device generator (utils.py:218-233)
            loc, std = mu[:, :self.action_dim], mu[:, self.action_dim:]
            if eval_mode:
                mu = np.tanh(loc)
            else:
                dev = self.cfg.device
                action = torch.tanh(torch.normal(torch.as_tensor(loc, device=dev), torch.as_tensor(std, device=dev)))
                if step < self.cfg.num_expl_steps:
                    action.uniform_(-1.0, 1.0)
                return action.cpu().numpy()[0]
        if eval_mode:
            action = mu
            if self.cfg.additional_metric:   # F(s, z, mu) against F(s, z, random action): diagnostic branch, parameter-view modules
                with torch.no_grad():
                    o, z = (torch.as_tensor(x, device=self.cfg.device) for x in (obs_np, z_np))
                    a = torch.as_tensor(mu, device=self.cfg.device)
                    Qs = [torch.min(*(torch.einsum("sd, sd -> s", Fk, z) for Fk in self.forward_net(o, z, act_)))
                          for act_ in (a, torch.zeros_like(a).uniform_(-1.0, 1.0))]
                    self.actor_success = (Qs[0] > Qs[1]).cpu().numpy().tolist()
        else:
            # TruncatedNormal(mu, std).sample() with clip=None (utils.py:176-185), drawn from the device generator like the reference
            stddev = M.schedule(self.cfg.stddev_schedule, step)
            eps = _standard_normal((1, self.action_dim), dtype=torch.float32, device=self.cfg.device).cpu().numpy() * np.float32(stddev)
            action = np.clip(mu + eps, -1.0 + 1e-6, 1.0 - 1e-6).astype(np.float32)
            if step < self.cfg.num_expl_steps:   # the reference draws the sample first, then overwrites it (same generator order)
                action = torch.empty((1, self.action_dim), dtype=torch.float32, device=self.cfg.device).uniform_(-1.0, 1.0).cpu().numpy()
        return action[0]

    def compute_z_correl(self, time_step: tp.Any, meta: MetaDict) -> float:
        """fb_ddpg.py:283-289 (F.normalize(z, 1) there is the L1 normalisation: p = 1, dim = 1)."""
        goal = time_step.goal if self.cfg.goal_space is not None else time_step.observation
        b = self.engine.infer_backward(np.asarray(goal, dtype=np.float32).reshape(1, -1))[0].astype(np.float64)
        z = np.asarray(meta["z"], dtype=np.float64)
        b, z = b / max(np.abs(b).sum(), 1e-12), z / max(np.abs(z).sum(), 1e-12)
        return float(np.dot(b, z))

    # -- the gradient step ---------------------------------------------------------------------------
    def _metrics_enabled(self) -> bool:
        c = self.cfg
        return bool(c.use_tb) or bool(c.use_wandb) or bool(c.use_hiplog)

    def _set_scalars(self, step: int, replay_discount: float = 1.0, replay_future: float = 1.0) -> None:
        c = self.cfg
        self.engine.set_scalars(M.schedule(c.stddev_schedule, step), c.stddev_clip, self.fb_opt.param_groups[0]["lr"],
                                self.fb_opt.param_groups[1]["lr"], self.actor_opt.param_groups[0]["lr"], c.fb_target_tau,
                                replay_discount, replay_future, 1.0)

    def _run(self, mask: int) -> None:
        """Enqueue the phases of `mask`; with >1 rank, split at the two exchange points (DESIGN.md "Multi-GPU")."""
        e, g = self.engine, bool(self.cfg.use_cuda_graph)
        if self.world == 1 or e.has_nccl or e.has_p2p:   # the exchange kernels / NCCL calls are launches of the plan
            e.run(mask, graph=g)
            self.last_update_launches = e.launch_count(mask)
            return
        import torch.distributed as dist
        seg1 = mask & (L.PHASE_SAMPLE | L.PHASE_MIX | L.PHASE_FB_FWD | L.RUN_HOST_BATCH)
        seg2 = mask & (L.PHASE_FB_LOSS | L.PHASE_FB_BWD)
        seg3 = mask & (L.PHASE_FB_ADAM | L.PHASE_ACTOR_FWD | L.PHASE_ACTOR_BWD)
        seg4 = mask & (L.PHASE_ACTOR_ADAM | L.PHASE_METRICS)
        local, glob = e.gather_block()
        if seg1:
            e.run(seg1, graph=g)
        if seg2:
            dist.all_gather_into_tensor(glob, local)   # [F1|F2|tF1|tF2|B|tB|discount] rows of every rank
            e.run(seg2, graph=g)
            dist.all_reduce(e.grad_fb)
        if seg3:
            e.run(seg3, graph=g)
            dist.all_reduce(e.grad_actor)
        if seg4:
            e.run(seg4, graph=g)
        self.last_update_launches = e.launch_count(mask)

    def update_fb(self, obs: torch.Tensor, action: torch.Tensor, discount: torch.Tensor, next_obs: torch.Tensor,
                  next_goal: torch.Tensor, z: torch.Tensor, step: int) -> tp.Dict[str, float]:
        """fb_ddpg.py:291-387 on explicit tensors: targets, F/B forward, loss, backward, fb_opt.step().  `z` is used as
        given (mixing belongs to update()); like the reference this does NOT move the target networks."""
        e = self.engine
        if self.world > 1:
            raise NotImplementedError("update_fb on explicit tensors is single-GPU; use update()")
        use_goal = self.cfg.goal_space is not None
        e.set_batch(obs, action, discount, next_obs, next_goal if use_goal else None, next_goal if use_goal else None)
        e.set_z(z)
        e.set_indices(mix_mask=np.zeros(e.cfg.batch, np.int32))
        if self.cfg.future_ratio > 0:   # z is used as given: no hindsight rows either (the mask of the last update() must not linger)
            e.set_future_mask(np.zeros(e.cfg.batch, np.int32))
        shape = (e.cfg.batch, self.action_dim)
        # update_fb's and update_actor's N(0,1) draws (utils.py:178), consumed from the device generator in that order
        e.set_noise(_standard_normal(shape, dtype=torch.float32, device=e.device),
                    _standard_normal(shape, dtype=torch.float32, device=e.device))
        tau, self.cfg.fb_target_tau = self.cfg.fb_target_tau, 0.0   # Adam only: soft updates are update()'s job
        try:
            self._set_scalars(step)
        finally:
            self.cfg.fb_target_tau = tau
        mask = L.PHASE_MIX | L.PHASE_FB_FWD | L.PHASE_FB_LOSS | L.PHASE_FB_BWD | L.PHASE_FB_ADAM
        metrics: tp.Dict[str, float] = {}
        if self._metrics_enabled():
            self._run(mask | L.PHASE_METRICS)
            metrics = self._fb_metrics(e.read_metrics())
        else:
            self._run(mask)
        return metrics

    def update_actor(self, obs: torch.Tensor, z: torch.Tensor, step: int) -> tp.Dict[str, float]:
        """fb_ddpg.py:389-421.  Must follow update_fb() on the same (obs, z): the actor forward on `obs` was batched
        with the one on `next_obs` there, and is reused here (the actor has not changed in between)."""
        e = self.engine
        self._run(L.PHASE_ACTOR_FWD | L.PHASE_ACTOR_BWD | L.PHASE_ACTOR_ADAM)
        if self.cfg.use_tb or self.cfg.use_wandb:
            e.run(L.PHASE_METRICS, graph=False)
            return self._actor_metrics(e.read_metrics())
        return {}

    def _fb_metrics(self, m: tp.Mapping[str, float]) -> tp.Dict[str, float]:
        """The keys update_fb reports (fb_ddpg.py:356-377) out of the metrics block."""
        out = {k: m[k] for k in L.METRIC_KEYS[:14]}
        if self.cfg.q_loss:
            out["q_loss"] = m["q_loss"]
        out["fb_opt_lr"] = self.fb_opt.param_groups[0]["lr"]
        return out

    def _actor_metrics(self, m: tp.Mapping[str, float]) -> tp.Dict[str, float]:
        """The keys update_actor reports (fb_ddpg.py:413-418)."""
        out = {k: m[k] for k in L.METRIC_KEYS[14:]}
        if self.cfg.additional_metric:
            out["q1_success"] = m["q1_success"]
        return out

    def update(self, replay_loader: tp.Any, step: int) -> tp.Dict[str, float]:
        """fb_ddpg.py:427-520: sample, z draw + mixing, update_fb, update_actor, target soft updates."""
        metrics: tp.Dict[str, float] = {}
        c, e = self.cfg, self.engine
        if step % c.update_every_steps != 0:
            return metrics
        B = e.cfg.batch
        fused = isinstance(replay_loader, ReplayBuffer)
        if c.future_ratio > 0 and float(getattr(replay_loader, "_future", 0.0)) >= 1.0:
            # the reference asserts `future_goal is not None` (fb_ddpg.py:463): a replay with future = 1 samples no future rows
            raise ValueError("agent.future_ratio > 0 needs a replay buffer created with future < 1 (it samples no future_obs / future_goal)")
        if fused:
            self._set_scalars(step, float(replay_loader._discount), float(replay_loader._future))
            key = (id(replay_loader), replay_loader._version, len(replay_loader))
            if key != self._replay_key:
                e.bind_replay(replay_loader.view())
                self._replay_key = key
        else:
            self._set_scalars(step)
        mask = L.PHASE_ALL & ~L.PHASE_METRICS
        use_goal = c.goal_space is not None
        if not fused:
            # a host replay (the reference's ReplayBuffer): its numpy fields go to the engine's packed batch block with one
            # asynchronous copy from pinned staging (instead of EpisodeBatch.to's one blocking copy per field)
            if self._prefetched != id(replay_loader):
                self._upload_host_batch(replay_loader, B)
            self._prefetched = None
        if c.rng_mode == "device":
            if not fused:
                mask |= L.RUN_HOST_BATCH   # FB_PHASE_SAMPLE draws z / noise / perm / mix mask on the device, no gather
        else:
            # the reference's RNG streams in the reference's order
            if fused:
                ep_idx, step_idx, future_idx = replay_loader.draw_indices(B)
            else:
                mask &= ~L.PHASE_SAMPLE
            z = self.sample_z(B, device=self.draw_device)
            perm = torch.randperm(B)
            mix = (np.random.uniform(size=B) < c.mix_ratio).astype(np.int32) if c.mix_ratio > 0 else np.zeros(B, np.int32)
            if c.rand_weight and c.mix_ratio > 0:
                # fb_ddpg.py:477-480: torch.rand(nmix, B) then torch.rand(nmix, 1), both on the CPU generator whatever the device;
                # scattered into the [B, B] block the library reads (row s = the weights of batch row s)
                mix_idxs = np.where(mix)[0]
                if getattr(self, "_mixw_stage", None) is None:
                    self._mixw_stage = (torch.zeros((B, B), dtype=torch.float32).pin_memory(), torch.zeros(B, dtype=torch.float32).pin_memory())
                w_full, u_full = self._mixw_stage
                torch.cuda.current_stream(e.device).synchronize()   # the previous step's upload has left the pinned block
                w_full[mix_idxs] = torch.rand(size=(mix_idxs.shape[0], B))
                u_full[mix_idxs] = torch.rand(mix_idxs.shape[0], 1)[:, 0]
                e.set_mix_weights(w_full, u_full)
            noise_fb = _standard_normal((B, self.action_dim), dtype=torch.float32, device=self.draw_device)
            noise_actor = _standard_normal((B, self.action_dim), dtype=torch.float32, device=self.draw_device)
            if fused:
                e.set_indices(ep_idx, step_idx, future_idx, perm, mix)
            else:
                e.set_indices(perm=perm, mix_mask=mix)
            if c.future_ratio > 0:   # hindsight rows (fb_ddpg.py:488-491): drawn after the mix mask, like the reference
                e.set_future_mask((np.random.uniform(size=B) < c.future_ratio).astype(np.int32))
            e.set_z(z)
            e.set_noise(noise_fb, noise_actor)
        want_metrics = self._metrics_enabled()
        self._run(mask | L.PHASE_METRICS if want_metrics else mask)
        if not fused and c.prefetch_host_batch:
            # the step is in flight: sample the next batch and enqueue its upload behind it (stream order keeps this step's
            # reads of the packed block ahead of the copy)
            self._upload_host_batch(replay_loader, B)
            self._prefetched = id(replay_loader)
        if want_metrics:
            m = e.read_metrics()
            if self.world > 1:
                m = self._reduce_metrics(m)
            metrics.update(self._fb_metrics(m))
            if c.use_tb or c.use_wandb:   # the actor block logs under a narrower condition (fb_ddpg.py:413)
                metrics.update(self._actor_metrics(m))
        return metrics

    def _upload_host_batch(self, replay_loader: tp.Any, B: int) -> None:
        # a host buffer in the reference's own layout: its index draws here (same numpy stream as its sample()), its row gathers in
        # the library, straight into the pinned block; any other object: its own sample() contract
        view = self._host_view(replay_loader)
        if view is not None:
            ep_idx, step_idx, future_idx = draw_sample_indices(replay_loader, B, exact_stream=self.cfg.rng_mode != "device")
            self.engine.upload_host_rows(view, ep_idx, step_idx, future_idx, float(replay_loader._discount))
            return
        batch = replay_loader.sample(B)
        use_goal = self.cfg.goal_space is not None
        if use_goal:
            assert batch.goal is not None and batch.next_goal is not None
        fut = self.cfg.future_ratio > 0
        self.engine.upload_batch(batch.obs, batch.action, batch.discount, batch.next_obs, batch.goal if use_goal else None,
                                 batch.next_goal if use_goal else None, batch.future_obs if fut else None,
                                 batch.future_goal if (fut and use_goal) else None)

    def _host_view(self, replay_loader: tp.Any) -> tp.Optional[HostStorageView]:
        if not self.native_host_sampling:
            return None
        cached = self._host_views.get(id(replay_loader))
        if cached is not None and cached.replay is replay_loader and cached.still_valid():
            return cached
        view = HostStorageView.adopt(replay_loader)
        self._host_views = {id(replay_loader): view} if view is not None else {}
        return view

    def _reduce_metrics(self, m: tp.Dict[str, float]) -> tp.Dict[str, float]:
        """Per-rank metric blocks -> global values: loss-type entries are partial sums over the rank's rows, the
        others are per-rank means (or replicated)."""
        return reduce_metrics(m, L.METRIC_KEYS + L.OPTIONAL_METRIC_KEYS, self.world, self.engine.device)
