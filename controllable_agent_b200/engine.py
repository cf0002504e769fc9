"""FBStepEngine — the Python owner of one libfb_b200 handle.

PyTorch is used for device memory and streams only: the engine allocates the flat parameter / gradient /
Adam-moment / target segments and the workspace as CUDA tensors, hands their raw pointers to the C ABI
(include/fb_b200.h) and exposes named views on them.  All arithmetic of the gradient step
(url_benchmark/agent/fb_ddpg.py:427-520) happens inside the library's sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import typing as tp

import numpy as np
import torch

from . import _lib as L


@dataclasses.dataclass
class EngineConfig:
    batch: int
    obs_dim: int
    action_dim: int
    z_dim: int
    goal_dim: int
    hidden_dim: int = 1024
    feature_dim: int = 512
    backward_hidden_dim: int = 526
    use_goal: bool = False
    rng_device: bool = False
    contract_mode: int = L.CONTRACT_TCGEN05   # batch x batch contraction: tcgen05 3xTF32 (default) or fp32 SIMT
    mlp_mode: int = L.MLP_TCGEN05             # wide Linear fwd / dX products: tcgen05 3xTF32 (default) or fp32 SIMT
    ortho_coef: float = 1.0
    mix_ratio: float = 0.5
    future_ratio: float = 0.0   # hindsight z (fb_ddpg.py:488-491)
    q_loss: bool = False        # the optional Q loss of update_fb (fb_ddpg.py:330-341)
    q_loss_coef: float = 0.01
    preprocess: bool = True     # False: one deep trunk instead of the two embeds (fb_modules.py:102-104,175-177)
    add_trunk: bool = False     # trunk Linear + ReLU between the embeds and the heads (fb_modules.py:96-100,169-173)
    rand_weight: bool = False   # mixed z rows = random weighted sums of B rows (fb_ddpg.py:475-482)
    boltzmann: bool = False     # DiagGaussianActor + SquashedNormal actions, actor loss mean(temp * log pi - Q) (fb_modules.py:129-151)
    temp: float = 1.0
    log_std_bounds: tp.Tuple[float, float] = (-5.0, 2.0)
    debug: bool = False         # backward_net / backward_target_net = nn.Identity (fb_ddpg.py:128-130); needs z_dim == goal_dim
    norm_z: bool = True         # sqrt(z_dim)-sphere projection of B's output / of z (fb_modules.py:227-229, fb_ddpg.py:228,483)
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    seed: int = 0
    global_batch: tp.Optional[int] = None   # multi-GPU: rows of all ranks; None = batch
    row_offset: int = 0                     # multi-GPU: global index of local row 0
    nccl: tp.Optional[tp.Tuple[bytes, int, int]] = None   # (128-byte unique id, world, rank): collectives inside the step graph
    p2p: tp.Optional[tp.Tuple[int, int]] = None           # (world, rank): exchange by the library's own kernels over NVLink peer memory
    fused: bool = False  # True: runs of consecutive launches inside one persistent kernel per segment (k_fused_stack, fused.cuh);
    #                      False (default, the faster path on B200: DESIGN.md section 4b): one kernel per launch on three lanes


def _ptr(t: tp.Optional[torch.Tensor]) -> tp.Optional[int]:
    return None if t is None else t.data_ptr()


def _on_device(fn: tp.Callable) -> tp.Callable:
    """Run a method with the engine's device current: the library launches on streams / with handles of THAT device, whatever
    device the caller's thread has selected (cfg.device='cuda:1' while device 0 is current)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self: "FBStepEngine", *args: tp.Any, **kwargs: tp.Any) -> tp.Any:
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapped


class FBStepEngine:
    def __init__(self, cfg: EngineConfig, device: tp.Union[str, torch.device] = "cuda", p2p_attach: str = "ipc") -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("controllable_agent_b200 runs on CUDA (sm_100a) only; there is no CPU path "
                               f"(got device={device!r})")
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available: controllable_agent_b200 has no CPU fallback")
        if self.device.index is None:   # "cuda" -> the indexed current device: tensor.device comparisons and stream lookups need the index
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = L.load()
        self.cfg = cfg
        c = L.fb_config(abi_version=L.FB_ABI_VERSION, batch=cfg.batch, global_batch=cfg.global_batch or cfg.batch,
                        row_offset=cfg.row_offset, obs_dim=cfg.obs_dim, action_dim=cfg.action_dim, z_dim=cfg.z_dim,
                        goal_dim=cfg.goal_dim, hidden_dim=cfg.hidden_dim, feature_dim=cfg.feature_dim,
                        backward_hidden_dim=cfg.backward_hidden_dim, use_goal=int(cfg.use_goal),
                        rng_device=int(cfg.rng_device), contract_mode=int(cfg.contract_mode), mlp_mode=int(cfg.mlp_mode), ortho_coef=cfg.ortho_coef, mix_ratio=cfg.mix_ratio,
                        future_ratio=cfg.future_ratio,
                        beta1=cfg.beta1, beta2=cfg.beta2, adam_eps=cfg.adam_eps, seed=cfg.seed,
                        q_loss=int(cfg.q_loss), q_loss_coef=cfg.q_loss_coef, no_norm_z=int(not cfg.norm_z), rand_weight=int(cfg.rand_weight), add_trunk=int(cfg.add_trunk), no_preprocess=int(not cfg.preprocess),
                        boltzmann=int(cfg.boltzmann), temp=float(cfg.temp), log_std_min=float(cfg.log_std_bounds[0]), log_std_max=float(cfg.log_std_bounds[1]),
                        fused_stacks=int(bool(cfg.fused)), debug_identity_b=int(bool(cfg.debug)))
        h = C.c_void_p()
        L.check(self.lib.fb_create(C.byref(c), C.byref(h)), "fb_create")
        self.h = h
        self.has_nccl = False
        self.has_p2p = False
        self._arena_bufs: tp.Optional["L.fb_buffers"] = None
        self._bound = False
        if cfg.nccl is not None:   # before fb_bind: the plan places the all-gather / all-reduce launches
            uid, world, rank = cfg.nccl
            buf = C.create_string_buffer(uid, 128)
            with torch.cuda.device(self.device):
                L.check(self.lib.fb_nccl_init(h, L.nccl_library_path(), buf, world, rank), "fb_nccl_init")
            self.has_nccl = True
        if cfg.p2p is not None:    # the library's own exchange kernels over peer memory: arena first, then the peers' arenas, then bind
            world, rank = cfg.p2p
            self._ipc_handle = C.create_string_buffer(64)
            self._arena_bufs = L.fb_buffers()
            with torch.cuda.device(self.device):
                L.check(self.lib.fb_p2p_create(h, world, rank, self._ipc_handle, C.byref(self._arena_bufs)), "fb_p2p_create")
            self.has_p2p = True
            if p2p_attach == "ipc":   # one process per GPU: handles travel through torch.distributed (plumbing)
                import torch.distributed as dist
                handles: tp.List[tp.Any] = [None] * world
                dist.all_gather_object(handles, self._ipc_handle.raw)
                blob = C.create_string_buffer(b"".join(handles), 64 * world)
                with torch.cuda.device(self.device):
                    L.check(self.lib.fb_p2p_attach(h, blob, None), "fb_p2p_attach")
                self._bind()
                dist.barrier()    # every arena exists and is zeroed before anybody signals into it
            # else: the caller attaches engines of this process to each other (attach_local) and that binds
        else:
            self._bind()

    class _DeviceBlock:
        """A device address range as a __cuda_array_interface__ object, so that torch can view memory the library allocated."""

        def __init__(self, ptr: int, n: int) -> None:
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

    @staticmethod
    def attach_local(engines: tp.Sequence["FBStepEngine"]) -> None:
        """Several ranks inside ONE process (tests: R engines on one device): exchange the arena addresses directly, then bind."""
        arenas = (C.c_void_p * len(engines))(*[e.lib.fb_p2p_arena(e.h) for e in engines])
        for e in engines:
            with torch.cuda.device(e.device):
                L.check(e.lib.fb_p2p_attach(e.h, None, arenas), "fb_p2p_attach")
                e._bind()

    def _bind(self) -> None:
        h, cfg = self.h, self.cfg
        with torch.cuda.device(self.device):
            n_fb, n_actor = self.lib.fb_flat_size(h, 0), self.lib.fb_flat_size(h, 1)
            z = lambda n: torch.zeros(n, dtype=torch.float32, device=self.device)  # noqa: E731
            if self._arena_bufs is not None:   # gradients and parameters live in the arena the peers can address
                ab = self._arena_bufs
                wrap = lambda ptr, n: torch.as_tensor(self._DeviceBlock(ptr, n), device=self.device)  # noqa: E731
                self.param_fb, self.grad_fb = wrap(ab.d_param_fb, n_fb), wrap(ab.d_grad_fb, n_fb)
                self.param_actor, self.grad_actor = wrap(ab.d_param_actor, n_actor), wrap(ab.d_grad_actor, n_actor)
            else:
                self.param_fb, self.grad_fb, self.param_actor, self.grad_actor = z(n_fb), z(n_fb), z(n_actor), z(n_actor)
            self.m_fb, self.v_fb, self.target_fb = z(n_fb), z(n_fb), z(n_fb)
            self.m_actor, self.v_actor = z(n_actor), z(n_actor)
            ws_bytes = self.lib.fb_workspace_bytes(h)
            self.workspace = torch.zeros(ws_bytes + 256, dtype=torch.uint8, device=self.device)
            ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256
            self._ws_shift = ws_ptr - self.workspace.data_ptr()
            bufs = L.fb_buffers(self.param_fb.data_ptr(), self.grad_fb.data_ptr(), self.m_fb.data_ptr(), self.v_fb.data_ptr(),
                                self.target_fb.data_ptr(), self.param_actor.data_ptr(), self.grad_actor.data_ptr(),
                                self.m_actor.data_ptr(), self.v_actor.data_ptr(), ws_ptr, ws_bytes)
            L.check(self.lib.fb_bind(h, C.byref(bufs), self._stream()), "fb_bind")
        self._bound = True
        cfg = self.cfg
        self.layout = {net: self._tensor_table(net) for net in (L.NET_FORWARD, L.NET_BACKWARD, L.NET_ACTOR)}
        self._scalars: tp.Optional[tp.Tuple[float, ...]] = None
        self._keepalive: tp.List[tp.Any] = []
        self._idx_stage: tp.Optional[torch.Tensor] = None   # pinned staging ring for host-provided index arrays
        self._idx_slot = 0
        self._idx_events: tp.List[tp.Optional[torch.cuda.Event]] = [None] * 16
        # host-batch staging (upload_batch): two pinned [batch, pitch] blocks in the packed row layout of the library
        offs, pitch = (C.c_int32 * 9)(), C.c_int32()
        L.check(self.lib.fb_batch_row_layout(cfg.obs_dim, cfg.action_dim, cfg.goal_dim if cfg.use_goal else 0, 0,
                                             int(cfg.future_ratio > 0), offs, C.byref(pitch)))
        self._row_offsets, self._row_pitch = list(offs), pitch.value
        self._row_stage: tp.Optional[torch.Tensor] = None
        self._row_events: tp.List[tp.Optional[torch.cuda.Event]] = [None, None]
        self._row_slot = 0
        mp = self.lib.fb_metrics_ptr(h)
        self._metrics = self._wrap(mp, L.METRIC_COUNT)
        self._metrics_host = torch.zeros(L.METRIC_COUNT, dtype=torch.float32).pin_memory()

    # -- plumbing ---------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self) -> None:
        if getattr(self, "h", None) is not None and self.h:
            self.lib.fb_destroy(self.h)
            self.h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _wrap(self, ptr: int, numel: int) -> torch.Tensor:
        """fp32 tensor view of `numel` floats at device address `ptr` inside the workspace."""
        off = ptr - self.workspace.data_ptr()
        assert 0 <= off and off + 4 * numel <= self.workspace.numel(), "pointer outside the workspace"
        return self.workspace[off:off + 4 * numel].view(torch.float32)

    def _tensor_table(self, net: int) -> tp.List[tp.Tuple[str, int, tp.Tuple[int, ...]]]:
        out = []
        for i in range(self.lib.fb_num_tensors(self.h, net)):
            off, rows, cols = C.c_size_t(), C.c_int(), C.c_int()
            name = C.create_string_buffer(64)
            L.check(self.lib.fb_tensor_info(self.h, net, i, C.byref(off), C.byref(rows), C.byref(cols), name, 64))
            shape = (rows.value, cols.value) if cols.value else (rows.value,)
            out.append((name.value.decode(), off.value, shape))
        return out

    def tensors(self, net: int, which: str = "param") -> "tp.OrderedDict[str, torch.Tensor]":
        """Named views (nn.Module registration order) of one net inside a flat segment.
        which: param | grad | m | v | target (target only for forward/backward nets)."""
        import collections
        seg = {"param": (self.param_fb, self.param_actor), "grad": (self.grad_fb, self.grad_actor),
               "m": (self.m_fb, self.m_actor), "v": (self.v_fb, self.v_actor), "target": (self.target_fb, None)}[which]
        flat = seg[1] if net == L.NET_ACTOR else seg[0]
        if flat is None:
            raise ValueError("the actor has no target network")
        out = collections.OrderedDict()
        for name, off, shape in self.layout[net]:
            n = int(np.prod(shape))
            out[name] = flat[off:off + n].view(*shape)
        return out

    def view(self, name: str) -> torch.Tensor:
        """Strided view of a named workspace matrix (test introspection)."""
        p, rows, cols, ld = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        L.check(self.lib.fb_workspace_view(self.h, name.encode(), C.byref(p), C.byref(rows), C.byref(cols), C.byref(ld)),
                f"fb_workspace_view({name})")
        flat = self._wrap(p.value, (rows.value - 1) * ld.value + cols.value)
        return torch.as_strided(flat, (rows.value, cols.value), (ld.value, 1))

    # -- per-step inputs -------------------------------------------------------------------------
    @_on_device
    def set_scalars(self, stddev: float, stddev_clip: float, lr_forward: float, lr_backward: float, lr_actor: float,
                    tau: float, replay_discount: float = 1.0, replay_future: float = 1.0, grad_scale: float = 1.0) -> None:
        key = (stddev, stddev_clip, lr_forward, lr_backward, lr_actor, tau, replay_discount, replay_future, grad_scale)
        if key == self._scalars:
            return
        s = L.fb_step_scalars(*key)
        L.check(self.lib.fb_set_step_scalars(self.h, C.byref(s), self._stream()), "fb_set_step_scalars")
        self._scalars = key

    def _dev_i32(self, x: tp.Any) -> tp.Optional[torch.Tensor]:
        if x is None:
            return None
        t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
        if not (t.device.type == "cuda" and t.device.index == self.device.index):   # through a pinned staging ring so that the upload is asynchronous
            if self._idx_stage is None:
                self._idx_stage = torch.empty((16, self.cfg.batch), dtype=torch.int32).pin_memory()
                self._idx_slot = 0
            slot = self._idx_slot % 16
            self._idx_slot += 1
            if self._idx_events[slot] is not None:
                self._idx_events[slot].synchronize()     # the upload that last used this slot has left the host buffer
            stage = self._idx_stage[slot]
            stage.copy_(t.reshape(-1))
            t = stage.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._idx_events[slot] = ev
        t = t.to(dtype=torch.int32).contiguous()
        assert t.numel() == self.cfg.batch
        return t

    def _dev_f32(self, x: tp.Any, cols: int) -> tp.Optional[torch.Tensor]:
        if x is None:
            return None
        t = torch.as_tensor(x)
        if not (t.device.type == "cuda" and t.device.index == self.device.index):   # host -> device: asynchronous when the source is pinned, ordered on the current stream
            t = t.to(device=self.device, dtype=torch.float32, non_blocking=True)
        t = t.to(dtype=torch.float32).contiguous()
        assert t.numel() == self.cfg.batch * cols, (tuple(t.shape), self.cfg.batch, cols)
        return t

    @_on_device
    def set_indices(self, ep_idx: tp.Any = None, step_idx: tp.Any = None, future_idx: tp.Any = None, perm: tp.Any = None,
                    mix_mask: tp.Any = None) -> None:
        ts = [self._dev_i32(x) for x in (ep_idx, step_idx, future_idx, perm, mix_mask)]
        L.check(self.lib.fb_set_indices(self.h, *[_ptr(t) for t in ts], self._stream()), "fb_set_indices")
        self._keepalive = ts  # copies are enqueued on the stream; keep the sources alive until the next call

    @_on_device
    def set_batch(self, obs: tp.Any, action: tp.Any, discount: tp.Any, next_obs: tp.Any, goal: tp.Any = None,
                  next_goal: tp.Any = None) -> None:
        c = self.cfg
        ts = [self._dev_f32(obs, c.obs_dim), self._dev_f32(action, c.action_dim), self._dev_f32(discount, 1),
              self._dev_f32(next_obs, c.obs_dim),
              self._dev_f32(goal, c.goal_dim) if c.use_goal else None,
              self._dev_f32(next_goal, c.goal_dim) if c.use_goal else None]
        L.check(self.lib.fb_set_batch(self.h, *[_ptr(t) for t in ts], self._stream()), "fb_set_batch")
        self._keepalive_batch = ts

    @_on_device
    def upload_batch(self, obs: tp.Any, action: tp.Any, discount: tp.Any, next_obs: tp.Any, goal: tp.Any = None,
                     next_goal: tp.Any = None, future_obs: tp.Any = None, future_goal: tp.Any = None) -> int:
        """Host arrays of one sampled batch (EpisodeBatch fields, replay_buffer.py:27-40) -> the step's packed batch
        block with ONE asynchronous host-to-device copy (the reference's EpisodeBatch.to issues one blocking pageable
        copy per field, replay_buffer.py:50-63).  Returns the bytes copied."""
        c = self.cfg
        if self._row_stage is None:
            self._row_stage = torch.zeros((2, c.batch, self._row_pitch), dtype=torch.float32).pin_memory()
        slot = self._row_slot
        self._row_slot ^= 1
        if self._row_events[slot] is not None:
            self._row_events[slot].synchronize()   # the copy that last read this block has completed
        rows = self._row_stage[slot].numpy()
        o = self._row_offsets   # obs, action, (reward, discount), next_obs, goal, next_goal, ...

        def put(off: int, x: tp.Any, dim: int) -> None:
            a = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
            rows[:, off:off + dim] = a.reshape(c.batch, dim)

        put(o[0], obs, c.obs_dim)
        put(o[1], action, c.action_dim)
        put(o[2] + 1, discount, 1)
        put(o[3], next_obs, c.obs_dim)
        if c.use_goal:
            put(o[4], goal, c.goal_dim)
            put(o[5], next_goal, c.goal_dim)
        if c.future_ratio > 0:   # hindsight inputs (EpisodeBatch.future_obs / future_goal)
            if future_obs is None or (c.use_goal and future_goal is None):
                raise ValueError("future_ratio > 0 needs batches with future_obs / future_goal (a replay buffer with future < 1)")
            put(o[7], future_obs, c.obs_dim)
            if c.use_goal:
                put(o[8], future_goal, c.goal_dim)
        L.check(self.lib.fb_upload_batch(self.h, self._row_stage[slot].data_ptr(), self._row_pitch, self._stream()), "fb_upload_batch")
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._row_events[slot] = ev
        return 4 * c.batch * self._row_pitch

    @_on_device
    def upload_host_rows(self, host: tp.Any, ep_idx: np.ndarray, step_idx: np.ndarray, future_idx: tp.Optional[np.ndarray],
                         replay_discount: float) -> int:
        """ReplayBuffer.sample's gathers + EpisodeBatch.to for a reference-layout HOST replay (replay.HostStorageView): the library
        copies the sampled rows of the host storage straight into the pinned block (fb_host_gather_rows), one asynchronous
        host-to-device copy follows.  Returns the bytes copied."""
        c = self.cfg
        if self._row_stage is None:
            self._row_stage = torch.zeros((2, c.batch, self._row_pitch), dtype=torch.float32).pin_memory()
        slot = self._row_slot
        self._row_slot ^= 1
        if self._row_events[slot] is not None:
            self._row_events[slot].synchronize()   # the copy that last read this block has completed
        want_future = c.future_ratio > 0
        if want_future and future_idx is None:
            raise ValueError("future_ratio > 0 needs a replay buffer created with future < 1 (it samples no future_obs / future_goal)")
        if bool(c.use_goal) != bool(host.c.goal):
            raise ValueError("agent.goal_space and the replay's `goal` storage disagree")
        ep = np.ascontiguousarray(ep_idx, dtype=np.int32)
        st = np.ascontiguousarray(step_idx, dtype=np.int32)
        fu = np.ascontiguousarray(future_idx, dtype=np.int32) if want_future else None
        assert ep.shape[0] == c.batch and st.shape[0] == c.batch
        L.check(self.lib.fb_host_gather_rows(C.byref(host.c), ep.ctypes.data, st.ctypes.data, fu.ctypes.data if fu is not None else None,
                                             c.batch, float(replay_discount), self._row_stage[slot].data_ptr(), self._row_pitch),
                "fb_host_gather_rows")
        L.check(self.lib.fb_upload_batch(self.h, self._row_stage[slot].data_ptr(), self._row_pitch, self._stream()), "fb_upload_batch")
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._row_events[slot] = ev
        return 4 * c.batch * self._row_pitch

    @_on_device
    def set_future_mask(self, mask: tp.Any) -> None:
        t = self._dev_i32(mask)
        L.check(self.lib.fb_set_future_mask(self.h, _ptr(t), self._stream()), "fb_set_future_mask")
        self._keepalive_future = t

    @_on_device
    def set_mix_weights(self, weight: tp.Any, row_scale: tp.Any) -> None:
        """rand_weight with host RNG: the [batch, batch] U(0,1) weight rows and [batch] row scales (fb_ddpg.py:477-480)."""
        B = self.cfg.batch
        w = torch.as_tensor(weight, dtype=torch.float32).to(self.device, non_blocking=True).contiguous()
        u = torch.as_tensor(row_scale, dtype=torch.float32).to(self.device, non_blocking=True).contiguous()
        assert w.numel() == B * B and u.numel() == B
        L.check(self.lib.fb_set_mix_weights(self.h, _ptr(w), _ptr(u), self._stream()), "fb_set_mix_weights")
        self._keepalive_mixw = (w, u)

    @_on_device
    def set_z(self, z: tp.Any) -> None:
        t = self._dev_f32(z, self.cfg.z_dim)
        L.check(self.lib.fb_set_z(self.h, _ptr(t), self._stream()), "fb_set_z")
        self._keepalive_z = t

    @_on_device
    def set_noise(self, noise_fb: tp.Any = None, noise_actor: tp.Any = None) -> None:
        a, b = self._dev_f32(noise_fb, self.cfg.action_dim), self._dev_f32(noise_actor, self.cfg.action_dim)
        L.check(self.lib.fb_set_noise(self.h, _ptr(a), _ptr(b), self._stream()), "fb_set_noise")
        self._keepalive_noise = (a, b)

    @_on_device
    def bind_replay(self, view: "L.fb_replay_view") -> None:
        L.check(self.lib.fb_bind_replay(self.h, C.byref(view), self._stream()), "fb_bind_replay")

    @_on_device
    def set_adam_steps(self, fb_step: int, actor_step: int) -> None:
        L.check(self.lib.fb_set_adam_steps(self.h, fb_step, actor_step, self._stream()), "fb_set_adam_steps")

    @_on_device
    def get_adam_steps(self) -> tp.Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        L.check(self.lib.fb_get_adam_steps(self.h, C.byref(a), C.byref(b), self._stream()), "fb_get_adam_steps")
        return a.value, b.value

    # -- the step --------------------------------------------------------------------------------
    @_on_device
    def run(self, mask: int = L.PHASE_ALL, graph: bool = False, fused: tp.Optional[bool] = None) -> None:
        if not (self.cfg.fused if fused is None else fused):
            mask |= L.RUN_UNFUSED
        L.check(self.lib.fb_run(self.h, mask, int(graph), self._stream()), f"fb_run(0x{mask:x})")

    @_on_device
    def prepare_graph(self, mask: int = L.PHASE_ALL, fused: tp.Optional[bool] = None) -> None:
        """Capture and instantiate the CUDA graph of `mask` without launching it."""
        if not (self.cfg.fused if fused is None else fused):
            mask |= L.RUN_UNFUSED
        L.check(self.lib.fb_run(self.h, mask, 2, self._stream()), f"fb_run(0x{mask:x}, capture only)")

    def launch_count(self, mask: int = L.PHASE_ALL, fused: tp.Optional[bool] = None) -> int:
        """Kernel launches one run of `mask` issues (fused execution: one per fused segment / stand-alone kernel)."""
        if not (self.cfg.fused if fused is None else fused):
            mask |= L.RUN_UNFUSED
        return self.lib.fb_launch_count(self.h, mask)

    @_on_device
    def profile_ops(self, mask: int = L.PHASE_ALL, reps: int = 5) -> tp.List[tp.Dict[str, tp.Any]]:
        """Per-launch CUDA-event timings of the phases in `mask` (runs the step `reps` times eagerly)."""
        cap = self.launch_count(mask, fused=False)
        ms, kind = (C.c_float * cap)(), (C.c_int32 * cap)()
        flops, nbytes = (C.c_double * cap)(), (C.c_double * cap)()
        n = self.lib.fb_profile_ops(self.h, mask, reps, self._stream(), ms, kind, flops, nbytes, cap)
        if n < 0:
            L.check(n, "fb_profile_ops")
        return [{"ms": ms[i], "kind": L.OP_KINDS[kind[i]], "flops": flops[i], "bytes": nbytes[i]} for i in range(n)]

    @_on_device
    def fused_profile(self, mask: int = L.PHASE_ALL, reps: int = 5) -> tp.List[tp.Dict[str, tp.Any]]:
        """Per-stage timings of the fused execution of `mask` (device timestamps between the grid barriers) and per-kernel
        timings of the launches that stay kernels of their own."""
        cap = 1024
        us, info = (C.c_float * cap)(), (C.c_int32 * (5 * cap))()
        n = self.lib.fb_fused_profile(self.h, mask, reps, self._stream(), us, info, cap)
        if n < 0:
            L.check(n, "fb_fused_profile")
        out = []
        for i in range(n):
            unit, stage, items, t, count = (info[5 * i + j] for j in range(5))
            what = L.FS_TYPES[t] if t >= 0 else ("staging batch" if t == -2 else "kernel:" + L.OP_KINDS[-1 - t])
            out.append({"us": us[i], "unit": unit, "stage": stage, "items": items, "first": what, "count": count})
        return out

    @_on_device
    def p2p_status(self) -> tp.Tuple[int, tp.List[int]]:
        """(error code, epochs of the 8 exchange barriers); raises nothing: 0 = every wait of this rank was answered."""
        code, ep = C.c_uint32(), (C.c_uint64 * 8)()
        L.check(self.lib.fb_p2p_status(self.h, C.byref(code), ep, self._stream()), "fb_p2p_status")
        return code.value, list(ep)

    def moment_slice(self, actor: bool) -> tp.Tuple[int, int]:
        """p2p exchange: the [first, first + count) float range of the flat segment whose Adam moments THIS rank owns
        (every other range of m / v is dead on this rank); without p2p: the whole segment."""
        if not self.has_p2p:
            n = (self.m_actor if actor else self.m_fb).numel()
            return 0, n
        a, b = C.c_size_t(), C.c_size_t()
        L.check(self.lib.fb_p2p_slice(self.h, int(actor), C.byref(a), C.byref(b)), "fb_p2p_slice")
        return a.value, b.value

    def full_moments(self) -> tp.Dict[str, torch.Tensor]:
        """Adam moments of the whole segments: with the p2p exchange the slices are all-gathered from their owners (collective)."""
        out = {"m_fb": self.m_fb, "v_fb": self.v_fb, "m_actor": self.m_actor, "v_actor": self.v_actor}
        if not self.has_p2p:
            return {k: v.clone() for k, v in out.items()}
        import torch.distributed as dist
        world = self.cfg.p2p[0]
        res = {}
        for k, t in out.items():
            first, count = self.moment_slice(k.endswith("actor"))
            per = -(-t.numel() // (4 * world)) * 4
            mine = torch.zeros(per, dtype=torch.float32, device=self.device)
            mine[:count] = t[first:first + count]
            full = torch.empty(per * world, dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(full, mine)
            res[k] = full[:t.numel()].clone()
        return res

    def gather_block(self) -> tp.Tuple[torch.Tensor, torch.Tensor]:
        """(local, global) packed [rows, pitch] blocks for the multi-GPU all-gather between FB_FWD and FB_LOSS."""
        n, pl, pg = C.c_int(), C.c_void_p(), C.c_void_p()
        L.check(self.lib.fb_gather_block(self.h, C.byref(n), C.byref(pl), C.byref(pg)), "fb_gather_block")
        gb = self.cfg.global_batch or self.cfg.batch
        return (self._wrap(pl.value, self.cfg.batch * n.value).view(self.cfg.batch, n.value),
                self._wrap(pg.value, gb * n.value).view(gb, n.value))

    # -- inference plans (per environment step / zero-shot inference; fb_ddpg.py:177-222,258-289) ----------------
    def _infer_io(self) -> tp.Dict[str, tp.Any]:
        if getattr(self, "_infer", None) is None:
            c = self.cfg
            names = ("infer_obs", "infer_z", "infer_mu", "infer_goal", "infer_b", "infer_goal_batch", "infer_reward", "infer_zsum")
            views = {n: self.view(n) for n in names}
            R = L.INFER_ROWS
            pin = lambda *shape: torch.zeros(shape, dtype=torch.float32).pin_memory()  # noqa: E731
            self._infer = {"v": views, "obs": pin(R, c.obs_dim), "z": pin(R, c.z_dim), "mu": pin(R, c.action_dim * (2 if c.boltzmann else 1)),
                           "goal": pin(R, c.goal_dim), "b": pin(R, c.z_dim), "zsum": pin(1, c.z_dim)}
        return self._infer

    @_on_device
    def infer_actor(self, obs: np.ndarray, z: np.ndarray, graph: bool = True) -> np.ndarray:
        """mu = tanh(policy(obs, z)) of the online actor for up to 8 rows (fb_modules.py:110-122): one small upload, the
        FB_PHASE_INFER_ACTOR graph, one small read-back.  cfg.boltzmann: the rows are [mu (pre-tanh) | std] of the
        DiagGaussianActor (fb_modules.py:141-151)."""
        io = self._infer_io()
        n = obs.shape[0]
        assert n <= L.INFER_ROWS and z.shape[0] == n
        io["obs"][:n] = torch.from_numpy(np.ascontiguousarray(obs, dtype=np.float32))
        io["z"][:n] = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float32))
        io["v"]["infer_obs"].copy_(io["obs"], non_blocking=True)
        io["v"]["infer_z"].copy_(io["z"], non_blocking=True)
        self.run(L.PHASE_INFER_ACTOR, graph=graph)
        io["mu"].copy_(io["v"]["infer_mu"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return io["mu"][:n].numpy().copy()

    @_on_device
    def infer_backward(self, goal: np.ndarray, graph: bool = True) -> np.ndarray:
        """sqrt(z_dim) * normalize(backward_net(goal)) for up to 8 rows (fb_modules.py:223-230)."""
        io = self._infer_io()
        n = goal.shape[0]
        assert n <= L.INFER_ROWS
        io["goal"][:n] = torch.from_numpy(np.ascontiguousarray(goal, dtype=np.float32))
        io["v"]["infer_goal"].copy_(io["goal"], non_blocking=True)
        self.run(L.PHASE_INFER_B, graph=graph)
        io["b"].copy_(io["v"]["infer_b"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return io["b"][:n].numpy().copy()

    @_on_device
    def infer_backward_weighted_sum(self, goal: tp.Any, reward: tp.Any, graph: bool = True) -> np.ndarray:
        """sum_i reward_i * backward_net(goal_i) over N rows (fb_ddpg.py:213-215), in chunks of `batch` rows (the last chunk
        padded with zero rewards)."""
        io = self._infer_io()
        B = self.cfg.batch
        goal = torch.as_tensor(goal, dtype=torch.float32)
        reward = torch.as_tensor(reward, dtype=torch.float32).reshape(-1, 1)
        N = goal.shape[0]
        assert reward.shape[0] == N
        vg, vr = io["v"]["infer_goal_batch"], io["v"]["infer_reward"]
        io["v"]["infer_zsum"].zero_()
        for i in range(0, N, B):
            n = min(B, N - i)
            if n < B:
                vr.zero_()
            vg[:n].copy_(goal[i:i + n], non_blocking=True)
            vr[:n].copy_(reward[i:i + n], non_blocking=True)
            self.run(L.PHASE_INFER_BN, graph=graph)
        io["zsum"].copy_(io["v"]["infer_zsum"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return io["zsum"][0].numpy().copy()

    def metrics_tensor(self) -> torch.Tensor:
        return self._metrics

    @_on_device
    def read_metrics(self) -> tp.Dict[str, float]:
        """One D2H copy of the metrics block (replaces the ~19 .item() syncs of fb_ddpg.py:357-377,414-418)."""
        self._metrics_host.copy_(self._metrics, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        vals = self._metrics_host.tolist()
        return {k: vals[i] for i, k in enumerate(L.METRIC_KEYS + L.OPTIONAL_METRIC_KEYS)}
