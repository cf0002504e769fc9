"""HBM-resident replay buffer with the API of url_benchmark.in_memory_replay_buffer.ReplayBuffer.

Storage: one packed fp32 row per (episode, time step) in device memory,
    [observation | action | reward discount 0 0 | goal | extra(meta keys)]      (every field 16-byte aligned)
so that `sample()` (in_memory_replay_buffer.py:139-190) is ONE gather kernel (fb_replay_gather) whose warps read
rows t-1 and t of an episode as contiguous 128-bit loads, instead of ~8 numpy fancy-index gathers followed by 6-9
pageable host->device copies (replay_buffer.py:50-63).  Fields the update path never reads (physics, step_type)
stay in host numpy arrays exactly as the reference keeps them.
"""
from __future__ import annotations

import collections
import ctypes as C
import dataclasses
import typing as tp

import numpy as np
import torch

from . import _lib as L

T = tp.TypeVar("T", np.ndarray, torch.Tensor)
B = tp.TypeVar("B", bound="EpisodeBatch")

# field names of url_benchmark.dmc.ExtendedGoalTimeStep (dmc.py:35-73): everything else stored by add() is "meta"
TIMESTEP_FIELDS = ("step_type", "reward", "discount", "observation", "physics", "goal", "action")
HOT_FIELDS = ("observation", "action", "reward", "discount", "goal")


@dataclasses.dataclass
class EpisodeBatch(tp.Generic[T]):
    """Same container as url_benchmark.replay_buffer.EpisodeBatch (replay_buffer.py:27-103)."""
    obs: T
    action: T
    reward: T
    next_obs: T
    discount: T
    meta: tp.Dict[str, T] = dataclasses.field(default_factory=dict)
    _physics: tp.Optional[T] = None
    goal: tp.Optional[T] = None
    next_goal: tp.Optional[T] = None
    future_obs: tp.Optional[T] = None
    future_goal: tp.Optional[T] = None

    def __post_init__(self) -> None:
        for name in ("reward", "discount"):
            if not isinstance(getattr(self, name), (np.ndarray, torch.Tensor)):
                raise TypeError(f"EpisodeBatch.{name} must be an array or a tensor")
        if not isinstance(self.meta, dict):
            raise TypeError("EpisodeBatch.meta must be a dict")

    # -- array fields, in declaration order (everything but `meta`) ---------------------------------
    @classmethod
    def _array_fields(cls) -> tp.List[str]:
        return [f.name for f in dataclasses.fields(cls) if f.name != "meta"]

    def to(self, device: tp.Union[str, torch.device]) -> "EpisodeBatch[torch.Tensor]":
        """Every array / tensor field (and every meta entry) as a tensor on `device`; absent optional fields stay None
        (replay_buffer.py:50-63).  Batches sampled from the HBM replay are already device tensors: nothing is copied."""
        def move(name: str, value: tp.Any) -> tp.Any:
            if value is None:
                return None
            if not isinstance(value, (torch.Tensor, np.ndarray)):
                raise RuntimeError(f"Not sure what to do with {name}: {value}")
            return torch.as_tensor(value, device=device)
        moved = {name: move(name, getattr(self, name)) for name in self._array_fields()}
        return EpisodeBatch(meta={k: torch.as_tensor(v, device=device) for k, v in self.meta.items()}, **moved)

    @classmethod
    def collate_fn(cls, batches: tp.List["EpisodeBatch[T]"]) -> "EpisodeBatch[torch.Tensor]":
        """Stack a list of batches field by field (replay_buffer.py:65-88): numpy batches are first moved to CPU tensors, a field must
        be present in all batches or in none."""
        if isinstance(batches[0].obs, np.ndarray):
            batches = [b.to("cpu") for b in batches]  # type: ignore
        stacked: tp.Dict[str, tp.Any] = {}
        for name in cls._array_fields():
            column = [getattr(b, name) for b in batches]
            if all(x is None for x in column):
                stacked[name] = None
            elif any(x is None for x in column):
                raise RuntimeError("Found a non-None value mixed with Nones")
            elif isinstance(column[0], torch.Tensor):
                stacked[name] = torch.stack(column)
            else:
                raise RuntimeError(f"Not sure what to do with {name}: {column}")
        meta = {k: torch.stack([b.meta[k] for b in batches]) for k in batches[0].meta}
        return EpisodeBatch(meta=meta, **stacked)

    def unpack(self) -> tp.Tuple[T, T, T, T, T]:
        """(obs, action, reward, discount, next_obs) — the tuple order DDPG-style agents destructure (replay_buffer.py:90-96)."""
        return self.obs, self.action, self.reward, self.discount, self.next_obs

    def with_no_reward(self: B) -> B:
        """A copy whose reward field is all zeros (replay_buffer.py:98-103)."""
        zero = torch.zeros_like(self.reward) if isinstance(self.reward, torch.Tensor) else np.zeros_like(self.reward)
        return dataclasses.replace(self, reward=zero)


def load_episode(fn: tp.Any) -> tp.Dict[str, np.ndarray]:
    """url_benchmark.replay_buffer.load_episode (replay_buffer.py:119-123): one `.npz` file = one episode, field -> `[len+1, ...]`."""
    with open(fn, "rb") as f:
        data = np.load(f)
        return {k: data[k] for k in data.keys()}


def relabel_episode(env: tp.Any, episode: tp.Dict[str, np.ndarray], goal_func: tp.Any) -> tp.Dict[str, np.ndarray]:
    """Recompute an episode's rewards (and, with `goal_func`, its goals) by replaying the stored MuJoCo states through the task
    (same contract as in_memory_replay_buffer.py:40-55).  Env stepping stays on the host, untouched: `env` is the reference's
    dm_control wrapper (needs `physics.reset_context`, `physics.set_state`, `task.get_reward`)."""
    states = np.asarray(episode["physics"])
    rewards = np.empty((states.shape[0], 1), dtype=np.float32)
    goals: tp.List[np.ndarray] = []
    for i, state in enumerate(states):
        with env.physics.reset_context():
            env.physics.set_state(state)
        rewards[i, 0] = env.task.get_reward(env.physics)
        if goal_func is not None:
            goals.append(np.asarray(goal_func(env), dtype=np.float32))
    out = dict(episode)
    out["reward"] = rewards
    if goals:
        out["goal"] = np.stack(goals).astype(np.float32)
    return out


def _round4(x: int) -> int:
    return (x + 3) // 4 * 4


def draw_sample_indices(replay: tp.Any, batch_size: int, exact_stream: bool = True) -> tp.Tuple[np.ndarray, np.ndarray, tp.Optional[np.ndarray]]:
    """(episode, step, future step) indices of one sampled batch: the draws of in_memory_replay_buffer.py:141-161 on the numpy GLOBAL
    generator, call for call, for any object with that class's attributes (this package's ReplayBuffer, the reference's own).
    exact_stream=False (the agent's rng_mode="device", where nothing else follows the reference's generator stream either): the same
    distributions through cheaper calls — a scalar bound instead of a per-row bound array when every episode has the same length (the
    array form costs 4x as much per draw), no consistency asserts."""
    if not isinstance(replay._future, float):
        assert isinstance(replay._future, bool)
        replay._future = float(replay._future)
    if replay._is_fixed_episode_length:
        ep_idx = np.random.randint(0, len(replay), size=batch_size)
    else:
        if replay._episodes_selection_probability is None:
            replay._episodes_selection_probability = replay._episodes_length / replay._episodes_length.sum()
        ep_idx = np.random.choice(np.arange(len(replay._episodes_length)), size=batch_size, p=replay._episodes_selection_probability)
    if not exact_stream and replay._is_fixed_episode_length:
        lengths = replay._episodes_length
        first = int(lengths[0])
        cache = getattr(replay, "_fb_uniform_length", None)     # (array identity, length) of the last check: episodes are appended rarely
        if cache is None or cache[0] is not lengths or cache[1] != first or cache[2] != len(replay):
            n = len(replay)
            uniform = bool(n > 0 and (lengths[:n] == first).all())
            replay._fb_uniform_length = cache = (lengths, first if uniform else -1, n)
        if cache[1] > 0:
            step_idx = np.random.randint(0, cache[1], size=batch_size) + 1
            future_idx = None
            if replay._future < 1:
                future_idx = np.minimum(step_idx + np.random.geometric(p=(1 - replay._future), size=batch_size), cache[1])
            return ep_idx, step_idx, future_idx
    eps_lengths = replay._episodes_length[ep_idx]
    step_idx = np.random.randint(0, eps_lengths) + 1
    assert (step_idx <= eps_lengths).all()
    future_idx = None
    if replay._future < 1:
        future_idx = step_idx + np.random.geometric(p=(1 - replay._future), size=batch_size)
        future_idx = np.clip(future_idx, 0, eps_lengths)
        assert (future_idx <= eps_lengths).all()
    return ep_idx, step_idx, future_idx


class HostStorageView:
    """A replay buffer that lives in HOST memory in the reference's layout (url_benchmark.in_memory_replay_buffer.ReplayBuffer, or any
    object carrying its attributes: `_storage` name -> fp32 [max_episodes, T + 1, dim], `_episodes_length`, `_discount`, `_future`, ...)
    as the C ABI sees it (fb_host_storage).  `agent.update(replay, step)` samples such a buffer without going through its Python
    `sample()`: the index draws are the reference's (draw_sample_indices, same numpy RNG stream), the row gathers run in the library
    (fb_host_gather_rows) straight into the pinned block that crosses PCIe."""
    FIELDS = ("observation", "action", "discount")

    def __init__(self, replay: tp.Any) -> None:
        st = replay._storage
        self.replay = replay
        self.arrays = {k: st[k] for k in ("observation", "action", "reward", "discount", "goal") if k in st}   # keeps them alive
        E, R = self.arrays["observation"].shape[:2]
        ptr = lambda k: self.arrays[k].ctypes.data if k in self.arrays else None  # noqa: E731
        self.c = L.fb_host_storage(observation=ptr("observation"), action=ptr("action"), reward=ptr("reward"), discount=ptr("discount"),
                                   goal=ptr("goal"), rows_per_episode=R, obs_dim=self.arrays["observation"].shape[2],
                                   action_dim=self.arrays["action"].shape[2],
                                   goal_dim=self.arrays["goal"].shape[2] if "goal" in self.arrays else 0, max_episodes=E)
        self.key = tuple((k, v.ctypes.data, v.shape) for k, v in self.arrays.items())

    @classmethod
    def adopt(cls, replay: tp.Any) -> tp.Optional["HostStorageView"]:
        """The view of `replay` if it is a reference-layout host buffer the library can read in place, else None (the caller then
        falls back to the object's own sample())."""
        st = getattr(replay, "_storage", None)
        need = ("_episodes_length", "_discount", "_future", "_is_fixed_episode_length", "_episodes_selection_probability")
        if not isinstance(st, dict) or not all(hasattr(replay, a) for a in need) or not hasattr(replay, "__len__"):
            return None
        if not all(k in st for k in cls.FIELDS):
            return None
        shape = None
        for k in ("observation", "action", "reward", "discount", "goal"):
            v = st.get(k)
            if v is None:
                continue
            if not (isinstance(v, np.ndarray) and v.dtype == np.float32 and v.ndim == 3 and v.flags["C_CONTIGUOUS"]):
                return None
            if shape is not None and v.shape[:2] != shape:
                return None
            shape = v.shape[:2]
            if k in ("reward", "discount") and v.shape[2] != 1:
                return None
        return cls(replay)

    def still_valid(self) -> bool:
        st = self.replay._storage
        if all(st.get(k) is v for k, v in self.arrays.items()):   # the very same array objects (the per-step check)
            return True
        return all(k in st and st[k].ctypes.data == p and st[k].shape == shp for k, p, shp in self.key)


class ReplayBuffer:
    """Drop-in for in_memory_replay_buffer.ReplayBuffer(max_episodes, discount, future, max_episode_length)."""

    def __init__(self, max_episodes: int, discount: float, future: float, max_episode_length: tp.Optional[int] = None,
                 device: tp.Union[str, torch.device, None] = None) -> None:
        self._max_episodes = max_episodes
        self._discount = discount
        assert 0 <= future <= 1
        self._future = future
        self._current_episode: tp.Dict[str, tp.List[np.ndarray]] = collections.defaultdict(list)
        self._idx = 0
        self._full = False
        self._num_transitions = 0
        self._collected_episodes = 0
        self._batch_names = set(TIMESTEP_FIELDS)
        self._episodes_length = np.zeros(max_episodes, dtype=np.int32)
        self._episodes_selection_probability: tp.Optional[np.ndarray] = None
        self._is_fixed_episode_length = True
        self._max_episode_length = max_episode_length
        self._device = torch.device(device) if device is not None else None
        self._init_device_state()

    def _init_device_state(self) -> None:
        self._host: tp.Dict[str, np.ndarray] = {}        # cold fields, reference layout [E, T+1, d]
        self._dims: tp.Dict[str, int] = {}               # hot field -> width
        self._extra: tp.List[tp.Tuple[str, int]] = []    # meta keys and widths, storage order
        self._offsets: tp.Dict[str, int] = {}
        self._rows: tp.Optional[torch.Tensor] = None     # [E, R, stride] fp32 on the device
        self._ep_len_dev: tp.Optional[torch.Tensor] = None
        self._rows_per_episode = 0
        self._row_stride = 0
        self._version = 0                                # bumped when the device storage is (re)allocated
        self._staging: tp.Optional[torch.Tensor] = None  # [2, R, stride] pinned: double-buffered episode uploads
        self._staging_events: tp.List[tp.Optional[torch.cuda.Event]] = [None, None]
        self._staging_slot = 0
        self._storage_cache: tp.Optional[tp.Tuple[int, tp.Dict[str, np.ndarray]]] = None   # (content version, host view)
        self._content = 0                                # bumped on every write to the device rows

    # -- bookkeeping identical to the reference ---------------------------------------------------
    def __len__(self) -> int:
        return self._max_episodes if self._full else self._idx

    @property
    def avg_episode_length(self) -> int:
        return round(self._episodes_length[:len(self)].mean())

    @property
    def device(self) -> torch.device:
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("controllable_agent_b200.ReplayBuffer keeps its storage in GPU memory; CUDA is unavailable")
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    # -- ingest ----------------------------------------------------------------------------------
    def add(self, time_step: tp.Any, meta: tp.Mapping[str, np.ndarray]) -> None:
        """Per-environment-step append (in_memory_replay_buffer.py:104-133); the finished episode is committed to HBM."""
        dtype = np.float32
        for key, value in meta.items():
            self._current_episode[key].append(value)
        for field in dataclasses.fields(time_step):
            value = time_step[field.name]
            if np.isscalar(value):
                value = np.full((1,), value, dtype=dtype)
            if isinstance(value, np.ndarray):
                self._current_episode[field.name].append(np.array(value, dtype=dtype))
        if time_step.last():
            episode = {name: np.array(values, dtype) for name, values in self._current_episode.items()}
            self._current_episode = collections.defaultdict(list)
            self.add_episode(episode)

    def add_episode(self, episode: tp.Mapping[str, np.ndarray]) -> None:
        """Commit one whole episode: `episode[name]` is `[len+1, dim]` (or `[len+1]`) including the dummy first row."""
        ep = {}
        for name, values in episode.items():
            v = np.asarray(values, dtype=np.float32)
            ep[name] = v.reshape(len(v), -1) if v.ndim != 2 else v
        rows = len(ep["discount"])
        if self._rows is None:
            self._allocate(ep, rows)
        if rows > self._rows_per_episode:
            raise ValueError(f"episode of {rows} rows does not fit storage rows_per_episode={self._rows_per_episode} "
                             "(pass max_episode_length for variable-length episodes)")
        self._upload_episode(ep, self._idx, rows)
        for name, v in ep.items():
            if name in self._offsets:
                continue
            if name not in self._host:
                self._host[name] = np.zeros((self._max_episodes, self._rows_per_episode) + v.shape[1:], dtype=np.float32)
            self._host[name][self._idx][:rows] = v
        n = rows - 1  # compensate for the dummy transition at the beginning
        self._episodes_length[self._idx] = n
        if n != self._episodes_length[self._idx - 1] and self._episodes_length[self._idx - 1] != 0:
            self._is_fixed_episode_length = False
        assert self._ep_len_dev is not None
        self._ep_len_dev[self._idx] = n
        self._collected_episodes += 1
        self._num_transitions += n
        self._idx = (self._idx + 1) % self._max_episodes
        self._full = self._full or self._idx == 0
        self._episodes_selection_probability = None

    def _layout(self, dims: tp.Mapping[str, int], extra: tp.Sequence[tp.Tuple[str, int]]) -> None:
        self._dims = {k: int(dims[k]) for k in HOT_FIELDS if k in dims}
        self._extra = [(k, int(d)) for k, d in extra]
        off = 0
        self._offsets = {}
        for name in ("observation", "action"):
            self._offsets[name] = off
            off += _round4(self._dims[name])
        self._offsets["reward"] = off
        self._offsets["discount"] = off + 1
        off += 4
        if "goal" in self._dims:
            self._offsets["goal"] = off
            off += _round4(self._dims["goal"])
        if self._extra:
            self._offsets["__extra__"] = off
            for name, d in self._extra:
                self._offsets[name] = off
                off += d
            off = _round4(off)
        self._row_stride = off

    def _allocate(self, ep: tp.Mapping[str, np.ndarray], rows: int) -> None:
        for name in ("observation", "action", "reward", "discount"):
            if name not in ep:
                raise KeyError(f"episode is missing the '{name}' field")
        self._layout({k: ep[k].shape[1] for k in HOT_FIELDS if k in ep},
                     [(k, v.shape[1]) for k, v in ep.items() if k not in self._batch_names])
        self._rows_per_episode = self._max_episode_length if self._max_episode_length is not None else rows
        self._rows = torch.zeros((self._max_episodes, self._rows_per_episode, self._row_stride), dtype=torch.float32,
                                 device=self.device)
        self._ep_len_dev = torch.zeros(self._max_episodes, dtype=torch.int32, device=self.device)
        self._staging = torch.zeros((2, self._rows_per_episode, self._row_stride), dtype=torch.float32).pin_memory()
        self._staging_events = [None, None]
        self._version += 1

    def _pack_host(self, ep: tp.Mapping[str, np.ndarray], out: np.ndarray) -> None:
        out[...] = 0
        for name in self._offsets:
            if name == "__extra__":
                continue
            v = ep[name]
            o = self._offsets[name]
            out[:len(v), o:o + v.shape[1]] = v

    def _upload_episode(self, ep: tp.Mapping[str, np.ndarray], slot: int, rows: int) -> None:
        assert self._rows is not None and self._staging is not None
        k = self._staging_slot
        self._staging_slot ^= 1
        ev = self._staging_events[k]
        if ev is not None:
            ev.synchronize()   # only the copy that last read THIS staging block (two episodes ago); the stream itself is never drained
        stage = self._staging[k, :rows]
        self._pack_host(ep, stage.numpy())
        self._rows[slot, :rows].copy_(stage, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._staging_events[k] = ev
        self._content += 1

    def load_storage(self, storage: tp.Mapping[str, tp.Any], episodes_length: tp.Optional[np.ndarray] = None,
                     n_episodes: tp.Optional[int] = None) -> None:
        """Bulk ingest of a reference-format `_storage` dict: name -> `[E, T+1, dim]` (numpy or torch, host or device)."""
        arrs = {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, dtype=np.float32))) for k, v in storage.items()}
        arrs = {k: (v.reshape(v.shape[0], v.shape[1], -1)).float() for k, v in arrs.items()}
        E, R = arrs["discount"].shape[:2]
        self._max_episodes = max(self._max_episodes, E) if self._rows is None else self._max_episodes
        if E > self._max_episodes:
            raise ValueError("storage holds more episodes than max_episodes")
        self._layout({k: arrs[k].shape[2] for k in HOT_FIELDS if k in arrs},
                     [(k, v.shape[2]) for k, v in arrs.items() if k not in self._batch_names])
        self._rows_per_episode = R
        self._rows = torch.zeros((self._max_episodes, R, self._row_stride), dtype=torch.float32, device=self.device)
        for name, o in self._offsets.items():
            if name == "__extra__":
                continue
            v = arrs[name]
            self._rows[:E, :, o:o + v.shape[2]] = v.to(self.device)
        self._host = {k: np.asarray(v.cpu().numpy()) for k, v in arrs.items() if k not in self._offsets}
        n = E if n_episodes is None else n_episodes
        if len(self._episodes_length) != self._max_episodes:
            self._episodes_length = np.zeros(self._max_episodes, dtype=np.int32)
        if episodes_length is None:
            self._episodes_length[:n] = R - 1
            self._episodes_length[n:] = 0
        else:
            self._episodes_length[:len(episodes_length)] = np.asarray(episodes_length, dtype=np.int32)
        lens = self._episodes_length[:n]
        self._is_fixed_episode_length = bool(len(lens) == 0 or lens.min() == lens.max())
        self._ep_len_dev = torch.as_tensor(self._episodes_length, dtype=torch.int32, device=self.device)
        self._staging = torch.zeros((2, R, self._row_stride), dtype=torch.float32).pin_memory()
        self._staging_events = [None, None]
        self._content += 1
        self._idx = n % self._max_episodes
        self._full = n >= self._max_episodes
        self._collected_episodes = n
        self._num_transitions = int(self._episodes_length.sum())
        self._episodes_selection_probability = None
        self._version += 1

    # -- the reference's `_storage` dict, materialised from HBM on demand (checkpoints, drivers that peek) ------
    @property
    def _storage(self) -> tp.Dict[str, np.ndarray]:
        out: tp.Dict[str, np.ndarray] = collections.OrderedDict()
        if self._rows is None:
            return out
        if self._storage_cache is not None and self._storage_cache[0] == self._content:
            return self._storage_cache[1]   # drivers read it repeatedly (train_offline.py:95, demo/main.py:108): one D2H per content version
        rows = self._rows.cpu().numpy()
        widths = dict(self._dims, reward=1, discount=1, **dict(self._extra))
        for name, o in self._offsets.items():
            if name != "__extra__":
                out[name] = np.ascontiguousarray(rows[:, :, o:o + widths[name]])
        out.update(self._host)
        self._storage_cache = (self._content, out)
        return out

    @_storage.setter
    def _storage(self, storage: tp.Mapping[str, np.ndarray]) -> None:
        self.load_storage(storage, n_episodes=len(self) if len(self) else None)

    def __getstate__(self) -> tp.Dict[str, tp.Any]:
        state = {k: v for k, v in self.__dict__.items()
                 if k not in ("_rows", "_ep_len_dev", "_staging", "_staging_events", "_staging_slot", "_storage_cache", "_content",
                              "_host", "_device", "_version")}
        state["_storage"] = dict(self._storage)   # reference pickle layout (pretrain.py:437-449)
        return state

    def __setstate__(self, state: tp.Dict[str, tp.Any]) -> None:
        storage = state.pop("_storage", {})
        self.__dict__.update(state)
        self._device = None
        if not hasattr(self, "_batch_names"):
            self._batch_names = set(TIMESTEP_FIELDS)
        n, idx, full = len(self), self._idx, self._full
        lens = getattr(self, "_episodes_length", None)
        self._init_device_state()
        if storage:
            if lens is None:  # pickles older than variable-length support (in_memory_replay_buffer.py:95-102)
                lens = np.zeros(len(storage["discount"]), dtype=np.int32)
                lens[:n] = storage["discount"].shape[1] - 1
                self._max_episode_length = None
            self._episodes_length = np.asarray(lens, dtype=np.int32)
            self.load_storage(storage, episodes_length=self._episodes_length, n_episodes=n)
            self._idx, self._full = idx, full

    @classmethod
    def from_reference(cls, other: tp.Any, device: tp.Union[str, torch.device, None] = None) -> "ReplayBuffer":
        """Adopt a url_benchmark.in_memory_replay_buffer.ReplayBuffer instance (e.g. an ExORL pickle)."""
        d = other.__dict__
        buf = cls(d["_max_episodes"], d["_discount"], float(d["_future"]), d.get("_max_episode_length"), device=device)
        n = len(other)
        lens = d.get("_episodes_length")
        if lens is None or not np.any(lens):  # load()-filled buffers leave lengths at 0
            lens = np.zeros(d["_max_episodes"], dtype=np.int32)
            lens[:n] = d["_storage"]["discount"].shape[1] - 1
        buf._episodes_length = np.asarray(lens, dtype=np.int32).copy()
        buf.load_storage(d["_storage"], episodes_length=buf._episodes_length, n_episodes=n)
        buf._idx, buf._full = d["_idx"], d["_full"]
        return buf

    # -- sampling --------------------------------------------------------------------------------
    def draw_indices(self, batch_size: int) -> tp.Tuple[np.ndarray, np.ndarray, tp.Optional[np.ndarray]]:
        """The index draws of in_memory_replay_buffer.py:146-161 on the numpy GLOBAL RNG, in the same order."""
        return draw_sample_indices(self, batch_size)

    def view(self) -> "L.fb_replay_view":
        """The C-ABI description of the device storage (fb_replay_view, include/fb_b200.h)."""
        if self._rows is None or self._ep_len_dev is None:
            raise RuntimeError("the replay buffer is empty")
        o = self._offsets
        return L.fb_replay_view(d_rows=self._rows.data_ptr(), d_episode_len=self._ep_len_dev.data_ptr(),
                                max_episodes=self._max_episodes, rows_per_episode=self._rows_per_episode,
                                row_stride=self._row_stride, n_episodes=len(self), off_obs=o["observation"],
                                off_action=o["action"], off_reward=o["reward"], off_discount=o["discount"],
                                off_goal=o.get("goal", -1), off_extra=o.get("__extra__", -1),
                                goal_dim=self._dims.get("goal", 0), extra_dim=sum(d for _, d in self._extra))

    def gather(self, ep_idx: tp.Any, step_idx: tp.Any, future_idx: tp.Any = None) -> EpisodeBatch:
        """One gather kernel for given indices -> EpisodeBatch of device tensors (views of one packed block)."""
        lib = L.load()
        dev = self.device
        as_i32 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.int32).to(dev, non_blocking=True)  # noqa: E731
        d_ep, d_step = as_i32(ep_idx), as_i32(step_idx)
        d_fut = as_i32(future_idx) if future_idx is not None else None
        batch = d_ep.numel()
        O, A = self._dims["observation"], self._dims["action"]
        G, X = self._dims.get("goal", 0), sum(d for _, d in self._extra)
        offs = (C.c_int32 * 9)()
        pitch = C.c_int32()
        L.check(lib.fb_batch_row_layout(O, A, G, X, int(d_fut is not None), offs, C.byref(pitch)))
        out = torch.empty((batch, pitch.value), dtype=torch.float32, device=dev)
        view = self.view()
        with torch.cuda.device(dev):
            L.check(lib.fb_replay_gather(C.byref(view), O, A, d_ep.data_ptr(), d_step.data_ptr(),
                                         d_fut.data_ptr() if d_fut is not None else None, batch, float(self._discount),
                                         out.data_ptr(), pitch.value, torch.cuda.current_stream(dev).cuda_stream),
                    "fb_replay_gather")
        o_obs, o_act, o_rd, o_nobs, o_goal, o_ngoal, o_extra, o_fobs, o_fgoal = list(offs)
        meta: tp.Dict[str, torch.Tensor] = {}
        x = o_extra
        for name, dim in self._extra:
            meta[name] = out[:, x:x + dim]
            x += dim
        return EpisodeBatch(obs=out[:, o_obs:o_obs + O], action=out[:, o_act:o_act + A], reward=out[:, o_rd:o_rd + 1],
                            discount=out[:, o_rd + 1:o_rd + 2], next_obs=out[:, o_nobs:o_nobs + O], meta=meta,
                            goal=out[:, o_goal:o_goal + G] if G else None,
                            next_goal=out[:, o_ngoal:o_ngoal + G] if G else None,
                            future_obs=out[:, o_fobs:o_fobs + O] if d_fut is not None else None,
                            future_goal=out[:, o_fgoal:o_fgoal + G] if (G and d_fut is not None) else None)

    def sample(self, batch_size: int, custom_reward: tp.Optional[tp.Any] = None, with_physics: bool = False) -> EpisodeBatch:
        """in_memory_replay_buffer.py:139-190: host index draws (numpy global RNG, reference order) + one gather kernel.
        The batch already lives on the device; `.to(device)` on it is a no-op."""
        ep_idx, step_idx, future_idx = self.draw_indices(batch_size)
        batch = self.gather(ep_idx, step_idx, future_idx)
        if custom_reward is not None or with_physics:
            phy = self._host["physics"][ep_idx, step_idx]
            if custom_reward is not None:
                reward = np.array([[custom_reward.from_physics(p)] for p in phy], dtype=np.float32)
                batch = dataclasses.replace(batch, reward=torch.as_tensor(reward, device=self.device))
            if with_physics:
                batch = dataclasses.replace(batch, _physics=torch.as_tensor(phy, device=self.device))
        return batch

    def load(self, env: tp.Any, replay_dir: tp.Any, relabel: bool = True, goal_func: tp.Any = None) -> None:
        """in_memory_replay_buffer.py:192-208: fill the buffer from a directory of per-episode `.npz` files (sorted by name) until it
        is full.  Episodes are committed through add_episode(), so — unlike the reference, which leaves `_episodes_length` at 0 after
        load() — the lengths, the fixed-length flag and `avg_episode_length` are correct afterwards.
        `relabel=True` recomputes rewards (and goals) through `env.physics` / `env.task` (relabel_episode above, the contract of
        in_memory_replay_buffer.py:40-55)."""
        import pathlib
        for eps_fn in sorted(pathlib.Path(replay_dir).glob("*.npz")):
            if self._full:
                break
            episode = load_episode(eps_fn)
            if relabel:
                episode = relabel_episode(env, episode, goal_func)   # MuJoCo state replay: stays on the host, untouched
            self.add_episode(episode)

    def relabel(self, custom_reward: tp.Any) -> None:
        """in_memory_replay_buffer.py:210-216: recompute rewards from stored physics (host loop, as in the reference)."""
        assert self._rows is not None
        o = self._offsets["reward"]
        physics = self._host["physics"]
        reward = np.array([[custom_reward.from_physics(p) for p in phy] for phy in physics], dtype=np.float32)   # [E, T+1]
        self._rows[:len(physics), :, o] = torch.as_tensor(reward).to(self.device, non_blocking=False)
        self._content += 1
        self._max_episodes = len(physics)
        self._full = True
