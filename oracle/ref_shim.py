"""TEST INFRASTRUCTURE ONLY — import the *unmodified* reference under a sys.modules shim.

The reference (`/root/reference/url_benchmark/agent/fb_ddpg.py` and friends) imports hydra,
omegaconf, dm_env, dm_control and `url_benchmark.dmc`, none of which are installed in this image
(and `url_benchmark/dmc.py:41` does not even import on Python >= 3.11).  This module pre-seeds
`sys.modules` with five tiny stand-ins so that the reference's own files

    url_benchmark/agent/fb_ddpg.py, fb_modules.py, ddpg.py,
    url_benchmark/utils.py, in_memory_replay_buffer.py, replay_buffer.py

import and run byte-for-byte.  It is used by `oracle/make_golden.py` to
produce the committed fixtures under `tests/golden/` and by the container-only tests that pin the
oracle restatement against the live reference.  `/root/reference` does not exist on the GPU box:
nothing that runs there may import this module (it raises if the tree is missing).
"""
from __future__ import annotations

import collections
import dataclasses
import enum
import os
import sys
import types
import typing as tp

import numpy as np

REFERENCE_ROOT = os.environ.get("FB_REFERENCE_ROOT", "/root/reference")

# goal-space sizes the reference obtains by instantiating a MuJoCo env (goals.py:218-221)
GOAL_SPACE_DIMS = {
    "simplified_walker": 3,
    "walker_pos_speed": 4,
    "walker_pos_speed_z": 6,
    "simplified_quadruped": 2,
    "quad_pos_speed": 7,
    "simplified_jaco": 3,
    "simplified_point_mass_maze": 2,
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "url_benchmark", "agent", "fb_ddpg.py"))


def _module(name: str, **attrs: tp.Any) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install() -> None:
    """Idempotently install the stubs and put the reference tree on sys.path."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}; the shim only works in the build container")
    if "url_benchmark.agent.fb_ddpg" in sys.modules:
        return

    # --- omegaconf -----------------------------------------------------------------------------
    if "omegaconf" not in sys.modules:
        _module("omegaconf", MISSING="???", II=lambda s: "${%s}" % s, SI=lambda s: s,
                DictConfig=dict, OmegaConf=types.SimpleNamespace())

    # --- hydra ---------------------------------------------------------------------------------
    if "hydra" not in sys.modules:
        class _ConfigStore:
            _inst: tp.Optional["_ConfigStore"] = None

            def __init__(self) -> None:
                self.repo: tp.Dict[tp.Tuple[tp.Optional[str], str], tp.Any] = {}

            @classmethod
            def instance(cls) -> "_ConfigStore":
                if cls._inst is None:
                    cls._inst = cls()
                return cls._inst

            def store(self, name: str, node: tp.Any, group: tp.Optional[str] = None, **_: tp.Any) -> None:
                self.repo[(group, name)] = node

        hydra = _module("hydra", main=lambda *a, **k: (lambda f: f))
        core = _module("hydra.core")
        cstore = _module("hydra.core.config_store", ConfigStore=_ConfigStore)
        hutils = _module("hydra.utils", instantiate=None)
        hydra.core, hydra.utils, core.config_store = core, hutils, cstore

    # --- dm_env --------------------------------------------------------------------------------
    if "dm_env" not in sys.modules:
        class StepType(enum.IntEnum):
            FIRST = 0
            MID = 1
            LAST = 2

        class _Array:
            def __init__(self, shape: tp.Any, dtype: tp.Any, name: tp.Optional[str] = None) -> None:
                self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name

        class _Bounded(_Array):
            def __init__(self, shape: tp.Any, dtype: tp.Any, minimum: tp.Any, maximum: tp.Any,
                         name: tp.Optional[str] = None) -> None:
                super().__init__(shape, dtype, name)
                self.minimum, self.maximum = minimum, maximum

        specs = _module("dm_env.specs", Array=_Array, BoundedArray=_Bounded, DiscreteArray=_Array)
        _module("dm_env", StepType=StepType, specs=specs,
                TimeStep=collections.namedtuple("TimeStep", "step_type reward discount observation"),
                Environment=object)

    # --- url_benchmark package skeleton (skip the real __init__ side effects) --------------------
    ub_root = os.path.join(REFERENCE_ROOT, "url_benchmark")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ub = sys.modules.get("url_benchmark") or _module("url_benchmark")
    ub.__path__ = [ub_root]  # type: ignore[attr-defined]

    # url_benchmark.dmc: the four TimeStep dataclasses of dmc.py:35-73 (field names are what the
    # replay buffer reads, in_memory_replay_buffer.py:82), with a default_factory for `physics`.
    StepType = sys.modules["dm_env"].StepType

    @dataclasses.dataclass
    class TimeStep:
        step_type: tp.Any
        reward: float
        discount: float
        observation: np.ndarray
        physics: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros((0,), np.float32), init=False)

        def first(self) -> bool:
            return self.step_type == StepType.FIRST

        def mid(self) -> bool:
            return self.step_type == StepType.MID

        def last(self) -> bool:
            return self.step_type == StepType.LAST

        def __getitem__(self, attr: str) -> tp.Any:
            return getattr(self, attr)

    @dataclasses.dataclass
    class GoalTimeStep(TimeStep):
        goal: np.ndarray

    @dataclasses.dataclass
    class ExtendedGoalTimeStep(GoalTimeStep):
        action: tp.Any

    @dataclasses.dataclass
    class ExtendedTimeStep(TimeStep):
        action: tp.Any

    dmc = _module("url_benchmark.dmc", TimeStep=TimeStep, GoalTimeStep=GoalTimeStep,
                  ExtendedGoalTimeStep=ExtendedGoalTimeStep, ExtendedTimeStep=ExtendedTimeStep,
                  EnvWrapper=object)
    ub.dmc = dmc  # type: ignore[attr-defined]

    goals = _module("url_benchmark.goals", get_goal_space_dim=lambda name: GOAL_SPACE_DIMS[name])
    ub.goals = goals  # type: ignore[attr-defined]

    # bare `url_benchmark.agent` so that agent/__init__.py (which imports all 18 agents) is skipped
    agent_pkg = _module("url_benchmark.agent")
    agent_pkg.__path__ = [os.path.join(ub_root, "agent")]  # type: ignore[attr-defined]
    ub.agent = agent_pkg  # type: ignore[attr-defined]


def load() -> types.SimpleNamespace:
    """Return the reference symbols the oracle / golden generator needs."""
    install()
    import importlib

    fb_ddpg = importlib.import_module("url_benchmark.agent.fb_ddpg")
    fb_modules = importlib.import_module("url_benchmark.agent.fb_modules")
    utils = importlib.import_module("url_benchmark.utils")
    imrb = importlib.import_module("url_benchmark.in_memory_replay_buffer")
    rb = importlib.import_module("url_benchmark.replay_buffer")
    dmc = sys.modules["url_benchmark.dmc"]
    return types.SimpleNamespace(
        fb_ddpg=fb_ddpg, fb_modules=fb_modules, utils=utils, in_memory_replay_buffer=imrb,
        replay_buffer=rb, dmc=dmc, StepType=sys.modules["dm_env"].StepType,
        FBDDPGAgent=fb_ddpg.FBDDPGAgent, FBDDPGAgentConfig=fb_ddpg.FBDDPGAgentConfig,
        ReplayBuffer=imrb.ReplayBuffer, EpisodeBatch=rb.EpisodeBatch)
