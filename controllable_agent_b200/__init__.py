"""B200-native FB-DDPG update path (drop-in for url_benchmark's agent=fb_ddpg); see DESIGN.md / INTEGRATION.md."""
from . import _lib  # noqa: F401
from .agent import FBDDPGAgent, FBDDPGAgentConfig  # noqa: F401
from .engine import EngineConfig, FBStepEngine  # noqa: F401
from .replay import EpisodeBatch, ReplayBuffer  # noqa: F401
