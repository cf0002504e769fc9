"""B200-native FB-DDPG update path (drop-in for url_benchmark's agent=fb_ddpg); see DESIGN.md."""
from . import _lib  # noqa: F401
