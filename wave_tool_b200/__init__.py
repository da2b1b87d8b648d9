"""wave_tool_b200 — B200-native (sm_100a) evaluator for the per-vertex ocean-surface path of
TexelBox/wave-tool.  The product is `libwavetool_b200.so` (C ABI in include/wavetool_b200.h);
this package holds its sources (csrc/), the in-tree build, a ctypes view for the harness
(api.py) and the synthetic workloads (scenes.py).

(The directory is spelled with an underscore because a hyphenated name cannot be imported.)
"""
from . import api, build, scenes  # noqa: F401

__all__ = ["api", "build", "scenes"]
