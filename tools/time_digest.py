import os, sys
sys.path.insert(0, "/root/repo")
import torch
from wave_tool_b200 import api, scenes
sys.path.insert(0, "/root/repo/tools")
from explore_sched import setup, time_it
ctx, waves, bufs = setup(4096, 64, True)
for flags, name in ((api.FLAG_SHADE | api.FLAG_FAST_PHASE, "fast"), (api.FLAG_SHADE, "exact")):
    ms0 = min(time_it(ctx, waves, bufs, 4096, flags) for _ in range(2))
    cks = torch.zeros(40 * 64, dtype=torch.uint8, device="cuda")
    ms1 = min(time_it(ctx, waves, bufs, 4096, flags, cks=cks) for _ in range(2))
    print(name, "no digest %.4f ms, with digest %.4f ms (+%.1f%%)" % (ms0, ms1, 100 * (ms1 / ms0 - 1)))
