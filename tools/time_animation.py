#!/usr/bin/env python
"""Development helper (GPU): times wt_run_animation (native driver) for a small-frame animation, e.g. cfg2.
    python tools/time_animation.py [L] [W] [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wave_tool_b200 import api, scenes

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
waves = scenes.seeded_waves(W, cfg=2)
ctx = api.Context(0)
ctx.set_waves(waves)
ctx.set_heightmap(scenes.procedural_heightmap(512))
n = L * L
RING = int(os.environ.get("RING", "4"))
pos = torch.empty((RING * n, 4), dtype=torch.float32, device="cuda")
nrm = torch.empty((RING * n, 4), dtype=torch.float32, device="cuda")
cks = torch.zeros(40 * N, dtype=torch.uint8, device="cuda")
anim = api.default_animation()
total = api.total_amplitude(waves)
sp = torch.cuda.current_stream().cuda_stream
for fan in (1, 2, 4):
  os.environ["WT_FANOUT"] = str(fan)
  api.lib().wt_debug_reload_env()
  for flags, name in ((api.FLAG_FAST_PHASE, "recurrence"), (0, "exact"), (api.FLAG_FAST_PHASE | api.FLAG_CHECKSUM, "recurrence+digest")):
      name = f"fanout {fan} " + name
      for rep in range(2):
          p = api.default_params()
          torch.cuda.synchronize()
          t0 = time.perf_counter()
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          e0.record()
          drawn = ctx.run_animation(scenes.camera(), p, anim, total, 1.0 / 60.0, N, L, pos.data_ptr(), nrm.data_ptr(), 0, RING,
                                    cks.data_ptr() if flags & api.FLAG_CHECKSUM else 0, flags=flags, stream=sp)
          t_host = time.perf_counter() - t0
          e1.record()
          torch.cuda.synchronize()
          t_all = time.perf_counter() - t0
      print(f"{name:20s} {L}^2 x {W} waves, {N} frames: host enqueue {t_host/N*1e6:7.2f} us/frame, device {e0.elapsed_time(e1)/N*1e3:7.2f} us/frame, "
            f"wall {t_all/N*1e6:7.2f} us/frame -> {n*W*N/t_all/1e9:8.1f} Gvw/s, {n*32*N/t_all/1e9:7.1f} GB/s", flush=True)
