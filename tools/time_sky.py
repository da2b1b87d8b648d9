#!/usr/bin/env python
"""Development helper (GPU): times wt_synthesize_skybox at the reference's CUBEMAP_LENGTH (2048)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from wave_tool_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
src = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rng = np.random.default_rng(0)
ctx = api.Context(0)
grad = rng.integers(0, 256, (4096, 4), dtype=np.uint8)
stars = rng.integers(0, 256, (6, src, src, 4), dtype=np.uint8)
clouds = rng.integers(0, 256, (6, src, src, 4), dtype=np.uint8)
ctx.set_sky_layers(grad, stars, clouds)
p = api.default_sky_params()
sp = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.synthesize_skybox(p, n, stream=sp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(20):
    p.timeOfDayInHours = 6.0 + 0.5 * k
    ctx.synthesize_skybox(p, n, stream=sp)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
out_b = 6 * n * n * 4
print("texels recomputed on the exact path in the last synthesis:", ctx.sky_exact_texels(), "of", 6 * n * n)
print(f"skybox synthesis {n}^2 faces from {src}^2 star/cloud cubemaps: {ms*1e3:.1f} us per cubemap (synthesis + atlas), "
      f"{(2*out_b + 6*(n+2)*(n+2)*4 + 2*6*src*src*4)/(ms*1e-3)/1e9:.0f} GB/s of algorithmic traffic")
