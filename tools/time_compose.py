#!/usr/bin/env python
"""Development helper (GPU): times wt_compose_frame (the screen-space layers pass) on a 4096^2 grid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from wave_tool_b200 import api, scenes

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W, H = 1920, 1080
waves = scenes.seeded_waves(64, cfg=3)
ctx = api.Context(0)
ctx.set_waves(waves)
ctx.set_heightmap(scenes.procedural_heightmap(512))
ctx.set_skybox(scenes.procedural_skybox(2048))
rng = np.random.default_rng(0)
ctx.set_screen_textures(rng.uniform(0.9, 1.0, (H, W)).astype(np.float32), rng.integers(0, 256, (H, W, 4), dtype=np.uint8), rng.integers(0, 256, (H, W, 4), dtype=np.uint8))
n = L * L
b = [torch.empty((n, 4), dtype=torch.float32, device="cuda") for _ in range(7)]
out = api.Outputs(*[t.data_ptr() for t in b[:6]], None)
cam = scenes.camera(aspect=W / H)
st, frame = api.build_frame(cam, scenes.animation_params(75), api.total_amplitude(waves))
sp = torch.cuda.current_stream().cuda_stream
cp = api.default_compose_params(cam, W, H)
for flags, name in ((api.FLAG_SHADE | api.FLAG_VARYINGS | api.FLAG_FAST_PHASE, "vertex stage + varyings"),):
    for _ in range(3):
        ctx.eval_frame_ex(frame, L, out, flags=flags, stream=sp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ctx.eval_frame_ex(frame, L, out, flags=flags, stream=sp)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/10:.4f} ms per {L}^2 frame ({n*96/ (e0.elapsed_time(e1)/10*1e-3)/1e9:.0f} GB/s written)")
for _ in range(3):
    ctx.compose_frame(frame, cp, L, out, b[6].data_ptr(), stream=sp)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ctx.compose_frame(frame, cp, L, out, b[6].data_ptr(), stream=sp)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"compose pass: {ms:.4f} ms per {L}^2 frame, {n*112/(ms*1e-3)/1e9:.0f} GB/s of algorithmic traffic (96 B read + 16 B written per vertex)")
lean = api.Outputs(b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), None, None, None, None)
for _ in range(3):
    ctx.compose_frame(frame, cp, L, lean, b[6].data_ptr(), stream=sp)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ctx.compose_frame(frame, cp, L, lean, b[6].data_ptr(), stream=sp)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"compose pass, varyings recomputed: {ms:.4f} ms per {L}^2 frame, {n*64/(ms*1e-3)/1e9:.0f} GB/s of algorithmic traffic (48 B read + 16 B written per vertex)")
