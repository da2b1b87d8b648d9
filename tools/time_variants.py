#!/usr/bin/env python
"""Development helper: times wt_eval_batch for (grid_len, waves, flags) combinations with CUDA events.
    python tools/time_variants.py 4096:64:shade,fast 4096:64:fast 1024:8:
Not a bench: bench.py is the measured contract."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wave_tool_b200 import api, scenes


def run(spec, steps=20, warm=3):
    parts = spec.split(":")
    L, W, fl = parts[:3]
    sky_n = int(parts[3]) if len(parts) > 3 else 2048
    L, W = int(L), int(W)
    names = [f for f in fl.split(",") if f]
    flags = 0
    for f in names:
        flags |= {"shade": api.FLAG_SHADE, "fast": api.FLAG_FAST_PHASE, "ref": api.FLAG_REFERENCE_ORDER, "stencil": api.FLAG_STENCIL_ORDER, "auto": api.FLAG_AUTO_PHASE, "cks": api.FLAG_CHECKSUM}[f]
    waves = scenes.seeded_waves(W, cfg=3) if W else scenes.seeded_waves(1, cfg=3)[:0]
    ctx = api.Context(0)
    ctx.set_waves(waves)
    ctx.set_heightmap(scenes.procedural_heightmap(512))
    if flags & api.FLAG_SHADE:
        ctx.set_skybox(scenes.procedural_skybox(sky_n))
    n = L * L
    dev = torch.device("cuda", 0)
    pos = torch.empty((2 * n, 4), dtype=torch.float32, device=dev)
    nrm = torch.empty((2 * n, 4), dtype=torch.float32, device=dev)
    shd = torch.empty((2 * n, 4), dtype=torch.float32, device=dev) if flags & api.FLAG_SHADE else None
    cks = torch.zeros((steps + warm) * 40, dtype=torch.uint8, device=dev)
    frames = scenes.frames_for_animation(scenes.camera(), waves, first=0, count=steps + warm)
    sp = torch.cuda.current_stream().cuda_stream
    arr = (api.Frame * len(frames))(*frames)

    def launch(lo, hi):
        sub = (api.Frame * (hi - lo))(*frames[lo:hi])
        ctx.eval_batch(sub, L, pos.data_ptr(), nrm.data_ptr(), shd.data_ptr() if shd is not None else 0, 2,
                       cks.data_ptr() if flags & api.FLAG_CHECKSUM else 0, flags=flags, stream=sp)
    launch(0, warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch(warm, warm + steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{spec:28s} {ms*1e3:9.1f} us/frame  {n*max(W,1)/ms/1e6:9.1f} Gvw/s  {n*(48 if flags&1 else 32)/ms/1e6:8.1f} GB/s out", flush=True)


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        run(spec)
