#!/usr/bin/env python
"""Development helper (GPU): where the end-to-end step of the bench workload (cfg3) spends its time — the kernel alone, the kernel
with its digest, and the host-driven step (wt_step_submit / wt_step_wait) at 1..4 frames in flight."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.common import Scene, reference_heightmap
from wave_tool_b200 import api, scenes

L, W = 4096, 64
waves = scenes.seeded_waves(W, cfg=3)
cam = scenes.camera()
sc = Scene(waves, cam, scenes.animation_params(3), L, reference_heightmap(), skybox=scenes.procedural_skybox(2048))
ctx = sc.make_context(0)
n = L * L
ring = 2
pos, nrm, shd = (torch.empty((ring * n, 4), dtype=torch.float32, device="cuda") for _ in range(3))
cks = torch.zeros(16, dtype=torch.int64, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
params = [scenes.animation_params(3 + k) for k in range(200)]
frames = [Scene(waves, cam, p, L, sc.heightmap).frame()[1] for p in params]
amp = api.total_amplitude(waves)
K = 200


def timed(fn, label):
    fn(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    fn(K)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / K * 1e3
    print(f"{label:60s} device {e0.elapsed_time(e1) / K:8.4f} ms/frame   wall {wall:8.4f} ms/frame", flush=True)


for name, fl in (("bit-faithful", 0), ("auto", api.FLAG_AUTO_PHASE)):
    f = api.FLAG_SHADE | fl
    timed(lambda c: [ctx.eval_frame(frames[k % K], L, pos.data_ptr(), nrm.data_ptr(), shd.data_ptr(), flags=f, stream=sp) for k in range(c)], f"{name}: kernel")
    timed(lambda c: [ctx.eval_frame(frames[k % K], L, pos.data_ptr(), nrm.data_ptr(), shd.data_ptr(), d_checksum=cks.data_ptr(), flags=f | api.FLAG_CHECKSUM, stream=sp)
                     for k in range(c)], f"{name}: kernel + digest (device record)")
    out = api.Checksum()
    for depth in (1, 2, 3, 4):
        def run(c, depth=depth):
            pending = []
            for k in range(c):
                slot = k % ring
                pending.append(ctx.step_submit(cam, params[k % K], amp, L, pos.data_ptr() + slot * n * 16, nrm.data_ptr() + slot * n * 16, shd.data_ptr() + slot * n * 16, flags=f, stream=sp))
                if len(pending) >= depth:
                    ctx.step_wait(pending.pop(0), out=out)
            for t in pending:
                ctx.step_wait(t, out=out)
        timed(run, f"{name}: wt_step_submit / wt_step_wait, {depth} in flight")
    t0 = time.perf_counter()
    for k in range(K):
        api.build_frame(cam, params[k], amp) if hasattr(api, "build_frame") else None
    print(f"{name}: host prologue alone {(time.perf_counter() - t0) / K * 1e6:.1f} us/frame")
