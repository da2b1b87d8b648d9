#!/usr/bin/env python
"""Development helper (GPU): times the production kernel under the experiment knobs
(WT_DYN = rows per dynamically queued work item, -1 = the library's policy; WT_DYN_WARPS = resident warps per
SM; WT_LIB = a tuning variant built by tools/build_variant.py) and can dump the per-work-item trace (WT_TRACE).
Not a bench.
    python tools/explore_sched.py --mode fast --configs -1:0,0:0,22:0 [--trace 22:0] [--tag base]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from wave_tool_b200 import api, scenes


def setup(L, W, shade, cfg=3):
    waves = scenes.seeded_waves(W, cfg=cfg)
    ctx = api.Context(0)
    ctx.set_waves(waves)
    ctx.set_heightmap(scenes.procedural_heightmap(512))
    if shade:
        ctx.set_skybox(scenes.procedural_skybox(2048))
    n = L * L
    dev = torch.device("cuda", 0)
    bufs = [torch.empty((2 * n, 4), dtype=torch.float32, device=dev) for _ in range(3 if shade else 2)]
    return ctx, waves, bufs


def time_it(ctx, waves, bufs, L, flags, steps=20, warm=3, cks=None):
    frames = scenes.frames_for_animation(scenes.camera(), waves, first=0, count=steps + warm)
    sp = torch.cuda.current_stream().cuda_stream
    shade = len(bufs) == 3

    def launch(lo, hi):
        ctx.eval_batch(frames[lo:hi], L, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr() if shade else 0, 2,
                       cks.data_ptr() if cks is not None else 0, flags=flags | (api.FLAG_CHECKSUM if cks is not None else 0), stream=sp)
    if warm:
        launch(0, warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch(warm, warm + steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def set_env(dyn, warps, trace=0):
    """dyn < 0 leaves the library's own policy in place"""
    if dyn >= 0:
        os.environ["WT_DYN"] = str(dyn)
    else:
        os.environ.pop("WT_DYN", None)
    os.environ["WT_TRACE"] = str(trace)
    if warps:
        os.environ["WT_DYN_WARPS"] = str(warps)
    else:
        os.environ.pop("WT_DYN_WARPS", None)
    api.lib().wt_debug_reload_env()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sched")
    ap.add_argument("--tag", default="base")
    ap.add_argument("--L", type=int, default=4096)
    ap.add_argument("--W", type=int, default=64)
    ap.add_argument("--mode", default="fast,exact")
    ap.add_argument("--noshade", action="store_true")
    ap.add_argument("--configs", default="-1:0")
    ap.add_argument("--trace", default="")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    L, W = args.L, args.W
    ctx, waves, bufs = setup(L, W, not args.noshade)
    for mode in args.mode.split(","):
        flags = (0 if args.noshade else api.FLAG_SHADE) | (api.FLAG_FAST_PHASE if mode == "fast" else 0)
        for c in args.configs.split(","):
            dyn, warps = (int(x) for x in c.split(":")[:2])
            set_env(dyn, warps)
            ms = min(time_it(ctx, waves, bufs, L, flags) for _ in range(2))
            cks = torch.zeros(40 * 4, dtype=torch.uint8, device="cuda")
            time_it(ctx, waves, bufs, L, flags, steps=1, warm=0, cks=cks)
            d = cks.cpu().numpy().view(api.CHECKSUM_DTYPE)[0]
            rec = {"tag": args.tag, "mode": mode, "L": L, "W": W, "dyn_rows": dyn, "warps_cap": warps, "ms": round(ms, 4),
                   "pos_bits": int(d["position_bits"]), "nrm_bits": int(d["normal_bits"])}
            print(json.dumps(rec), flush=True)
        for c in [x for x in args.trace.split(",") if x]:
            dyn, warps = (int(x) for x in c.split(":")[:2])
            set_env(dyn, warps, trace=1)
            time_it(ctx, waves, bufs, L, flags, steps=2, warm=1)
            np.save(f"{args.out}_trace_{args.tag}_{mode}_{dyn}_{warps}.npy", ctx.read_trace())
            set_env(dyn, warps)


if __name__ == "__main__":
    main()
