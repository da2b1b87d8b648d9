#!/usr/bin/env python
"""bench.py — vertex-waves/s of the fused water-surface kernel (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one frame of the workload per GPU (frame sharding: global frame f runs on rank f mod N,
no data-path collective; weak scaling).  Default workload = BASELINE.json configs[2], the one the
metric ("ms per 4096^2 grid frame", ">=70% of roofline at a 4096^2 grid with 64 waves") is quoted on:
4096x4096 projected grid, 64 seeded Gerstner waves + heightmap detail, finite-difference normals and the
Fresnel / sky-cubemap / sun-specular terms, default camera.  Synthetic data (seeded waves, procedural
heightmap and cubemap).  Rank 0 prints ONE JSON line.

Timing: W >= 3 untimed warm-up steps, then EXACTLY K steps bracketed by barrier + synchronize, timed
with CUDA events on the stream the kernels are launched on, max over ranks.  Every step writes
805 MB of fresh output (>> the 126 MB L2) into a 2-deep ring, so no step can be served from L2.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (grid_len, n_waves, cfg seed, shade, description)
    "cfg3": (4096, 64, 3, True, "configs[2]: 4096x4096 grid, 64 waves, normals + Fresnel/sky-cubemap/sun-specular terms"),
    "cfg2": (1024, 8, 2, False, "configs[1]: 1024x1024 projected grid, 8 Gerstner waves + heightmap noise, animation frames"),
    "cfg4": (8192, 256, 4, False, "configs[3]: 8192x8192 grid, 256 waves, frame batch sharded by frame"),
    "cfg5": (16384, 32, 5, False, "configs[4]: 16384x16384 grid, 32 waves, sharded by grid rows (strong scaling)"),
}
ROW_SHARDED = {"cfg5"}
TRANSCENDENTALS_PER_VERTEX_WAVE = 2      # 1 sin + 1 cos (water-grid.vert:78-81), SURVEY.md §8d
FLOP_PER_VERTEX_WAVE = 15
_REAL_STDOUT = sys.stdout
SKYBOX_N = 2048                           # reference: 2048^2 faces (render-engine.h:191)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05] or self.rows
        for _, line in rows:
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(L, W, cfg, seconds_target=12.0, max_seconds=40.0):
    """The reference's own CPU implementation of the path on the host cores: the reference
    water-grid.vert shader text compiled as C++ (oracle/_ref, kind 'reference') when it was built,
    else the oracle port (kind 'port').  Bounded sample: bands of whole rows of consecutive animation
    frames of the SAME workload until about `seconds_target` seconds of wall time have been spent."""
    from oracle import oracle as orc                      # cpu_baseline leg: the one place bench.py may use oracle/
    from tests.common import Scene
    from wave_tool_b200 import scenes
    global _CPU_ARM
    if "_CPU_ARM" not in globals():
        _CPU_ARM = (orc.Oracle(), orc.Ref(), scenes.procedural_heightmap(512))     # built once: the sample times the path, not the setup
    O, R, hm = _CPU_ARM
    # all the host threads this process may use — torchrun exports OMP_NUM_THREADS=1, which would otherwise
    # pin the reference arm to one core (the oracle and oracle/_ref share one OpenMP runtime)
    try:
        O.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        O.set_num_threads(os.cpu_count() or 1)
    waves = scenes.seeded_waves(W, cfg=cfg)
    use_ref = R.available and L <= 4096 and W <= R.max_waves()
    cores = O.num_threads()
    rows_total, t_spent, chunk, frame, row = 0, 0.0, 8, 75, 0
    sc = Scene(waves, scenes.camera(), scenes.animation_params(frame), L, hm)
    st, u, wu = sc.oracle_uniforms(O)
    while t_spent < seconds_target:
        hi = min(row + chunk, L)
        t0 = time.perf_counter()
        if use_ref:
            R.vert_run(u, wu, hm, id_range=(row * L, hi * L))
        else:
            O.eval(u, wu, hm, rows=(row, hi))
        dt = time.perf_counter() - t0
        t_spent += dt
        rows_total += hi - row
        row = hi
        if row >= L:                                      # next animation frame
            frame, row = frame + 1, 0
            sc = Scene(waves, scenes.camera(), scenes.animation_params(frame), L, hm)
            st, u, wu = sc.oracle_uniforms(O)
        if t_spent + 2 * dt > max_seconds:
            break
        if dt < 1.0:
            chunk = min(chunk * 2, 512)
    vw = rows_total * L * W
    return {"value": vw / t_spent, "unit": "vertex-waves/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{rows_total} rows ({rows_total / L:.2f} frames) of the {L}x{L} grid with {W} waves, vertex stage "
                      f"(positions + 4-neighbour normals: 5 evaluations per vertex as the shader does), {t_spent:.1f} s wall, "
                      f"OpenMP over vertices on {cores} threads, g++ -O2 -ffp-contract=off",
            "seconds": t_spent, "vertex_waves": vw}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L, W, cfg, shade, desc = WORKLOADS[args.workload]
    vals = []
    per_step = max(0.2, min(12.0, 60.0 / max(1, args.steps + args.warmup)))   # the whole run stays near a minute whatever K
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_reference_rate(L, W, cfg, seconds_target=per_step, max_seconds=per_step * 2.5)
        if i >= args.warmup:
            vals.append(base)
    total_vw = sum(b["vertex_waves"] for b in vals)
    total_s = sum(b["seconds"] for b in vals)
    value = total_vw / total_s
    cb = dict(base)
    cb["value"] = value
    cb.pop("seconds", None); cb.pop("vertex_waves", None)
    line = {"impl": "reference", "metric": "vertex_waves_per_sec", "value": value, "unit": "vertex-waves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "grid_len": L, "waves": W, "note": "each step is a bounded row sample of one frame on the host cores"},
            "cpu_baseline": cb, "e2e": {"value": value, "unit": "vertex-waves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ms_per_frame_extrapolated": 1e3 * (L * L * W) / value}
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from wave_tool_b200 import api, scenes, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    L, W, cfg, shade, desc = WORKLOADS[args.workload]
    K, Wm = args.steps, max(3, args.warmup)
    flags = api.FLAG_SHADE if shade else 0
    recurrence = not args.exact_phase
    if recurrence:
        flags |= api.FLAG_FAST_PHASE

    # ---- setup (untimed): context, wave table, textures, output ring ----------------------------
    waves = scenes.seeded_waves(W, cfg=cfg)
    cam = scenes.camera()
    ctx = api.Context(local_rank)
    ctx.set_waves(waves)
    ctx.set_heightmap(scenes.procedural_heightmap(512))
    if shade:
        ctx.set_skybox(scenes.procedural_skybox(SKYBOX_N))
    # frame sharding (weak scaling): every rank evaluates whole frames, frame f on rank f mod N;
    # row sharding (cfg5, strong scaling): every rank evaluates its band of rows of the SAME frames
    row_sharded = args.workload in ROW_SHARDED
    rows = shard.row_shard(L, rank, n_gpus) if row_sharded else (0, L)
    n = (rows[1] - rows[0]) * L                                         # vertices this rank evaluates per frame
    ring = 2
    pos = torch.empty((ring * n, 4), dtype=torch.float32, device=dev)
    nrm = torch.empty((ring * n, 4), dtype=torch.float32, device=dev)
    shd = torch.empty((ring * n, 4), dtype=torch.float32, device=dev) if shade else None
    if row_sharded:
        my_frames = list(range(Wm + K))
        frames = scenes.frames_for_animation(cam, waves, first=0, count=Wm + K, stride=1)
    else:
        total_frames = (Wm + K) * n_gpus
        my_frames = shard.frame_shard(total_frames, rank, n_gpus)      # global frame indices of this rank
        frames = scenes.frames_for_animation(cam, waves, first=rank, count=Wm + K, stride=n_gpus)
    cks = torch.zeros((Wm + K) * 40, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    # roofline denominators measured on this very device (untimed)
    mufu_pairs = ctx.measure_peak(0)
    fma_flops = ctx.measure_peak(1)
    fma2_flops = ctx.measure_peak(3)
    store_bps = ctx.measure_peak(2)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launch(lo, hi, with_cks=False):
        arr = (api.Frame * (hi - lo))(*frames[lo:hi])
        ctx.eval_batch(arr, L, pos.data_ptr(), nrm.data_ptr(), shd.data_ptr() if shade else 0, ring,
                       cks.data_ptr() + 40 * lo if with_cks else 0, flags=flags | (api.FLAG_CHECKSUM if with_cks else 0), rows=rows, stream=sp)

    # ---- device-timed region: W warm-up steps, then exactly K steps ------------------------------
    launch(0, Wm)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.15 if sampler else 0.0)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record(stream)
    launch(Wm, Wm + K)
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = ctx.launch_count() - l0
    ms_local = e0.elapsed_time(e1)
    ms = shard.max_over_ranks(ms_local, dev) if world > 1 else ms_local
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    vw_per_frame = n * W                                               # this rank's share of one step
    frames_total = K if row_sharded else n_gpus * K                    # whole frames the job evaluated in the timed region
    value = frames_total * L * L * W / (ms * 1e-3)
    ms_per_step = ms / K

    # ---- end-to-end through the C ABI with HOST inputs/outputs -----------------------------------
    # per step ONE call, wt_step_host: camera + params (host structs) -> wt_build_frame (host prologue) -> fused kernel
    # with the digest -> the 40-byte result copied to pinned host memory, stream synchronised every step.
    total_amp = api.total_amplitude(waves)
    h2d = C.sizeof(api.Frame) + 32 * W            # the uniforms + wave table every launch carries to the device (kernel parameters)
    d2h = 40

    e2e_params = [scenes.animation_params(my_frames[min(Wm + k, len(my_frames) - 1)]) for k in range(K)]   # host inputs of the timed steps
    e2e_out = api.Checksum()

    def e2e_step(k):
        # ONE C-ABI call per step: host camera + parameters in, the frame's digest (host) out, synchronous
        slot = k % ring
        ctx.step_host(cam, e2e_params[k % K], total_amp, L, pos.data_ptr() + slot * n * 16, nrm.data_ptr() + slot * n * 16,
                      (shd.data_ptr() + slot * n * 16) if shade else 0, flags=flags, rows=rows, stream=sp, out=e2e_out)

    def e2e_run(count, in_flight):
        """count steps; with in_flight = 2 the host submits frame k+1 (ring slot (k+1) % 2) before it waits for frame k's digest —
        every step still carries its host inputs in and its digest out, the two directions just overlap like in a render loop"""
        pending = []
        for k in range(count):
            if in_flight == 1:
                e2e_step(k)
                continue
            slot = k % ring
            pending.append(ctx.step_submit(cam, e2e_params[k % K], total_amp, L, pos.data_ptr() + slot * n * 16, nrm.data_ptr() + slot * n * 16,
                                           (shd.data_ptr() + slot * n * 16) if shade else 0, flags=flags, rows=rows, stream=sp))
            if len(pending) >= in_flight:
                ctx.step_wait(pending.pop(0), out=e2e_out)
        for t in pending:
            ctx.step_wait(t, out=e2e_out)

    e2e = {}
    for in_flight in (1, 2):
        e2e_run(3, in_flight)
        barrier()
        t0 = time.perf_counter()
        e2e_run(K, in_flight)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_s = shard.max_over_ranks(e2e_s, dev) if world > 1 else e2e_s
        e2e[in_flight] = frames_total * L * L * W / e2e_s
    e2e_value = e2e[2]

    # full read-back variant: every output byte copied to pinned host memory each step (PCIe-bound)
    rb = None
    if not args.no_readback and rank == 0 and world == 1 and n <= (1 << 25):
        hp = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        hn = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        hs = torch.empty((n, 4), dtype=torch.float32).pin_memory() if shade else None
        t0 = time.perf_counter()
        reps = min(K, 3)
        for k in range(reps):
            e2e_step(k)
            hp.copy_(pos[:n], non_blocking=True); hn.copy_(nrm[:n], non_blocking=True)
            if shade:
                hs.copy_(shd[:n], non_blocking=True)
            stream.synchronize()
        dt = time.perf_counter() - t0
        rb = {"value": reps * vw_per_frame / dt, "unit": "vertex-waves/s", "d2h_bytes_per_step": n * 16 * (3 if shade else 2), "n_gpus": 1}

    # ---- the other phase mode on the same frames (untimed region of the bench; rank 0, N=1) + deviation ----
    other = None
    if rank == 0:
        oflags = flags ^ api.FLAG_FAST_PHASE
        f_last = frames[Wm + K - 1]
        slot_main = (Wm + K - 1) % ring                       # where the timed run left its last frame
        slot_other = 1 - slot_main
        ctx.eval_frame(f_last, L, pos.data_ptr() + slot_main * n * 16, nrm.data_ptr() + slot_main * n * 16,
                       (shd.data_ptr() + slot_main * n * 16) if shade else 0, flags=flags, rows=rows, stream=sp)
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(K, 10))
        for w_ in range(2):
            ctx.eval_frame(f_last, L, pos.data_ptr() + slot_other * n * 16, nrm.data_ptr() + slot_other * n * 16,
                           (shd.data_ptr() + slot_other * n * 16) if shade else 0, flags=oflags, rows=rows, stream=sp)
        o0.record(stream)
        for r_ in range(reps):
            ctx.eval_frame(f_last, L, pos.data_ptr() + slot_other * n * 16, nrm.data_ptr() + slot_other * n * 16,
                           (shd.data_ptr() + slot_other * n * 16) if shade else 0, flags=oflags, rows=rows, stream=sp)
        o1.record(stream)
        torch.cuda.synchronize()
        oms = o0.elapsed_time(o1) / reps
        pa, pb = pos[slot_main * n:(slot_main + 1) * n, :3], pos[slot_other * n:(slot_other + 1) * n, :3]
        na, nb = nrm[slot_main * n:(slot_main + 1) * n, :3], nrm[slot_other * n:(slot_other + 1) * n, :3]
        dpos = float((pa - pb).abs().max().item())
        dn = (na - nb).norm(dim=1)
        dn = torch.minimum(dn, (na + nb).norm(dim=1))          # the view-facing flip is sign-ambiguous at grazing incidence
        other = {"ms_per_frame": oms, "vertex_waves_per_sec": vw_per_frame / (oms * 1e-3),
                 "phase": "bit-faithful fp32 phase, MUFU pair per vertex-wave" if recurrence else "phase recurrence",
                 "max_position_deviation_between_modes": dpos, "position_tolerance": 1e-4 * float(total_amp),
                 "max_normal_deviation_rad_between_modes": float(dn.max().item()), "normal_tolerance_rad": 1e-3,
                 "checked_on": "all %d vertices of the last timed frame, on the device" % n}
        if L > 8192:
            other["note"] = ("on a 16384^2 grid the near-field vertex spacing is ~100 ulp of the fp32 coordinates: the reference's own fp32 "
                             "evaluation sits up to ~3e-3 rad from its fp64 twin there, and the two phase modes round differently; "
                             "tests/test_gpu_parity.py::test_cfg5_edge_cameras_16384 gates each mode against that measured noise floor")

    # gather the per-frame digests of the timed frames (the only collective; after the timed region)
    launch(Wm, Wm + K, with_cks=True)
    torch.cuda.synchronize()
    local_digests = cks.cpu().numpy().view(api.CHECKSUM_DTYPE)[Wm:Wm + K]
    if world > 1:
        allr = shard.all_gather_checksums(local_digests, dev)
        digest_of_digests = shard.combine_checksums(allr.ravel())
    else:
        digest_of_digests = shard.combine_checksums(local_digests)

    if rank == 0:
        peaks = measured_peaks()
        kernel_ms = ms_local / K                                    # one evaluation kernel per step
        out_bytes = n * 16 * (3 if shade else 2)
        hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload + ("" if recurrence else "_exact"))
        except Exception:
            pass
        pipes = None
        try:
            pipes = json.load(open(os.path.join(ROOT, "profiles", "pipes.json"))).get(args.workload + ("" if recurrence else "_exact"))
        except Exception:
            pass
        alg = {"vertex_waves": vw_per_frame, "transcendentals": vw_per_frame * TRANSCENDENTALS_PER_VERTEX_WAVE,
               "flop": vw_per_frame * FLOP_PER_VERTEX_WAVE, "bytes_written": out_bytes}
        trans_per_s = alg["transcendentals"] / (kernel_ms * 1e-3)
        flop_per_s = alg["flop"] / (kernel_ms * 1e-3)
        sfu_peak = 2.0 * mufu_pairs                                  # measured MUFU-only micro-kernel, this device, this run
        roof_sfu = {"bound": "sfu", "achieved": trans_per_s / 1e9, "peak": sfu_peak / 1e9, "unit": "Gtranscendental/s",
                    "frac": trans_per_s / sfu_peak, "traffic": traffic,
                    "peak_source": "of measured: MUFU-only micro-kernel (sin+cos) on this device in this run (wt_measure_peak 0); "
                                   "nominal 16/clk/SM x 148 SM x clocks.max.sm = %.0f" % (16 * 148 * ((peaks or {}).get("sm_max_mhz", 1965.0)) * 1e6 / 1e9),
                    "algorithmic_per_launch": alg}
        roof_fp32 = {"bound": "fp32", "achieved": flop_per_s / 1e12, "peak": fma_flops / 1e12, "unit": "TFLOP/s",
                     "frac": flop_per_s / fma_flops, "traffic": traffic,
                     "peak_source": "of measured: FFMA-only micro-kernel on this device in this run (wt_measure_peak 1; packed FFMA2: %.1f); "
                                    "MEASURED_PEAKS.json has no FP32 figure; nominal 128 FMA/clk/SM = 74.4" % (fma2_flops / 1e12),
                     "algorithmic_per_launch": alg,
                     "note": "15 FLOP per vertex-wave (SURVEY 8d) is the algorithmic count; the recurrence mode EXECUTES ~11.6 FLOP + 0.5 MUFU per "
                             "vertex-wave at 64 waves (row 1 of a step by one complex rotation, the others by the three-term recurrence, "
                             "1 FMA per value), so frac can exceed the FMA-pipe utilisation reported in pipes_ncu"}
        roof_hbm = {"bound": "hbm", "achieved": out_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": out_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                    "peak_source": "of measured: MEASURED_PEAKS.json hbm_gbs (copy)" if peaks else "of fallback: 6650 GB/s",
                    "store_stream_measured_gbs": store_bps / 1e9, "algorithmic_per_launch": alg}
        t_sfu = alg["transcendentals"] / sfu_peak if not recurrence else 0.25 * alg["transcendentals"] / sfu_peak
        t_fp32, t_hbm = alg["flop"] / fma_flops, out_bytes / (hbm_peak * 1e9)
        if t_hbm >= max(t_sfu, t_fp32):
            primary, secondary = roof_hbm, [roof_fp32 if recurrence else roof_sfu]
        elif recurrence:
            primary, secondary = roof_fp32, [roof_sfu, roof_hbm]
        else:
            primary, secondary = roof_sfu, [roof_fp32, roof_hbm]
        if pipes:
            primary = dict(primary)
            primary["pipes_ncu"] = pipes          # what the binding pipe was actually doing (ncu capture of the same kernel, profiles/)
        line = {
            "metric": "vertex_waves_per_sec", "value": value, "unit": "vertex-waves/s", "n_gpus": n_gpus, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if row_sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "grid_len": L, "waves": W, "shade": shade, "frames_per_step_per_gpu": 1,
                       "parallelism": (f"row-sharded x{n_gpus} (rank g evaluates rows [g L/N, (g+1) L/N) of every frame; the +-1-row stencil halo is recomputed), "
                                       "no data-path collective") if row_sharded else
                                      f"frame-sharded x{n_gpus} (frame f on rank f mod N), no data-path collective",
                       "l2": "each step writes %d MB of fresh output into a 2-deep ring (> 126 MB L2)" % (out_bytes // 2 ** 20),
                       "phase": "phase recurrence (WT_FLAG_FAST_PHASE): one MUFU seed per wave every 8 rows (12 above 64 waves), row 1 by complex rotation, "
                                "rows 2.. by the three-term recurrence x[j+1] = 2 cos(dk) x[j] - x[j-1]"
                                if recurrence else "bit-faithful fp32 phase, MUFU sin/cos pair per vertex-wave"},
            "ms_per_frame": ms_per_step,
            "roofline": primary,
            "roofline_secondary": secondary,
            "e2e": {"value": e2e_value, "unit": "vertex-waves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "per step: host camera/params -> wt_step_submit (wt_build_frame, fused kernel + digest written by its last warp into mapped host memory) "
                            "-> wt_step_wait reads the step's 40-byte digest on the host; 2 frames in flight (frame k+1 is submitted before frame k's digest is awaited)",
                    "synchronous_value": e2e[1], "synchronous_what": "the same with one wt_step_host call per step (submit + wait back to back, nothing overlapped)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "digest_of_timed_frames": {k: (int(v) if not isinstance(v, float) else v) for k, v in digest_of_digests.items()},
        }
        if other:
            line["other_phase_mode"] = other
        if rb:
            line["e2e_full_readback"] = rb
        if n_gpus == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_rate(L, W, cfg)
            cb.pop("seconds", None); cb.pop("vertex_waves", None)
            line["cpu_baseline"] = cb
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the real stdout."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    global _REAL_STDOUT
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, ...) goes to stderr
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=list(WORKLOADS), default="cfg3",
                    help="cfg3 is the bench line (BASELINE.json's metric is quoted on it); the others are extra measurements")
    ap.add_argument("--exact-phase", action="store_true",
                    help="time the bit-faithful-phase wave loop (one MUFU pair per vertex-wave) instead of the phase recurrence")
    ap.add_argument("--fast-phase", action="store_true", help="(default; kept for compatibility)")
    ap.add_argument("--no-readback", action="store_true", help="skip the full output read-back (PCIe-bound) variant of e2e")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
