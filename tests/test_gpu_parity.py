"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs (and against the committed golden vectors of the reference's own code).

Tolerances are BASELINE.json's: positions within 1e-4 x total wave amplitude, normals within 1e-3 rad
(tests/common.py).  Everything that involves no transcendental is required to be BIT-EXACT."""
import os

import numpy as np
import pytest

from tests.common import GOLDEN, Scene, angle_between, compare, deviation_stats, noise_aware_gate, numpy_checksum, reference_heightmap
from wave_tool_b200 import api, scenes

pytestmark = pytest.mark.gpu

# the two phase modes of the production kernel: bit-faithful fp32 phase + MUFU pair per vertex-wave, and
# MUFU seed + phase recurrence (WT_FLAG_FAST_PHASE).  Both must meet the same tolerances.
MODES = {"exact": 0, "recurrence": api.FLAG_FAST_PHASE}


def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def gpu_eval(ctx, frame, L, flags=0, rows=None, checksum=False):
    """Evaluate through wt_eval_frame into torch-owned device buffers; return numpy copies."""
    torch = torch_mod()
    r0, r1 = rows if rows is not None else (0, L)
    n = (r1 - r0) * L
    pos = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    nrm = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    shd = torch.empty((n, 4), dtype=torch.float32, device="cuda") if flags & api.FLAG_SHADE else None
    cks = torch.zeros(40, dtype=torch.uint8, device="cuda") if checksum else None
    f = flags | (api.FLAG_CHECKSUM if checksum else 0)
    st = ctx.eval_frame(frame, L, pos.data_ptr(), nrm.data_ptr(), shd.data_ptr() if shd is not None else 0,
                        cks.data_ptr() if cks is not None else 0, flags=f, rows=(r0, r1), stream=torch.cuda.current_stream().cuda_stream)
    assert st == api.WT_OK
    torch.cuda.synchronize()
    out = {"pos": pos.cpu().numpy(), "normal": nrm.cpu().numpy(), "shade": shd.cpu().numpy() if shd is not None else None}
    if checksum:
        out["checksum"] = cks.cpu().numpy().view(api.CHECKSUM_DTYPE)[0]
    return out


def rows_to_ids(rows, L):
    return np.concatenate([np.arange(r * L, (r + 1) * L, dtype=np.int64) for r in rows])


def take_rows(full, rows, L, row_begin=0):
    idx = np.concatenate([np.arange((r - row_begin) * L, (r - row_begin + 1) * L) for r in rows])
    return {k: (v[idx] if isinstance(v, np.ndarray) and v.ndim == 2 else v) for k, v in full.items()}


@pytest.fixture(scope="module")
def hm():
    return reference_heightmap()


# --------------------------------------------------------------------------------------------------
# configs[0]: the reference default scene (513^2 grid, the 4 hard-coded wave slots)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("t", [0.0, 1.5])
def test_cfg1_default_scene_full_grid(oracle, hm, t, mode):
    p = scenes.animation_params(int(round(t * 60)))
    sc = Scene(scenes.default_waves(), scenes.camera(), p, 513, hm)
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    got = gpu_eval(ctx, frame, 513, flags=MODES[mode])
    want = sc.oracle_eval(oracle)
    stats = compare(got, want, sc, label=f"cfg1 t={t}")
    print("cfg1", t, stats)
    # the committed golden vectors of the reference's own shader (bit-exact with the oracle on CPU)
    z = np.load(os.path.join(GOLDEN, "vert_default.npz"))
    ids = z["ids"]
    gold = {"pos": z[f"world_t{t}"], "normal": np.concatenate([z[f"out_t{t}"]["normal"], np.zeros((len(ids), 1), np.float32)], 1)}
    compare({"pos": got["pos"][ids], "normal": got["normal"][ids]}, gold, sc, label="cfg1 golden")


def test_cfg1_reference_order_kernel(oracle, hm):
    """The straight transliteration (5 evaluations per vertex, libdevice sincosf, IEEE sqrt/div) must
    sit much closer to the oracle than the tolerance: only sinf/cosf implementations differ."""
    sc = Scene(scenes.default_waves(), scenes.camera(), scenes.animation_params(90), 513, hm)
    st, frame = sc.frame()
    ctx = sc.make_context()
    got = gpu_eval(ctx, frame, 513, flags=api.FLAG_REFERENCE_ORDER)
    want = sc.oracle_eval(oracle)
    # one ulp of |x| ~ 87 is 7.6e-6: the running fp32 sum P += A*(...) can round the other way when a
    # sin/cos value differs in its last bit between libdevice and glibc
    stats = compare(got, want, sc, pos_tol=8e-6, normal_tol=2e-4, label="cfg1 reference-order")
    print("cfg1 reference-order", stats)


def test_no_transcendentals_means_bit_exact(oracle, hm):
    """With zero waves the path is bilinear placement + heightmap bump + bounce only — no sin/cos — so
    positions must match the oracle BIT FOR BIT (same fp32 operations in the same order), on every
    projector edge case."""
    for name, kw in scenes.EDGE_CAMERAS.items():
        sc = Scene(np.zeros((0, 6), np.float32), scenes.camera(**kw), scenes.animation_params(17), 257, hm)
        st, frame = sc.frame()
        if st == api.WT_NO_GRID:
            assert name == "looking_up_no_grid"
            continue
        ctx = sc.make_context()
        got = gpu_eval(ctx, frame, 257)
        want = sc.oracle_eval(oracle)
        assert np.array_equal(got["pos"].view(np.uint32), want["pos"].view(np.uint32)), name
        ang = angle_between(got["normal"][:, :3], want["normal"][:, :3])
        flipped = angle_between(got["normal"][:, :3], -want["normal"][:, :3])
        assert np.minimum(ang, flipped).max() < 2e-6, name      # rsqrt vs 1/sqrt: a few ulp of length only
        assert (ang > 1e-3).sum() <= 2, name


# --------------------------------------------------------------------------------------------------
# configs[1]: 1024^2 grid, 8 Gerstner waves + heightmap noise, animation
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", list(MODES))
def test_cfg2_animation_frames_and_batch(oracle, hm, mode):
    torch = torch_mod()
    fl = MODES[mode]
    L, waves, cam = 1024, scenes.seeded_waves(8, cfg=2), scenes.camera()
    frame_ids = [0, 1, 499, 999]
    frames = [api.build_frame(cam, scenes.animation_params(f), api.total_amplitude(waves))[1] for f in frame_ids]
    sc = Scene(waves, cam, scenes.animation_params(0), L, hm)
    ctx = sc.make_context()
    singles = []
    for f, frame in zip(frame_ids, frames):
        sc.params = scenes.animation_params(f)
        got = gpu_eval(ctx, frame, L, flags=fl, checksum=True)
        want = sc.oracle_eval(oracle)
        stats = compare(got, want, sc, label=f"cfg2 frame {f} {mode}")
        print("cfg2 frame", f, stats)
        assert tuple(got["checksum"])[:3] == numpy_checksum(got["pos"], got["normal"], None)[:3]
        assert got["checksum"]["min_y"] == got["pos"][:, 1].min() and got["checksum"]["max_y"] == got["pos"][:, 1].max()
        assert got["checksum"]["vertices"] == L * L
        singles.append(got)
    # the golden vectors of the reference's own shader for this configuration (t = 0.5 -> frame 30)
    z = np.load(os.path.join(GOLDEN, "vert_seeded8.npz"))
    sc.params = scenes.animation_params(30)
    got = gpu_eval(ctx, sc.frame()[1], L, flags=fl)
    gold = {"pos": z["world"], "normal": np.concatenate([z["out"]["normal"], np.zeros((len(z["ids"]), 1), np.float32)], 1)}
    compare({"pos": got["pos"][z["ids"]], "normal": got["normal"][z["ids"]]}, gold, sc, label="cfg2 golden")
    # wt_eval_batch: ring of 2 buffers + per-frame digests == the single-frame calls, bit for bit
    n = L * L
    ring_p = torch.empty((2 * n, 4), dtype=torch.float32, device="cuda")
    ring_n = torch.empty((2 * n, 4), dtype=torch.float32, device="cuda")
    cks = torch.zeros(40 * len(frames), dtype=torch.uint8, device="cuda")
    b0, l0 = ctx.batched_frame_count(), ctx.launch_count()
    ctx.eval_batch(frames, L, ring_p.data_ptr(), ring_n.data_ptr(), 0, 2, cks.data_ptr(), flags=api.FLAG_CHECKSUM | fl,
                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    # small frames are evaluated in grouped launches: 4 frames on a 2-deep ring = 2 groups of 2 (+ 1 launch each to arm the digests)
    assert ctx.batched_frame_count() - b0 == len(frames) and ctx.launch_count() - l0 == 4
    digests = cks.cpu().numpy().view(api.CHECKSUM_DTYPE)
    for k, s in enumerate(singles):
        assert digests[k] == s["checksum"], k
    last = len(frames) - 1
    assert np.array_equal(ring_p[(last % 2) * n:(last % 2 + 1) * n].cpu().numpy(), singles[last]["pos"])
    assert np.array_equal(ring_n[((last - 1) % 2) * n:((last - 1) % 2 + 1) * n].cpu().numpy(), singles[last - 1]["normal"])


# --------------------------------------------------------------------------------------------------
# configs[2]: 4096^2 grid, 64 waves, normals + Fresnel / sky cubemap / sun specular
# --------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg3(hm):
    sc = Scene(scenes.seeded_waves(64, cfg=3), scenes.camera(), scenes.animation_params(75), 4096, hm, skybox=scenes.procedural_skybox(512))
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    return sc, frame, ctx


@pytest.mark.parametrize("mode", list(MODES))
def test_cfg3_shaded_vs_oracle_on_sampled_rows(oracle, cfg3, mode):
    sc, frame, ctx = cfg3
    L = sc.L
    rows = [0, 1, 2, 117, 118, 119, 120, 121, 1500, 2047, 2048, 3333, 4093, 4094, 4095]
    got = gpu_eval(ctx, frame, L, flags=api.FLAG_SHADE | MODES[mode], checksum=True)
    want = sc.oracle_eval(oracle, shade=True, ids=rows_to_ids(rows, L))
    stats = compare(take_rows(got, rows, L), want, sc, shade=True, label="cfg3")
    print("cfg3", stats)
    assert tuple(got["checksum"])[:3] == numpy_checksum(got["pos"], got["normal"], got["shade"])[:3]
    # columns too: the strip seams of the tile kernel (every column of a few rows was covered above;
    # here a full-height slice of columns around likely seams)
    cols = np.array([0, 1, 2, 30, 31, 32, 60, 61, 62, 63, 64, 123, 124, 125, 2046, 2047, 2048, 4029, 4030, 4031, 4093, 4094, 4095])
    ids = (np.arange(L)[:, None] * L + cols[None, :]).ravel()
    want = sc.oracle_eval(oracle, shade=True, ids=ids)
    sub = {k: got[k][ids] for k in ("pos", "normal", "shade")}
    compare(sub, want, sc, shade=True, label="cfg3 columns")


# --------------------------------------------------------------------------------------------------
# the validity domain of the recurrence kernel over the reference's time slider (program.cpp:433: 0 .. 3600 s)
# --------------------------------------------------------------------------------------------------
ENVELOPE_CASES = {
    # name: (grid_len, waves, rows compared against the oracle (near field first))
    "cfg1": (513, lambda: scenes.default_waves(), None),
    "cfg2": (1024, lambda: scenes.seeded_waves(8, cfg=2), None),
    "cfg3": (4096, lambda: scenes.seeded_waves(64, cfg=3), list(range(0, 12)) + list(range(97, 4096, 333)) + [4094, 4095]),
    "cfg4": (8192, lambda: scenes.seeded_waves(256, cfg=4), [0, 1, 2, 3, 4097, 8190, 8191]),
}


@pytest.mark.parametrize("t", [0.0, 17.0, 600.0, 3600.0])
@pytest.mark.parametrize("name", list(ENVELOPE_CASES))
def test_phase_envelope_over_the_time_slider(oracle, hm, name, t):
    """Every BASELINE scene at the start, in the middle and at the top of waveAnimationTimeInSeconds' slider, on all three
    kernels and on WT_FLAG_AUTO_PHASE:
      * each kernel meets the north-star tolerances wherever wt_fast_phase_report says it is inside its validity domain
        (recurrence: `valid`; bit-faithful: `shared_stencil_valid`; the two five-evaluation kernels — reference-order with
        libdevice sincosf, stencil-order with the production trigonometry — always);
      * WT_FLAG_AUTO_PHASE meets them at EVERY t, and runs exactly the kernel the report recommends;
      * at the top of the slider the recurrence kernel AND the bit-faithful kernel really do leave the tolerance — the first
        because it does not share the reference's rounding of k = w d + phi t (ulp(36000) = 4e-3 rad), the second because
        the shader re-evaluates the stencil neighbours at uv +- du and those re-evaluations flip roundings of k no
        shared-stencil kernel reproduces — which is why the third tier exists.
    On the 8192^2 grid one ulp of the fp32 coordinates already tilts a finite-difference normal by more than 1e-3 rad
    (report.coordinate_ulp_normal_error_rad): there the gate is that bound, and at most 1e-4 of the vertices
    may exceed the plain tolerance."""
    torch = torch_mod()
    L, waves, rows = ENVELOPE_CASES[name]
    p = api.default_params()
    p.waveAnimationTimeInSeconds = t
    sc = Scene(waves(), scenes.camera(), p, L, hm)
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    rep = ctx.fast_phase_valid(frame, L)
    fine = L >= 8192
    if rows is None:
        want = sc.oracle_eval(oracle)
    else:
        want = sc.oracle_eval(oracle, ids=rows_to_ids(rows, L))

    def evaluate(fl):
        if fine:      # row bands (2.15 GB per full frame)
            parts = [gpu_eval(ctx, frame, L, flags=fl, rows=(max(r - 1, 0), min(r + 2, L))) for r in rows]
            return {k: np.concatenate([pp[k][(r - max(r - 1, 0)) * L:(r - max(r - 1, 0) + 1) * L] for pp, r in zip(parts, rows)]) for k in ("pos", "normal")}
        got = gpu_eval(ctx, frame, L, flags=fl)
        return got if rows is None else take_rows(got, rows, L)

    def passes(stats):
        if not fine:
            return stats["pos_margin"] >= 1.0 and stats["normal_margin"] >= 1.0
        return (stats["pos_margin"] >= 1.0 and stats["normal_max"] <= max(1e-3, rep.coordinate_ulp_normal_error_rad)
                and stats["normal_over_tol"] <= 1e-4 * stats["n"])

    results = {}
    for mode, fl, inside in (("exact", 0, rep.shared_stencil_valid), ("recurrence", api.FLAG_FAST_PHASE, rep.valid),
                             ("reference_order", api.FLAG_REFERENCE_ORDER, 1), ("stencil_order", api.FLAG_STENCIL_ORDER, 1), ("auto", api.FLAG_AUTO_PHASE, 1)):
        got = evaluate(fl)
        ran = ctx.last_phase_mode()
        stats, perr, ang, band = deviation_stats(got, want, sc)
        results[mode] = stats
        print("envelope", name, t, mode, "ran", api.PHASE_NAMES[ran], {k: round(stats[k], 3) for k in ("pos_margin", "normal_margin")}, "over", stats["normal_over_tol"],
              "inside" if inside else "outside", "recommended", api.PHASE_NAMES[rep.recommended])
        assert ran == {"exact": api.PHASE_BIT_FAITHFUL, "recurrence": api.PHASE_RECURRENCE, "reference_order": api.PHASE_REFERENCE_ORDER,
                       "stencil_order": api.PHASE_STENCIL_ORDER, "auto": rep.recommended}[mode]
        if inside:
            assert passes(stats), (name, t, mode, stats)
            assert stats["flip_band"] <= max(2, stats["n"] // 10000)
    if t >= 3600.0 and name != "cfg1":
        assert rep.recommended == api.PHASE_STENCIL_ORDER
        assert results["recurrence"]["normal_margin"] < 1.0 and results["exact"]["normal_margin"] < 1.0, "both shared-stencil kernels were expected to leave the tolerance at the top of the slider"
    if t <= 17.0:
        assert rep.recommended == (api.PHASE_RECURRENCE if name == "cfg2" else api.PHASE_BIT_FAITHFUL)


@pytest.mark.parametrize("frame_index", [22, 202])
def test_cfg3_full_grid_vs_oracle_on_the_bench_scene(oracle, frame_index):
    """ALL 16.7 M vertices of a frame bench.py times (its exact scene: the seeded 64-wave set, the reference heightmap
    fixture, the 2048^2 procedural skybox, default camera; frame 22 is the last timed frame of the driver's --steps 20 run,
    frame 202 of the default --steps 200 run) against the oracle, positions, normals and the shading terms:
      * the bit-faithful kernel and WT_FLAG_AUTO_PHASE (which must pick it here) inside the north-star tolerances;
      * the recurrence kernel is OUTSIDE its validity domain on this scene (wt_fast_phase_valid) — its deviation is
        measured and bounded at twice the tolerance, not claimed as parity."""
    L = 4096
    sky = scenes.procedural_skybox(2048)
    sc = Scene(scenes.seeded_waves(64, cfg=3), scenes.camera(), scenes.animation_params(frame_index), L, scenes.reference_heightmap(), skybox=sky)
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    assert ctx.fast_phase_valid(frame, L).valid == 0
    want = sc.oracle_eval(oracle, shade=True)
    for mode, fl in (("exact", 0), ("auto", api.FLAG_AUTO_PHASE), ("recurrence", api.FLAG_FAST_PHASE)):
        got = gpu_eval(ctx, frame, L, flags=api.FLAG_SHADE | fl)
        if mode == "recurrence":
            stats, perr, ang, band = deviation_stats(got, want, sc)
            print("cfg3 bench scene, full grid", frame_index, mode, stats)
            assert stats["pos_margin"] >= 1.0 and stats["normal_margin"] >= 0.5
        else:
            assert ctx.last_phase_mode() == 0
            stats = compare(got, want, sc, shade=True, oracle=oracle, label=f"cfg3 bench scene frame {frame_index} {mode}")
            print("cfg3 bench scene, full grid", frame_index, mode, stats)
            assert stats["pos_margin"] >= 1.2 and stats["normal_margin"] >= 1.2


@pytest.mark.parametrize("mode", list(MODES))
def test_cfg3_full_grid_vs_reference_order_kernel(cfg3, mode):
    """Full 4096^2 coverage: the production kernel against the reference-order kernel (itself checked against
    the oracle above and in cfg1) — every vertex, every strip/chunk seam."""
    sc, frame, ctx = cfg3
    a = gpu_eval(ctx, frame, sc.L, flags=api.FLAG_SHADE | MODES[mode])
    b = gpu_eval(ctx, frame, sc.L, flags=api.FLAG_SHADE | api.FLAG_REFERENCE_ORDER)
    stats = compare(a, b, sc, shade=True, label="cfg3 tile vs reference-order")
    print("cfg3 full", stats)


@pytest.mark.parametrize("mode", list(MODES))
def test_row_sharded_equals_unsharded_bit_exact(cfg3, mode):
    """Row sharding: every vertex is a pure function of (col,row) — 8 row shards give
    the very same bits as one evaluation, and their digests fold into the unsharded digest."""
    from wave_tool_b200 import shard
    sc, frame, ctx = cfg3
    L = 1024     # same scene, smaller grid (memory/time)
    fl = api.FLAG_SHADE | MODES[mode]
    full = gpu_eval(ctx, frame, L, flags=fl, checksum=True)
    parts = []
    for r in range(8):
        b, e = shard.row_shard(L, r, 8)
        part = gpu_eval(ctx, frame, L, flags=fl, rows=(b, e), checksum=True)
        # every vertex is a pure function of (col,row): the very same bits whatever the shard.  In the recurrence mode
        # too: the marching steps (and with them the re-seed rows) sit on the CANONICAL layout of the whole grid
        # (wt_ctx.h), not on the shard's first row
        for k in ("pos", "normal", "shade"):
            assert np.array_equal(part[k].view(np.uint32), full[k][b * L:e * L].view(np.uint32)), (r, k)
        parts.append(part["checksum"])
    # uneven cuts, incl. one-row shards and cuts inside a marching step
    for b, e in [(0, 1), (1, 2), (2, 9), (9, 500), (500, 501), (501, 1023), (1023, 1024)]:
        part = gpu_eval(ctx, frame, L, flags=fl, rows=(b, e))
        for k in ("pos", "normal", "shade"):
            assert np.array_equal(part[k].view(np.uint32), full[k][b * L:e * L].view(np.uint32)), (b, e, k)
    c = shard.combine_checksums(np.array(parts))
    f = full["checksum"]
    assert (c["position_bits"], c["normal_bits"], c["shade_bits"], c["vertices"]) == (f["position_bits"], f["normal_bits"], f["shade_bits"], f["vertices"])
    assert c["min_y"] == f["min_y"] and c["max_y"] == f["max_y"]


def stencil_flags(L):
    """numpy restatement of the fix-up table (wt_device.cu:ensure_fixup_tables): bit 0 / 1 where the shader's neighbour
    argument fl(fl(i du) -/+ du) is not the neighbour's own fl((i -/+ 1) du)."""
    f32 = np.float32
    du = f32(1.0) / f32(L - 1)
    i = np.arange(L, dtype=np.float32)
    u = (i * du).astype(f32)
    lo, hi = ((i - f32(1)) * du).astype(f32), ((i + f32(1)) * du).astype(f32)
    return ((u - du).astype(f32) != lo).astype(np.uint8) | (((u + du).astype(f32) != hi).astype(np.uint8) << 1)


def words_digest(out, shade):
    s = lambda a, ncomp: int(a[:, :ncomp].view(np.uint32).astype(np.uint64).sum(dtype=np.uint64))
    return s(out["pos"], 3), s(out["normal"], 4), (s(out["shade"], 4) if shade else 0)


@pytest.mark.parametrize("L,nw,cfg,shade,t", [(1024, 8, 2, False, 3600.0), (700, 64, 3, True, 3600.0), (4096, 64, 3, True, 3600.0), (513, 4, 1, True, 900.0)])
def test_stencil_fixup_equals_the_five_evaluation_kernel(hm, L, nw, cfg, shade, t):
    """WT_FLAG_STENCIL_ORDER = the bit-faithful kernel + a fix-up pass over the rows / columns whose stencil arguments differ
    from their neighbours' own, against the one-kernel form that evaluates all five points of every vertex
    (wt_debug_full_stencil).  Late in the animation, where the plain bit-faithful kernel is far off on exactly those rows and
    columns: positions identical; normals and shading terms bit-identical on the flagged rows / columns (same arithmetic, same
    neighbour positions), untouched elsewhere, where the two forms differ only by their normalisation arithmetic; the digest
    follows the corrected outputs; row shards reproduce the unsharded bits."""
    from wave_tool_b200 import shard
    waves = scenes.default_waves() if cfg == 1 else scenes.seeded_waves(nw, cfg=cfg)
    p = api.default_params()
    p.waveAnimationTimeInSeconds = t
    sc = Scene(waves, scenes.camera(), p, L, hm, skybox=scenes.procedural_skybox(64) if shade else None)
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    fl = api.FLAG_SHADE if shade else 0
    plain = gpu_eval(ctx, frame, L, flags=fl)
    ctx.stencil_form(2)
    fixed = gpu_eval(ctx, frame, L, flags=fl | api.FLAG_STENCIL_ORDER, checksum=True)
    assert api.PHASE_NAMES[ctx.last_phase_mode()] == "stencil-order"
    ctx.stencil_form(1)
    five = gpu_eval(ctx, frame, L, flags=fl | api.FLAG_STENCIL_ORDER, checksum=True)
    ctx.stencil_form(2)
    flags = stencil_flags(L)
    flagged = ((flags[None, :] != 0) | (flags[:, None] != 0)).reshape(-1)       # [row, col]
    print(f"stencil fix-up L={L}: {int((flags != 0).sum())} flagged indices, {flagged.mean():.4f} of the vertices redone")
    assert flagged.mean() < 1 and (flagged.any() or L == 513)      # du = 1 / 512 is a power of two: no argument ever differs
    u32 = lambda a: a.view(np.uint32)
    assert np.array_equal(u32(fixed["pos"]), u32(plain["pos"])) and np.array_equal(u32(five["pos"]), u32(plain["pos"]))
    keys = ("normal", "shade") if shade else ("normal",)
    for k in keys:
        assert np.array_equal(u32(fixed[k])[flagged], u32(five[k])[flagged]), k            # redone with the five-evaluation arithmetic
        assert np.array_equal(u32(fixed[k])[~flagged], u32(plain[k])[~flagged]), k          # not touched
    # the two forms agree everywhere up to the normalisation arithmetic (MUFU rsqrt vs IEEE sqrt / div) ...
    either = lambda a, b: np.minimum(angle_between(a, b), angle_between(a, -b))    # the view-facing flip may go either way at |dot| ~ 0
    ang = either(fixed["normal"][:, :3], five["normal"][:, :3])
    assert ang.max() < 2e-6, ang.max()
    # ... while the plain bit-faithful kernel, which shares positions between stencils, is off on the flagged vertices by now
    off = either(plain["normal"][:, :3], five["normal"][:, :3])
    print(f"  plain vs five-evaluation normals: max {off.max():.2e} rad on flagged vertices, {off[~flagged].max():.2e} elsewhere")
    assert off[~flagged].max() < 2e-6
    if t >= 3600.0 and flagged.any():
        assert off[flagged].max() > 1e-4
    # digest = the words actually in the buffers, for both forms
    for out in (fixed, five):
        c = out["checksum"]
        assert (int(c["position_bits"]), int(c["normal_bits"]), int(c["shade_bits"])) == words_digest(out, shade)
        assert int(c["vertices"]) == L * L
    # row shards: same bits, digests fold
    parts = []
    for b, e in [(0, 1), (1, 130), (130, L - 1), (L - 1, L)]:
        part = gpu_eval(ctx, frame, L, flags=fl | api.FLAG_STENCIL_ORDER, rows=(b, e), checksum=True)
        for k in ("pos",) + keys:
            assert np.array_equal(u32(part[k]), u32(fixed[k][b * L:e * L])), (b, e, k)
        parts.append(part["checksum"])
    c = shard.combine_checksums(np.array(parts))
    f = fixed["checksum"]
    assert (c["position_bits"], c["normal_bits"], c["shade_bits"], c["vertices"]) == (f["position_bits"], f["normal_bits"], f["shade_bits"], f["vertices"])


# --------------------------------------------------------------------------------------------------
# configs[3]: 8192^2 grid, 256 waves (row bands only — 2.15 GB per full frame)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", list(MODES))
def test_cfg4_8192_grid_256_waves_row_bands(oracle, hm, mode):
    L = 8192
    sc = Scene(scenes.seeded_waves(256, cfg=4), scenes.camera(), scenes.animation_params(1023), L, hm)
    st, frame = sc.frame()
    ctx = sc.make_context()
    for band in [(0, 6), (4090, 4100), (8186, 8192)]:
        got = gpu_eval(ctx, frame, L, rows=band, flags=MODES[mode])
        rows = [band[0], band[1] - 1]
        want = sc.oracle_eval(oracle, ids=rows_to_ids(rows, L))
        stats = compare(take_rows(got, rows, L, row_begin=band[0]), want, sc, label=f"cfg4 rows {band}")
        print("cfg4", band, stats)


# --------------------------------------------------------------------------------------------------
# configs[4]: 16384^2 grid, 32 waves, horizon-grazing / high-altitude projector edge cases
# --------------------------------------------------------------------------------------------------
FINE_GRID_MODES = {"exact": 0, "auto": api.FLAG_AUTO_PHASE, "reference_order": api.FLAG_REFERENCE_ORDER, "stencil_order": api.FLAG_STENCIL_ORDER,
                   "recurrence": api.FLAG_FAST_PHASE}


def fine_grid_check(oracle, sc, frame, ctx, L, bands, mode, label):
    """Row bands of a very fine grid against the oracle, PER VERTEX.  On a 16384^2 grid the near-field vertex
    spacing is ~100 ulp of the fp32 coordinates: ONE flipped rounding of the shader's running fp32 sum moves a vertex by
    ulp(coord) and tilts its neighbours' finite-difference normals by up to ulp(coord) / (2 x spacing) — more than the
    1e-3 rad tolerance — and any evaluation whose sin / cos are not bit-identical to the reference's libm flips such a
    rounding at a few vertices per million (so does the fp32 oracle against its fp64 twin, far more often).  Per vertex:
        err_i <= max(tolerance, 2 noise_i, one-ulp bound_i)
    with noise_i = the fp32 oracle's own distance from fp64 at that vertex and one-ulp bound_i = ulp(|x_i|, |z_i|) over the
    distance to its nearest stencil neighbour; and, as a population, the fraction of vertices over the PLAIN north-star
    tolerances may not exceed the oracle's own fraction against fp64.  Asserted for the bit-faithful kernel,
    WT_FLAG_AUTO_PHASE and the reference-order kernel (t = 0.18 s: all three are inside their validity domains); the
    recurrence kernel forced outside its domain is bounded at 3x the oracle's own noise and reported."""
    fl = FINE_GRID_MODES[mode]
    rep = ctx.fast_phase_valid(frame, L)
    tol = 1e-4 * sc.total_amplitude
    for band in bands:
        got = gpu_eval(ctx, frame, L, rows=band, flags=fl)
        ran = ctx.last_phase_mode()
        r = band[0] + 1
        ids3 = rows_to_ids([r - 1, r, r + 1], L)                   # the row and its two stencil rows (their own positions)
        want3 = sc.oracle_eval(oracle, ids=ids3)
        want = {k: (v[L:2 * L] if isinstance(v, np.ndarray) else v) for k, v in want3.items()}
        truth = sc.oracle_eval(oracle, ids=rows_to_ids([r], L), f64=True)
        sub = take_rows(got, [r], L, row_begin=band[0])
        g = noise_aware_gate(sub, want, truth, sc)
        # per-vertex one-ulp bound from the oracle's own geometry
        P = want3["pos"][:, :3].astype(np.float64).reshape(3, L, 3)
        c = P[1]
        d_lr = np.linalg.norm(np.diff(c[:, [0, 2]], axis=0), axis=1)
        d_lr = np.minimum(np.r_[d_lr, d_lr[-1]], np.r_[d_lr[0], d_lr])
        d_ud = np.minimum(np.linalg.norm((P[2] - c)[:, [0, 2]], axis=1), np.linalg.norm((c - P[0])[:, [0, 2]], axis=1))
        ulp = np.spacing(np.maximum(np.abs(c[:, 0]), np.abs(c[:, 2])).astype(np.float32)).astype(np.float64)
        one_ulp = ulp / np.maximum(np.minimum(d_lr, d_ud), 1e-30)
        g_err = np.abs(sub["pos"][:, :3].astype(np.float64) - want["pos"][:, :3]).max(axis=1)
        o_err = np.abs(want["pos"][:, :3].astype(np.float64) - truth["pos"][:, :3]).max(axis=1)
        g_ang = np.minimum(angle_between(sub["normal"][:, :3], want["normal"][:, :3]), angle_between(sub["normal"][:, :3], -want["normal"][:, :3]))
        o_ang = np.minimum(angle_between(want["normal"][:, :3], truth["normal"][:, :3].astype(np.float32)),
                           angle_between(want["normal"][:, :3], -truth["normal"][:, :3].astype(np.float32)))
        pos_fail = int((g_err > np.maximum(np.maximum(tol, 2.0 * o_err), 2.0 * ulp)).sum())
        nrm_fail = int((g_ang > np.maximum(np.maximum(1e-3, 2.0 * o_ang), one_ulp)).sum())
        print("fine grid", label, mode, "ran", api.PHASE_NAMES[ran], band, dict(pos_fail=pos_fail, normal_fail=nrm_fail, normal_max=float(g_ang.max()),
              one_ulp_max=float(one_ulp.max()), frac_over=g["normal_frac_over_tol"], oracle_frac_over=g["oracle_normal_frac_over_tol"]))
        inside = {api.PHASE_RECURRENCE: rep.valid, api.PHASE_BIT_FAITHFUL: rep.shared_stencil_valid, api.PHASE_REFERENCE_ORDER: 1, api.PHASE_STENCIL_ORDER: 1}[ran]
        if inside:
            assert pos_fail == 0 and nrm_fail == 0, (label, band, mode, g)
            assert g["pos_frac_over_tol"] <= g["oracle_pos_frac_over_tol"] + 1e-12, (label, band, g)
            assert g["normal_frac_over_tol"] <= g["oracle_normal_frac_over_tol"] + 1e-12, (label, band, g)
        else:
            # outside its domain the recurrence kernel sits about as far from the reference as the reference sits from fp64
            assert g["pos_max"] <= max(tol, 3.0 * g["oracle_pos_noise_max"]), (label, band, g)
            assert g["normal_max"] <= max(1e-3, 3.0 * g["oracle_normal_noise_max"]), (label, band, g)


@pytest.mark.parametrize("mode", list(FINE_GRID_MODES))
@pytest.mark.parametrize("name", list(scenes.EDGE_CAMERAS))
def test_cfg5_edge_cameras_16384(oracle, hm, name, mode):
    L = 16384
    sc = Scene(scenes.seeded_waves(32, cfg=5), scenes.camera(**scenes.EDGE_CAMERAS[name]), scenes.animation_params(11), L, hm)
    st, frame = sc.frame()
    st_o, _, _ = sc.oracle_uniforms(oracle)
    assert st == st_o
    if st == api.WT_NO_GRID:      # render-engine.cpp:894: surfaced as a status, not an error
        assert name == "looking_up_no_grid"
        return
    ctx = sc.make_context()
    fine_grid_check(oracle, sc, frame, ctx, L, [(0, 4), (8190, 8194), (16380, 16384)], mode, f"cfg5 {name}")


# --------------------------------------------------------------------------------------------------
# boundary behaviour
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", list(MODES))
def test_ragged_and_tiny_grids(oracle, hm, mode):
    for L in (2, 3, 5, 31, 61, 62, 63, 64, 65, 124, 125, 127, 130, 255, 257, 1000):
        sc = Scene(scenes.seeded_waves(8, cfg=2), scenes.camera(), scenes.animation_params(3), L, hm)
        st, frame = sc.frame()
        ctx = sc.make_context()
        got = gpu_eval(ctx, frame, L, flags=MODES[mode])
        want = sc.oracle_eval(oracle)
        compare(got, want, sc, label=f"L={L}")


@pytest.mark.parametrize("mode", list(FINE_GRID_MODES))
def test_maximum_grid_length_row_bands(oracle, hm, mode):
    """grid_len = 32768, the largest the ABI accepts (1 G vertices, 34 GB per full frame): row bands at the bottom, the
    middle and the top against the oracle, with the same per-vertex noise-aware gate as cfg5 (the vertex spacing is half
    of cfg5's); grid_len = 32769 is refused."""
    torch = torch_mod()
    L = 32768
    sc = Scene(scenes.seeded_waves(8, cfg=2), scenes.camera(pos=(0.0, 30.0, 70.0), pitch=-25.0), scenes.animation_params(4), L, hm)
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    fine_grid_check(oracle, sc, frame, ctx, L, [(0, 3), (16383, 16386), (32765, 32768)], mode, "L=32768")
    b = torch.empty((8, 4), dtype=torch.float32, device="cuda")
    with pytest.raises(api.WaveToolError):
        ctx.eval_frame(frame, 32769, b.data_ptr(), b.data_ptr(), rows=(0, 1))


@pytest.mark.parametrize("mode", list(MODES))
def test_wave_count_tiers(oracle, hm, mode):
    for W in (0, 1, 4, 16, 17, 64, 65, 256, 257, 1000):
        waves = scenes.seeded_waves(W, cfg=9) if W else np.zeros((0, 6), np.float32)
        sc = Scene(waves, scenes.camera(), scenes.animation_params(5), 96, hm)
        st, frame = sc.frame()
        ctx = sc.make_context()
        got = gpu_eval(ctx, frame, 96, flags=MODES[mode])
        want = sc.oracle_eval(oracle)
        tol = None if W else 1e-12
        compare(got, want, sc, pos_tol=tol, label=f"W={W}")


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("n", [1, 2, 3, 16, 61, 256])
def test_sky_footprint_on_noise_cubemap(oracle, hm, n, mode):
    """texture(skybox, R) with a maximal-contrast cubemap: the kernel's sky colour must equal the oracle's
    GL sampler (face selection, per-face CLAMP_TO_EDGE, exact fp32 bilinear weights) evaluated on the kernel's
    own position and normal.  A wrong texel order, tile origin or border of the gather atlas would be off by
    O(1); the only legitimate outliers are vertices whose reflection vector sits on a cube edge (non-seamless
    sampling: the face is decided by the last bit)."""
    sky = scenes.noise_skybox(n, seed=100 + n)
    sc = Scene(scenes.seeded_waves(8, cfg=2), scenes.camera(pos=(0.0, 9.0, 70.0), pitch=-20.0), scenes.animation_params(11), 320, hm, skybox=sky)
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    got = gpu_eval(ctx, frame, 320, flags=api.FLAG_SHADE | MODES[mode])
    st, u, _ = sc.oracle_uniforms(oracle)
    want_shade, _ = oracle.shade(u, sky, got["pos"], got["normal"])
    err = np.abs(got["shade"][:, :3].astype(np.float64) - want_shade[:, :3]).max(axis=1)
    # a 1e-7 relative difference of R moves the sample by ~1e-7 n texels; colour slope <= 1 per texel
    tol = 1e-3
    bad = err > tol
    print("noise sky", n, mode, float(np.median(err)), float(err.max()), int(bad.sum()))
    assert bad.sum() <= max(3, len(err) // 2000), f"n={n}: {int(bad.sum())} vertices off by more than {tol} (max {err.max():.3f})"
    assert got["shade"][:, :3].std() > 0.05 or n == 1     # the cubemap really was sampled


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("shade", [False, True])
def test_varyings_clip_ndc_yaw_normal(oracle, hm, mode, shade):
    """WT_FLAG_VARYINGS (wt_eval_frame_ex): gl_Position, xyPositionNDCSpaceHeight0, heightmap_colour.r and
    normalVecInViewSpaceOnlyYaw (water-grid.vert:141,150-154,193) against the oracle's copies of the same
    varyings, on a camera that is neither axis-aligned nor level so that every matrix entry matters."""
    torch = torch_mod()
    L = 384
    sky = scenes.procedural_skybox(64) if shade else None
    sc = Scene(scenes.seeded_waves(16, cfg=7), scenes.camera(pos=(3.0, 11.0, 60.0), yaw=-75.0, pitch=-17.0, aspect=1.6),
               scenes.animation_params(23), L, hm, skybox=sky)
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    n = L * L
    bufs = {k: torch.empty((n, 4), dtype=torch.float32, device="cuda") for k in ("pos", "normal", "shade", "clip", "va", "vb")}
    out = api.Outputs(bufs["pos"].data_ptr(), bufs["normal"].data_ptr(), bufs["shade"].data_ptr() if shade else None,
                      bufs["clip"].data_ptr(), bufs["va"].data_ptr(), bufs["vb"].data_ptr(), None)
    flags = api.FLAG_VARYINGS | MODES[mode] | (api.FLAG_SHADE if shade else 0)
    assert ctx.eval_frame_ex(frame, L, out, flags=flags, stream=torch.cuda.current_stream().cuda_stream) == api.WT_OK
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in bufs.items()}
    want = sc.oracle_eval(oracle, shade=shade, extra=True)
    stats = compare({"pos": got["pos"], "normal": got["normal"], "shade": got["shade"]}, want, sc, shade=shade, label="varyings scene")
    # the same launch without the flag writes the very same position / normal bits
    plain = gpu_eval(ctx, frame, L, flags=flags & ~api.FLAG_VARYINGS)
    assert np.array_equal(plain["pos"].view(np.uint32), got["pos"].view(np.uint32))
    assert np.array_equal(plain["normal"].view(np.uint32), got["normal"].view(np.uint32))
    ex = want["extra"].astype(np.float64)
    clip_w, ndc_w, yaw_w, hm_w = ex[:, 0:4], ex[:, 4:6], ex[:, 9:12], ex[:, 12]
    perr = np.abs(got["pos"][:, :3].astype(np.float64) - want["pos"][:, :3]).max(axis=1)
    vp = np.array(frame.viewProjection[:], np.float64).reshape(4, 4).T            # row r = vp[r, :]
    gain = np.abs(vp[:, :3]).sum(axis=1)                                           # |d clip_r / d position|_1
    clip_err = np.abs(got["clip"].astype(np.float64) - clip_w)
    assert np.all(clip_err <= perr[:, None] * gain[None, :] + 4e-6 * (np.abs(clip_w) + np.abs(vp[:, 3])[None, :] + 1.0)), clip_err.max()
    # ndc0 = clip0.xy / clip0.w: first-order error bound from the position error, plus a few ulp
    w0 = clip_w[:, 3] - vp[3, 1] * want["pos"][:, 1].astype(np.float64)            # w of the height-0 point
    ndc_err = np.abs(got["va"][:, :2].astype(np.float64) - ndc_w)
    bound = (perr[:, None] * (gain[None, :2] + np.abs(ndc_w) * gain[3]) / np.abs(w0)[:, None]) + 8e-6 * (1.0 + np.abs(ndc_w))
    assert np.all(ndc_err <= bound), float((ndc_err - bound).max())
    hm_err = np.abs(got["va"][:, 2].astype(np.float64) - hm_w)
    assert np.all(hm_err <= 5.0 * perr + 1e-6), hm_err.max()
    assert np.all(got["va"][:, 3] == 0.0) and np.all(got["vb"][:, 3] == 0.0)
    ang = np.minimum(angle_between(got["vb"][:, :3], yaw_w), angle_between(got["vb"][:, :3], -yaw_w))
    nang = np.minimum(angle_between(got["normal"][:, :3], want["normal"][:, :3]), angle_between(got["normal"][:, :3], -want["normal"][:, :3]))
    assert np.all(ang <= nang + 2e-6), float((ang - nang).max())                   # a rotation: no angle is amplified
    assert np.abs(np.linalg.norm(got["vb"][:, :3].astype(np.float64), axis=1) - 1.0).max() < 1e-6
    print("varyings", mode, shade, stats, float(clip_err.max()), float(ndc_err.max()), float(hm_err.max()))


def test_varyings_with_the_stencil_fixup(hm):
    """WT_FLAG_VARYINGS with WT_FLAG_STENCIL_ORDER (what WT_FLAG_AUTO_PHASE picks late in an animation): gl_Position and the
    height-0 NDC position are the production kernel's, normalVecInViewSpaceOnlyYaw follows the redone normal on the rows /
    columns the fix-up pass touches and is the production kernel's elsewhere."""
    torch = torch_mod()
    L = 1024
    p = api.default_params()
    p.waveAnimationTimeInSeconds = 3600.0
    sc = Scene(scenes.seeded_waves(8, cfg=2), scenes.camera(pos=(3.0, 11.0, 60.0), yaw=-75.0, pitch=-17.0, aspect=1.6), p, L, hm, skybox=scenes.procedural_skybox(32))
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    n = L * L

    def run(flags):
        bufs = {k: torch.empty((n, 4), dtype=torch.float32, device="cuda") for k in ("pos", "normal", "shade", "clip", "va", "vb")}
        out = api.Outputs(bufs["pos"].data_ptr(), bufs["normal"].data_ptr(), bufs["shade"].data_ptr(), bufs["clip"].data_ptr(), bufs["va"].data_ptr(), bufs["vb"].data_ptr(), None)
        assert ctx.eval_frame_ex(frame, L, out, flags=flags, stream=torch.cuda.current_stream().cuda_stream) == api.WT_OK
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in bufs.items()}
    base = api.FLAG_SHADE | api.FLAG_VARYINGS
    plain, fixed, auto = run(base), run(base | api.FLAG_STENCIL_ORDER), run(base | api.FLAG_AUTO_PHASE)
    assert api.PHASE_NAMES[ctx.last_phase_mode()] == "stencil-order"             # AUTO at t = 3600 s on this grid
    flags = stencil_flags(L)
    flagged = ((flags[None, :] != 0) | (flags[:, None] != 0)).reshape(-1)
    u32 = lambda a: a.view(np.uint32)
    for k in ("pos", "clip", "va"):
        assert np.array_equal(u32(fixed[k]), u32(plain[k])), k
    for k in ("normal", "shade", "vb"):
        assert np.array_equal(u32(fixed[k])[~flagged], u32(plain[k])[~flagged]), k
        assert np.array_equal(u32(fixed[k]), u32(auto[k])), k
    assert not np.array_equal(u32(fixed["normal"])[flagged], u32(plain["normal"])[flagged])
    # the yaw-only view-space normal of the redone vertices is the rotation of their (new) normal
    m = np.array(frame.viewMatOnlyYaw[:], np.float64).reshape(4, 4).T[:3, :3]
    want = fixed["normal"][flagged, :3].astype(np.float64) @ m.T
    want /= np.linalg.norm(want, axis=1, keepdims=True)
    assert angle_between(fixed["vb"][flagged, :3], want).max() < 2e-6
    assert np.all(fixed["vb"][:, 3] == 0.0)


def test_varyings_argument_checks(hm):
    torch = torch_mod()
    sc = Scene(scenes.default_waves(), scenes.camera(), scenes.animation_params(0), 64, hm)
    st, frame = sc.frame()
    ctx = sc.make_context()
    b = torch.empty((64 * 64, 4), dtype=torch.float32, device="cuda")
    with pytest.raises(api.WaveToolError):
        ctx.eval_frame_ex(frame, 64, api.Outputs(b.data_ptr(), b.data_ptr(), None, None, None, None, None), flags=api.FLAG_VARYINGS)
    with pytest.raises(api.WaveToolError):
        ctx.eval_frame_ex(frame, 64, api.Outputs(b.data_ptr(), b.data_ptr(), None, b.data_ptr(), b.data_ptr(), b.data_ptr(), None),
                          flags=api.FLAG_VARYINGS | api.FLAG_REFERENCE_ORDER)


def test_run_animation_is_the_frame_by_frame_loop(hm):
    """wt_run_animation (native driver: advance state, build frame, launch) must produce, frame for frame, the
    digests of the same loop driven from Python through wt_animation_advance / wt_build_frame / wt_eval_frame —
    unsharded and frame-sharded (stride 2, two start offsets)."""
    torch = torch_mod()
    L, n_frames, dt = 200, 9, 1.0 / 60.0
    waves = scenes.seeded_waves(8, cfg=2)
    sc = Scene(waves, scenes.camera(), api.default_params(), L, hm)
    ctx = sc.make_context()
    anim = api.default_animation()
    anim.isAnimatingTimeOfDay = 1
    total = sc.total_amplitude
    n = L * L
    ring = [torch.empty((2 * n, 4), dtype=torch.float32, device="cuda") for _ in range(3)]
    sky = scenes.procedural_skybox(32)
    ctx.set_skybox(sky)
    flags = api.FLAG_SHADE | api.FLAG_CHECKSUM
    # reference loop
    p = api.default_params()
    want = []
    for k in range(2 * n_frames):
        st, fr = api.build_frame(scenes.camera(), p, total)
        assert st == api.WT_OK
        want.append(tuple(gpu_eval(ctx, fr, L, flags=api.FLAG_SHADE, checksum=True)["checksum"]))
        api.animation_advance(p, anim, dt)
    for first, stride, count in ((0, 1, 2 * n_frames), (0, 2, n_frames), (1, 2, n_frames)):
        p = api.default_params()
        for _ in range(first):
            api.animation_advance(p, anim, dt)
        cks = torch.zeros(40 * count, dtype=torch.uint8, device="cuda")
        drawn = ctx.run_animation(scenes.camera(), p, anim, total, dt, count, L, ring[0].data_ptr(), ring[1].data_ptr(), ring[2].data_ptr(), 2,
                                  cks.data_ptr(), flags=flags, frame_stride=stride, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert drawn == count
        got = cks.cpu().numpy().view(api.CHECKSUM_DTYPE)
        for k in range(count):
            assert tuple(got[k]) == want[first + k * stride], (first, stride, k)
    # a camera that looks away from the water draws nothing and leaves zero digests (render-engine.cpp:894)
    p = api.default_params()
    cks = torch.full((40 * 3,), 7, dtype=torch.uint8, device="cuda")
    up = scenes.camera(**scenes.EDGE_CAMERAS["looking_up_no_grid"])
    drawn = ctx.run_animation(up, p, anim, total, dt, 3, L, ring[0].data_ptr(), ring[1].data_ptr(), ring[2].data_ptr(), 2, cks.data_ptr(), flags=flags)
    torch.cuda.synchronize()
    assert drawn == 0 and int(cks.cpu().numpy().view(api.CHECKSUM_DTYPE)["vertices"].sum()) == 0


@pytest.mark.parametrize("mode", list(MODES))
def test_compose_screen_space_layers(oracle, hm, mode):
    """wt_compose_frame (water-grid.frag:66-69,76-89,113-162 per vertex) against the oracle's compose_vertex — itself
    bit-exact with the reference's fragment shader text (tests/test_oracle_vs_ref.py) — fed the kernel's own
    vertex-stage outputs: the same fp32 operations in the same order, so the colours agree to the last bit
    except where a discontinuity (nearest-texel depth fetch, the 0.999 skybox test) sits within rounding."""
    from tests.common import compose_uniforms, screen_inputs
    torch = torch_mod()
    L, w, h = 320, 96, 64
    sky = scenes.noise_skybox(32, seed=9)
    p = api.default_params()
    p.timeOfDayInHours = 10.5
    sc = Scene(scenes.seeded_waves(16, cfg=7), scenes.camera(pos=(3.0, 11.0, 60.0), yaw=-75.0, pitch=-17.0, aspect=1.5), p, L, hm, skybox=sky)
    p.waveAnimationTimeInSeconds = 2.25
    st, frame = sc.frame()
    assert st == api.WT_OK
    ctx = sc.make_context()
    depth, refl, refr = screen_inputs(w, h, seed=21)
    ctx.set_screen_textures(depth, refl, refr)
    n = L * L
    bufs = {k: torch.empty((n, 4), dtype=torch.float32, device="cuda") for k in ("pos", "normal", "shade", "clip", "va", "vb", "colour")}
    out = api.Outputs(bufs["pos"].data_ptr(), bufs["normal"].data_ptr(), bufs["shade"].data_ptr(), bufs["clip"].data_ptr(), bufs["va"].data_ptr(),
                      bufs["vb"].data_ptr(), None)
    sp = torch.cuda.current_stream().cuda_stream
    ctx.eval_frame_ex(frame, L, out, flags=api.FLAG_SHADE | api.FLAG_VARYINGS | MODES[mode], stream=sp)
    for over in (dict(), dict(tintDeltaDepthThreshold=0.0, softEdgesDeltaDepthThreshold=0.0), dict(fogDepthRadiusNear=0.1, fogDepthRadiusFar=0.5, waterClarity=1.0)):
        cp = api.default_compose_params(sc.cam, w, h)
        for k, v in over.items():
            setattr(cp, k, v)
        ctx.compose_frame(frame, cp, L, out, bufs["colour"].data_ptr(), stream=sp)
        torch.cuda.synchronize()
        got = {k: v.cpu().numpy() for k, v in bufs.items()}
        cu = compose_uniforms(sc, oracle, w, h, **over)
        want = oracle.compose(cu, depth, refl, refr, got["pos"], got["normal"], got["shade"], got["clip"], got["va"], got["vb"])
        ok = np.isfinite(want).all(axis=1) & np.isfinite(got["colour"]).all(axis=1)
        assert ok.mean() > 0.999
        same = (got["colour"][ok].view(np.uint32) == want[ok].view(np.uint32)).all(axis=1)
        err = np.abs(got["colour"][ok].astype(np.float64) - want[ok]).max(axis=1)
        print("compose", mode, over, "bit-equal %.5f" % same.mean(), "max err %.2e" % err.max())
        assert same.mean() > 0.999
        assert np.sort(err)[int(0.9999 * len(err))] <= 1e-6
        assert 0.0 <= got["colour"][ok].min() and got["colour"][ok].max() <= 2.0 and got["colour"][ok][:, :3].std() > 0.02
        # without the varyings buffers the pass recomputes gl_Position / NDC / yaw normal from position and normal: the same
        # colours except where a last-bit difference of a screen coordinate crosses a depth texel or the 0.999 test
        lean = api.Outputs(bufs["pos"].data_ptr(), bufs["normal"].data_ptr(), bufs["shade"].data_ptr(), None, None, None, None)
        col2 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        ctx.compose_frame(frame, cp, L, lean, col2.data_ptr(), stream=sp)
        torch.cuda.synchronize()
        d2 = np.abs(col2.cpu().numpy()[ok].astype(np.float64) - got["colour"][ok]).max(axis=1)
        assert np.sort(d2)[int(0.999 * len(d2))] <= 2e-5 and np.median(d2) <= 1e-6, (np.median(d2), np.sort(d2)[int(0.999 * len(d2))])


def test_compose_argument_checks(hm):
    torch = torch_mod()
    sc = Scene(scenes.default_waves(), scenes.camera(), scenes.animation_params(0), 64, hm)
    st, frame = sc.frame()
    ctx = sc.make_context()
    b = torch.empty((64 * 64, 4), dtype=torch.float32, device="cuda")
    full = api.Outputs(*([b.data_ptr()] * 6), None)
    cp = api.default_compose_params(sc.cam, 8, 8)
    with pytest.raises(api.WaveToolError) as e:
        ctx.compose_frame(frame, cp, 64, full, b.data_ptr())
    assert e.value.status == api.WT_ERR_NOT_READY
    ctx.set_screen_textures(np.ones((8, 8), np.float32), np.zeros((8, 8, 4), np.uint8), np.zeros((8, 8, 4), np.uint8))
    with pytest.raises(api.WaveToolError):
        ctx.compose_frame(frame, cp, 64, api.Outputs(b.data_ptr(), b.data_ptr(), None, None, None, None, None), b.data_ptr())      # no shade
    with pytest.raises(api.WaveToolError):
        ctx.compose_frame(frame, cp, 64, api.Outputs(b.data_ptr(), b.data_ptr(), b.data_ptr(), b.data_ptr(), None, None, None), b.data_ptr())   # 1 of 3 varyings
    assert ctx.compose_frame(frame, cp, 64, full, b.data_ptr()) == api.WT_OK
    assert ctx.compose_frame(frame, cp, 64, api.Outputs(b.data_ptr(), b.data_ptr(), b.data_ptr(), None, None, None, None), b.data_ptr()) == api.WT_OK


@pytest.mark.parametrize("hours,cp,oc,layers", [(9.0, 0.6, 0.0, (1, 1, 1)), (12.0, 0.9, 0.7, (1, 1, 1)), (18.5, 0.3, 1.0, (0, 1, 1)),
                                                 (2.0, 0.5, 0.3, (1, 1, 0)), (6.2, 1.0, 0.5, (1, 0, 1)), (21.0, 0.6, 0.0, (0, 1, 0))])
def test_skybox_synthesis_matches_oracle(oracle, hm, hours, cp, oc, layers):
    """wt_synthesize_skybox (render-engine.cpp:306-442: stars, sky sphere, clouds, fog quad blended into RGBA8) against
    the oracle's per-texel restatement, whose layer functions are bit-exact with the reference's sky shader texts.
    Bytes must agree exactly; the only legitimate differences are texels where powf (CUDA vs glibc, <= 2 ulp apart)
    lands within rounding of a .5 boundary of the 8-bit conversion: at most 1 LSB on a vanishing fraction."""
    check_sky_synthesis(oracle, hm, hours, cp, oc, layers, 96, 24, 20)


@pytest.mark.parametrize("n,fog_alpha", [(64, 0.1), (128, 0.0), (64, 1.0), (2, 0.35), (1, 0.1), (512, 0.1)])
@pytest.mark.parametrize("hours,layers", [(9.0, (1, 1, 1)), (2.0, (1, 1, 1)), (18.6, (1, 0, 1))])
def test_skybox_synthesis_same_size_sources(oracle, hm, n, fog_alpha, hours, layers):
    """The reference's own configuration — star / cloud cubemaps of the target's power-of-two size (2048^2 assets,
    CUBEMAP_LENGTH 2048): the kernel reads them texel for texel instead of through the cube-map sampler, tabulates the fog
    quad per destination byte and writes the gather atlas (borders included) itself.  Same bytes as the oracle's general
    path; fog alpha 0 / 1 exercise the identity and the opaque blend."""
    check_sky_synthesis(oracle, hm, hours, 0.6, 0.2, layers, n, n, n, fog_noon=(1.0, 0.8, 0.9, fog_alpha))


def check_sky_synthesis(oracle, hm, hours, cp, oc, layers, n, stars_n, clouds_n, fog_noon=None):
    from tests.common import sky_sources, sky_uniforms
    gradient, stars, clouds = sky_sources(seed=int(hours * 10), stars_n=stars_n, clouds_n=clouds_n)
    sc = Scene(scenes.default_waves(), scenes.camera(), scenes.animation_params(0), 64, hm)
    ctx = sc.make_context()
    ctx.set_sky_layers(gradient, stars if layers[0] else None, clouds if layers[2] else None)
    sp = api.default_sky_params()
    sp.timeOfDayInHours, sp.cloudProportion, sp.overcastStrength = hours, cp, oc
    sp.drawStars, sp.drawSkysphere, sp.drawClouds = layers
    kw = {}
    if fog_noon is not None:
        sp.fogColourFarAtNoon[:] = fog_noon
        kw["fog_noon"] = fog_noon
    ctx.synthesize_skybox(sp, n)
    got = ctx.get_skybox()
    su = sky_uniforms(oracle, hours, cloud_proportion=cp, overcast=oc, layers=layers, **kw)
    want = oracle.sky_synthesize(su, gradient, stars, clouds, n)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    print("sky synth", hours, layers, n, "texels differing:", int((diff.max(axis=-1) > 0).sum()), "of", 6 * n * n, "max", int(diff.max()))
    assert diff.max() <= 1
    assert (diff.max(axis=-1) > 0).sum() <= max(2e-3 * 6 * n * n, 1 if n > 2 else 0)
    if n >= 16 and (fog_noon is None or fog_noon[3] < 1.0):
        assert got[..., :3].std() > 5                  # not a constant image (an opaque fog quad makes it one)
    # the synthesized cubemap is what the surface kernel samples afterwards: shade a frame with it and compare with
    # an evaluation whose skybox was uploaded through wt_set_skybox_rgba8 from the same bytes
    st, frame = sc.frame()
    a = gpu_eval(ctx, frame, 64, flags=api.FLAG_SHADE)
    ctx2 = sc.make_context()
    ctx2.set_skybox(got)
    b = gpu_eval(ctx2, frame, 64, flags=api.FLAG_SHADE)
    assert np.array_equal(a["shade"].view(np.uint32), b["shade"].view(np.uint32))


def test_step_host_round_trip(oracle, hm):
    """wt_step_host = wt_build_frame + wt_eval_frame(+digest) + read-back in one call: same digest as the separate calls,
    and WT_NO_GRID with a zero digest when the camera looks away from the water."""
    torch = torch_mod()
    L = 300
    sc = Scene(scenes.seeded_waves(8, cfg=2), scenes.camera(), scenes.animation_params(13), L, hm, skybox=scenes.procedural_skybox(32))
    ctx = sc.make_context()
    st, frame = sc.frame()
    n = L * L
    b = [torch.empty((n, 4), dtype=torch.float32, device="cuda") for _ in range(3)]
    for flags in (api.FLAG_SHADE, api.FLAG_SHADE | api.FLAG_FAST_PHASE, 0, api.FLAG_STENCIL_ORDER, api.FLAG_SHADE | api.FLAG_REFERENCE_ORDER):
        want = gpu_eval(ctx, frame, L, flags=flags, checksum=True)["checksum"]
        st, got = ctx.step_host(sc.cam, sc.params, sc.total_amplitude, L, b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr() if flags & 1 else 0,
                                flags=flags, stream=torch.cuda.current_stream().cuda_stream)
        assert st == api.WT_OK and got.as_tuple() == tuple(want)
    st, got = ctx.step_host(scenes.camera(**scenes.EDGE_CAMERAS["looking_up_no_grid"]), sc.params, sc.total_amplitude, L, b[0].data_ptr(), b[1].data_ptr())
    assert st == api.WT_NO_GRID and got.vertices == 0
    # pipelined: four frames in flight, digests come back per ticket in any order; a fifth unwaited ticket is refused later
    ring = [torch.empty((4 * n, 4), dtype=torch.float32, device="cuda") for _ in range(2)]
    params = [scenes.animation_params(20 + k) for k in range(6)]
    want = []
    for p_ in params:
        st, fr = api.build_frame(sc.cam, p_, sc.total_amplitude)
        want.append(tuple(gpu_eval(ctx, fr, L, flags=api.FLAG_FAST_PHASE, checksum=True)["checksum"]))
    tickets = [ctx.step_submit(sc.cam, params[k], sc.total_amplitude, L, ring[0].data_ptr() + (k % 4) * n * 16, ring[1].data_ptr() + (k % 4) * n * 16,
                               flags=api.FLAG_FAST_PHASE) for k in range(4)]
    for k in (2, 0, 3, 1):
        st, got = ctx.step_wait(tickets[k])
        assert st == api.WT_OK and got.as_tuple() == want[k], k
    for k in (4, 5):
        tickets.append(ctx.step_submit(sc.cam, params[k], sc.total_amplitude, L, ring[0].data_ptr() + (k % 4) * n * 16, ring[1].data_ptr() + (k % 4) * n * 16,
                                       flags=api.FLAG_FAST_PHASE))
    assert ctx.step_wait(tickets[5])[1].as_tuple() == want[5] and ctx.step_wait(tickets[4])[1].as_tuple() == want[4]
    with pytest.raises(api.WaveToolError):
        ctx.step_wait(tickets[0])            # that record has been reused since
    with pytest.raises(api.WaveToolError):
        ctx.step_wait(999)


def test_auto_phase_through_batches_and_steps(hm):
    """WT_FLAG_AUTO_PHASE is resolved per FRAME wherever frames are evaluated — wt_eval_frame, wt_eval_batch (incl. the fan-out
    over internal streams), wt_step_host — and always to the kernel wt_fast_phase_report recommends: the digests equal those of
    the same frames evaluated with that kernel's explicit flag.  The frames span all three tiers."""
    torch = torch_mod()
    L = 1024
    waves = scenes.seeded_waves(8, cfg=2)
    sc = Scene(waves, scenes.camera(), api.default_params(), L, hm)
    ctx = sc.make_context()
    explicit = {api.PHASE_RECURRENCE: api.FLAG_FAST_PHASE, api.PHASE_BIT_FAITHFUL: 0, api.PHASE_STENCIL_ORDER: api.FLAG_STENCIL_ORDER}
    frames, params, want, tiers = [], [], [], []
    for t in (0.0, 2.5, 200.0, 300.0, 3000.0, 3600.0):      # this scene: recurrence to 180 s, bit-faithful to 220 s, stencil-order beyond
        p = api.default_params()
        p.waveAnimationTimeInSeconds = t
        st, fr = api.build_frame(sc.cam, p, sc.total_amplitude)
        assert st == api.WT_OK
        tier = ctx.fast_phase_valid(fr, L).recommended
        got = gpu_eval(ctx, fr, L, flags=explicit[tier], checksum=True)
        assert ctx.last_phase_mode() == tier
        auto = gpu_eval(ctx, fr, L, flags=api.FLAG_AUTO_PHASE, checksum=True)
        assert ctx.last_phase_mode() == tier and tuple(auto["checksum"]) == tuple(got["checksum"])
        frames.append(fr); params.append(p); want.append(tuple(got["checksum"])); tiers.append(tier)
    assert set(tiers) == {api.PHASE_RECURRENCE, api.PHASE_BIT_FAITHFUL, api.PHASE_STENCIL_ORDER}, tiers
    # the stencil-order kernel evaluates the vertex itself with the bit-faithful kernel's arithmetic: identical POSITION bits
    # (the normals differ: it re-evaluates the four neighbours at the shader's uv +- du instead of sharing them)
    a = gpu_eval(ctx, frames[-1], L, flags=0)
    b = gpu_eval(ctx, frames[-1], L, flags=api.FLAG_STENCIL_ORDER)
    assert np.array_equal(a["pos"].view(np.uint32), b["pos"].view(np.uint32))
    n = L * L
    ring = [torch.empty((4 * n, 4), dtype=torch.float32, device="cuda") for _ in range(2)]
    cks = torch.zeros(40 * len(frames), dtype=torch.uint8, device="cuda")
    before = ctx.phase_mode_counts()
    ctx.eval_batch(frames, L, ring[0].data_ptr(), ring[1].data_ptr(), 0, 4, cks.data_ptr(), flags=api.FLAG_CHECKSUM | api.FLAG_AUTO_PHASE,
                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    after = ctx.phase_mode_counts()
    got = cks.cpu().numpy().view(api.CHECKSUM_DTYPE)
    for k in range(len(frames)):
        assert tuple(got[k]) == want[k], (k, tiers[k])
    assert [after[m] - before[m] for m in range(4)] == [tiers.count(m) for m in range(4)]
    for k, p in enumerate(params):
        st, d = ctx.step_host(sc.cam, p, sc.total_amplitude, L, ring[0].data_ptr(), ring[1].data_ptr(), flags=api.FLAG_AUTO_PHASE,
                              stream=torch.cuda.current_stream().cuda_stream)
        assert st == api.WT_OK and d.as_tuple() == want[k], (k, tiers[k])


def test_cpp_host_program_matches_python_driver(tmp_path):
    """examples/wt_headless.cpp (C++ host code on the C ABI, wt_run_animation) prints the same per-frame digests as
    the same animation driven through the ctypes harness."""
    import subprocess
    from tests.test_abi import build_example
    torch = torch_mod()
    exe = build_example(tmp_path)
    L, frames = 200, 5
    res = subprocess.run([exe, str(L), str(frames)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = [ln.split() for ln in res.stdout.splitlines() if ln.startswith("frame ")]
    assert len(lines) == frames
    hw = 64
    yy, xx = np.meshgrid(np.arange(hw), np.arange(hw), indexing="ij")
    v = (96 + ((xx * 7 + yy * 13) % 64)).astype(np.uint8)
    hm = np.stack([v, v, v, np.full_like(v, 255)], -1)
    waves = scenes.default_waves()
    ctx = api.Context(0)
    ctx.set_waves(waves)
    ctx.set_heightmap(hm)
    n = L * L
    pos = torch.empty((2 * n, 4), dtype=torch.float32, device="cuda")
    nrm = torch.empty((2 * n, 4), dtype=torch.float32, device="cuda")
    cks = torch.zeros(40 * frames, dtype=torch.uint8, device="cuda")
    p = api.default_params()
    ctx.run_animation(api.default_camera(1.0), p, api.default_animation(), api.total_amplitude(waves), 1.0 / 60.0, frames, L,
                      pos.data_ptr(), nrm.data_ptr(), 0, 2, cks.data_ptr(), flags=api.FLAG_CHECKSUM)
    torch.cuda.synchronize()
    want = cks.cpu().numpy().view(api.CHECKSUM_DTYPE)
    for f, ln in enumerate(lines):
        assert int(ln[3]) == int(want[f]["position_bits"]) and int(ln[5]) == int(want[f]["normal_bits"]) and int(ln[11]) == n


def test_stream_ordering_on_a_non_default_stream(hm):
    """Every entry point is stream-ordered on the caller's stream: a batch of small frames (fanned out over internal
    streams, joined back with events) followed by work on the SAME user stream sees the finished results, with no
    device-wide synchronisation in between; two contexts on two streams do not disturb each other."""
    torch = torch_mod()
    L, frames = 128, 12
    waves = scenes.seeded_waves(8, cfg=2)
    sc = Scene(waves, scenes.camera(), api.default_params(), L, hm)
    ctxs = [sc.make_context(), sc.make_context()]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    n = L * L
    fr = scenes.frames_for_animation(scenes.camera(), waves, first=0, count=frames)
    results = []
    for ctx, st in zip(ctxs, streams):
        with torch.cuda.stream(st):
            pos = torch.empty((4 * n, 4), dtype=torch.float32, device="cuda")
            nrm = torch.empty((4 * n, 4), dtype=torch.float32, device="cuda")
            cks = torch.zeros(40 * frames, dtype=torch.uint8, device="cuda")
            ctx.eval_batch(fr, L, pos.data_ptr(), nrm.data_ptr(), 0, 4, cks.data_ptr(), flags=api.FLAG_CHECKSUM, stream=st.cuda_stream)
            snap = pos.clone()                      # ordered after the batch on the same stream
            results.append((pos, snap, cks))
    for st in streams:
        st.synchronize()
    ref_ctx = sc.make_context()
    for pos, snap, cks in results:
        assert torch.equal(pos, snap)
        got = cks.cpu().numpy().view(api.CHECKSUM_DTYPE)
        for f in range(frames):
            want = gpu_eval(ref_ctx, fr[f], L, checksum=True)
            assert tuple(got[f]) == tuple(want["checksum"]), f
            if f >= frames - 4:                      # the ring holds the last four frames
                slot = f % 4
                assert np.array_equal(pos[slot * n:(slot + 1) * n].cpu().numpy().view(np.uint32), want["pos"].view(np.uint32))


@pytest.mark.parametrize("mode", list(MODES))
def test_large_grid_digest_is_shard_invariant(hm, mode):
    """Size-independent property at a BASELINE-size grid (8192^2, the two-level dynamic work queue in play): in BOTH phase
    modes every vertex is a pure function of (col, row), so the digest of the whole frame equals the fold of the digests
    of an uneven set of row shards and of the 8 row shards of a box, and min/max of y fold the same way."""
    from wave_tool_b200 import shard
    torch = torch_mod()
    L = 8192
    sc = Scene(scenes.seeded_waves(8, cfg=4), scenes.camera(), scenes.animation_params(21), L, hm)
    st, frame = sc.frame()
    ctx = sc.make_context()
    pos = torch.empty((L * L, 4), dtype=torch.float32, device="cuda")
    nrm = torch.empty((L * L, 4), dtype=torch.float32, device="cuda")
    cks = torch.zeros(40, dtype=torch.uint8, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream

    def digest(rows):
        ctx.eval_frame(frame, L, pos.data_ptr(), nrm.data_ptr(), 0, cks.data_ptr(), flags=api.FLAG_CHECKSUM | MODES[mode], rows=rows, stream=sp)
        torch.cuda.synchronize()
        return cks.cpu().numpy().view(api.CHECKSUM_DTYPE)[0].copy()

    full = digest((0, L))
    assert int(full["vertices"]) == L * L
    cuts = [0, 1, 700, 701, 4096, 8000, 8191, L]
    parts = np.array([digest((a, b)) for a, b in zip(cuts[:-1], cuts[1:])])
    c = shard.combine_checksums(parts)
    assert (c["position_bits"], c["normal_bits"], c["vertices"]) == (full["position_bits"], full["normal_bits"], full["vertices"])
    assert c["min_y"] == full["min_y"] and c["max_y"] == full["max_y"]
    c8 = shard.combine_checksums(np.array([digest(shard.row_shard(L, r, 8)) for r in range(8)]))
    assert (c8["position_bits"], c8["normal_bits"], c8["vertices"]) == (full["position_bits"], full["normal_bits"], full["vertices"])
    # and the frame itself is sane: y within the displaceable amplitude around the base plane
    amp = sc.total_amplitude + 1.0 + 0.1
    assert -amp <= full["min_y"] < full["max_y"] <= amp


def test_skybox_synthesis_argument_checks(hm):
    sc = Scene(scenes.default_waves(), scenes.camera(), scenes.animation_params(0), 64, hm)
    ctx = sc.make_context()
    with pytest.raises(api.WaveToolError) as e:
        ctx.synthesize_skybox(api.default_sky_params(), 16)
    assert e.value.status == api.WT_ERR_NOT_READY
    ctx.set_sky_layers(np.full((8, 4), 200, np.uint8))
    ctx.synthesize_skybox(api.default_sky_params(), 16)         # stars and clouds absent: sky sphere + fog only
    assert ctx.get_skybox().shape == (6, 16, 16, 4)
    with pytest.raises(api.WaveToolError):
        ctx.synthesize_skybox(api.default_sky_params(), 0)


@pytest.mark.parametrize("scale", [1.0, 60.0])
def test_heightmap_sample_scale_extremes(oracle, hm, scale):
    """heightmapSampleScale at the top of its slider (1.0, program.cpp:410) and far beyond (a CTRL-click is only clamped from
    below): the bit-faithful kernel floors the sample coordinates with a magic-number add while they are far inside +-2^22
    texels and with FRND / F2I beyond (scale 60: ~3e6 texels).  With zero waves the positions must be bit-exact either way."""
    p = scenes.animation_params(9)
    p.heightmapSampleScale = scale
    for waves, exact in ((np.zeros((0, 6), np.float32), True), (scenes.seeded_waves(8, cfg=2), False)):
        sc = Scene(waves, scenes.camera(), p, 257, hm)
        st, frame = sc.frame()
        ctx = sc.make_context()
        want = sc.oracle_eval(oracle)
        for fl in (0, api.FLAG_FAST_PHASE):
            got = gpu_eval(ctx, frame, 257, flags=fl)
            if exact:
                assert np.array_equal(got["pos"].view(np.uint32), want["pos"].view(np.uint32)), (scale, fl)
            else:
                # at these scales one heightmap texel spans 1/512 .. 1/30000 of a world unit: a 1e-5 difference in the displaced xz
                # (inside the position tolerance) lands on a different part of the texture, so y is only comparable statistically
                perr = np.abs(got["pos"][:, :3].astype(np.float64) - want["pos"][:, :3]).max(axis=1)
                hxz = np.abs(got["pos"][:, [0, 2]].astype(np.float64) - want["pos"][:, [0, 2]]).max(axis=1)
                assert hxz.max() <= 1e-4 * sc.total_amplitude                      # the horizontal position is unaffected by the heightmap
                if scale <= 1.0:
                    assert np.median(perr) <= 1e-5 and (perr > 1e-3).mean() < 1e-3
                else:      # one texel = 3e-5 world units: y decorrelates completely, only its range is checkable
                    assert np.abs(got["pos"][:, 1]).max() <= sc.total_amplitude + 1.0 + 0.1 + 1e-3


def test_non_power_of_two_heightmap(oracle):
    hm = scenes.procedural_heightmap(200)[:150]      # 200 x 150 texels
    sc = Scene(scenes.seeded_waves(8, cfg=2), scenes.camera(), scenes.animation_params(3), 300, hm)
    st, frame = sc.frame()
    ctx = sc.make_context()
    compare(gpu_eval(ctx, frame, 300), sc.oracle_eval(oracle), sc, label="npot heightmap")


def test_error_statuses(hm):
    torch = torch_mod()
    ctx = api.Context(0)
    buf = torch.empty((64, 4), dtype=torch.float32, device="cuda")
    st, frame = api.build_frame(scenes.camera(), api.default_params(), 0.16)
    with pytest.raises(api.WaveToolError) as e:
        ctx.eval_frame(frame, 8, buf.data_ptr(), buf.data_ptr())
    assert e.value.status == api.WT_ERR_NOT_READY
    ctx.set_waves(scenes.default_waves())
    with pytest.raises(api.WaveToolError) as e:      # no heightmap -> the reference draws no water (program.cpp:721)
        ctx.eval_frame(frame, 8, buf.data_ptr(), buf.data_ptr())
    assert e.value.status == api.WT_ERR_NOT_READY
    ctx.set_heightmap(hm)
    for bad in [dict(grid_len=1), dict(grid_len=8, rows=(5, 5)), dict(grid_len=8, rows=(0, 9))]:
        with pytest.raises(api.WaveToolError) as e:
            ctx.eval_frame(frame, bad["grid_len"], buf.data_ptr(), buf.data_ptr(), rows=bad.get("rows"))
        assert e.value.status == api.WT_ERR_INVALID_ARGUMENT
    with pytest.raises(api.WaveToolError):            # shading requested without a shade buffer
        ctx.eval_frame(frame, 8, buf.data_ptr(), buf.data_ptr(), flags=api.FLAG_SHADE)
    bad_wave = scenes.default_waves()
    bad_wave[0, 4] = 3.0                               # |D| != 1 (render-engine.h:152-154)
    with pytest.raises(api.WaveToolError):
        ctx.set_waves(bad_wave)
    with pytest.raises(api.WaveToolError):
        api.Context(10 ** 6)
    assert ctx.eval_frame(frame, 8, buf.data_ptr(), buf.data_ptr()) == api.WT_OK


def test_host_buffer_entry_point(oracle, hm):
    sc = Scene(scenes.default_waves(), scenes.camera(), scenes.animation_params(0), 513, hm, skybox=scenes.procedural_skybox(64))
    st, frame = sc.frame()
    ctx = sc.make_context()
    out = ctx.eval_frame_host(frame, 513, flags=api.FLAG_SHADE, want_checksum=True)
    want = sc.oracle_eval(oracle, shade=True)
    compare(out, want, sc, shade=True, label="host entry point")
    assert out["checksum"].as_tuple()[:3] == numpy_checksum(out["pos"], out["normal"], out["shade"])[:3]
    n0 = ctx.launch_count()
    ctx.eval_frame_host(frame, 513, flags=api.FLAG_SHADE, want_checksum=True)
    assert ctx.launch_count() == n0 + 1        # ONE launch per frame: the digest is finalised by the kernel's last warp
