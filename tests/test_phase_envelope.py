"""Validity domain of WT_FLAG_FAST_PHASE (the recurrence kernel) — CPU part.

The reference rounds every wave's phase to fp32 at every vertex (water-grid.vert:78), so its result carries a
deterministic noise of ~ulp(k) per wave; the bit-faithful kernel reproduces it, the recurrence kernel cannot.  The
oracle's `smooth_phase` variant models exactly that: phase and sin/cos in double, everything else (products, the running
fp32 sum in wave order) as the shader.  Its distance from the plain fp32 oracle is the deviation from the reference any
smooth-phase evaluation has, before the kernel's own MUFU / recurrence noise (measured on the GPU: x1.27 on normals,
profiles/parity_margins.json).  wt_fast_phase_report must bound that deviation wherever it says "valid", and must say
"invalid" towards the top of the reference's time slider (program.cpp:433: 3600 s), where WT_FLAG_AUTO_PHASE runs the
bit-faithful kernel instead."""
import numpy as np
import pytest

from tests.common import Scene, angle_between, reference_heightmap
from wave_tool_b200 import api, scenes

KERNEL_OWN_NOISE = 1.27      # recurrence kernel vs the smooth-phase model, measured on the GPU (normals, cfg3)

CASES = {
    "cfg1": (513, lambda: scenes.default_waves(), list(range(0, 513, 8))),
    "cfg2": (1024, lambda: scenes.seeded_waves(8, cfg=2), list(range(0, 1024, 32))),
    "cfg3": (4096, lambda: scenes.seeded_waves(64, cfg=3), list(range(0, 10)) + list(range(100, 4096, 450))),
}


def model_deviation(oracle, name, t, cam=None):
    L, waves, rows = CASES[name]
    p = api.default_params()
    p.waveAnimationTimeInSeconds = t
    sc = Scene(waves(), cam or scenes.camera(), p, L, reference_heightmap())
    st, frame = sc.frame()
    assert st == api.WT_OK
    ids = np.concatenate([np.arange(r * L, (r + 1) * L, dtype=np.int64) for r in rows])
    ref = sc.oracle_eval(oracle, ids=ids)
    st_, u, wu = sc.oracle_uniforms(oracle)
    smooth = oracle.eval(u, wu, sc.heightmap, ids=ids, smooth_phase=True)
    perr = np.abs(ref["pos"][:, :3].astype(np.float64) - smooth["pos"][:, :3]).max()
    ang = np.minimum(angle_between(ref["normal"][:, :3], smooth["normal"][:, :3]), angle_between(ref["normal"][:, :3], -smooth["normal"][:, :3])).max()
    return sc, frame, float(perr), float(ang), api.fast_phase_report(sc.waves, frame, L)


@pytest.mark.parametrize("t", [0.0, 17.0, 600.0, 3600.0])
@pytest.mark.parametrize("name", list(CASES))
def test_report_bounds_the_smooth_phase_model(oracle, name, t):
    sc, frame, perr, ang, rep = model_deviation(oracle, name, t)
    print(name, t, "model pos %.2e est %.2e | normal %.2e est %.2e | valid %d" % (perr, rep.est_position_error, ang, rep.est_normal_error_rad, rep.valid))
    assert abs(rep.position_tolerance - 1e-4 * sc.total_amplitude) <= 1e-9 and abs(rep.normal_tolerance_rad - 1e-3) < 1e-9
    # the estimate is an upper bound of what the model shows on the sampled rows (which include the near field)
    assert perr <= rep.est_position_error * 1.05
    with_kernel = min(ang * KERNEL_OWN_NOISE, ang + 3e-4)      # the kernel's own noise is absolute: it only matters while the deviation is small
    assert ang <= rep.est_normal_error_rad and with_kernel <= rep.est_normal_error_rad * 1.05
    if rep.valid:
        assert perr <= rep.position_tolerance and with_kernel <= rep.normal_tolerance_rad
    # the domain: the 1024^2 animation scene is inside it at the start of an animation; the 4096^2 scene under the default
    # (horizon-grazing) camera never is — its near-field vertices are 2.6e-3 apart, ~350 ulp of their z coordinate, and the
    # reference's own rounding noise tilts the finite-difference normals by up to ~1e-3 rad there; nothing is at the top
    # of the slider
    if t <= 17.0:
        assert rep.valid == (1 if name == "cfg2" else 0)
    if t >= 600.0:
        assert rep.valid == 0
    # the bit-faithful (shared-stencil) kernel: inside its domain for the first minutes of an animation on every scene,
    # outside at the top of the slider, where only a kernel that evaluates the shader's own neighbour arguments follows it — except
    # on the 513^2 grid: du = 1/512 is a power of two, no stencil argument differs from the neighbour's own, the kernel is exact there
    assert rep.shared_stencil_valid == (1 if t <= 17.0 or name == "cfg1" else 0 if t >= 3600.0 else rep.shared_stencil_valid)
    assert rep.recommended == (api.PHASE_RECURRENCE if rep.valid else api.PHASE_BIT_FAITHFUL if rep.shared_stencil_valid else api.PHASE_STENCIL_ORDER)
    assert rep.coordinate_ulp_normal_error_rad > 0.0


def test_smooth_phase_really_leaves_the_tolerance_at_large_t(oracle):
    """Why the switch exists: at t = 3600 s ANY smooth-phase evaluation of the cfg3 scene is outside both north-star
    tolerances against the reference — by the reference's own rounding of k = w d + phi t to ulp(36000) = 2e-3 rad."""
    sc, frame, perr, ang, rep = model_deviation(oracle, "cfg3", 3600.0)
    assert ang > 10e-3 and perr > 1e-4 * sc.total_amplitude and rep.valid == 0


def test_report_is_monotone_in_time_and_resolution(oracle):
    waves = scenes.seeded_waves(64, cfg=3)
    prev = None
    for t in (0.0, 10.0, 100.0, 1000.0, 3600.0):
        p = api.default_params()
        p.waveAnimationTimeInSeconds = t
        st, frame = api.build_frame(scenes.camera(), p, api.total_amplitude(waves))
        rep = api.fast_phase_report(waves, frame, 4096)
        if prev is not None:
            assert rep.est_normal_error_rad >= prev.est_normal_error_rad and rep.est_position_error >= prev.est_position_error
            assert rep.max_phase > prev.max_phase
        prev = rep
    p = api.default_params()
    st, frame = api.build_frame(scenes.camera(), p, api.total_amplitude(waves))
    e = [api.fast_phase_report(waves, frame, L) for L in (512, 1024, 4096, 16384)]
    assert all(a.est_normal_error_rad < b.est_normal_error_rad for a, b in zip(e[:-1], e[1:]))      # finer grid, noisier finite differences
    assert all(abs(a.min_vertex_spacing / b.min_vertex_spacing - (Lb - 1) / (La - 1)) < 1e-3
               for a, b, La, Lb in zip(e[:-1], e[1:], (512, 1024, 4096), (1024, 4096, 16384)))
    assert e[0].valid == 1 and e[1].valid == 1 and e[2].valid == 0 and e[-1].valid == 0
    # the same waves seen from above (vertex spacing 1.8e-2 instead of 2.6e-3): inside the domain at 4096^2
    st, high = api.build_frame(scenes.camera(**scenes.EDGE_CAMERAS["high_looking_down"]), p, api.total_amplitude(waves))
    assert api.fast_phase_report(waves, high, 4096).valid == 1
    # no waves: nothing to deviate
    assert api.fast_phase_report(np.zeros((0, 6), np.float32), frame, 4096).valid == 1


def test_stencil_flag_rate_matches_its_restatement():
    """wt_stencil_flag_rate: where the shader's neighbour argument fl(fl(i du) -/+ du) is not the neighbour's own fl((i -/+ 1) du)
    (water-grid.vert:123-125,164-167) — the table the fix-up pass of WT_FLAG_STENCIL_ORDER works from — against numpy."""
    f32 = np.float32
    for L, expect in ((513, 0.0), (1025, 0.0), (1024, None), (4096, None), (8192, None), (700, None), (2, None), (3, None)):
        du = f32(1.0) / f32(L - 1)
        i = np.arange(L, dtype=np.float32)
        u = (i * du).astype(f32)
        sides = int(((u - du).astype(f32) != ((i - f32(1)) * du).astype(f32)).sum() + ((u + du).astype(f32) != ((i + f32(1)) * du).astype(f32)).sum())
        got = api.stencil_flag_rate(L)
        assert abs(got - sides / (2.0 * L)) < 1e-6, (L, got, sides)
        if expect is not None:
            assert got == expect
    assert 0.001 < api.stencil_flag_rate(4096) < 0.005 and 0.15 < api.stencil_flag_rate(8192) < 0.25
