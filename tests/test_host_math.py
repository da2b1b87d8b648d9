"""The PRODUCT's host prologue (C ABI: wt_camera_matrices / wt_project_grid / wt_build_frame / scalar
rules, wave_tool_b200/csrc/wt_host.cpp) against the golden vectors produced by the reference's own
code.  Bit-exact.  No GPU needed: this is the part the reference itself runs on the CPU."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.common import GOLDEN, Scene
from wave_tool_b200 import api, scenes


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_projector_and_matrices_match_reference_bits():
    z = np.load(os.path.join(GOLDEN, "projector.npz"))
    for i in range(len(z["names"])):
        cam = api.Camera.from_buffer_copy(z["cameras"][i].tobytes())
        st, corners = api.project_grid(cam, float(z["D"][i]))
        assert st == z["status"][i], z["names"][i]
        if st == api.WT_OK:
            assert np.array_equal(bits(corners).ravel(), bits(z["corners"][i]).ravel()), z["names"][i]
        m = api.camera_matrices(cam)
        mats = np.concatenate([m["view"], m["proj"], m["view_proj"], m["view_only_yaw"]])
        assert np.array_equal(bits(mats), bits(z["matrices"][i])), z["names"][i]
        assert np.array_equal(bits(m["forward"]), bits(z["forward"][i]))


def test_scalar_rules_match_reference_bits():
    z = np.load(os.path.join(GOLDEN, "scalars.npz"))
    L = api.lib()
    for (q, w, a), n, expect in zip(z["q_in"], z["q_count"], z["q_out"]):
        wave = api.Wave(a, w, 0.0, q, 1.0, 0.0)
        got = np.float32(L.wt_steepness_Q_i(C.byref(wave), int(n)))
        assert bits(got) == bits(expect)
    for h, expect in zip(z["hours"], z["sun"]):
        out = np.zeros(3, np.float32)
        L.wt_sun_position(float(h), out.ctypes.data)
        assert np.array_equal(bits(out), bits(expect))
    for p, expect in zip(z["phases"], z["bounce"]):
        assert bits(np.float32(L.wt_bounce_displacement(0.1, float(p)))) == bits(expect)
    waves = np.zeros((256, 6), np.float32)
    waves[:, 0] = z["amps"]
    assert bits(np.float32(api.total_amplitude(waves))) == bits(z["amp_total"])


def test_defaults_are_the_reference_defaults():
    p = api.default_params()
    assert (p.heightmapDisplacementScale, p.verticalBounceWaveAmplitude, p.timeOfDayInHours, p.sunShininess) == (1.0, np.float32(0.1), 9.0, 50.0)
    assert abs(p.heightmapSampleScale - 0.02) < 1e-9
    c = api.default_camera(1.0)
    assert (c.fov_deg, c.z_far, c.yaw_deg, c.pitch_deg, tuple(c.pos)) == (72.0, 100.0, -90.0, 0.0, (0.0, 4.0, 70.0))
    w = api.default_waves()
    assert w.shape == (4, 6) and np.float32(api.total_amplitude(w)) == np.float32(np.float32(0.06) + np.float32(0.1))
    assert abs(np.hypot(w[1, 4], w[1, 5]) - 1.0) < 1e-6


def test_build_frame_equals_oracle_uniforms(oracle, heightmap):
    """wt_build_frame (product) and the oracle derive the draw call's uniforms independently from the
    same camera/params: every float must agree bit for bit."""
    for name, kw in scenes.EDGE_CAMERAS.items():
        p = scenes.animation_params(37)
        sc = Scene(scenes.seeded_waves(32, cfg=5), scenes.camera(**kw), p, 64, heightmap)
        st, f = sc.frame()
        st_o, u, _ = sc.oracle_uniforms(oracle)
        assert st == st_o, name
        if st != api.WT_OK:
            assert st == api.WT_NO_GRID
            continue
        pairs = [(f.bottomLeftGridPointInWorld, u.bottomLeft), (f.bottomRightGridPointInWorld, u.bottomRight),
                 (f.topLeftGridPointInWorld, u.topLeft), (f.topRightGridPointInWorld, u.topRight),
                 (f.viewProjection, u.viewProjection), (f.viewMatOnlyYaw, u.viewMatOnlyYaw), (f.cameraPosition, u.cameraPosition),
                 (f.sunPosition, u.sunPosition)]
        for a, b in pairs:
            assert np.array_equal(bits(a[:]), bits(b[:])), name
        for a, b in [(f.waveAnimationTimeInSeconds, u.waveAnimationTimeInSeconds), (f.verticalBounceWaveDisplacement, u.verticalBounceWaveDisplacement),
                     (f.heightmapDisplacementScale, u.heightmapDisplacementScale), (f.heightmapSampleScale, u.heightmapSampleScale),
                     (f.sunShininess, u.sunShininess), (f.sunStrength, u.sunStrength)]:
            assert bits(np.float32(a)) == bits(np.float32(b)), name


def test_null_arguments_are_reported_not_crashed():
    L = api.lib()
    assert L.wt_project_grid(None, C.c_float(1.0), None) == api.WT_ERR_INVALID_ARGUMENT
    assert b"NULL" in L.wt_last_error(None)
    assert L.wt_build_frame(None, None, C.c_float(1.0), None) == api.WT_ERR_INVALID_ARGUMENT


def test_unorm8_reciprocal_trick_is_exact():
    """The kernel converts RGBA8 texels with q = c*r; q += fma(-q,255,c)*r (wt_kernels.cuh:unorm8);
    emulate that in numpy (fp64 holds the fp32 FMA exactly) and compare with IEEE c/255."""
    c = np.arange(256, dtype=np.float32)
    r = np.float32(1) / np.float32(255)
    q = (c * r).astype(np.float32)
    rem = (c.astype(np.float64) - q.astype(np.float64) * 255.0).astype(np.float32)
    q2 = (q.astype(np.float64) + rem.astype(np.float64) * np.float64(r)).astype(np.float32)
    assert np.array_equal(q2, c / np.float32(255))


def test_grid_indices_match_reference_tessellation():
    """wt_grid_indices against a literal restatement of the tessellation loop of src/program.cpp:686-717."""
    for L in (2, 3, 5, 17, 513):
        grid = np.arange(L * L, dtype=np.uint32).reshape(L, L)      # gridIndices[row][col] = counterIndex (:695-700)
        want = []
        for row in range(L - 1):                                      # :704-716
            for col in range(L - 1):
                want += [grid[row, col], grid[row + 1, col + 1], grid[row + 1, col],
                         grid[row, col], grid[row, col + 1], grid[row + 1, col + 1]]
        got = api.grid_indices(L)
        assert got.dtype == np.uint32 and len(got) == 6 * (L - 1) ** 2
        assert np.array_equal(got, np.array(want, np.uint32))
    with pytest.raises(api.WaveToolError):
        api.grid_indices(1)


def test_animation_advance_matches_reference_state_machine():
    """wt_animation_advance against a literal fp32 restatement of program.cpp:220-225 and :452-464."""
    f = np.float32
    for tod_on, period, sph in ((0, 3.0, 1.0), (1, 0.7, 2.5), (1, 0.0, 0.0)):
        p = api.default_params()
        anim = api.default_animation()
        anim.isAnimatingTimeOfDay = tod_on
        anim.animationSpeedVerticalBounceWavePhasePeriodInSeconds = period
        anim.animationSpeedTimeOfDayInSecondsPerHour = sph
        t, ph, tod = f(p.waveAnimationTimeInSeconds), f(p.verticalBounceWavePhase), f(p.timeOfDayInHours)
        for k in range(2000):
            dt = f(1.0) / f(60.0 if k % 3 else 144.0)              # deltaTimeInSeconds = 1 / Framerate (:453)
            api.animation_advance(p, anim, float(dt))
            if tod_on and f(sph) > 0:                               # :220-225
                tod = f(tod + f(f(f(1.0) / f(sph)) * dt))
                tod = f(tod - f(f(24.0) * np.floor(f(tod / f(24.0)))))
            t = f(t + dt)                                           # :455-457
            if t < 0:
                t = f(0)
            if f(period) > 0:                                       # :459-463
                ph = f(ph + f(f(f(1.0) / f(period)) * dt))
                ph = f(ph - f(f(1.0) * np.floor(f(ph / f(1.0)))))
            assert f(p.waveAnimationTimeInSeconds) == t and f(p.verticalBounceWavePhase) == ph and f(p.timeOfDayInHours) == tod, k
    d = api.default_animation()
    assert (d.animationSpeedTimeOfDayInSecondsPerHour, d.animationSpeedVerticalBounceWavePhasePeriodInSeconds, d.isAnimatingTimeOfDay, d.isAnimatingWaves) == (1.0, 3.0, 0, 1)


# --------------------------------------------------------------------------------------------------
# parameter surface: the reference's sliders, their force-clamp code, the physical wave constructor
# --------------------------------------------------------------------------------------------------
def test_wave_from_wavelength_speed_conventions():
    """render-engine.h:139-140: 'w = 2/L (roughly)', 'phi = S x w'.  The physical convention (w = 2 pi / L) reproduces the
    seeded benchmark sets (wave_tool_b200/scenes.py) to fp32 rounding."""
    import ctypes as C
    w = api.wave_from_wavelength_speed(0.25, 8.0, 0.5, (3.0, 4.0), 2.0, api.WAVENUMBER_GPU_GEMS)
    assert (w.amplitude_A, w.frequency_w, w.phaseConstant_phi, w.steepness_Q) == (0.25, 0.25, 0.5, 0.5)
    assert abs(w.dir_x - 0.6) < 1e-7 and abs(w.dir_z - 0.8) < 1e-7
    w = api.wave_from_wavelength_speed(0.25, 8.0, 0.5, (0.0, -2.0), 2.0, api.WAVENUMBER_PHYSICAL)
    assert abs(w.frequency_w - 2 * np.pi / 8) < 1e-7 and abs(w.phaseConstant_phi - 2 * 2 * np.pi / 8) < 2e-7 and (w.dir_x, w.dir_z) == (0.0, -1.0)
    rows = scenes.seeded_waves(64, cfg=3)
    for A, wn, phi, Q, dx, dz in rows[::7]:
        lam = 2 * np.pi / float(wn)
        speed = float(phi) / float(wn)                      # deep-water phase speed sqrt(g lambda / 2 pi)
        assert abs(speed - np.sqrt(9.8 * lam / (2 * np.pi))) < 1e-5 * speed
        got = api.wave_from_wavelength_speed(float(A), lam, float(Q), (float(dx), float(dz)), speed)
        want = np.array([A, wn, phi, Q, dx, dz], np.float32)
        have = np.array([got.amplitude_A, got.frequency_w, got.phaseConstant_phi, got.steepness_Q, got.dir_x, got.dir_z], np.float32)
        assert np.all(np.abs(have - want) <= 4 * np.spacing(np.abs(want)) + 1e-7), (have, want)
        # and the result is a wave wt_set_waves' validation (render-engine.h:147-154) accepts: unit direction, ranges
        assert abs(np.hypot(got.dir_x, got.dir_z) - 1.0) < 1e-6
    for bad in [dict(wavelength=0.0), dict(amplitude=-1.0), dict(speed=-0.1), dict(steepness=1.5), dict(direction=(0.0, 0.0))]:
        kw = dict(amplitude=0.1, wavelength=4.0, steepness=0.5, direction=(1.0, 0.0), speed=1.0)
        kw.update(bad)
        with pytest.raises(api.WaveToolError):
            api.wave_from_wavelength_speed(**kw)


def test_ui_ranges_and_force_clamps_at_the_range_ends():
    """program.cpp:201-479: slider bounds and the asymmetric 'force-clamp (handle CTRL + LEFT_CLICK)' code — glm::clamp on
    some fields, only `if (v < 0) v = 0` on others (a CTRL-click can exceed their slider maximum)."""
    L = api.lib()
    r = api.ui_ranges()
    assert len(r) == api.UI_RANGE_COUNT
    assert r["waveAnimationTimeInSeconds"] == (0.0, 3600.0, True, False)         # :433
    assert r["sunShininess"] == (0.0, 200.0, True, False) and r["sunStrength"] == (0.0, 1.0, True, True)   # :242, :247
    assert r["phaseConstant_phi"] == (0.0, 10.0, True, False) and r["steepness_Q"] == (0.0, 1.0, True, True)   # :384, :388
    assert r["amplitude_A"] == (0.0, 1.0, True, False) and r["frequency_w"] == (0.0, 1.0, True, False)
    assert r["timeOfDayInHours"] == (0.0, 24.0, True, True) and r["xzDirection_D"] == (-1.0, 1.0, True, True)
    # params: both ends of every range + beyond
    p = api.default_params()
    p.timeOfDayInHours, p.sunShininess, p.sunStrength = 25.0, 200.0, 1.0
    p.heightmapSampleScale, p.heightmapDisplacementScale, p.verticalBounceWaveAmplitude = 1.0, 1.0, 1.0
    p.verticalBounceWavePhase, p.waveAnimationTimeInSeconds = 1.0, 3600.0
    L.wt_params_clamp_ui(C.byref(p))
    assert (p.timeOfDayInHours, p.sunShininess, p.sunStrength, p.heightmapSampleScale, p.heightmapDisplacementScale,
            p.verticalBounceWaveAmplitude, p.verticalBounceWavePhase, p.waveAnimationTimeInSeconds) == (24.0, 200.0, 1.0, 1.0, 1.0, 1.0, 1.0, 3600.0)
    p.timeOfDayInHours, p.sunShininess, p.sunStrength = -1.0, -5.0, 1.5
    p.heightmapSampleScale, p.heightmapDisplacementScale, p.verticalBounceWaveAmplitude = -0.1, 3.0, -2.0
    p.verticalBounceWavePhase, p.waveAnimationTimeInSeconds = -0.5, 7200.0
    L.wt_params_clamp_ui(C.byref(p))
    assert (p.timeOfDayInHours, p.sunShininess, p.sunStrength, p.heightmapSampleScale, p.heightmapDisplacementScale,
            p.verticalBounceWaveAmplitude, p.verticalBounceWavePhase, p.waveAnimationTimeInSeconds) == (0.0, 0.0, 1.0, 0.0, 3.0, 0.0, 0.0, 7200.0)
    p.sunShininess = 250.0                                                         # above the slider: NOT clamped by the reference (:244)
    L.wt_params_clamp_ui(C.byref(p))
    assert p.sunShininess == 250.0
    # waves
    w = api.Wave(1.0, 1.0, 10.0, 1.0, 1.0, 1.0)
    L.wt_wave_clamp_ui(C.byref(w))
    assert (w.amplitude_A, w.frequency_w, w.phaseConstant_phi, w.steepness_Q) == (1.0, 1.0, 10.0, 1.0)
    assert abs(w.dir_x - np.float32(1 / np.sqrt(np.float32(2)))) < 1e-7 and w.dir_x == w.dir_z              # normalised (:398)
    w = api.Wave(-1.0, -0.5, -3.0, 1.7, 0.0, 0.0)
    L.wt_wave_clamp_ui(C.byref(w))
    assert (w.amplitude_A, w.frequency_w, w.phaseConstant_phi, w.steepness_Q, w.dir_x, w.dir_z) == (0.0, 0.0, 0.0, 1.0, 0.0, 1.0)   # :397 zero vector -> (0,1)
    w = api.Wave(2.0, 3.0, 12.0, -0.2, -5.0, 0.0)
    L.wt_wave_clamp_ui(C.byref(w))
    assert (w.amplitude_A, w.frequency_w, w.phaseConstant_phi, w.steepness_Q, w.dir_x, w.dir_z) == (2.0, 3.0, 12.0, 0.0, -1.0, 0.0)
    # a clamped wave always passes the GerstnerWave constructor's constraints (render-engine.h:147-154) ...
    rng = np.random.default_rng(3)
    for _ in range(200):
        v = rng.uniform(-2, 12, size=6)
        w = api.Wave(*[float(x) for x in v])
        L.wt_wave_clamp_ui(C.byref(w))
        assert w.amplitude_A >= 0 and w.frequency_w >= 0 and w.phaseConstant_phi >= 0 and 0 <= w.steepness_Q <= 1
        assert abs(np.hypot(w.dir_x, w.dir_z) - 1) <= 1e-3
    # animation, sky, compose
    a = api.default_animation()
    a.animationSpeedTimeOfDayInSecondsPerHour, a.animationSpeedVerticalBounceWavePhasePeriodInSeconds = -1.0, 60.0
    L.wt_animation_clamp_ui(C.byref(a))
    assert (a.animationSpeedTimeOfDayInSecondsPerHour, a.animationSpeedVerticalBounceWavePhasePeriodInSeconds) == (0.0, 60.0)
    s = api.default_sky_params()
    s.cloudProportion, s.overcastStrength, s.sunHorizonDarkness, s.sunShininess, s.sunStrength, s.timeOfDayInHours = 1.5, -0.5, 1.0, 201.0, 2.0, 24.0
    s.fogColourFarAtNoon[:] = [2.0, -1.0, 0.5, 1.0]
    L.wt_sky_params_clamp_ui(C.byref(s))
    assert (s.cloudProportion, s.overcastStrength, s.sunHorizonDarkness, s.sunShininess, s.sunStrength, s.timeOfDayInHours) == (1.0, 0.0, 1.0, 201.0, 1.0, 24.0)
    assert list(s.fogColourFarAtNoon) == [1.0, 0.0, 0.5, 1.0]
    c = api.default_compose_params(api.default_camera(), 8, 8)
    c.fogDepthRadiusFar, c.fogDepthRadiusNear = 0.25, 0.75                          # the FAR slider moved below near: near follows (:259-264)
    L.wt_compose_params_clamp_ui(C.byref(c), 0)
    assert (c.fogDepthRadiusFar, c.fogDepthRadiusNear) == (0.25, 0.25)
    c.fogDepthRadiusFar, c.fogDepthRadiusNear = 0.25, 0.75                          # the NEAR slider moved above far: far follows (:266-271)
    L.wt_compose_params_clamp_ui(C.byref(c), 1)
    assert (c.fogDepthRadiusFar, c.fogDepthRadiusNear) == (0.75, 0.75)
    c.tintDeltaDepthThreshold, c.waterClarity, c.softEdgesDeltaDepthThreshold = 1.0, 0.0, 7.0
    L.wt_compose_params_clamp_ui(C.byref(c), 0)
    assert (c.tintDeltaDepthThreshold, c.waterClarity, c.softEdgesDeltaDepthThreshold) == (1.0, 0.0, 1.0)


def test_division_by_a_uniform_divisor_is_ieee_division():
    """wt_kernels.cuh:div_by_uniform — x / c as q = fl(x rc), q' = fma(fma(-q, c, x), rc, q) with rc = fl(1 / c) (the compose
    pass uses it for the divisors that are the same for every vertex of a launch).  Emulated with exact rational arithmetic
    (one rounding per fp32 operation) against correctly rounded division, on the operand ranges the pass feeds it."""
    from fractions import Fraction
    f32 = np.float32

    def rn(fr):            # round a Fraction to the nearest fp32 (ties to even), via float64 only when that is exact enough
        d = float(fr)      # correctly rounded to fp64
        r = f32(d)
        # double rounding: fix up by comparing the two neighbouring fp32 values exactly
        lo, hi = np.nextafter(r, f32(-np.inf)), np.nextafter(r, f32(np.inf))
        best = min((abs(Fraction(float(c)) - fr), c) for c in (lo, r, hi) if np.isfinite(c))
        cands = [c for c in (lo, r, hi) if np.isfinite(c) and abs(Fraction(float(c)) - fr) == best[0]]
        if len(cands) == 1:
            return cands[0]
        return [c for c in cands if (np.frombuffer(f32(c).tobytes(), np.uint32)[0] & 1) == 0][0]

    rng = np.random.default_rng(11)
    divisors = [f32(v) for v in (100.0, 1920.0, 1080.0, 1024.0, 0.1, 0.05, 1.0, 0.3, 3.0, 7.0, 96.0, 64.0, 0.999, 1e-3, 2.4)] + \
               [f32(v) for v in rng.uniform(1e-3, 4e3, size=40)]
    bad = 0
    for c in divisors:
        rc = f32(1.0) / c
        xs = np.concatenate([rng.uniform(0.0, 2.0 * float(c), size=150), rng.uniform(0.0, 1.0, size=100), [0.0, float(c), 1.0]]).astype(f32)
        for x in xs:
            q = rn(Fraction(float(x)) * Fraction(float(rc)))
            rem = rn(Fraction(float(x)) - Fraction(float(q)) * Fraction(float(c)))
            q2 = rn(Fraction(float(q)) + Fraction(float(rem)) * Fraction(float(rc)))
            want = rn(Fraction(float(x)) / Fraction(float(c)))
            bad += int(q2 != want)
    assert bad == 0, bad


def test_fog_quad_table_is_the_gl_blend_per_byte():
    """wt_debug_fog_table — the table the skybox synthesis applies for the fog quad — against a numpy restatement of the GL
    blend (GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA on an RGBA8 attachment: operands clamped, Cs As + Cd (1 - As) on all four
    channels, round to nearest on the way back to unorm8; GL 4.1 sections 4.1.7, 2.1.5), one fp32 operation at a time."""
    f32 = np.float32
    d = np.arange(256, dtype=np.float32)
    for fog in ((1.0, 1.0, 1.0, 0.1), (0.3, 0.8, 0.9, 0.35), (0.0, 0.0, 0.0, 1.0), (0.7, 0.2, 0.1, 0.0), (1.5, -0.2, 0.5, 0.999), (0.25, 0.5, 0.75, 0.5)):
        src = np.clip(np.array(fog, f32), f32(0), f32(1))
        a_s = src[3]
        oma = f32(1.0) - a_s
        want = np.zeros((256, 4), np.uint8)
        for c in range(4):
            t0 = f32(src[c] * a_s)
            t1 = ((d / f32(255.0)).astype(f32) * oma).astype(f32)
            v = np.clip((t0 + t1).astype(f32), f32(0), f32(1))
            want[:, c] = np.floor(((v * f32(255.0)).astype(f32) + f32(0.5)).astype(f32)).astype(np.uint8)
        got = api.fog_table(fog)
        assert np.array_equal(got, want), fog
    assert np.array_equal(api.fog_table((0.4, 0.5, 0.6, 0.0)), np.repeat(np.arange(256, dtype=np.uint8)[:, None], 4, 1))   # alpha 0: identity
