"""Host-side check of the arithmetic behind the recurrence mode (WT_FLAG_FAST_PHASE) of surface_warp_kernel
(wave_tool_b200/csrc/wt_kernels.cuh, MODE 1 wave loop): one seed e^{i k0} per wave and step, row 1 by one complex
rotation with e^{i dk}, rows 2.. by the three-term recurrence x[j+1] = 2 cos(dk) x[j] - x[j-1].

The kernel itself is checked against the oracle on the GPU (tests/test_gpu_parity.py).  This file pins the ERROR
BOUND DESIGN.md §4 states for the scheme, in an fp32 emulation of the very instruction sequence (every FMA rounded
once), for the two step lengths the kernel uses (8 rows up to 64 waves, 12 rows above) and phase steps from 1e-7 rad
(a wave travelling along the grid rows) to pi (a wave sampled at its Nyquist rate by the far rows of the grid)."""
import numpy as np
import pytest

f32 = np.float32


def fma(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 values is exact in fp64; one rounding to fp32 at the end."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def mul(a, b):
    return (a.astype(np.float64) * b.astype(np.float64)).astype(f32)


def chain(k0, dk, rows, seed_err=0.0):
    """cos/sin of k0 + j dk for j < rows the way the kernel's MODE 1 loop produces them.  The seeds are the exact
    values rounded to fp32 (+ seed_err, standing in for the MUFU approximation error)."""
    c = (np.cos(k0) + seed_err).astype(f32)
    s = (np.sin(k0) - seed_err).astype(f32)
    cd = np.cos(dk).astype(f32)
    sd = np.sin(dk).astype(f32)
    out_c, out_s = [c], [s]
    cp, sp = c, s
    c = fma(cp, cd, mul(-sp, sd))
    s = fma(sp, cd, mul(cp, sd))
    out_c.append(c); out_s.append(s)
    tc = (cd + cd).astype(f32)
    for _ in range(2, rows):
        cn, sn = fma(tc, c, -cp), fma(tc, s, -sp)
        cp, sp, c, s = c, s, cn, sn
        out_c.append(c); out_s.append(s)
    return np.stack(out_c), np.stack(out_s)


def phases(n, seed):
    rng = np.random.default_rng(seed)
    k0 = rng.uniform(-np.pi, np.pi, n)
    # log-uniform magnitudes from 1e-7 to pi, both signs, plus the exact end points
    dk = np.exp(rng.uniform(np.log(1e-7), np.log(np.pi), n)) * rng.choice([-1.0, 1.0], n)
    dk[:4] = [0.0, np.pi, -np.pi, 2.0 ** -12]
    # what the kernel is handed is an fp32 angle; the truth is evaluated on that same angle
    return k0.astype(f32).astype(np.float64), dk.astype(f32).astype(np.float64)


@pytest.mark.parametrize("rows,bound", [(8, 2.5e-6), (12, 6.0e-6)])
def test_three_term_chain_stays_within_the_design_bound(rows, bound):
    k0, dk = phases(200000, seed=rows)
    c, s = chain(k0, dk, rows)
    j = np.arange(rows)[:, None]
    err = np.maximum(np.abs(c - np.cos(k0[None, :] + j * dk[None, :])), np.abs(s - np.sin(k0[None, :] + j * dk[None, :])))
    per_row = err.max(axis=1)
    print("max error per row", per_row)
    assert per_row.max() <= bound
    # the error grows with the distance from the seed, not faster than quadratically
    assert per_row[1] <= 2.5e-7 and per_row[-1] >= per_row[1]


def test_seed_error_is_amplified_at_most_linearly():
    """An error e of the seed (MUFU: ~2^-21) reaches row n as at most ~ (2n - 1) e: U_{n-1}(cos dk) e_1 - U_{n-2}(cos dk) e_0."""
    rows, e = 12, 2.0 ** -21
    k0, dk = phases(100000, seed=3)
    c0, s0 = chain(k0, dk, rows)
    c1, s1 = chain(k0, dk, rows, seed_err=e)
    growth = np.maximum(np.abs(c1.astype(np.float64) - c0), np.abs(s1.astype(np.float64) - s0)).max(axis=1) / e
    print("growth of a seed error per row", growth)
    assert np.all(growth <= 2.0 * np.arange(rows) + 2.0)


def test_reference_order_accumulation_is_what_parity_needs():
    """Why every row's terms are added in the reference's order, starting from the grid position: the
    reference's own sequential fp32 sum is further from the exact sum than the parity bar allows once the position is
    large against the amplitudes, so only the same order of additions lands on the same roundings."""
    rng = np.random.default_rng(11)
    W, pos = 1000, f32(68.0)
    amp = rng.uniform(0.2, 1.0, W); amp /= amp.sum()            # sum(A) = 1: the parity bar is 1e-4
    worst_seq, worst_sum = 0.0, 0.0
    for trial in range(200):
        terms = (amp * np.cos(rng.uniform(-np.pi, np.pi, W))).astype(f32)
        exact = float(pos) + terms.astype(np.float64).sum()
        acc = pos
        for t in terms:
            acc = f32(acc + t)                                   # the shader: position += displacement, wave by wave
        same_order = pos
        for t in (terms.astype(np.float64) * (1.0 + 1e-7)).astype(f32):   # a kernel whose terms differ in the last bit
            same_order = f32(same_order + t)
        summed_first = f32(pos + f32(terms.astype(np.float64).sum()))    # e.g. the symmetric-pair form: sum, then add
        worst_seq = max(worst_seq, abs(float(same_order) - float(acc)))
        worst_sum = max(worst_sum, abs(float(summed_first) - float(acc)))
    print("vs the reference's fp32 result: same order", worst_seq, " summed first", worst_sum)
    assert worst_seq <= 1e-4          # following the reference's order stays inside the bar ...
    assert worst_sum > worst_seq      # ... summing first (more exact!) is further from the reference


# --------------------------------------------------------------------------------------------------
# The whole MODE 1 wave loop, emulated in fp32 for a few grid columns, against the ORACLE (x and z: the heightmap
# detail and the bounce only touch y).  No GPU: this pins the algorithm — seed row, phase step, three-term chain and
# the reference-order accumulation — to the reference's own arithmetic; the kernel is pinned to the same oracle on
# the GPU (tests/test_gpu_parity.py).
# --------------------------------------------------------------------------------------------------
def _emulate_column(frame, waves, L, col, R):
    """x, z of every row of one grid column the way surface_warp_kernel<MODE 1> computes them (steps of R rows)."""
    bl = np.array(frame.bottomLeftGridPointInWorld[:], f32); br = np.array(frame.bottomRightGridPointInWorld[:], f32)
    tl = np.array(frame.topLeftGridPointInWorld[:], f32); tr = np.array(frame.topRightGridPointInWorld[:], f32)
    t = f32(frame.waveAnimationTimeInSeconds)
    A, w, phi, Q, Dx, Dz = [waves[:, i].astype(f32) for i in range(6)]
    W = len(A)
    Qi = (Q / (w * A * f32(max(W, 4)))).astype(f32)             # render-engine.cpp:1073
    cX, cZ = ((A * Qi).astype(f32) * Dx).astype(f32), ((A * Qi).astype(f32) * Dz).astype(f32)
    wDx, wDz = (w * Dx).astype(f32), (w * Dz).astype(f32)
    ph = np.remainder((phi * t).astype(f32).astype(np.float64) + np.pi, 2 * np.pi) - np.pi
    ph = ph.astype(f32)
    du = f32(1.0) / f32(L - 1)

    def mix(a, b, x):
        return f32(f32(a * f32(f32(1) - x)) + f32(b * x))
    u = f32(col) * du
    m1 = (mix(bl[0], br[0], u), mix(bl[2], br[2], u))
    m2 = (mix(tl[0], tr[0], u), mix(tl[2], tr[2], u))
    ex, ez = f32(f32(m2[0] - m1[0]) * du), f32(f32(m2[1] - m1[1]) * du)
    out = np.zeros((L, 2), f32)
    for row_a in range(0, L, R):
        rows = [r for r in range(row_a, min(row_a + R, L))]
        gx = np.array([mix(m1[0], m2[0], f32(r) * du) for r in rows], f32)
        gz = np.array([mix(m1[1], m2[1], f32(r) * du) for r in rows], f32)
        k0 = fma(wDx, np.full(W, gx[0], f32), fma(wDz, np.full(W, gz[0], f32), ph))
        n = np.rint(k0.astype(np.float64) / (2 * np.pi))
        r0 = (k0.astype(np.float64) - n * 2 * np.pi).astype(f32)       # Cody-Waite: exact to fp32 rounding
        dk = fma(wDx, np.full(W, ex, f32), mul(wDz, np.full(W, ez, f32)))
        c, s = chain(r0.astype(np.float64), dk.astype(np.float64), len(rows)) if len(rows) > 1 else (np.cos(r0)[None].astype(f32), np.sin(r0)[None].astype(f32))
        for j in range(len(rows)):
            x, z = gx[j], gz[j]
            for i in range(W):                                          # wave order, starting from the grid position
                x = f32(np.float64(cX[i]) * np.float64(c[j][i]) + np.float64(x))
                z = f32(np.float64(cZ[i]) * np.float64(c[j][i]) + np.float64(z))
            out[rows[j]] = (x, z)
    return out


@pytest.mark.parametrize("W,R", [(16, 8), (64, 8), (256, 12)])
def test_emulated_recurrence_loop_matches_the_oracle(W, R):
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import Scene
    from oracle import oracle as orc
    from wave_tool_b200 import scenes
    L = 96
    waves = scenes.seeded_waves(W, cfg=9)
    sc = Scene(waves, scenes.camera(), scenes.animation_params(5), L, scenes.procedural_heightmap(64))
    st, frame = sc.frame()
    assert st == 0
    O = orc.Oracle()
    tol = 1e-4 * sc.total_amplitude
    for col in (0, 58, 95):
        got = _emulate_column(frame, sc.waves, L, col, R)
        want = sc.oracle_eval(O, ids=np.arange(L) * L + col)["pos"]
        err = np.abs(got.astype(np.float64) - want[:, [0, 2]].astype(np.float64)).max()
        print(f"W={W} R={R} col={col}: max |x,z error| = {err:.2e} (bar {tol:.1e})")
        assert err <= tol
