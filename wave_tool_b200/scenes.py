"""Synthetic workloads of the water path: the reference's default scene and the seeded scenes of
BASELINE.json's configs.  Pure numpy, deterministic, no asset or oracle dependency —
used by tests/ and bench.py on both the CPU container and the GPU box.

All parameters are produced as float32 ONCE here, so every consumer (CUDA path, oracle, reference
build) reads identical bits.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import api

MASK64 = (1 << 64) - 1


class SplitMix64:
    """splitmix64 (Steele, Lea, Flood 2014) — tiny, portable, reproducible."""

    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def uniform(self) -> float:   # [0,1) with 53 bits
        return (self.next_u64() >> 11) * (1.0 / (1 << 53))


def seeded_waves(count: int, cfg: int, wind_deg: float = 30.0) -> np.ndarray:
    """(count,6) float32 rows [A, w, phi, Q, Dx, Dz].  SURVEY.md §8d: direction within +-60 deg of the
    wind, wavelength log-uniform in [2,64], w = 2pi/lambda, phi = w*sqrt(9.8*lambda/2pi) (deep-water
    dispersion), A proportional to lambda with sum(A) = 1, Q uniform in [0.2,1]."""
    rng = SplitMix64(0x57415645 + cfg)
    lam = np.empty(count)
    ang = np.empty(count)
    q = np.empty(count)
    for i in range(count):
        ang[i] = math.radians(wind_deg + (rng.uniform() * 2.0 - 1.0) * 60.0)
        lam[i] = 2.0 * (32.0 ** rng.uniform())
        q[i] = 0.2 + 0.8 * rng.uniform()
    w = 2.0 * math.pi / lam
    phi = w * np.sqrt(9.8 * lam / (2.0 * math.pi))
    amp = lam / lam.sum()
    rows = np.stack([amp, w, phi, q, np.cos(ang), np.sin(ang)], axis=1).astype(np.float32)
    return rows


def default_waves() -> np.ndarray:
    """The reference's hard-coded set (render-engine.cpp:46-49), from the library's own host code."""
    return api.default_waves()


def procedural_heightmap(n: int = 512, seed: int = 7) -> np.ndarray:
    """(n,n,4) uint8 grey noise texture (R=G=B, A=255): a few octaves of tileable sinusoids plus hash
    noise.  Stands in for the reference's assets/textures/noise/waves/waves3/00.png where the real
    asset is not available (GPU box, bench)."""
    y, x = np.mgrid[0:n, 0:n].astype(np.float64) / n
    rng = SplitMix64(seed)
    h = np.zeros((n, n))
    for octave in range(1, 6):
        for _ in range(3):
            fx, fy = int(rng.uniform() * 4 * octave) + 1, int(rng.uniform() * 4 * octave) + 1
            ph = rng.uniform() * 2 * math.pi
            h += np.sin(2 * math.pi * (fx * x + fy * y) + ph) / octave
    h = (h - h.min()) / (h.max() - h.min())
    grey = np.clip(np.round(45 + h * (207 - 45)), 0, 255).astype(np.uint8)   # the asset's R range is [45,207]
    return np.stack([grey, grey, grey, np.full_like(grey, 255)], axis=-1).copy()


def procedural_skybox(n: int = 2048) -> np.ndarray:
    """(6,n,n,4) uint8 cubemap, rgb = 0.5 + 0.5*direction quantised to RGBA8, faces in
    GL order +x,-x,+y,-y,+z,-z with the GL (s,t) orientation of each face."""
    c = (np.arange(n, dtype=np.float64) + 0.5) / n * 2.0 - 1.0   # texel-centre coordinate in [-1,1]
    sc, tc = np.meshgrid(c, c)                                    # sc varies along s (columns), tc along t (rows)
    one = np.ones_like(sc)
    # inverse of the GL major-axis table: (sc, tc, ma) -> direction
    dirs = [(one, -tc, -sc), (-one, -tc, sc), (sc, one, tc), (sc, -one, -tc), (sc, -tc, one), (-sc, -tc, -one)]
    faces = np.empty((6, n, n, 4), np.uint8)
    for f, (dx, dy, dz) in enumerate(dirs):
        ln = np.sqrt(dx * dx + dy * dy + dz * dz)
        for ch, d in enumerate((dx, dy, dz)):
            faces[f, :, :, ch] = np.clip(np.round((0.5 + 0.5 * d / ln) * 255.0), 0, 255).astype(np.uint8)
        faces[f, :, :, 3] = 255
    return faces


def noise_skybox(n: int, seed: int = 1) -> np.ndarray:
    """(6,n,n,4) uint8 cubemap of independent uniform random texels: maximal contrast between neighbouring
    texels and across face edges, so any mix-up of the bilinear footprint (texel order, face, clamp) shows."""
    return np.random.default_rng(seed).integers(0, 256, size=(6, n, n, 4), dtype=np.uint8)


def camera(pos=(0.0, 4.0, 70.0), yaw=-90.0, pitch=0.0, fov=72.0, aspect=1.0, z_near=0.1, z_far=100.0) -> api.Camera:
    """Camera ctor arguments; defaults are the reference's (render-engine.cpp:52, camera.cpp:40-45)."""
    return api.Camera(fov, aspect, z_near, z_far, (C.c_float * 3)(*pos), yaw, pitch)


# projector edge cases of BASELINE.json configs[4] / SURVEY.md §8d cfg 5.  D denotes the displaceable
# amplitude of the seeded scene (sum(A)=1 + heightmap 1 + bounce 0.1 = 2.1).
EDGE_CAMERAS = {
    "default_horizon": dict(),                                    # pitch 0: SAFE_EPSILON path (:920-926)
    "pitch_plus_0.05": dict(pitch=0.05),
    "pitch_minus_0.05": dict(pitch=-0.05),
    "inside_lift_zone": dict(pos=(0.0, 2.6, 70.0)),               # D < y < D+1: projector lifted (:912-917)
    "inside_slab": dict(pos=(0.0, 1.0, 70.0)),                    # camera inside the displaceable volume
    "underwater": dict(pos=(0.0, -3.0, 70.0)),                    # projector mirrored (:915)
    "high_looking_down": dict(pos=(0.0, 60.0, 70.0), pitch=-30.0),
    "birds_eye": dict(pos=(0.0, 95.0, 70.0), pitch=-89.0),
    "looking_up_no_grid": dict(pos=(0.0, 60.0, 70.0), pitch=60.0),   # frustum misses the slab -> WT_NO_GRID (:894)
    "yawed": dict(yaw=-47.0, pitch=-12.0, pos=(10.0, 7.5, -20.0)),
}


def animation_params(frame_index: int, fps: float = 60.0, bounce_period_s: float = 3.0) -> api.Params:
    """Per-frame RenderEngine fields for an animation: t_f = f/fps, bounce phase = frac(t_f/period)
    (program.cpp:452-464), everything else at the reference defaults."""
    p = api.default_params()
    t = np.float32(frame_index) / np.float32(fps)
    p.waveAnimationTimeInSeconds = float(t)
    p.verticalBounceWavePhase = float(np.float32(math.fmod(float(t) / bounce_period_s, 1.0)))
    return p


def frames_for_animation(cam: api.Camera, waves: np.ndarray, first: int, count: int, stride: int = 1, fps: float = 60.0):
    """wt_frame records for frames first, first+stride, ... (frame sharding passes stride = n_gpus)."""
    total = api.total_amplitude(waves)
    out = []
    for k in range(count):
        st, fr = api.build_frame(cam, animation_params(first + k * stride, fps), total)
        if st != api.WT_OK:
            raise RuntimeError("camera does not see the water volume")
        out.append(fr)
    return out
