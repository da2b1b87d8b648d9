"""Generates the committed golden fixtures from the REFERENCE ITSELF.

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py

It needs oracle/_ref/*.so (make -C oracle ref), i.e. the reference's camera.cpp / render-engine.h / GLM
and its water-grid.vert shader text compiled as C++.  The reference ships no tests or golden vectors
for this path, so these outputs of the reference code are what pins the oracle; the
GPU box has no /root/reference and uses the files written here.

Outputs (tests/golden/):
  heightmap_waves3_00.npz   the reference heightmap asset as uploaded to GL (R plane, rows flipped)
  projector.npz             camera matrices + projected-grid corners for every edge camera
  scalars.npz               Q_i / sun / bounce / total-amplitude rules
  vert_default.npz          water-grid.vert outputs, default scene (513^2, 4 wave slots), t in {0, 1.5}
  vert_seeded8.npz          water-grid.vert outputs, 1024^2 grid, 8 seeded waves, t = 0.5
  sky_layers.npz            skysphere.frag / skybox-clouds.frag outputs on random directions, four times of day
  frag_default.npz          water-grid.frag outputs (final colour) for vertices of the default scene at t = 1.5, fed the
                            vertex stage's own varyings + synthetic screen-space inputs (stored alongside)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from wave_tool_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"


def load_reference_heightmap():
    from PIL import Image
    img = np.array(Image.open(os.path.join(REF_ROOT, "assets/textures/noise/waves/waves3/00.png")))
    assert img.ndim == 2 and img.dtype == np.uint8          # 8-bit grey; stbi expands to (g,g,g,255)
    return np.flipud(img).copy()                             # stbi_set_flip_vertically_on_load(true), render-engine.cpp:1310


def sample_ids(L, step):
    rows = sorted(set(list(range(0, L, step)) + [0, 1, L - 2, L - 1]))
    cols = sorted(set(list(range(0, L, step)) + [0, 1, L - 2, L - 1]))
    return np.array([r * L + c for r in rows for c in cols], np.int64)


def ref_uniforms(R, cam, waves, params, L):
    total = R.total_amplitude(waves[:, 0])
    D = np.float32(np.float32(total + np.float32(params["hm_disp"])) + np.float32(params["bounce_amp"]))
    st, g = R.project_grid(cam, D)
    assert st == 0
    u = orc.Uniforms()
    c = np.array(g.corners, np.float32).reshape(4, 4)
    u.bottomLeft[:] = c[0]; u.topLeft[:] = c[1]; u.bottomRight[:] = c[2]; u.topRight[:] = c[3]
    u.cameraPosition[:] = cam.pos[:]
    u.gerstnerWaveCount = len(waves)
    u.gridLength = L
    u.heightmapDisplacementScale = params["hm_disp"]
    u.heightmapSampleScale = params["hm_scale"]
    u.verticalBounceWaveDisplacement = float(R.bounce_displacement(params["bounce_amp"], params["bounce_phase"]))
    u.viewMatOnlyYaw[:] = g.view_only_yaw[:]
    u.viewProjection[:] = g.view_proj[:]
    u.waveAnimationTimeInSeconds = params["t"]
    wu = np.zeros(len(waves), orc.WAVE_UNIFORM_DTYPE)
    for i, (A, w, phi, Q, dx, dz) in enumerate(waves):
        wu[i] = (A, w, phi, R.steepness_Q_i(Q, w, A, len(waves)), dx, dz)
    return u, wu


def run_vert(R, u, wu, hm4, ids):
    """Two passes of the compiled shader: real viewProjection (all outputs), identity (world position)."""
    out = R.vert_run(u, wu, hm4, ids=ids)
    vp = np.array(u.viewProjection[:], np.float32)
    u.viewProjection[:] = np.eye(4, dtype=np.float32).ravel()
    world = R.vert_run(u, wu, hm4, ids=ids)["gl_Position"].copy()
    u.viewProjection[:] = vp
    return out, world


def main():
    R = orc.Ref()
    assert R.available, "build oracle/_ref first: make -C oracle ref"
    hm = load_reference_heightmap()
    np.savez_compressed(os.path.join(HERE, "heightmap_waves3_00.npz"), r_gl_rows=hm)
    hm4 = np.stack([hm, hm, hm, np.full_like(hm, 255)], -1).copy()

    # ---- projector / camera ----------------------------------------------------------------------
    names, cams, Ds, status, corners, mats, fwd = [], [], [], [], [], [], []
    for name, kw in scenes.EDGE_CAMERAS.items():
        for D in (np.float32(1.26), np.float32(2.1), np.float32(0.35)):
            cam = scenes.camera(**kw)
            ocam = orc.Camera.from_buffer_copy(bytes(cam))
            st, g = R.project_grid(ocam, D)
            names.append(name); cams.append(np.frombuffer(bytes(cam), np.float32).copy()); Ds.append(D); status.append(st)
            corners.append(np.array(g.corners, np.float32)); fwd.append(np.array(g.cam_forward, np.float32))
            mats.append(np.concatenate([np.array(m, np.float32) for m in (g.view, g.proj, g.view_proj, g.view_only_yaw)]))
    np.savez_compressed(os.path.join(HERE, "projector.npz"), names=np.array(names), cameras=np.array(cams), D=np.array(Ds),
                        status=np.array(status), corners=np.array(corners), matrices=np.array(mats), forward=np.array(fwd))

    # ---- scalar rules ----------------------------------------------------------------------------
    rng = np.random.default_rng(1234)
    q_in = rng.random((64, 3)).astype(np.float32) * np.array([1.0, 3.0, 0.5], np.float32)
    q_in[:4, 1] = 0.0   # w*A == 0 -> symbolic 0 (render-engine.cpp:1073)
    counts = rng.integers(1, 257, 64)
    q_out = np.array([R.steepness_Q_i(q, w, a, int(n)) for (q, w, a), n in zip(q_in, counts)], np.float32)
    hours = np.linspace(0, 24, 49).astype(np.float32)
    sun = np.array([R.sun_position(h) for h in hours], np.float32)
    phases = np.linspace(0, 1, 33).astype(np.float32)
    bounce = np.array([R.bounce_displacement(0.1, p) for p in phases], np.float32)
    amps = rng.random(256).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "scalars.npz"), q_in=q_in, q_count=counts, q_out=q_out, hours=hours, sun=sun,
                        phases=phases, bounce=bounce, amps=amps, amp_total=np.float32(R.total_amplitude(amps)))

    # ---- vertex shader, default scene --------------------------------------------------------------
    waves = scenes.default_waves()
    cam = orc.Camera.from_buffer_copy(bytes(scenes.camera()))
    ids = sample_ids(513, 8)
    rec = {}
    for t in (0.0, 1.5):
        params = dict(hm_disp=1.0, hm_scale=0.02, bounce_amp=0.1, bounce_phase=float(np.float32(np.fmod(t / 3.0, 1.0))), t=t)
        u, wu = ref_uniforms(R, cam, waves, params, 513)
        out, world = run_vert(R, u, wu, hm4, ids)
        rec[f"out_t{t}"] = out
        rec[f"world_t{t}"] = world
    np.savez_compressed(os.path.join(HERE, "vert_default.npz"), ids=ids, **rec)

    # ---- vertex shader, 8 seeded waves on a 1024^2 grid --------------------------------------------
    waves8 = scenes.seeded_waves(8, cfg=2)
    ids8 = sample_ids(1024, 16)
    params = dict(hm_disp=1.0, hm_scale=0.02, bounce_amp=0.1, bounce_phase=float(np.float32(np.fmod(0.5 / 3.0, 1.0))), t=0.5)
    u, wu = ref_uniforms(R, cam, waves8, params, 1024)
    out, world = run_vert(R, u, wu, hm4, ids8)
    np.savez_compressed(os.path.join(HERE, "vert_seeded8.npz"), ids=ids8, out=out, world=world, waves=waves8)
    # ---- fragment shader on the default scene's vertex outputs (t = 1.5) -----------------------------
    rng = np.random.default_rng(2024)
    w, h = 96, 64
    depth = rng.uniform(0.90, 1.0, size=(h, w)).astype(np.float32)
    depth[rng.uniform(size=(h, w)) < 0.33] = 1.0
    refl = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    refr = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    refl[..., 3][rng.uniform(size=(h, w)) < 0.4] = 0
    sky = rng.integers(0, 256, size=(6, 16, 16, 4), dtype=np.uint8)
    sel = np.arange(0, len(ids), 7)
    vo, world = rec["out_t1.5"][sel], rec["world_t1.5"][sel]
    f = np.float32
    fu = orc.RefFragUniforms()
    sun = R.sun_position(9.0)
    sy = f(min(max(float(sun[1]), 0.0), 1.0))
    fu.fogColourFarAtCurrentTime[:] = [float(sy)] * 3 + [0.1]          # render-engine.cpp:295 with fogColourFarAtNoon (1,1,1,0.1)
    fu.fogDepthRadiusFar, fu.fogDepthRadiusNear, fu.softEdgesDeltaDepthThreshold = 1.0, 0.0, 0.05
    fu.sunPosition[:] = sun
    fu.sunShininess, fu.sunStrength, fu.tintDeltaDepthThreshold, fu.waterClarity = 50.0, 1.0, 0.1, 0.3
    fu.viewportWidthHeight[:] = [float(w), float(h)]
    fu.zFar, fu.zNear = 100.0, 0.1
    fi = np.zeros(len(sel), orc.REF_FRAG_IN_DTYPE)
    clip = vo["gl_Position"]
    nd = (clip[:, :3] / clip[:, 3:4]).astype(f)
    fi["gl_FragCoord"][:, 0] = (nd[:, 0] + f(1)) * (f(w) / f(2))        # viewport transform, GL 4.1 section 2.14.1
    fi["gl_FragCoord"][:, 1] = (nd[:, 1] + f(1)) * (f(h) / f(2))
    fi["gl_FragCoord"][:, 2] = (nd[:, 2] + f(1)) / f(2)
    fi["gl_FragCoord"][:, 3] = f(1) / clip[:, 3]
    for k in ("heightmap_colour", "normal", "normalVecInViewSpaceOnlyYaw", "viewVecRaw", "xyPositionNDCSpaceHeight0"):
        fi[k] = vo[k]
    colour = R.frag_run(fu, depth, refl, refr, sky, fi)
    np.savez_compressed(os.path.join(HERE, "frag_default.npz"), depth=depth, reflections=refl, refractions=refr, sky=sky, frag_in=fi, world=world,
                        uniforms=np.frombuffer(bytes(fu), np.uint8).copy(), colour=colour)
    # ---- sky layer shaders (skysphere.frag, skybox-clouds.frag) on random directions ------------------
    rng = np.random.default_rng(77)
    G = 64
    tt = np.linspace(0.0, 1.0, G)
    gradient = np.stack([40 + 200 * tt, 90 + 120 * np.sin(3 * tt) ** 2, 255 - 180 * tt, np.full_like(tt, 255)], 1).round().astype(np.uint8)
    clouds = rng.integers(0, 256, size=(6, 12, 12, 4), dtype=np.uint8)
    dirs = (rng.normal(size=(1500, 3)) * rng.uniform(0.2, 3.0, size=(1500, 1))).astype(np.float32)
    cases, sphere_out, cloud_out = [], [], []
    for hours, cp, oc, dark, shin in ((9.0, 0.6, 0.0, 0.25, 50.0), (12.0, 0.9, 0.7, 0.0, 1.0), (18.5, 0.3, 1.0, 1.0, 200.0), (2.0, 0.0, 0.3, 0.5, 0.0)):
        su = orc.SkyUniforms()
        su.sunPosition[:] = R.sun_position(hours)
        su.sunHorizonDarkness, su.sunShininess, su.sunStrength = dark, shin, 1.0
        su.oneMinusCloudProportion = float(np.float32(1.0) - np.float32(cp))
        su.overcastStrength = oc
        cases.append((hours, cp, oc, dark, shin))
        sphere_out.append(R.sky_layer_run("skysphere", su, gradient, None, clouds, dirs))
        cloud_out.append(R.sky_layer_run("clouds", su, gradient, None, clouds, dirs))
    np.savez_compressed(os.path.join(HERE, "sky_layers.npz"), gradient=gradient, clouds=clouds, dirs=dirs, cases=np.array(cases, np.float32),
                        skysphere=np.array(sphere_out), cloud_layer=np.array(cloud_out))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
