"""Direct bit-for-bit comparison of the oracle restatement with the reference's own code compiled
headless (oracle/_ref).  Runs wherever oracle/_ref exists (build container: rebuilt from
/root/reference; GPU box: the prebuilt .so travels with the snapshot); skipped otherwise."""
import numpy as np

from oracle import oracle as orc
from tests.common import Scene
from wave_tool_b200 import api, scenes


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_random_cameras_bit_exact(oracle, ref):
    rng = np.random.default_rng(7)
    n_grid = 0
    for _ in range(300):
        cam = scenes.camera(pos=(rng.uniform(-200, 200), rng.uniform(-20, 120), rng.uniform(-200, 200)),
                            yaw=rng.uniform(-360, 720), pitch=rng.uniform(-89, 89), fov=rng.uniform(15, 105),
                            aspect=rng.choice([1.0, 16 / 9, 0.75]))
        ocam = orc.Camera.from_buffer_copy(bytes(cam))
        D = np.float32(rng.uniform(0.05, 4.0))
        st_o, g_o = oracle.project_grid(ocam, D)
        st_r, g_r = ref.project_grid(ocam, D)
        assert st_o == st_r
        assert bytes(g_o) == bytes(g_r)
        n_grid += st_o == 0
    assert 50 < n_grid < 300      # both outcomes exercised


def test_vertex_shader_random_scenes_bit_exact(oracle, ref, heightmap):
    rng = np.random.default_rng(11)
    for case in range(6):
        W = [1, 3, 4, 8, 37, 256][case]
        L = [2, 17, 513, 640, 1024, 4096][case]
        waves = scenes.default_waves()[:W] if W <= 4 else scenes.seeded_waves(W, cfg=case)
        p = api.default_params()
        p.waveAnimationTimeInSeconds = float(rng.uniform(0, 20))
        p.verticalBounceWavePhase = float(rng.uniform(0, 1))
        p.heightmapSampleScale = float(rng.choice([0.02, 0.13, 1.7]))
        p.heightmapDisplacementScale = float(rng.choice([1.0, 0.25, 0.0]))
        cam = scenes.camera(**list(scenes.EDGE_CAMERAS.values())[case])
        sc = Scene(waves, cam, p, L, heightmap)
        st, u, wu = sc.oracle_uniforms(oracle)
        assert st == 0
        ids = np.unique(np.concatenate([rng.integers(0, L * L, 400), [0, L - 1, L * (L - 1), L * L - 1]]))
        o = oracle.eval(u, wu, heightmap, ids=ids, extra=True)
        r = ref.vert_run(u, wu, heightmap, ids=ids)
        assert np.array_equal(bits(o["normal"][:, :3]), bits(r["normal"])), case
        assert np.array_equal(bits(o["extra"][:, 0:4]), bits(r["gl_Position"])), case
        assert np.array_equal(bits(o["extra"][:, 6:9]), bits(r["viewVecRaw"])), case
        assert np.array_equal(bits(o["extra"][:, 9:12]), bits(r["normalVecInViewSpaceOnlyYaw"])), case
        assert np.array_equal(bits(o["extra"][:, 12:16]), bits(r["heightmap_colour"])), case
        u.viewProjection[:] = np.eye(4, dtype=np.float32).ravel()
        rw = ref.vert_run(u, wu, heightmap, ids=ids)
        assert np.array_equal(bits(o["pos"]), bits(rw["gl_Position"])), case


def test_integer_vertex_id_semantics_match_shader_up_to_4096(oracle, ref, heightmap):
    """water-grid.vert:125 derives (col,row) with a float mod(); exact for gridLength <= 4096 (2^24 ids).
    The oracle/kernels use exact integers — identical there, and well defined beyond."""
    L = 4096
    sc = Scene(scenes.seeded_waves(8, cfg=2), scenes.camera(), api.default_params(), L, heightmap)
    st, u, wu = sc.oracle_uniforms(oracle)
    ids = np.array([L * L - 1, L * L - L, L * 4095 + 4094, 16777215, 12345678, 4095, 4096, 8191], np.int64)
    u.viewProjection[:] = np.eye(4, dtype=np.float32).ravel()
    o = oracle.eval(u, wu, heightmap, ids=ids)
    r = ref.vert_run(u, wu, heightmap, ids=ids)
    assert np.array_equal(bits(o["pos"]), bits(r["gl_Position"]))


def test_fragment_shader_bit_exact(oracle, ref, heightmap):
    """The oracle's fragment restatement — shade_vertex (water-grid.frag:71-73,95-111) + compose_vertex
    (:66-69,76-89,113-162) — against the reference's own water-grid.frag text compiled as C++
    (oracle/_ref/libwt_ref_frag.so), fed the same un-interpolated varyings and the same gl_FragCoord."""
    from tests.common import compose_uniforms, screen_inputs
    assert ref.frag is not None
    f = np.float32
    rng = np.random.default_rng(3)
    for case, (name, kw) in enumerate(list(scenes.EDGE_CAMERAS.items())[:5] + [("tilted", dict(pos=(3.0, 11.0, 60.0), yaw=-75.0, pitch=-17.0, aspect=1.6))]):
        p = api.default_params()
        p.waveAnimationTimeInSeconds = float(rng.uniform(0, 20))
        p.timeOfDayInHours = float([9.0, 12.0, 17.5, 5.0, 20.0, 13.0][case])
        p.sunShininess = float([50.0, 1.0, 200.0, 0.0, 8.0, 50.0][case])
        sky = scenes.noise_skybox(16, seed=case)
        L = 192
        sc = Scene(scenes.seeded_waves(8, cfg=2), scenes.camera(**kw), p, L, heightmap, skybox=sky)
        st, u, wu = sc.oracle_uniforms(oracle)
        if st != 0:
            continue
        w, h = [(96, 64), (33, 17), (128, 128), (7, 5), (64, 96), (96, 64)][case]
        depth, refl, refr = screen_inputs(w, h, seed=case)
        over = [dict(), dict(tintDeltaDepthThreshold=0.0), dict(softEdgesDeltaDepthThreshold=0.0), dict(fogDepthRadiusNear=0.2, fogDepthRadiusFar=0.6),
                dict(waterClarity=1.0), dict()][case]
        cu = compose_uniforms(sc, oracle, w, h, **over)
        ids = np.unique(rng.integers(0, L * L, 3000))
        o = oracle.eval(u, wu, heightmap, sky, ids=ids, shade=True, extra=True)
        ex = o["extra"]
        colour_o = oracle.compose(cu, depth, refl, refr, o["pos"], o["normal"], o["shade"], ex[:, 0:4], ex[:, 4:8], np.concatenate([ex[:, 9:12], ex[:, :1]], 1))
        # the fragment inputs a rasteriser would hand over, un-interpolated; gl_FragCoord by the viewport transform
        fi = np.zeros(len(ids), orc.REF_FRAG_IN_DTYPE)
        clip = ex[:, 0:4]
        nd = (clip[:, :3] / clip[:, 3:4]).astype(f)
        fi["gl_FragCoord"][:, 0] = (nd[:, 0] + f(1)) * (f(w) / f(2))
        fi["gl_FragCoord"][:, 1] = (nd[:, 1] + f(1)) * (f(h) / f(2))
        fi["gl_FragCoord"][:, 2] = (nd[:, 2] + f(1)) / f(2)
        fi["gl_FragCoord"][:, 3] = f(1) / clip[:, 3]
        fi["heightmap_colour"] = ex[:, 12:16]
        fi["normal"] = o["normal"][:, :3]
        fi["normalVecInViewSpaceOnlyYaw"] = ex[:, 9:12]
        fi["viewVecRaw"] = ex[:, 6:9]
        fi["xyPositionNDCSpaceHeight0"] = ex[:, 4:6]
        fu = orc.RefFragUniforms()
        for k_ in ("fogColourFarAtCurrentTime", "viewportWidthHeight", "sunPosition"):
            getattr(fu, k_)[:] = getattr(cu, k_)[:]
        for k_ in ("fogDepthRadiusFar", "fogDepthRadiusNear", "softEdgesDeltaDepthThreshold", "tintDeltaDepthThreshold", "waterClarity", "zFar", "zNear"):
            setattr(fu, k_, getattr(cu, k_))
        fu.sunShininess, fu.sunStrength = u.sunShininess, u.sunStrength
        colour_r = ref.frag_run(fu, depth, refl, refr, sky, fi)
        ok = np.isfinite(colour_r).all(axis=1)
        assert ok.mean() > 0.99
        same = bits(colour_o[ok]) == bits(colour_r[ok])
        assert same.all(), (name, int((~same).sum()), float(np.abs(colour_o[ok] - colour_r[ok]).max()))


def test_sky_layer_shaders_bit_exact(oracle, ref):
    """The oracle's restatement of skysphere.frag / skybox-clouds.frag / skybox-stars.frag against the reference's own
    shader texts compiled as C++ (oracle/_ref/libwt_ref_sky.so), on random directions and times of day."""
    from tests.common import sky_sources, sky_uniforms
    assert ref.sky is not None
    rng = np.random.default_rng(17)
    gradient, stars, clouds = sky_sources(seed=3)
    dirs = rng.normal(size=(4000, 3)).astype(np.float32) * rng.uniform(0.2, 3.0, size=(4000, 1)).astype(np.float32)
    for hours, cp, oc, dark, shin in ((9.0, 0.6, 0.0, 0.25, 50.0), (12.0, 0.9, 0.7, 0.0, 1.0), (18.5, 0.3, 1.0, 1.0, 200.0), (2.0, 0.0, 0.3, 0.5, 0.0), (6.0, 1.0, 0.5, 0.25, 8.0)):
        su = sky_uniforms(oracle, hours, cloud_proportion=cp, overcast=oc, darkness=dark, shininess=shin)
        for which in ("skysphere", "clouds"):
            o = oracle.sky_layer(which, su, gradient, stars, clouds, dirs)
            r = ref.sky_layer_run(which, su, gradient, stars, clouds, dirs)
            ok = np.isfinite(r).all(axis=1)
            assert ok.mean() > 0.99
            assert np.array_equal(bits(o[ok]), bits(r[ok])), (which, hours, float(np.abs(o[ok] - r[ok]).max()))
        r = ref.sky_layer_run("stars", su, gradient, stars, clouds, dirs)
        assert np.all(r[:, 3] == 1.0) and r[:, :3].std() > 0.01
