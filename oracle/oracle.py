"""TEST INFRASTRUCTURE — ctypes loader for the CPU oracle.  Not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It loads

  * oracle/libwt_oracle.so        — our plain-C++ restatement (oracle/wt_oracle.cpp), fp32 + fp64 twin
  * oracle/_ref/libwt_ref_*.so    — the reference's own camera.cpp / render-engine.h / GLM and its
                                    water-grid.vert text compiled as C++ (oracle/Makefile); optional,
                                    present wherever `make -C oracle ref` ran (it needs /root/reference)

and exposes both behind the same numpy-level interface so tests can compare them bit for bit.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class Camera(C.Structure):
    """Camera ctor arguments (reference src/camera.h:49, src/render-engine.cpp:52)."""
    _fields_ = [("fov_deg", C.c_float), ("aspect", C.c_float), ("z_near", C.c_float), ("z_far", C.c_float),
                ("pos", C.c_float * 3), ("yaw_deg", C.c_float), ("pitch_deg", C.c_float)]


def _grid_struct(ft):
    class Grid(C.Structure):
        _fields_ = [("has_grid", C.c_int), ("n_intersections", C.c_int), ("corners", ft * 16),
                    ("view", ft * 16), ("proj", ft * 16), ("view_proj", ft * 16), ("view_only_yaw", ft * 16),
                    ("cam_forward", ft * 3), ("aimpoint_1", ft * 3), ("aimpoint_2", ft * 3), ("aimpoint_final", ft * 3),
                    ("projector_pitch_deg", ft), ("projector_y", ft), ("range", ft * 4)]
    return Grid


GridF32 = _grid_struct(C.c_float)
GridF64 = _grid_struct(C.c_double)


class Uniforms(C.Structure):
    """Water-program uniforms in scope (water-grid.vert:37-64, water-grid.frag:39-47)."""
    _fields_ = [("bottomLeft", C.c_float * 4), ("bottomRight", C.c_float * 4), ("topLeft", C.c_float * 4), ("topRight", C.c_float * 4),
                ("cameraPosition", C.c_float * 3), ("gerstnerWaveCount", C.c_uint), ("gridLength", C.c_uint),
                ("heightmapDisplacementScale", C.c_float), ("heightmapSampleScale", C.c_float),
                ("verticalBounceWaveDisplacement", C.c_float),
                ("viewMatOnlyYaw", C.c_float * 16), ("viewProjection", C.c_float * 16),
                ("waveAnimationTimeInSeconds", C.c_float),
                ("sunPosition", C.c_float * 3), ("sunShininess", C.c_float), ("sunStrength", C.c_float)]


class RefVertUniforms(C.Structure):
    _fields_ = [("bottomLeft", C.c_float * 4), ("bottomRight", C.c_float * 4), ("topLeft", C.c_float * 4), ("topRight", C.c_float * 4),
                ("cameraPosition", C.c_float * 3), ("gerstnerWaveCount", C.c_uint), ("gridLength", C.c_uint),
                ("heightmapDisplacementScale", C.c_float), ("heightmapSampleScale", C.c_float),
                ("verticalBounceWaveDisplacement", C.c_float),
                ("viewMatOnlyYaw", C.c_float * 16), ("viewProjection", C.c_float * 16),
                ("waveAnimationTimeInSeconds", C.c_float)]


class ComposeUniforms(C.Structure):
    """Uniforms of the screen-space layers + fog of water-grid.frag (:34-55)."""
    _fields_ = [("fogColourFarAtCurrentTime", C.c_float * 4), ("fogDepthRadiusFar", C.c_float), ("fogDepthRadiusNear", C.c_float),
                ("softEdgesDeltaDepthThreshold", C.c_float), ("tintDeltaDepthThreshold", C.c_float), ("waterClarity", C.c_float),
                ("zNear", C.c_float), ("zFar", C.c_float), ("viewportWidthHeight", C.c_float * 2),
                ("cameraPosition", C.c_float * 3), ("sunPosition", C.c_float * 3)]


class SkyUniforms(C.Structure):
    """Inputs of the skybox synthesis (render-engine.cpp:289-297, 333-424)."""
    _fields_ = [("sunPosition", C.c_float * 3), ("sunHorizonDarkness", C.c_float), ("sunShininess", C.c_float), ("sunStrength", C.c_float),
                ("oneMinusCloudProportion", C.c_float), ("overcastStrength", C.c_float), ("fogColourFarAtCurrentTime", C.c_float * 4),
                ("draw_stars", C.c_int), ("draw_skysphere", C.c_int), ("draw_clouds", C.c_int)]


class SkyTextures(C.Structure):
    _fields_ = [("gradient_rgba", C.c_void_p), ("gradient_len", C.c_int), ("stars", C.c_void_p * 6), ("stars_n", C.c_int),
                ("clouds", C.c_void_p * 6), ("clouds_n", C.c_int)]


def sky_textures(gradient, stars, clouds):
    """(SkyTextures, keep-alive list) from a (G,4) gradient and optional (6,n,n,4) star / cloud cubemaps."""
    g = np.ascontiguousarray(gradient, np.uint8).reshape(-1, 4)
    tx = SkyTextures()
    tx.gradient_rgba, tx.gradient_len = g.ctypes.data, len(g)
    keep = [g]
    for name, faces in (("stars", stars), ("clouds", clouds)):
        if faces is not None:
            a = np.ascontiguousarray(faces, np.uint8)
            keep.append(a)
            for f in range(6):
                getattr(tx, name)[f] = a[f].ctypes.data
            setattr(tx, name + "_n", a.shape[1])
    return tx, keep


class ScreenTextures(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("reflections", C.c_void_p), ("refractions", C.c_void_p), ("w", C.c_int), ("h", C.c_int)]


class RefFragUniforms(C.Structure):
    _fields_ = [("fogColourFarAtCurrentTime", C.c_float * 4), ("fogDepthRadiusFar", C.c_float), ("fogDepthRadiusNear", C.c_float),
                ("softEdgesDeltaDepthThreshold", C.c_float), ("sunPosition", C.c_float * 3), ("sunShininess", C.c_float),
                ("sunStrength", C.c_float), ("tintDeltaDepthThreshold", C.c_float), ("viewportWidthHeight", C.c_float * 2),
                ("waterClarity", C.c_float), ("zFar", C.c_float), ("zNear", C.c_float)]


REF_FRAG_IN_DTYPE = np.dtype([("gl_FragCoord", "f4", 4), ("heightmap_colour", "f4", 4), ("normal", "f4", 3),
                              ("normalVecInViewSpaceOnlyYaw", "f4", 3), ("viewVecRaw", "f4", 3), ("xyPositionNDCSpaceHeight0", "f4", 2)])


class Textures(C.Structure):
    _fields_ = [("heightmap_rgba", C.c_void_p), ("hm_w", C.c_int), ("hm_h", C.c_int),
                ("sky_faces", C.c_void_p * 6), ("sky_n", C.c_int)]


# numpy dtype of one shader-side wave record: A, w, phi, Q_i, Dx, Dz
WAVE_UNIFORM_DTYPE = np.dtype([("A", "f4"), ("w", "f4"), ("phi", "f4"), ("Q_i", "f4"), ("Dx", "f4"), ("Dz", "f4")])
REF_VERT_OUT_DTYPE = np.dtype([("gl_Position", "f4", 4), ("heightmap_colour", "f4", 4), ("normal", "f4", 3),
                               ("normalVecInViewSpaceOnlyYaw", "f4", 3), ("viewVecRaw", "f4", 3),
                               ("xyPositionNDCSpaceHeight0", "f4", 2)])


def build(ref: bool = True) -> None:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref) with oracle/Makefile."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE, "-s"] + targets, check=True)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


class Oracle:
    """Our restatement.  fp32 is the parity oracle, fp64 the truth twin."""

    def __init__(self):
        path = os.path.join(HERE, "libwt_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        L = self.lib
        L.wto_total_amplitude.restype = C.c_float
        L.wto_steepness_Q_i.restype = C.c_float
        L.wto_steepness_Q_i.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint]
        L.wto_bounce_displacement.restype = C.c_float
        L.wto_bounce_displacement.argtypes = [C.c_float, C.c_float]
        L.wto_sun_position.argtypes = [C.c_float, c_float_p]
        L.wto_project_grid_f32.argtypes = [C.POINTER(Camera), C.c_float, C.POINTER(GridF32)]
        L.wto_project_grid_f64.argtypes = [C.POINTER(Camera), C.c_double, C.POINTER(GridF64)]
        common = [C.POINTER(Uniforms), C.c_void_p, C.POINTER(Textures), C.c_void_p, C.c_longlong,
                  C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong]
        L.wto_eval_f32.argtypes = common + [C.c_void_p] * 4
        L.wto_eval_f64.argtypes = common + [C.c_void_p] * 4
        L.wto_shade_f32.argtypes = [C.POINTER(Uniforms), C.POINTER(Textures), C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.wto_compose_f32.argtypes = [C.POINTER(ComposeUniforms), C.POINTER(ScreenTextures), C.c_longlong] + [C.c_void_p] * 7
        L.wto_sky_synthesize.argtypes = [C.POINTER(SkyUniforms), C.POINTER(SkyTextures), C.c_int, C.c_void_p]
        L.wto_skysphere_colour_f32.argtypes = [C.POINTER(SkyUniforms), C.POINTER(SkyTextures), C.c_longlong, C.c_void_p, C.c_void_p]
        L.wto_clouds_colour_f32.argtypes = [C.POINTER(SkyUniforms), C.POINTER(SkyTextures), C.c_longlong, C.c_void_p, C.c_void_p]

    # --- host-side scalar rules -------------------------------------------------------------
    def num_threads(self) -> int:
        return int(self.lib.wto_num_threads())

    def set_num_threads(self, n: int) -> None:
        self.lib.wto_set_num_threads(int(n))

    def total_amplitude(self, amplitudes) -> np.float32:
        a = np.ascontiguousarray(amplitudes, dtype=np.float32)
        return np.float32(self.lib.wto_total_amplitude(a.ctypes.data_as(c_float_p), C.c_uint(a.size)))

    def steepness_Q_i(self, Q, w, A, count) -> np.float32:
        return np.float32(self.lib.wto_steepness_Q_i(float(Q), float(w), float(A), int(count)))

    def bounce_displacement(self, amplitude, phase) -> np.float32:
        return np.float32(self.lib.wto_bounce_displacement(float(amplitude), float(phase)))

    def sun_position(self, hours) -> np.ndarray:
        out = np.zeros(3, np.float32)
        self.lib.wto_sun_position(float(hours), out.ctypes.data_as(c_float_p))
        return out

    def wave_uniforms(self, waves: np.ndarray) -> np.ndarray:
        """(W,6) reference-native rows [A, w, phi, Q, Dx, Dz] -> shader-side records with Q_i
        (src/render-engine.cpp:1073, Count() == W)."""
        waves = np.asarray(waves, np.float32).reshape(-1, 6)
        out = np.zeros(len(waves), WAVE_UNIFORM_DTYPE)
        for i, (A, w, phi, Q, dx, dz) in enumerate(waves):
            out[i] = (A, w, phi, self.steepness_Q_i(Q, w, A, len(waves)), dx, dz)
        return out

    # --- projected grid -----------------------------------------------------------------------
    def project_grid(self, cam: Camera, displaceable_amplitude, f64: bool = False):
        if f64:
            g = GridF64()
            st = self.lib.wto_project_grid_f64(C.byref(cam), C.c_double(float(displaceable_amplitude)), C.byref(g))
        else:
            g = GridF32()
            st = self.lib.wto_project_grid_f32(C.byref(cam), C.c_float(float(displaceable_amplitude)), C.byref(g))
        return st, g

    # --- per-vertex evaluation ----------------------------------------------------------------
    def eval(self, u: Uniforms, wave_uniforms: np.ndarray, heightmap_rgba: np.ndarray, sky_faces=None,
             ids=None, rows=None, cols=None, shade: bool = False, extra: bool = False, f64: bool = False, clean_phase: bool = False,
             smooth_phase: bool = False):
        """Returns dict(pos (n,4), normal (n,4), shade (n,4)|None, extra (n,16)|None).
        clean_phase (with f64): the time offset of every wave is the shader's fp32 product fl(phi t), everything else exact —
        the phase a smooth evaluation (the recurrence kernel) aims at.
        smooth_phase (fp32): phase and sin/cos in double from the same fp32 inputs, all else (products, the running fp32 sum in
        wave order) as the shader: models a kernel that does not reproduce the reference's rounding of the phase."""
        L = int(u.gridLength)
        wu = np.ascontiguousarray(wave_uniforms)
        assert wu.dtype == WAVE_UNIFORM_DTYPE and len(wu) >= u.gerstnerWaveCount
        hm = np.ascontiguousarray(heightmap_rgba, np.uint8)
        assert hm.ndim == 3 and hm.shape[2] == 4
        tx = Textures()
        tx.heightmap_rgba = hm.ctypes.data
        tx.hm_h, tx.hm_w = hm.shape[0], hm.shape[1]
        keep = [hm, wu]
        if sky_faces is not None:
            sky = np.ascontiguousarray(sky_faces, np.uint8)
            assert sky.ndim == 4 and sky.shape[0] == 6 and sky.shape[1] == sky.shape[2] and sky.shape[3] == 4
            for f in range(6):
                tx.sky_faces[f] = sky[f].ctypes.data
            tx.sky_n = sky.shape[1]
            keep.append(sky)
        if ids is not None:
            ids_a = np.ascontiguousarray(ids, np.int64)
            n = ids_a.size
            ids_p, r0, r1, c0, c1 = ids_a.ctypes.data, 0, 0, 0, 0
        else:
            r0, r1 = rows if rows is not None else (0, L)
            c0, c1 = cols if cols is not None else (0, L)
            n = (r1 - r0) * (c1 - c0)
            ids_p = None
        dt = np.float64 if f64 else np.float32
        pos = np.empty((n, 4), dt)
        nrm = np.empty((n, 4), dt)
        shd = np.empty((n, 4), dt) if shade else None
        ext = np.empty((n, 16), dt) if extra else None
        fn = self.lib.wto_eval_f64 if f64 else self.lib.wto_eval_f32
        self.lib.wto_set_time_offset_fp32(2 if (smooth_phase and not f64) else (1 if (f64 and clean_phase) else 0))
        st = fn(C.byref(u), wu.ctypes.data, C.byref(tx), ids_p, n, r0, r1, c0, c1, pos.ctypes.data, nrm.ctypes.data,
                shd.ctypes.data if shade else None, ext.ctypes.data if extra else None)
        if st != 0:
            raise RuntimeError(f"oracle eval failed: {st}")
        return {"pos": pos, "normal": nrm, "shade": shd, "extra": ext}


    def sky_synthesize(self, su: SkyUniforms, gradient, stars, clouds, n: int) -> np.ndarray:
        """The skybox cubemap of render-engine.cpp:306-442 -> (6,n,n,4) uint8."""
        tx, keep = sky_textures(gradient, stars, clouds)
        out = np.empty((6, n, n, 4), np.uint8)
        self.lib.wto_sky_synthesize(C.byref(su), C.byref(tx), n, out.ctypes.data)
        return out

    def sky_layer(self, which: str, su: SkyUniforms, gradient, stars, clouds, dirs: np.ndarray) -> np.ndarray:
        tx, keep = sky_textures(gradient, stars, clouds)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        out = np.empty((len(d), 4), np.float32)
        fn = self.lib.wto_skysphere_colour_f32 if which == "skysphere" else self.lib.wto_clouds_colour_f32
        fn(C.byref(su), C.byref(tx), len(d), d.ctypes.data, out.ctypes.data)
        return out

    def compose(self, cu: ComposeUniforms, depth: np.ndarray, reflections: np.ndarray, refractions: np.ndarray,
                pos4, nrm4, shade4, clip4, vary_a, vary_b) -> np.ndarray:
        """water-grid.frag:66-69,76-89,113-162 at caller-supplied vertices -> colour (n,4).  nrm4.w carries
        fresnel_f_theta, shade4 = (sky.rgb, sun), vary_a = (ndc0.xy, ., .), vary_b = (yaw normal, .)."""
        d = np.ascontiguousarray(depth, np.float32)
        rl = np.ascontiguousarray(reflections, np.uint8)
        rr = np.ascontiguousarray(refractions, np.uint8)
        assert d.ndim == 2 and rl.shape == d.shape + (4,) and rr.shape == rl.shape
        st = ScreenTextures(d.ctypes.data, rl.ctypes.data, rr.ctypes.data, d.shape[1], d.shape[0])
        arrs = [np.ascontiguousarray(a, np.float32) for a in (pos4, nrm4, shade4, clip4, vary_a, vary_b)]
        out = np.empty((len(arrs[0]), 4), np.float32)
        self.lib.wto_compose_f32(C.byref(cu), C.byref(st), len(out), *[a.ctypes.data for a in arrs], out.ctypes.data)
        return out

    def shade(self, u: Uniforms, sky_faces, pos4: np.ndarray, nrm4: np.ndarray):
        """water-grid.frag:71-73,95-111 at caller-supplied vertices -> (shade (n,4), fresnel (n,))."""
        tx = Textures()
        keep = None
        if sky_faces is not None:
            keep = np.ascontiguousarray(sky_faces, np.uint8)
            for f in range(6):
                tx.sky_faces[f] = keep[f].ctypes.data
            tx.sky_n = keep.shape[1]
        p = np.ascontiguousarray(pos4, np.float32)
        n = np.ascontiguousarray(nrm4, np.float32)
        shade = np.empty((len(p), 4), np.float32)
        fres = np.empty(len(p), np.float32)
        self.lib.wto_shade_f32(C.byref(u), C.byref(tx), len(p), p.ctypes.data, n.ctypes.data, shade.ctypes.data, fres.ctypes.data)
        return shade, fres


class Ref:
    """The reference itself (camera.cpp / render-engine.h / GLM / water-grid.vert as C++)."""

    def __init__(self):
        self.proj = _load(os.path.join(HERE, "_ref", "libwt_ref_proj.so"))
        self.vert = _load(os.path.join(HERE, "_ref", "libwt_ref_vert.so"))
        if self.proj is not None:
            P = self.proj
            P.wtref_total_amplitude.restype = C.c_float
            P.wtref_steepness_Q_i.restype = C.c_float
            P.wtref_steepness_Q_i.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint]
            P.wtref_bounce_displacement.restype = C.c_float
            P.wtref_bounce_displacement.argtypes = [C.c_float, C.c_float]
            P.wtref_sun_position.argtypes = [C.c_float, c_float_p]
            P.wtref_project_grid.argtypes = [C.POINTER(Camera), C.c_float, C.POINTER(GridF32)]
        self.frag = _load(os.path.join(HERE, "_ref", "libwt_ref_frag.so"))
        if self.frag is not None:
            self.frag.wtref_frag_set_uniforms.argtypes = [C.POINTER(RefFragUniforms), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                          C.POINTER(C.c_void_p), C.c_int]
            self.frag.wtref_frag_run.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
        self.sky = _load(os.path.join(HERE, "_ref", "libwt_ref_sky.so"))
        if self.sky is not None:
            self.sky.wtref_skysphere_run.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p]
            self.sky.wtref_clouds_run.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
            self.sky.wtref_stars_run.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
        if self.vert is not None:
            V = self.vert
            V.wtref_vert_set_uniforms.argtypes = [C.POINTER(RefVertUniforms), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            V.wtref_vert_run.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]

    @property
    def available(self) -> bool:
        return self.proj is not None and self.vert is not None

    def max_waves(self) -> int:
        return int(self.vert.wtref_max_waves())

    def total_amplitude(self, amplitudes) -> np.float32:
        a = np.ascontiguousarray(amplitudes, dtype=np.float32)
        return np.float32(self.proj.wtref_total_amplitude(a.ctypes.data_as(c_float_p), C.c_uint(a.size)))

    def steepness_Q_i(self, Q, w, A, count) -> np.float32:
        return np.float32(self.proj.wtref_steepness_Q_i(float(Q), float(w), float(A), int(count)))

    def bounce_displacement(self, amplitude, phase) -> np.float32:
        return np.float32(self.proj.wtref_bounce_displacement(float(amplitude), float(phase)))

    def sun_position(self, hours) -> np.ndarray:
        out = np.zeros(3, np.float32)
        self.proj.wtref_sun_position(float(hours), out.ctypes.data_as(c_float_p))
        return out

    def project_grid(self, cam: Camera, displaceable_amplitude):
        g = GridF32()
        st = self.proj.wtref_project_grid(C.byref(cam), C.c_float(float(displaceable_amplitude)), C.byref(g))
        return st, g

    def sky_layer_run(self, which: str, su: SkyUniforms, gradient, stars, clouds, dirs: np.ndarray) -> np.ndarray:
        """Runs the compiled skysphere.frag / skybox-clouds.frag / skybox-stars.frag main() on directions -> colour (n,4)."""
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        out = np.empty((len(d), 4), np.float32)
        sun = np.array(su.sunPosition[:], np.float32)
        if which == "skysphere":
            g = np.ascontiguousarray(gradient, np.uint8).reshape(-1, 4)
            self.sky.wtref_skysphere_run(g.ctypes.data, len(g), su.sunHorizonDarkness, sun.ctypes.data, su.sunShininess, su.sunStrength,
                                         d.ctypes.data, len(d), out.ctypes.data)
        else:
            a = np.ascontiguousarray(clouds if which == "clouds" else stars, np.uint8)
            ptrs = (C.c_void_p * 6)(*[a[f].ctypes.data for f in range(6)])
            if which == "clouds":
                self.sky.wtref_clouds_run(ptrs, a.shape[1], su.oneMinusCloudProportion, su.overcastStrength, sun.ctypes.data, d.ctypes.data, len(d), out.ctypes.data)
            else:
                self.sky.wtref_stars_run(ptrs, a.shape[1], d.ctypes.data, len(d), out.ctypes.data)
        return out

    def frag_run(self, fu: RefFragUniforms, depth, reflections, refractions, sky_faces, frag_in: np.ndarray) -> np.ndarray:
        """Runs the compiled water-grid.frag main() on REF_FRAG_IN_DTYPE records -> colour (n,4)."""
        d = np.ascontiguousarray(depth, np.float32)
        rl = np.ascontiguousarray(reflections, np.uint8)
        rr = np.ascontiguousarray(refractions, np.uint8)
        sky = np.ascontiguousarray(sky_faces, np.uint8)
        ptrs = (C.c_void_p * 6)(*[sky[f].ctypes.data for f in range(6)])
        self.frag.wtref_frag_set_uniforms(C.byref(fu), d.ctypes.data, rl.ctypes.data, rr.ctypes.data, d.shape[1], d.shape[0], ptrs, sky.shape[1])
        fi = np.ascontiguousarray(frag_in)
        assert fi.dtype == REF_FRAG_IN_DTYPE
        out = np.empty((len(fi), 4), np.float32)
        self.frag.wtref_frag_run(fi.ctypes.data, len(fi), out.ctypes.data)
        return out

    def vert_run(self, u: Uniforms, wave_uniforms: np.ndarray, heightmap_rgba: np.ndarray, ids=None, id_range=None):
        """Runs the compiled shader main() and returns the REF_VERT_OUT_DTYPE records.  World-space
        position is not a shader output; callers obtain it by passing viewProjection = identity, for
        which gl_Position == position exactly."""
        ru = RefVertUniforms()
        for name, _ in RefVertUniforms._fields_:
            setattr(ru, name, getattr(u, name))
        wu = np.ascontiguousarray(wave_uniforms)
        hm = np.ascontiguousarray(heightmap_rgba, np.uint8)
        st = self.vert.wtref_vert_set_uniforms(C.byref(ru), wu.ctypes.data, hm.ctypes.data, hm.shape[1], hm.shape[0])
        if st != 0:
            raise RuntimeError("reference shader: too many waves for the compiled array bound")
        if ids is not None:
            ids_a = np.ascontiguousarray(ids, np.int32)
            out = np.empty(ids_a.size, REF_VERT_OUT_DTYPE)
            self.vert.wtref_vert_run(ids_a.ctypes.data, 0, ids_a.size, out.ctypes.data)
        else:
            b, e = id_range
            out = np.empty(e - b, REF_VERT_OUT_DTYPE)
            self.vert.wtref_vert_run(None, b, e - b, out.ctypes.data)
        return out
