"""ctypes view of the C ABI (include/wavetool_b200.h) plus a small convenience class.

This is harness code for tests/ and bench.py: the product is the shared library.  Nothing here
computes surface values on the CPU, and nothing here touches oracle/ — if the CUDA library is
missing or no B200 is visible, device calls raise (there is no fallback path).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

WT_OK, WT_NO_GRID = 0, 1
WT_ERR_INVALID_ARGUMENT, WT_ERR_CUDA, WT_ERR_NO_DEVICE, WT_ERR_NOT_READY, WT_ERR_DEGENERATE = -1, -2, -3, -4, -5
WT_MAX_WAVES = 1000
FLAG_SHADE, FLAG_CHECKSUM, FLAG_FAST_PHASE, FLAG_REFERENCE_ORDER, FLAG_VARYINGS, FLAG_AUTO_PHASE, FLAG_STENCIL_ORDER = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40
WAVENUMBER_GPU_GEMS, WAVENUMBER_PHYSICAL = 0, 1
PHASE_BIT_FAITHFUL, PHASE_RECURRENCE, PHASE_REFERENCE_ORDER, PHASE_STENCIL_ORDER = 0, 1, 2, 3
PHASE_NAMES = {0: "bit-faithful", 1: "recurrence", 2: "reference-order", 3: "stencil-order"}
UI_RANGE_COUNT = 23

# every symbol include/wavetool_b200.h declares (tests/test_abi.py checks the library exports them all)
EXPORTED_SYMBOLS = [
    "wt_version", "wt_last_error", "wt_device_count",
    "wt_default_params", "wt_default_camera", "wt_default_waves", "wt_camera_matrices",
    "wt_total_amplitude", "wt_steepness_Q_i", "wt_sun_position", "wt_bounce_displacement",
    "wt_project_grid", "wt_build_frame",
    "wt_create", "wt_destroy", "wt_set_waves", "wt_set_heightmap_rgba8", "wt_set_skybox_rgba8",
    "wt_eval_frame", "wt_eval_batch", "wt_eval_frame_host", "wt_launch_count", "wt_measure_peak",
    "wt_debug_read_trace", "wt_debug_reload_env", "wt_debug_layout", "wt_eval_frame_ex", "wt_grid_indices",
    "wt_default_animation", "wt_animation_advance", "wt_run_animation",
    "wt_default_compose_params", "wt_set_screen_textures", "wt_compose_frame",
    "wt_step_host", "wt_step_submit", "wt_step_wait", "wt_default_sky_params", "wt_set_sky_layers", "wt_synthesize_skybox", "wt_get_skybox_rgba8",
    "wt_wave_from_wavelength_speed", "wt_wave_clamp_ui", "wt_params_clamp_ui", "wt_animation_clamp_ui", "wt_sky_params_clamp_ui",
    "wt_compose_params_clamp_ui", "wt_ui_ranges", "wt_fast_phase_report", "wt_fast_phase_valid", "wt_last_phase_mode", "wt_phase_mode_counts", "wt_debug_sky_exact_texels", "wt_batched_frame_count", "wt_debug_full_stencil", "wt_stencil_flag_rate", "wt_debug_fog_table",
]


class Wave(C.Structure):
    _fields_ = [("amplitude_A", C.c_float), ("frequency_w", C.c_float), ("phaseConstant_phi", C.c_float),
                ("steepness_Q", C.c_float), ("dir_x", C.c_float), ("dir_z", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("fov_deg", C.c_float), ("aspect", C.c_float), ("z_near", C.c_float), ("z_far", C.c_float),
                ("pos", C.c_float * 3), ("yaw_deg", C.c_float), ("pitch_deg", C.c_float)]


class Params(C.Structure):
    _fields_ = [("heightmapDisplacementScale", C.c_float), ("heightmapSampleScale", C.c_float),
                ("verticalBounceWaveAmplitude", C.c_float), ("verticalBounceWavePhase", C.c_float),
                ("waveAnimationTimeInSeconds", C.c_float), ("timeOfDayInHours", C.c_float),
                ("sunShininess", C.c_float), ("sunStrength", C.c_float)]


class Frame(C.Structure):
    _fields_ = [("bottomLeftGridPointInWorld", C.c_float * 4), ("bottomRightGridPointInWorld", C.c_float * 4),
                ("topLeftGridPointInWorld", C.c_float * 4), ("topRightGridPointInWorld", C.c_float * 4),
                ("viewProjection", C.c_float * 16), ("viewMatOnlyYaw", C.c_float * 16),
                ("cameraPosition", C.c_float * 3), ("waveAnimationTimeInSeconds", C.c_float),
                ("verticalBounceWaveDisplacement", C.c_float), ("heightmapDisplacementScale", C.c_float),
                ("heightmapSampleScale", C.c_float), ("sunShininess", C.c_float),
                ("sunPosition", C.c_float * 3), ("sunStrength", C.c_float)]


class Checksum(C.Structure):
    _fields_ = [("position_bits", C.c_uint64), ("normal_bits", C.c_uint64), ("shade_bits", C.c_uint64),
                ("min_y", C.c_float), ("max_y", C.c_float), ("vertices", C.c_uint64)]

    def as_tuple(self):
        return (self.position_bits, self.normal_bits, self.shade_bits, self.min_y, self.max_y, self.vertices)


class Animation(C.Structure):
    _fields_ = [("animationSpeedTimeOfDayInSecondsPerHour", C.c_float), ("animationSpeedVerticalBounceWavePhasePeriodInSeconds", C.c_float),
                ("isAnimatingTimeOfDay", C.c_int), ("isAnimatingWaves", C.c_int)]


class ComposeParams(C.Structure):
    _fields_ = [("fogColourFarAtNoon", C.c_float * 4), ("fogDepthRadiusFar", C.c_float), ("fogDepthRadiusNear", C.c_float),
                ("softEdgesDeltaDepthThreshold", C.c_float), ("tintDeltaDepthThreshold", C.c_float), ("waterClarity", C.c_float),
                ("zNear", C.c_float), ("zFar", C.c_float), ("viewportWidth", C.c_float), ("viewportHeight", C.c_float)]


class SkyParams(C.Structure):
    _fields_ = [("timeOfDayInHours", C.c_float), ("cloudProportion", C.c_float), ("overcastStrength", C.c_float), ("sunHorizonDarkness", C.c_float),
                ("sunShininess", C.c_float), ("sunStrength", C.c_float), ("fogColourFarAtNoon", C.c_float * 4),
                ("drawStars", C.c_int), ("drawSkysphere", C.c_int), ("drawClouds", C.c_int)]


class PhaseReport(C.Structure):
    _fields_ = [("est_position_error", C.c_float), ("position_tolerance", C.c_float), ("est_normal_error_rad", C.c_float),
                ("normal_tolerance_rad", C.c_float), ("min_vertex_spacing", C.c_float), ("max_phase", C.c_float), ("valid", C.c_int),
                ("est_shared_stencil_normal_error_rad", C.c_float), ("shared_stencil_valid", C.c_int),
                ("coordinate_ulp_normal_error_rad", C.c_float), ("recommended", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class UiRange(C.Structure):
    _fields_ = [("field", C.c_char_p), ("slider_min", C.c_float), ("slider_max", C.c_float), ("clamped_below", C.c_int), ("clamped_above", C.c_int)]


class Outputs(C.Structure):
    _fields_ = [("position", C.c_void_p), ("normal", C.c_void_p), ("shade", C.c_void_p), ("clip_position", C.c_void_p),
                ("varyings_a", C.c_void_p), ("varyings_b", C.c_void_p), ("checksum", C.c_void_p)]


CHECKSUM_DTYPE = np.dtype([("position_bits", "u8"), ("normal_bits", "u8"), ("shade_bits", "u8"),
                           ("min_y", "f4"), ("max_y", "f4"), ("vertices", "u8")])
assert CHECKSUM_DTYPE.itemsize == C.sizeof(Checksum) == 40

_lib = None


class WaveToolError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"wavetool_b200 status {status}: {msg}")
        self.status = status


def lib(build_if_missing: bool = True) -> C.CDLL:
    """Loads wave_tool_b200/libwavetool_b200.so (building it in-tree first if it is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(_build.LIB_PATH):
        raise WaveToolError(WT_ERR_NO_DEVICE, f"{_build.LIB_PATH} is missing: the CUDA extension is required (no CPU fallback)")
    # development knob: WT_LIB points at a tuning variant built by tools/build_variant.py
    L = C.CDLL(os.environ.get("WT_LIB") or _build.LIB_PATH)
    L.wt_version.restype = C.c_char_p
    L.wt_last_error.restype = C.c_char_p
    L.wt_last_error.argtypes = [C.c_void_p]
    L.wt_default_params.argtypes = [C.POINTER(Params)]
    L.wt_default_camera.argtypes = [C.POINTER(Camera), C.c_float]
    L.wt_default_waves.argtypes = [C.POINTER(Wave)]
    L.wt_camera_matrices.argtypes = [C.POINTER(Camera)] + [C.c_void_p] * 5
    L.wt_total_amplitude.restype = C.c_float
    L.wt_total_amplitude.argtypes = [C.c_void_p, C.c_uint]
    L.wt_steepness_Q_i.restype = C.c_float
    L.wt_steepness_Q_i.argtypes = [C.POINTER(Wave), C.c_uint]
    L.wt_sun_position.argtypes = [C.c_float, C.c_void_p]
    L.wt_bounce_displacement.restype = C.c_float
    L.wt_bounce_displacement.argtypes = [C.c_float, C.c_float]
    L.wt_project_grid.argtypes = [C.POINTER(Camera), C.c_float, C.c_void_p]
    L.wt_build_frame.argtypes = [C.POINTER(Camera), C.POINTER(Params), C.c_float, C.POINTER(Frame)]
    L.wt_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.wt_destroy.argtypes = [C.c_void_p]
    L.wt_set_waves.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    L.wt_set_heightmap_rgba8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.wt_set_skybox_rgba8.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int]
    L.wt_eval_frame.argtypes = [C.c_void_p, C.POINTER(Frame), C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_uint, C.c_void_p]
    L.wt_eval_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_uint, C.c_void_p, C.c_uint, C.c_void_p]
    L.wt_eval_frame_host.argtypes = [C.c_void_p, C.POINTER(Frame), C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_uint]
    L.wt_launch_count.restype = C.c_uint64
    L.wt_launch_count.argtypes = [C.c_void_p]
    L.wt_batched_frame_count.restype = C.c_uint64
    L.wt_batched_frame_count.argtypes = [C.c_void_p]
    L.wt_measure_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.wt_debug_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    L.wt_debug_reload_env.restype = None
    L.wt_debug_layout.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_uint]
    L.wt_debug_full_stencil.argtypes = [C.c_void_p, C.c_int]
    L.wt_stencil_flag_rate.argtypes = [C.c_uint]
    L.wt_stencil_flag_rate.restype = C.c_float
    L.wt_eval_frame_ex.argtypes = [C.c_void_p, C.POINTER(Frame), C.c_uint, C.c_uint, C.c_uint, C.POINTER(Outputs), C.c_uint, C.c_void_p]
    L.wt_grid_indices.argtypes = [C.c_uint, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.wt_default_compose_params.argtypes = [C.POINTER(ComposeParams), C.POINTER(Camera), C.c_uint, C.c_uint]
    L.wt_default_compose_params.restype = None
    L.wt_set_screen_textures.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.wt_compose_frame.argtypes = [C.c_void_p, C.POINTER(Frame), C.POINTER(ComposeParams), C.c_uint, C.c_uint, C.c_uint, C.POINTER(Outputs),
                                   C.c_void_p, C.c_void_p]
    L.wt_step_host.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Params), C.c_float, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_uint, C.POINTER(Checksum), C.c_void_p]
    L.wt_step_submit.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Params), C.c_float, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_uint)]
    L.wt_step_wait.argtypes = [C.c_void_p, C.c_uint, C.POINTER(Checksum)]
    L.wt_wave_from_wavelength_speed.argtypes = [C.c_float] * 6 + [C.c_int, C.POINTER(Wave)]
    L.wt_wave_clamp_ui.argtypes = [C.POINTER(Wave)]
    L.wt_params_clamp_ui.argtypes = [C.POINTER(Params)]
    L.wt_animation_clamp_ui.argtypes = [C.POINTER(Animation)]
    L.wt_sky_params_clamp_ui.argtypes = [C.POINTER(SkyParams)]
    L.wt_compose_params_clamp_ui.argtypes = [C.POINTER(ComposeParams), C.c_int]
    L.wt_ui_ranges.argtypes = [C.POINTER(UiRange)]
    L.wt_ui_ranges.restype = None
    L.wt_fast_phase_report.argtypes = [C.c_void_p, C.c_uint, C.POINTER(Frame), C.c_uint, C.POINTER(PhaseReport)]
    L.wt_fast_phase_valid.argtypes = [C.c_void_p, C.POINTER(Frame), C.c_uint, C.POINTER(PhaseReport)]
    L.wt_last_phase_mode.argtypes = [C.c_void_p]
    L.wt_phase_mode_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.wt_phase_mode_counts.restype = None
    L.wt_default_sky_params.argtypes = [C.POINTER(SkyParams)]
    L.wt_default_sky_params.restype = None
    L.wt_set_sky_layers.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.wt_synthesize_skybox.argtypes = [C.c_void_p, C.POINTER(SkyParams), C.c_int, C.c_void_p]
    L.wt_get_skybox_rgba8.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.wt_debug_sky_exact_texels.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.wt_default_animation.argtypes = [C.POINTER(Animation)]
    L.wt_default_animation.restype = None
    L.wt_animation_advance.argtypes = [C.POINTER(Params), C.POINTER(Animation), C.c_float]
    L.wt_run_animation.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Params), C.POINTER(Animation), C.c_float, C.c_float, C.c_uint, C.c_uint,
                                   C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p,
                                   C.POINTER(C.c_uint)]
    _lib = L
    return L


def _check(status, ctx=None):
    if status < 0:
        raise WaveToolError(status, (lib().wt_last_error(ctx) or b"").decode())
    return status


# ------------------------------------------------------------------------------------------------
# host math wrappers
# ------------------------------------------------------------------------------------------------
def default_params() -> Params:
    p = Params()
    lib().wt_default_params(C.byref(p))
    return p


def default_camera(aspect: float = 1.0) -> Camera:
    c = Camera()
    lib().wt_default_camera(C.byref(c), aspect)
    return c


def waves_array(rows) -> np.ndarray:
    """(W,6) float32 rows [A, w, phi, Q, Dx, Dz] — the memory layout of wt_wave[W]."""
    a = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 6)
    return a


def default_waves() -> np.ndarray:
    w = (Wave * 4)()
    lib().wt_default_waves(w)
    return np.frombuffer(w, dtype=np.float32).reshape(4, 6).copy()


def total_amplitude(waves: np.ndarray) -> float:
    w = waves_array(waves)
    return float(lib().wt_total_amplitude(w.ctypes.data, len(w)))


def camera_matrices(cam: Camera):
    out = [np.zeros(16, np.float32) for _ in range(4)] + [np.zeros(3, np.float32)]
    _check(lib().wt_camera_matrices(C.byref(cam), *[o.ctypes.data for o in out]))
    return {"view": out[0], "proj": out[1], "view_proj": out[2], "view_only_yaw": out[3], "forward": out[4]}


def project_grid(cam: Camera, displaceable_amplitude: float):
    corners = np.zeros((4, 4), np.float32)
    st = _check(lib().wt_project_grid(C.byref(cam), C.c_float(displaceable_amplitude), corners.ctypes.data))
    return st, corners


def default_compose_params(cam: Camera, viewport_w: int, viewport_h: int) -> ComposeParams:
    p = ComposeParams()
    lib().wt_default_compose_params(C.byref(p), C.byref(cam), viewport_w, viewport_h)
    return p


def default_sky_params() -> SkyParams:
    p = SkyParams()
    lib().wt_default_sky_params(C.byref(p))
    return p


def default_animation() -> Animation:
    a = Animation()
    lib().wt_default_animation(C.byref(a))
    return a


def animation_advance(params: Params, anim: Animation, dt: float) -> None:
    """In place: one UI frame of the reference's animation state machine (program.cpp:220-225, 452-464)."""
    _check(lib().wt_animation_advance(C.byref(params), C.byref(anim), C.c_float(dt)))


def grid_indices(grid_len: int) -> np.ndarray:
    """The water mesh index buffer (program.cpp:686-717) as a (6 (L-1)^2,) uint32 array."""
    n = C.c_uint64()
    _check(lib().wt_grid_indices(grid_len, None, 0, C.byref(n)))
    out = np.empty(n.value, np.uint32)
    _check(lib().wt_grid_indices(grid_len, out.ctypes.data, n.value, C.byref(n)))
    return out


def wave_from_wavelength_speed(amplitude, wavelength, steepness, direction, speed, convention=WAVENUMBER_PHYSICAL) -> Wave:
    w = Wave()
    _check(lib().wt_wave_from_wavelength_speed(amplitude, wavelength, steepness, direction[0], direction[1], speed, convention, C.byref(w)))
    return w


def ui_ranges() -> dict:
    arr = (UiRange * UI_RANGE_COUNT)()
    lib().wt_ui_ranges(arr)
    return {r.field.decode(): (r.slider_min, r.slider_max, bool(r.clamped_below), bool(r.clamped_above)) for r in arr}


def fast_phase_report(waves: np.ndarray, frame: Frame, grid_len: int) -> PhaseReport:
    """wt_fast_phase_report (host only): the estimated deviation of the recurrence kernel from the reference on this frame."""
    w = waves_array(waves)
    rep = PhaseReport()
    _check(lib().wt_fast_phase_report(w.ctypes.data, len(w), C.byref(frame), grid_len, C.byref(rep)))
    return rep


def fog_table(fog_rgba) -> np.ndarray:
    """wt_debug_fog_table (host only): (256, 4) uint8 — the byte each channel holds after the fog quad, per byte before it."""
    f = (C.c_float * 4)(*[float(v) for v in fog_rgba])
    out = np.zeros(256, np.uint32)
    _check(lib().wt_debug_fog_table(f, out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out.view(np.uint8).reshape(256, 4)


def stencil_flag_rate(grid_len: int) -> float:
    """wt_stencil_flag_rate (host only): fraction of the stencil arguments that differ from their neighbours' own uv."""
    return float(lib().wt_stencil_flag_rate(grid_len))


def debug_layout(grid_len: int, rows_per_step: int, phase_mode: int, rows=None) -> np.ndarray:
    """(n,2) int32 canonical row chunks [begin, end) of the work items that meet `rows` (wt_debug_layout; host only)."""
    r0, r1 = rows if rows is not None else (0, grid_len)
    cap = 4 * (grid_len // 4 + 16)
    out = np.zeros((cap, 2), np.int32)
    n = _check(lib().wt_debug_layout(grid_len, rows_per_step, phase_mode, r0, r1, out.ctypes.data, cap))
    assert n <= cap
    return out[:n].copy()


def build_frame(cam: Camera, params: Params, total_wave_amplitude: float):
    f = Frame()
    st = _check(lib().wt_build_frame(C.byref(cam), C.byref(params), C.c_float(total_wave_amplitude), C.byref(f)))
    return st, f


# ------------------------------------------------------------------------------------------------
# device context
# ------------------------------------------------------------------------------------------------
class Context:
    """Owns a wt_ctx.  Raises WaveToolError(WT_ERR_NO_DEVICE) when no B200 is visible."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self._lib = lib()
        _check(self._lib.wt_create(C.byref(self._h), device))
        self.device = device
        self.n_waves = 0

    def close(self):
        if self._h:
            self._lib.wt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_waves(self, waves: np.ndarray):
        w = waves_array(waves)
        _check(self._lib.wt_set_waves(self._h, w.ctypes.data, len(w)), self._h)
        self.n_waves = len(w)

    def set_heightmap(self, rgba: np.ndarray):
        a = np.ascontiguousarray(rgba, np.uint8)
        assert a.ndim == 3 and a.shape[2] == 4
        _check(self._lib.wt_set_heightmap_rgba8(self._h, a.ctypes.data, a.shape[1], a.shape[0]), self._h)

    def set_skybox(self, faces: np.ndarray):
        a = np.ascontiguousarray(faces, np.uint8)
        assert a.ndim == 4 and a.shape[0] == 6 and a.shape[1] == a.shape[2] and a.shape[3] == 4
        ptrs = (C.c_void_p * 6)(*[a[f].ctypes.data for f in range(6)])
        _check(self._lib.wt_set_skybox_rgba8(self._h, ptrs, a.shape[1]), self._h)

    def eval_frame(self, frame: Frame, grid_len: int, d_pos: int, d_normal: int, d_shade: int = 0, d_checksum: int = 0,
                   flags: int = 0, rows=None, stream: int = 0):
        """Device-pointer entry point (ints are raw device addresses, e.g. torch.Tensor.data_ptr())."""
        r0, r1 = rows if rows is not None else (0, grid_len)
        return _check(self._lib.wt_eval_frame(self._h, C.byref(frame), grid_len, r0, r1, d_pos, d_normal, d_shade or None,
                                              d_checksum or None, flags, stream or None), self._h)

    def eval_frame_ex(self, frame: Frame, grid_len: int, outputs: "Outputs", flags: int = 0, rows=None, stream: int = 0):
        """wt_eval_frame_ex: every output stream of the vertex stage (Outputs holds raw device addresses)."""
        r0, r1 = rows if rows is not None else (0, grid_len)
        return _check(self._lib.wt_eval_frame_ex(self._h, C.byref(frame), grid_len, r0, r1, C.byref(outputs), flags, stream or None), self._h)

    def step_host(self, cam: Camera, params: Params, total_wave_amplitude: float, grid_len: int, d_pos: int, d_normal: int, d_shade: int = 0,
                  flags: int = 0, rows=None, stream: int = 0, out: "Checksum | None" = None):
        """wt_step_host: host camera/params in, the frame's digest (host) out; returns (status, Checksum)."""
        r0, r1 = rows if rows is not None else (0, grid_len)
        cks = out if out is not None else Checksum()
        st = _check(self._lib.wt_step_host(self._h, C.byref(cam), C.byref(params), C.c_float(total_wave_amplitude), grid_len, r0, r1,
                                           d_pos, d_normal, d_shade or None, flags, C.byref(cks), stream or None), self._h)
        return st, cks

    def step_submit(self, cam: Camera, params: Params, total_wave_amplitude: float, grid_len: int, d_pos: int, d_normal: int, d_shade: int = 0,
                    flags: int = 0, rows=None, stream: int = 0) -> int:
        """wt_step_submit: enqueue one frame from host camera/params; returns the ticket for step_wait."""
        r0, r1 = rows if rows is not None else (0, grid_len)
        t = C.c_uint()
        _check(self._lib.wt_step_submit(self._h, C.byref(cam), C.byref(params), C.c_float(total_wave_amplitude), grid_len, r0, r1,
                                        d_pos, d_normal, d_shade or None, flags, stream or None, C.byref(t)), self._h)
        return t.value

    def step_wait(self, ticket: int, out: "Checksum | None" = None):
        cks = out if out is not None else Checksum()
        st = _check(self._lib.wt_step_wait(self._h, ticket, C.byref(cks)), self._h)
        return st, cks

    def set_sky_layers(self, gradient: np.ndarray, stars=None, clouds=None):
        g = np.ascontiguousarray(gradient, np.uint8).reshape(-1, 4)
        keep, ptrs, sizes = [g], [], []
        for faces in (stars, clouds):
            if faces is None:
                ptrs.append(None); sizes.append(0)
            else:
                a = np.ascontiguousarray(faces, np.uint8)
                keep.append(a)
                arr = (C.c_void_p * 6)(*[a[f].ctypes.data for f in range(6)])
                keep.append(arr)
                ptrs.append(C.cast(arr, C.c_void_p)); sizes.append(a.shape[1])
        _check(self._lib.wt_set_sky_layers(self._h, g.ctypes.data, len(g), ptrs[0], sizes[0], ptrs[1], sizes[1]), self._h)

    def synthesize_skybox(self, params: "SkyParams", n: int, stream: int = 0):
        return _check(self._lib.wt_synthesize_skybox(self._h, C.byref(params), n, stream or None), self._h)

    def stencil_form(self, form: int):
        """wt_debug_full_stencil: 0 = automatic, 1 = one kernel with five evaluations per vertex, 2 = bit-faithful kernel + fix-up pass."""
        _check(self._lib.wt_debug_full_stencil(self._h, int(form)), self._h)

    def sky_exact_texels(self) -> int:
        v = C.c_uint64()
        _check(self._lib.wt_debug_sky_exact_texels(self._h, C.byref(v)), self._h)
        return int(v.value)

    def get_skybox(self) -> np.ndarray:
        n = C.c_int()
        _check(self._lib.wt_get_skybox_rgba8(self._h, None, C.byref(n)), self._h)
        out = np.empty((6, n.value, n.value, 4), np.uint8)
        _check(self._lib.wt_get_skybox_rgba8(self._h, out.ctypes.data, C.byref(n)), self._h)
        return out

    def set_screen_textures(self, depth: np.ndarray, reflections: np.ndarray, refractions: np.ndarray):
        d = np.ascontiguousarray(depth, np.float32)
        rl = np.ascontiguousarray(reflections, np.uint8)
        rr = np.ascontiguousarray(refractions, np.uint8)
        assert d.ndim == 2 and rl.shape == d.shape + (4,) and rr.shape == rl.shape
        _check(self._lib.wt_set_screen_textures(self._h, d.ctypes.data, rl.ctypes.data, rr.ctypes.data, d.shape[1], d.shape[0]), self._h)

    def compose_frame(self, frame: Frame, params: "ComposeParams", grid_len: int, vertex_stage: "Outputs", d_colour: int, rows=None, stream: int = 0):
        r0, r1 = rows if rows is not None else (0, grid_len)
        return _check(self._lib.wt_compose_frame(self._h, C.byref(frame), C.byref(params), grid_len, r0, r1, C.byref(vertex_stage), d_colour,
                                                 stream or None), self._h)

    def run_animation(self, cam: Camera, params: Params, anim: Animation, total_wave_amplitude: float, dt: float, n_frames: int, grid_len: int,
                      d_pos_ring: int, d_normal_ring: int, d_shade_ring: int = 0, ring_depth: int = 2, d_checksums: int = 0, flags: int = 0,
                      frame_stride: int = 1, rows=None, stream: int = 0) -> int:
        """wt_run_animation: params is advanced in place; returns the number of frames that were drawn."""
        r0, r1 = rows if rows is not None else (0, grid_len)
        n = C.c_uint()
        _check(self._lib.wt_run_animation(self._h, C.byref(cam), C.byref(params), C.byref(anim), C.c_float(total_wave_amplitude), C.c_float(dt),
                                          frame_stride, n_frames, grid_len, r0, r1, d_pos_ring, d_normal_ring, d_shade_ring or None, ring_depth,
                                          d_checksums or None, flags, stream or None, C.byref(n)), self._h)
        return n.value

    def eval_batch(self, frames, grid_len: int, d_pos_ring: int, d_normal_ring: int, d_shade_ring: int = 0, ring_depth: int = 2,
                   d_checksums: int = 0, flags: int = 0, rows=None, stream: int = 0):
        r0, r1 = rows if rows is not None else (0, grid_len)
        arr = (Frame * len(frames))(*frames) if not isinstance(frames, C.Array) else frames
        return _check(self._lib.wt_eval_batch(self._h, arr, len(arr), grid_len, r0, r1, d_pos_ring, d_normal_ring, d_shade_ring or None,
                                              ring_depth, d_checksums or None, flags, stream or None), self._h)

    def eval_frame_host(self, frame: Frame, grid_len: int, flags: int = 0, rows=None, want_pos=True, want_normal=True,
                        want_checksum=False):
        """Host-buffer entry point: returns numpy arrays (n,4) copied back from the device."""
        r0, r1 = rows if rows is not None else (0, grid_len)
        n = (r1 - r0) * grid_len
        pos = np.empty((n, 4), np.float32) if want_pos else None
        nrm = np.empty((n, 4), np.float32) if want_normal else None
        shd = np.empty((n, 4), np.float32) if (flags & FLAG_SHADE) else None
        cks = Checksum() if want_checksum else None
        _check(self._lib.wt_eval_frame_host(self._h, C.byref(frame), grid_len, r0, r1,
                                            pos.ctypes.data if want_pos else None, nrm.ctypes.data if want_normal else None,
                                            shd.ctypes.data if shd is not None else None,
                                            C.byref(cks) if cks is not None else None, flags), self._h)
        return {"pos": pos, "normal": nrm, "shade": shd, "checksum": cks}

    def fast_phase_valid(self, frame: Frame, grid_len: int) -> PhaseReport:
        rep = PhaseReport()
        _check(self._lib.wt_fast_phase_valid(self._h, C.byref(frame), grid_len, C.byref(rep)), self._h)
        return rep

    def last_phase_mode(self) -> int:
        return int(self._lib.wt_last_phase_mode(self._h))

    def phase_mode_counts(self):
        """(bit-faithful, recurrence, reference-order, stencil-order) evaluation launches of this context so far (indexed by PHASE_*)."""
        out = (C.c_uint64 * 4)()
        self._lib.wt_phase_mode_counts(self._h, out)
        return tuple(int(v) for v in out)

    def batched_frame_count(self) -> int:
        return int(self._lib.wt_batched_frame_count(self._h))

    def launch_count(self) -> int:
        return int(self._lib.wt_launch_count(self._h))

    def read_trace(self, max_items: int = 1 << 20) -> np.ndarray:
        """WT_TRACE=1 diagnostic: (items, 4) uint64 rows {smid, cta, t_start_ns, t_end_ns} of the last launch."""
        buf = np.zeros((max_items, 4), np.uint64)
        n = _check(self._lib.wt_debug_read_trace(self._h, buf.ctypes.data, max_items), self._h)
        return buf[:n].copy()

    def measure_peak(self, kind: int) -> float:
        v = C.c_double()
        _check(self._lib.wt_measure_peak(self._h, kind, C.byref(v)), self._h)
        return v.value
