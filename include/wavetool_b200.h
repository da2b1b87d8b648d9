/*
 * wavetool_b200.h — C ABI of the B200-native water-surface evaluator.
 *
 * Drop-in boundary for ONE path of TexelBox/wave-tool: "set the water program's uniforms, then
 * glDrawElements" (reference src/render-engine.cpp:1045-1122), fed by the projected-grid builder
 * (src/render-engine.cpp:797-1042).  The reference has no plugin/FFI API; every entry point below
 * cites the reference code it replaces.  Plain C: pointers and sizes only, no C++/torch types, no
 * exceptions or stdout across the boundary.  Every call returns an int status:
 *     WT_OK (0)        success
 *     WT_NO_GRID (1)   the frustum misses the displaceable volume: the reference draws no water this
 *                      frame (render-engine.cpp:894) — not an error
 *     < 0              error; wt_last_error(ctx) (or wt_last_error(NULL) for ctx-less calls) describes it
 *
 * Ownership: the caller owns every host array (copied during the call) and every output device
 * buffer; a wt_ctx owns its wave table, textures and scratch.  A ctx is bound to one CUDA device and
 * is not thread-safe: one ctx per GPU per host thread.  Device work is stream-ordered on the stream
 * the caller passes (a cudaStream_t, passed as void*); no hidden synchronisation except in the
 * *_host convenience calls, which say so.
 *
 * There is no CPU fallback: device entry points fail with WT_ERR_NO_DEVICE when no CUDA device is
 * usable.  The host-math entry points (wt_camera_matrices, wt_project_grid, wt_build_frame, ...) are
 * pure CPU code by design — they are the reference's own CPU prologue (4 points per frame).
 */
#ifndef WAVETOOL_B200_H_
#define WAVETOOL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WT_OK 0
#define WT_NO_GRID 1
#define WT_ERR_INVALID_ARGUMENT (-1)
#define WT_ERR_CUDA (-2)
#define WT_ERR_NO_DEVICE (-3)
#define WT_ERR_NOT_READY (-4)   /* waves / heightmap / skybox not set before an evaluation that needs them */
#define WT_ERR_DEGENERATE (-5)  /* a projector ray is parallel to the base plane (reference asserts, render-engine.cpp:1034) */

#define WT_MAX_WAVES 1000u      /* reference: GerstnerWave::MAX_COUNT = 4 (render-engine.h:132); generalised here */

/* evaluation flags */
#define WT_FLAG_SHADE 0x1u            /* also evaluate the vertex-evaluable fragment terms (water-grid.frag:71-73,95-111) */
#define WT_FLAG_CHECKSUM 0x2u         /* fold the outputs into the wt_checksum record passed to the call */
#define WT_FLAG_FAST_PHASE 0x4u       /* pre-folded 2-FMA wave phase seeded once per wave every 8 or 12 rows, per-row recurrence in between.  NOT
                                         bit-faithful to the reference's fp32 phase: it meets the tolerances only inside the domain
                                         wt_fast_phase_valid describes (small phi t, vertex spacing not too fine) — prefer WT_FLAG_AUTO_PHASE */
#define WT_FLAG_REFERENCE_ORDER 0x8u  /* straight transliteration: 5 full evaluations per vertex, libdevice sincosf */
#define WT_FLAG_VARYINGS 0x10u        /* also write the vertex stage's other outputs (wt_eval_frame_ex): gl_Position,
                                         xyPositionNDCSpaceHeight0, heightmap_colour.r, normalVecInViewSpaceOnlyYaw */
#define WT_FLAG_STENCIL_ORDER 0x40u    /* the shader's own stencil — the four neighbours evaluated at uv +- du exactly as water-grid.vert:164-182 forms them —
                                         with the production kernels' trigonometry (Cody-Waite + MUFU): reproduces the reference at any animation
                                         time.  Those arguments differ from the neighbours' own uv only at the grid indices i where
                                         fl(fl(i du) +- du) != fl((i +- 1) du) (0.3 % of them at 4096, 19 % at 8192): the bit-faithful kernel runs first,
                                         then a fix-up pass re-evaluates the stencils of those rows and columns and redoes their normals and shading
                                         terms (+2 % at 4096^2; positions are the bit-faithful kernel's).  On grid sizes where most indices are flagged (an odd size
                                         like 700) one kernel evaluates all five points of every vertex instead (never with WT_FLAG_VARYINGS, which only the
                                         production kernel writes: the fix-up pass then also redoes normalVecInViewSpaceOnlyYaw) */
#define WT_FLAG_AUTO_PHASE 0x20u      /* per frame, the fastest kernel whose deviation from the reference wt_fast_phase_report bounds inside the
                                         tolerances: the recurrence kernel, else the bit-faithful kernel, else (late in an animation, where
                                         the shader's own re-evaluation of the stencil neighbours becomes noisy) the stencil-order kernel.
                                         All three run on the GPU; wt_last_phase_mode tells which one a launch used */

typedef struct wt_ctx wt_ctx;

/* geometry::GerstnerWave — reference src/render-engine.h:130-167 (the reference-native 5-tuple). */
typedef struct wt_wave {
    float amplitude_A;        /* >= 0, height of crest above the equilibrium plane */
    float frequency_w;        /* >= 0, ~2/L */
    float phaseConstant_phi;  /* >= 0, speed * w */
    float steepness_Q;        /* in [0,1]; the shader-side Q_i is derived by wt_set_waves (render-engine.cpp:1073) */
    float dir_x, dir_z;       /* xzDirection_D, unit length +-1e-3 */
} wt_wave;

/* Camera constructor arguments — reference src/camera.h:49, src/render-engine.cpp:52.
 * yaw/pitch are stored raw, as the constructor does (camera.cpp:47-54); defaults -90 / 0. */
typedef struct wt_camera {
    float fov_deg, aspect, z_near, z_far;
    float pos[3];
    float yaw_deg, pitch_deg;
} wt_camera;

/* The RenderEngine public fields the path reads — reference src/render-engine.h:224-238. */
typedef struct wt_params {
    float heightmapDisplacementScale;   /* default 1.0  */
    float heightmapSampleScale;         /* default 0.02 */
    float verticalBounceWaveAmplitude;  /* default 0.1  */
    float verticalBounceWavePhase;      /* default 0.0, in [0,1] */
    float waveAnimationTimeInSeconds;   /* default 0.0  */
    float timeOfDayInHours;             /* default 9.0  */
    float sunShininess;                 /* default 50.0 */
    float sunStrength;                  /* default 1.0  */
} wt_params;

/* The animation controls of the RenderEngine — reference src/render-engine.h:218-219,226-227. */
typedef struct wt_animation {
    float animationSpeedTimeOfDayInSecondsPerHour;                /* default 1.0, >= 0 */
    float animationSpeedVerticalBounceWavePhasePeriodInSeconds;   /* default 3.0, >= 0 */
    int isAnimatingTimeOfDay;                                     /* default 0 */
    int isAnimatingWaves;                                         /* default 1 */
} wt_animation;

/* One draw call's uniform set, as uploaded at render-engine.cpp:1057-1110 (in-scope subset;
 * names follow the GLSL uniforms of water-grid.vert:37-64 / water-grid.frag:39-47).
 * Matrices are column-major like glm::value_ptr. */
typedef struct wt_frame {
    float bottomLeftGridPointInWorld[4];
    float bottomRightGridPointInWorld[4];
    float topLeftGridPointInWorld[4];
    float topRightGridPointInWorld[4];
    float viewProjection[16];
    float viewMatOnlyYaw[16];
    float cameraPosition[3];
    float waveAnimationTimeInSeconds;
    float verticalBounceWaveDisplacement;
    float heightmapDisplacementScale;
    float heightmapSampleScale;
    float sunShininess;
    float sunPosition[3];
    float sunStrength;
} wt_frame;

/* Order-independent digest of one evaluation (sums of IEEE-754 bit patterns, mod 2^64). */
typedef struct wt_checksum {
    uint64_t position_bits;   /* sum over vertices of bits(x)+bits(y)+bits(z) */
    uint64_t normal_bits;     /* sum over vertices of bits(nx)+bits(ny)+bits(nz)+bits(w) */
    uint64_t shade_bits;      /* sum over vertices of the four shade words (0 without WT_FLAG_SHADE) */
    float min_y, max_y;       /* extrema of position.y */
    uint64_t vertices;        /* number of vertices folded in */
} wt_checksum;

/* ---------------------------------------------------------------------------------------------
 * Library / error reporting
 * ------------------------------------------------------------------------------------------- */
const char* wt_version(void);
const char* wt_last_error(const wt_ctx* ctx);   /* ctx may be NULL: last error of a ctx-less call on this thread */
int wt_device_count(void);                      /* number of usable CUDA devices (0 on a CPU-only host) */

/* ---------------------------------------------------------------------------------------------
 * Host math — the reference's CPU prologue, bit-faithful fp32 (same operation order as the
 * reference + GLM 0.9.9.8).  No GPU needed.
 * ------------------------------------------------------------------------------------------- */

/* RenderEngine defaults (render-engine.h:224-238) and the default camera (render-engine.cpp:52). */
void wt_default_params(wt_params* out);
void wt_default_camera(wt_camera* out, float aspect);
/* The reference's hard-coded wave set (render-engine.cpp:46-49); writes 4 waves. */
void wt_default_waves(wt_wave out[4]);

/* Camera::getViewMat/getProjectionMat and the per-frame products of render-engine.cpp:278-285.
 * Any output pointer may be NULL. */
int wt_camera_matrices(const wt_camera* cam, float view[16], float proj[16], float view_proj[16],
                       float view_only_yaw[16], float forward[3]);

/* GerstnerWave::TotalAmplitude() as the constructors accumulate it (render-engine.h:156-157). */
float wt_total_amplitude(const wt_wave* waves, unsigned count);
/* steepness_Q_i of render-engine.cpp:1073 with Count() == count. */
float wt_steepness_Q_i(const wt_wave* wave, unsigned count);
/* sunPosition of render-engine.cpp:289-291. */
void wt_sun_position(float time_of_day_hours, float out[3]);
/* verticalBounceWaveDisplacement of render-engine.cpp:299-300. */
float wt_bounce_displacement(float amplitude, float phase);

/* The reference's per-UI-frame animation step: time of day (program.cpp:220-225), wave time and bounce phase
 * (program.cpp:452-464), advanced by delta_time_s (the reference passes 1 / ImGui framerate).  Accumulates in
 * fp32 exactly like the reference (so frame f is NOT f * dt rounded once). */
void wt_default_animation(wt_animation* out);
int wt_animation_advance(wt_params* params, const wt_animation* anim, float delta_time_s);

/* Projected-grid builder, render-engine.cpp:797-1042.  displaceable_amplitude is
 * TotalAmplitude() + heightmapDisplacementScale + verticalBounceWaveAmplitude (:803).
 * corners[0..3] = bottomLeft, topLeft, bottomRight, topRight (the order of waterGridCornerPoints,
 * :1012-1015).  Returns WT_OK, WT_NO_GRID (:894) or WT_ERR_DEGENERATE (:1034). */
int wt_project_grid(const wt_camera* cam, float displaceable_amplitude, float corners[4][4]);

/* Everything RenderEngine::render() computes on the CPU for the water draw, in one call:
 * matrices (:278-285), sun (:289-291), bounce (:299-300), projected grid (:797-1042) -> wt_frame. */
int wt_build_frame(const wt_camera* cam, const wt_params* params, float total_wave_amplitude, wt_frame* out);

/* The wave-parameter interface in physical terms (north_star: "per-wave amplitude, wavelength, steepness, direction and
 * speed") on top of the reference-native 5-tuple — render-engine.h:138-142: "w = 2/L (roughly), where L = wavelength",
 * "phi = S x w, where S = speed" (GPU Gems ch.1).  convention picks the wavenumber: WT_WAVENUMBER_GPU_GEMS w = 2/L (the
 * reference's comment), WT_WAVENUMBER_PHYSICAL w = 2 pi / L (what the seeded benchmark scenes use).  The direction is
 * normalised like the UI does (program.cpp:398). */
#define WT_WAVENUMBER_GPU_GEMS 0
#define WT_WAVENUMBER_PHYSICAL 1
int wt_wave_from_wavelength_speed(float amplitude, float wavelength, float steepness, float dir_x, float dir_z, float speed,
                                  int convention, wt_wave* out);

/* The reference's parameter surface: what its ImGui panel does to a value after an edit ("force-clamp (handle CTRL +
 * LEFT_CLICK)", program.cpp:201-479).  Note the asymmetry the reference has: some fields are clamped at both ends
 * (glm::clamp), others only from below (`if (v < 0) v = 0`) — a CTRL-click can push those past the slider's maximum. */
int wt_wave_clamp_ui(wt_wave* wave);                 /* program.cpp:376-401: A, w, phi >= 0; Q in [0,1]; D clamped to [-1,1]^2, normalised, 0 -> (0,1) */
int wt_params_clamp_ui(wt_params* params);           /* :201-204, :242-250, :410-437 */
int wt_animation_clamp_ui(wt_animation* anim);       /* :212-215, :444-447 */
/* the SliderFloat bounds themselves, in UI order */
typedef struct wt_ui_range {
    const char* field;          /* RenderEngine / GerstnerWave field name */
    float slider_min, slider_max;
    int clamped_below, clamped_above;   /* which ends the force-clamp code enforces */
} wt_ui_range;
#define WT_UI_RANGE_COUNT 23
void wt_ui_ranges(wt_ui_range out[WT_UI_RANGE_COUNT]);

/* Validity domain of WT_FLAG_FAST_PHASE.  The reference rounds every wave's phase to fp32 at every vertex
 * (water-grid.vert:78), i.e. to ulp(w D.x + phi t): the bit-faithful kernel reproduces those roundings, the recurrence
 * kernel evaluates a cleaner phase and therefore differs from the REFERENCE by about that noise — per wave A_i ulp(k_i) in
 * position, and that over twice the vertex spacing in the finite-difference normal.  The report estimates both deviations
 * for one frame (a statistical bound: variances of the unshared roundings summed over the waves; the constants are the
 * upper envelope of deviations measured over grids 1024..16384, 8..256 waves, t = 0..3600 s and the edge cameras,
 * profiles/parity_margins.json) and compares them with the north-star
 * tolerances with a margin of WT_PHASE_MARGIN.  Host only; returns 1 (valid), 0 (use the bit-faithful kernel) or < 0. */
typedef struct wt_phase_report {
    float est_position_error, position_tolerance;      /* recurrence kernel; world units; tolerance = 1e-4 x total amplitude */
    float est_normal_error_rad, normal_tolerance_rad;  /* recurrence kernel; radians; tolerance = 1e-3 */
    float min_vertex_spacing;                          /* smallest distance between neighbouring grid vertices */
    float max_phase;                                   /* largest |k| any wave reaches on this frame */
    int valid;                                         /* 1: the recurrence kernel is inside its validity domain */
    /* The bit-faithful kernel reproduces k and the running sums bit for bit but, like the recurrence kernel, shares every
     * displaced position between the stencils of its neighbours, where the shader re-evaluates them at uv +- du
     * (water-grid.vert:164-182).  Once ulp(k) is large (late in an animation) that re-evaluation flips
     * roundings of k and the SHADER's normals carry noise of A_i ulp(k_i) / spacing that only an evaluation at exactly the
     * shader's arguments reproduces (WT_FLAG_STENCIL_ORDER / WT_FLAG_REFERENCE_ORDER).  The arguments differ from the
     * neighbours' own only at some grid indices (see WT_FLAG_STENCIL_ORDER) and at none when 1 / (grid_len - 1) is a power of
     * two (513, 1025, 2049 ...): the estimate is 0 there, the bit-faithful kernel is valid at any animation time. */
    float est_shared_stencil_normal_error_rad;
    int shared_stencil_valid;                          /* 1: the bit-faithful kernel is inside ITS validity domain */
    /* What one ulp of the fp32 coordinates does to a finite-difference normal on this grid: the floor for ANY evaluation whose
     * sin / cos are not bit-identical to the reference's libm.  Above the tolerance on very fine grids. */
    float coordinate_ulp_normal_error_rad;
    int recommended;                                   /* WT_PHASE_*: what WT_FLAG_AUTO_PHASE runs on this frame */
} wt_phase_report;
#define WT_PHASE_BIT_FAITHFUL 0
#define WT_PHASE_RECURRENCE 1
#define WT_PHASE_REFERENCE_ORDER 2
#define WT_PHASE_STENCIL_ORDER 3
#define WT_PHASE_MARGIN 1.0f        /* valid iff estimate x margin <= tolerance */
#define WT_PHASE_MARGIN_STENCIL 1.5f /* the same for the shared-stencil term: the measured deviation reaches 0.97 of the estimate (16384^2 seen from
                                       above, t = 600 s) and the next tier costs the bit-faithful kernel + 5 % at 4096^2, so it is taken early */
#define WT_PHASE_K_POSITION 4.5     /* extreme of the position noise in modelled sigmas */
#define WT_PHASE_K_NORMAL 3.7       /* height-noise term of the normal estimate: envelope 2.85 x 1.27 (the kernel's own MUFU / recurrence noise) */
#define WT_PHASE_K_FLIP 0.35        /* flipped-rounding term: envelope 0.274 x 1.27 */
#define WT_PHASE_K_STENCIL 1.0      /* shared-stencil term: two coinciding flips of A_i ulp(k_i) over the 2-spacing baseline */
int wt_fast_phase_report(const wt_wave* waves, unsigned count, const wt_frame* frame, unsigned grid_len, wt_phase_report* out);

/* ---------------------------------------------------------------------------------------------
 * Device context
 * ------------------------------------------------------------------------------------------- */
int wt_create(wt_ctx** out, int cuda_device);
void wt_destroy(wt_ctx* ctx);

/* Replaces the per-wave uniform upload loop (render-engine.cpp:1067-1081): derives Q_i with
 * Count() == count and stages the table for the kernels.  count in [0, WT_MAX_WAVES]. */
int wt_set_waves(wt_ctx* ctx, const wt_wave* waves, unsigned count);
/* wt_fast_phase_report on the context's wave set: 1 = the recurrence kernel is inside its validity domain on this
 * frame, 0 = not (WT_FLAG_AUTO_PHASE then runs the bit-faithful kernel), < 0 error.  report may be NULL. */
int wt_fast_phase_valid(const wt_ctx* ctx, const wt_frame* frame, unsigned grid_len, wt_phase_report* report);
/* Kernel of the most recent evaluation launch of this context: WT_PHASE_BIT_FAITHFUL (0), WT_PHASE_RECURRENCE (1),
 * WT_PHASE_REFERENCE_ORDER (2), WT_PHASE_STENCIL_ORDER (3); -1 none yet. */
int wt_last_phase_mode(const wt_ctx* ctx);
/* Evaluation launches of this context so far by kernel: out[WT_PHASE_*]. */
void wt_phase_mode_counts(const wt_ctx* ctx, uint64_t out[4]);

/* Replaces load2DTexture + create2DTexture (render-engine.cpp:1308-1322, texture.cpp:54-72):
 * tightly packed RGBA8, rows ALREADY in GL order (i.e. flipped as stbi_set_flip_vertically_on_load
 * does), sampled GL_LINEAR / GL_MIRRORED_REPEAT / LOD 0 with exact fp32 weights. */
int wt_set_heightmap_rgba8(wt_ctx* ctx, const uint8_t* rgba, int width, int height);

/* Replaces the skybox cubemap binding (render-engine.cpp:90-110, :1092-1094): six n x n RGBA8 faces
 * in GL order +x,-x,+y,-y,+z,-z; GL_LINEAR, GL_CLAMP_TO_EDGE, non-seamless. */
int wt_set_skybox_rgba8(wt_ctx* ctx, const uint8_t* const faces[6], int n);

/* ---------------------------------------------------------------------------------------------
 * Evaluation — replaces glDrawElements on the water program (render-engine.cpp:1115) for the
 * vertex stage (water-grid.vert:114-194) and, with WT_FLAG_SHADE, the vertex-evaluable fragment
 * terms (water-grid.frag:71-73,95-111).
 *
 * Vertex id layout is the shader's (water-grid.vert:115-125): id = row * grid_len + col, row 0 =
 * bottom.  Rows [row_begin,row_end) are evaluated (row sharding; the +-1-row normal stencil is
 * recomputed locally); output element k = (row - row_begin) * grid_len + col of each buffer:
 *     d_position  float4  (x, y, z, 1)                          world-space `position`
 *     d_normal    float4  (nx, ny, nz, fresnel_f_theta | 0)      `normal` after the view-facing flip
 *     d_shade     float4  (sky.r, sky.g, sky.b, sun specular)    only with WT_FLAG_SHADE (else may be NULL)
 * d_checksum (device pointer to ONE wt_checksum, zero-initialised by the call) is only touched with
 * WT_FLAG_CHECKSUM.  `frame` is a HOST pointer; it is consumed before the call returns.
 * ------------------------------------------------------------------------------------------- */
int wt_eval_frame(wt_ctx* ctx, const wt_frame* frame, unsigned grid_len, unsigned row_begin, unsigned row_end,
                  void* d_position, void* d_normal, void* d_shade, void* d_checksum, unsigned flags, void* cuda_stream);

/* Every output stream of the vertex stage.  position/normal are mandatory; shade with WT_FLAG_SHADE; the three
 * varyings buffers with WT_FLAG_VARYINGS; checksum with WT_FLAG_CHECKSUM.  All device pointers, one float4 per
 * vertex, element k = (row - row_begin) * grid_len + col:
 *     clip_position  gl_Position = viewProjection * position                         (water-grid.vert:150)
 *     varyings_a     (xyPositionNDCSpaceHeight0.x, .y, heightmap_colour.r, 0)         (:153-154, :141)
 *     varyings_b     (normalVecInViewSpaceOnlyYaw.xyz, 0)                             (:193)
 * viewVecRaw (:158) is cameraPosition - position and is not stored.  heightmap_colour is a debug varying of
 * which the shader consumes .r only (:110); g, b, a are not produced. */
typedef struct wt_outputs {
    void* position;
    void* normal;
    void* shade;
    void* clip_position;
    void* varyings_a;
    void* varyings_b;
    void* checksum;
} wt_outputs;

/* wt_eval_frame with the full set of output streams. */
int wt_eval_frame_ex(wt_ctx* ctx, const wt_frame* frame, unsigned grid_len, unsigned row_begin, unsigned row_end,
                     const wt_outputs* outputs, unsigned flags, void* cuda_stream);

/* The index buffer glDrawElements consumes — replaces the tessellation loop of src/program.cpp:686-717: two
 * counter-clockwise triangles per grid cell, (bl, tr, tl) and (bl, br, tr), cells row-major from the bottom-left.
 * Writes 6 (grid_len-1)^2 indices to `out` (HOST memory; may be NULL to query) and the count to *count. */
int wt_grid_indices(unsigned grid_len, uint32_t* out, uint64_t capacity, uint64_t* count);

/* ---------------------------------------------------------------------------------------------
 * Compose pass — the screen-space layers of water-grid.frag evaluated per vertex (:66-69,76-89,113-162: depth
 * delta, local reflection / refraction lookups, tint, Fresnel mix, soft edges, fog) on the outputs of an evaluation
 * with WT_FLAG_SHADE | WT_FLAG_VARYINGS.  The fragment inputs a rasteriser would interpolate are taken at the vertex.
 * ------------------------------------------------------------------------------------------- */
/* The RenderEngine fields / uniforms the layers read — reference src/render-engine.h:221-223,229,234-235 and the
 * uniform uploads src/render-engine.cpp:1096,1100,1107,1109-1110. */
typedef struct wt_compose_params {
    float fogColourFarAtNoon[4];          /* default (1,1,1,0.1); darkened by clamp(sunPosition.y,0,1) per frame (render-engine.cpp:295) */
    float fogDepthRadiusFar;              /* default 1.0 */
    float fogDepthRadiusNear;             /* default 0.0 */
    float softEdgesDeltaDepthThreshold;   /* default 0.05 */
    float tintDeltaDepthThreshold;        /* default 0.1 */
    float waterClarity;                   /* default 0.3 */
    float zNear, zFar;                    /* the camera's clip distances */
    float viewportWidth, viewportHeight;
} wt_compose_params;
void wt_default_compose_params(wt_compose_params* out, const wt_camera* cam, unsigned viewport_width, unsigned viewport_height);
/* program.cpp:254-272 (fog colour; the two fog radii push each other: edited_fog_near says which slider moved), :466-479 */
int wt_compose_params_clamp_ui(wt_compose_params* params, int edited_fog_near);

/* Replaces the three FBO attachments the water program samples (render-engine.cpp:128-188, bound at :1060,:1087-1088):
 * depth (float, window-space, GL_NEAREST), local reflections and refractions (RGBA8, GL_LINEAR), all CLAMP_TO_EDGE,
 * w x h texels, rows in GL order.  HOST pointers, copied during the call. */
int wt_set_screen_textures(wt_ctx* ctx, const float* depth, const uint8_t* reflections_rgba, const uint8_t* refractions_rgba,
                           int width, int height);

/* colour (float4 per vertex: rgb after fog, a = edge hardness) from the vertex stage's outputs in `vertex_stage`:
 * position, normal and shade (an evaluation with WT_FLAG_SHADE) are required; clip_position, varyings_a and varyings_b
 * (WT_FLAG_VARYINGS) are read when all three are given, else recomputed from position and normal — 64 instead of
 * 112 bytes of traffic per vertex, and the faster form (0.27 vs 0.35 ms per 4096^2 frame on a B200).  The reflection /
 * refraction footprints are fetched by the texture unit (exact unorm8 conversion, exact fp32 weights in the blend).
 * Stream-ordered. */
int wt_compose_frame(wt_ctx* ctx, const wt_frame* frame, const wt_compose_params* params, unsigned grid_len, unsigned row_begin,
                     unsigned row_end, const wt_outputs* vertex_stage, void* d_colour, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Skybox synthesis — replaces the cubemap render pass of src/render-engine.cpp:306-442 that produces the `skybox`
 * the water samples: per texel the clear colour, then stars (skybox-stars.frag), the sky sphere (skysphere.frag), clouds
 * (skybox-clouds.frag) and the fog quad, blended in that order with GL_SRC_ALPHA / GL_ONE_MINUS_SRC_ALPHA into RGBA8.
 * ------------------------------------------------------------------------------------------- */
typedef struct wt_sky_params {          /* RenderEngine fields, src/render-engine.h:220-221,228,230-233 */
    float timeOfDayInHours;             /* default 9.0 */
    float cloudProportion;              /* default 0.6 */
    float overcastStrength;             /* default 0.0 */
    float sunHorizonDarkness;           /* default 0.25 */
    float sunShininess, sunStrength;    /* defaults 50.0, 1.0 */
    float fogColourFarAtNoon[4];        /* default (1,1,1,0.1) */
    int drawStars, drawSkysphere, drawClouds;   /* the m_isVisible switches of the three sky objects; default 1 */
} wt_sky_params;
void wt_default_sky_params(wt_sky_params* out);
int wt_sky_params_clamp_ui(wt_sky_params* params);   /* program.cpp:201-204, :227-257 */

/* The three source textures: the 1-D sky gradient (program.cpp:644; RGBA8, GL_LINEAR, CLAMP_TO_EDGE) and the star and
 * cloud cubemaps (six n x n RGBA8 faces in GL order, rows as loaded — cubemaps are not flipped, render-engine.cpp:1332).
 * stars / clouds may be NULL (layer not drawn).  HOST pointers, copied during the call. */
int wt_set_sky_layers(wt_ctx* ctx, const uint8_t* gradient_rgba, int gradient_len, const uint8_t* const stars[6], int stars_n,
                      const uint8_t* const clouds[6], int clouds_n);
/* Synthesizes the 6 x n x n skybox into the context (exactly as if wt_set_skybox_rgba8 had been called with the result).
 * One kernel: all layers, the RGBA8 faces and the gather atlas the surface kernel samples.  Star / cloud cubemaps of the
 * target's own power-of-two size (the reference's configuration: 2048^2 assets, CUBEMAP_LENGTH 2048 — program.cpp:616,656,
 * render-engine.h:191) are read texel for texel, which is what the cube-map sampler returns there; other sizes go through the
 * bilinear sampler.  Same bytes either way. */
int wt_synthesize_skybox(wt_ctx* ctx, const wt_sky_params* params, int n, void* cuda_stream);
/* Diagnostic: how many texels of the most recent wt_synthesize_skybox were recomputed on the exact path (IEEE division, libdevice
 * powf) because the approximate path (MUFU lg2 / ex2, reciprocal multiply) could not exclude a different byte.  Synchronises. */
int wt_debug_sky_exact_texels(wt_ctx* ctx, uint64_t* out);
/* Diagnostic (host only): the per-byte table the synthesis kernel applies for the fog quad (screen-space-quad.frag:45 blended
 * with GL_SRC_ALPHA / GL_ONE_MINUS_SRC_ALPHA into RGBA8): byte c of out[d] = the byte channel c holds after the quad when it
 * held d before, for the fog colour / alpha given (already scaled by the sun height as render-engine.cpp:295 does). */
int wt_debug_fog_table(const float fog_rgba[4], uint32_t out[256]);
/* Copies the context's current skybox faces (6 * n * n RGBA8) to HOST memory; out may be NULL to query n.  Synchronises. */
int wt_get_skybox_rgba8(wt_ctx* ctx, uint8_t* out, int* n_out);

/* A batch of n_frames independent frames (animation; frame sharding takes a subset per GPU).
 * Frame f writes slot (f % ring_depth) of each ring buffer (slot stride = (row_end-row_begin)*grid_len
 * float4) and, with WT_FLAG_CHECKSUM, d_checksums[f].  Stream-ordered, returns without synchronising.
 * Small frames without shading terms / extra varyings are grouped: up to min(ring_depth / 2, 16) consecutive frames run as ONE
 * persistent launch over one work queue, the groups of the two halves of the ring on two internal streams forked off and joined
 * back into cuda_stream (1024^2 x 8 waves: 7.5 us per frame on a ring of 8 or more). */
int wt_eval_batch(wt_ctx* ctx, const wt_frame* frames, unsigned n_frames, unsigned grid_len, unsigned row_begin,
                  unsigned row_end, void* d_position_ring, void* d_normal_ring, void* d_shade_ring, unsigned ring_depth,
                  void* d_checksums, unsigned flags, void* cuda_stream);

/* Headless animation runner — the loop Program::start() drives (program.cpp:94: advance the animation state, render):
 * starting from *params (updated in place), for k = 0 .. n_frames-1 it evaluates the frame with the CURRENT state
 * into ring slot (k % ring_depth) / d_checksums[k], then advances the state `frame_stride` times by delta_time_s.
 * Frame sharding over G GPUs: rank r first advances its own copy of the state r times, then runs with
 * frame_stride = G.  A frame whose frustum misses the water (WT_NO_GRID) launches nothing and leaves a zero
 * checksum; *n_evaluated (may be NULL) counts the frames that were drawn.  Stream-ordered, no synchronisation. */
int wt_run_animation(wt_ctx* ctx, const wt_camera* cam, wt_params* params, const wt_animation* anim, float total_wave_amplitude,
                     float delta_time_s, unsigned frame_stride, unsigned n_frames, unsigned grid_len, unsigned row_begin,
                     unsigned row_end, void* d_position_ring, void* d_normal_ring, void* d_shade_ring, unsigned ring_depth,
                     void* d_checksums, unsigned flags, void* cuda_stream, unsigned* n_evaluated);

/* One frame from HOST inputs to a HOST result (synchronous on `cuda_stream`): wt_build_frame (camera + parameters ->
 * uniforms, projected grid), the evaluation into the caller's device buffers, and the frame's digest copied back to
 * *h_checksum through pinned memory.  Returns WT_NO_GRID (zero digest, nothing launched) when the frustum misses the
 * water.  The per-frame round trip of an application that keeps the vertex data on the GPU, as the reference does. */
int wt_step_host(wt_ctx* ctx, const wt_camera* cam, const wt_params* params, float total_wave_amplitude, unsigned grid_len,
                 unsigned row_begin, unsigned row_end, void* d_position, void* d_normal, void* d_shade, unsigned flags,
                 wt_checksum* h_checksum, void* cuda_stream);

/* The same round trip split in two so that frames overlap: wt_step_submit builds and enqueues a frame and returns a
 * ticket; wt_step_wait blocks until that frame's digest has arrived in host memory (WT_NO_GRID: zero digest).  Up to 4
 * frames may be in flight — the host prepares frame k+1 while the GPU evaluates frame k, as a render loop does.
 * Ordering: a submitted frame sees everything enqueued on `cuda_stream` before the call; its outputs are complete when
 * wt_step_wait returns for its ticket (NOT merely in `cuda_stream` order): consecutive frames run on two internal streams so
 * that one frame's tail overlaps the next frame's head.  A frame that writes the position buffer of a pending frame waits
 * for that frame first, so reusing buffers is safe; give consecutive frames different buffers (a ring of 2) to let them overlap. */
int wt_step_submit(wt_ctx* ctx, const wt_camera* cam, const wt_params* params, float total_wave_amplitude, unsigned grid_len,
                   unsigned row_begin, unsigned row_end, void* d_position, void* d_normal, void* d_shade, unsigned flags,
                   void* cuda_stream, unsigned* ticket);
int wt_step_wait(wt_ctx* ctx, unsigned ticket, wt_checksum* h_checksum);

/* Host-buffer convenience (synchronous): evaluates into ctx-owned device scratch and copies the
 * requested outputs to HOST memory.  Any of h_position/h_normal/h_shade/h_checksum may be NULL. */
int wt_eval_frame_host(wt_ctx* ctx, const wt_frame* frame, unsigned grid_len, unsigned row_begin, unsigned row_end,
                       float* h_position, float* h_normal, float* h_shade, wt_checksum* h_checksum, unsigned flags);

/* ---------------------------------------------------------------------------------------------
 * Measurement helpers (used by bench.py; they launch real kernels on the ctx's device)
 * ------------------------------------------------------------------------------------------- */
/* Number of kernels this ctx has launched so far (evaluation + helper kernels). */
uint64_t wt_launch_count(const wt_ctx* ctx);
/* Frames this ctx has evaluated inside GROUPED launches: wt_eval_batch / wt_run_animation put consecutive small frames (<= 4 M
 * vertices, <= 64 waves, no shading terms / extra varyings, same kernel) that write different ring slots into one persistent
 * launch — one work queue across the frames, one launch per group instead of one per frame. */
uint64_t wt_batched_frame_count(const wt_ctx* ctx);
/* Micro-benchmarks for the roofline denominators: kind 0 = MUFU (sin+cos pairs/s),
 * 1 = FP32 FMA (FLOP/s), 2 = HBM float4 store stream (bytes/s), 3 = packed FP32 FFMA2 (FLOP/s).
 * Synchronous; result in *per_second. */
int wt_measure_peak(wt_ctx* ctx, int kind, double* per_second);
/* Diagnostic: with the environment variable WT_TRACE=1 the production kernel records, per work item
 * (one warp's strip x row chunk), {SM id, CTA, start, end} in %globaltimer nanoseconds.  Copies the
 * records of the most recent launch (4 x uint64 each) to `out`; returns the number of items copied
 * (0 when tracing is off) or < 0.  Synchronises the device. */
int wt_debug_read_trace(wt_ctx* ctx, uint64_t* out, unsigned max_items);
/* Diagnostic: re-reads the WT_* development knobs (WT_DYN, WT_DYN_WARPS, WT_TRACE, WT_DBG, WT_EXTRA_SMEM) from
 * the environment.  Only libraries compiled with -DWT_DEV_KNOBS (tools/build_variant.py) have knobs; the production
 * library ignores the environment and this call does nothing. */
void wt_debug_reload_env(void);
/* Diagnostic (host only, no GPU): the canonical work layout of a grid — the row chunks [begin, end) the production
 * kernel's work items cover for rows [row_begin, row_end), for an instantiation that marches rows_per_step rows per step
 * in phase mode 0 (bit-faithful) or 1 (recurrence).  The layout is a function of grid_len, rows_per_step and the mode
 * only: a chunk that meets the row range is listed with its FULL canonical extent, which is what makes the results of
 * row shards bit-identical to the unsharded evaluation.  Writes (begin, end) pairs to out[0 .. 2 capacity) in launch
 * order and returns the number of chunks (which may exceed capacity: nothing is written then), or < 0. */
int wt_debug_layout(unsigned grid_len, int rows_per_step, int phase_mode, unsigned row_begin, unsigned row_end, int* out, unsigned capacity);
/* Host only: the fraction of the stencil arguments of a grid_len-grid that differ from their neighbours' own uv —
 * fl(fl(i du) -/+ du) != fl((i -/+ 1) du), du = 1 / (grid_len - 1) (water-grid.vert:123-125,164-167): 0.003 at 4096, 0.04 at
 * 1024, 0.19 at 8192, 0.11 at 16384, 0 where du is a power of two.  What the fix-up pass of WT_FLAG_STENCIL_ORDER re-evaluates. */
float wt_stencil_flag_rate(unsigned grid_len);
/* Diagnostic: which form WT_FLAG_STENCIL_ORDER runs — 0 (default): the cheaper one for the grid size (the fix-up form unless
 * more than 30 % of the stencil arguments are flagged, e.g. grid_len 700); 1: one kernel with five full evaluations per vertex
 * (4.5x the bit-faithful kernel); 2: the bit-faithful kernel + fix-up pass.  The tests hold the two against each other. */
int wt_debug_full_stencil(wt_ctx* ctx, int form);

#ifdef __cplusplus
}
#endif
#endif /* WAVETOOL_B200_H_ */
