// CPU prologue of the water path: camera matrices, sun/bounce scalars and the projected-grid
// builder.  The reference runs exactly this on the CPU too (4 points per frame) — it is not a
// "CPU fallback" of the kernel.  fp32 throughout, operation order as in the reference + GLM so that
// the four corner points handed to the kernel are the ones the reference would upload.
//
// Replaces: Camera (src/camera.cpp:47-146), RenderEngine::render's per-frame scalars
// (src/render-engine.cpp:278-300) and its projected-grid block (src/render-engine.cpp:797-1042),
// geometry::Plane/Line + utils::linePlaneIntersection (src/render-engine.h:53-124).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/wavetool_b200.h"
#include "wt_host_math.h"
#include "wt_internal.h"

namespace wt {

namespace {

// Camera state after construction / setRotation / translate (src/camera.cpp)
struct CameraState {
    float fov, aspect, zn, zf, yaw, pitch;
    float3h pos, forward, right, up;
    mat4h view, proj;

    explicit CameraState(const wt_camera& c)   // camera.cpp:47-54
        : fov(c.fov_deg), aspect(c.aspect), zn(c.z_near), zf(c.z_far), yaw(c.yaw_deg), pitch(c.pitch_deg),
          pos{c.pos[0], c.pos[1], c.pos[2]} {
        rebuild_basis();
        proj = perspective_rh_no(to_radians(fov), aspect, zn, zf);   // camera.cpp:140-142
    }
    void rebuild_basis() {   // camera.cpp:130-138
        const float3h world_up{0.0f, 1.0f, 0.0f};
        float yr = to_radians(yaw), pr = to_radians(pitch);
        forward = unit(float3h{std::cos(yr) * std::cos(pr), std::sin(pr), std::sin(yr) * std::cos(pr)});
        right = unit(cross3(forward, world_up));
        up = unit(cross3(right, forward));
        rebuild_view();
    }
    void rebuild_view() { view = look_at_rh(pos, add(pos, forward), up); }   // camera.cpp:144-146
    void set_rotation(float yaw_deg, float pitch_deg) {                       // camera.cpp:93-98
        yaw = floor_mod(yaw_deg, 360.0f);
        pitch = clampf(pitch_deg, -89.0f, 89.0f);
        rebuild_basis();
    }
    void lift(float dy) { pos.y = pos.y + dy; pos.x = pos.x + 0.0f; pos.z = pos.z + 0.0f; rebuild_view(); }   // translate({0,dy,0}), camera.cpp:100-104
};

// plane y = d with normal (0,1,0); utils::linePlaneIntersection (render-engine.h:111-124) on the
// infinite line through p0,p1.  The general dot products are kept (0*x + 1*y + 0*z) so NaN/Inf
// propagate exactly as in the reference.
bool intersect_y_plane(float3h& hit, float3h p0, float3h p1, float d) {
    const float3h n{0.0f, 1.0f, 0.0f};
    float3h dir = unit(sub(p1, p0));
    float denom = dot3(n, dir);
    if (0.0f == denom) return false;
    float t = (d - dot3(n, p0)) / denom;
    hit = add(p0, scale(dir, t));
    return true;
}

void store(float dst[16], const mat4h& m) { std::memcpy(dst, m.m, sizeof(float) * 16); }

}  // namespace

int project_grid_impl(const wt_camera& cam_in, float D, float corners[4][4], GridDebug* dbg) {
    CameraState cam(cam_in);
    mat4h view_proj = matmul(cam.proj, cam.view);       // render-engine.cpp:284
    mat4h inv_view_proj = invert(view_proj);             // :286

    // :815-838 — NDC cube (xy padded by 1.2) to world space; index = 4*right + 2*top + far
    float3h fr[8];
    for (int i = 0; i < 8; ++i) {
        float4h ndc{(i & 4) ? 1.0f : -1.0f, (i & 2) ? 1.0f : -1.0f, (i & 1) ? 1.0f : -1.0f, 1.0f};
        ndc.x *= 1.2f;
        ndc.y *= 1.2f;
        float4h w = divide_w(transform(inv_view_proj, ndc));
        fr[i] = {w.x, w.y, w.z};
    }
    // :842-891 — frustum edges against y = +D and y = -D, then frustum corners inside the slab
    static const unsigned char edge[12][2] = {{0, 1}, {0, 2}, {0, 4}, {1, 3}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 6}, {5, 7}, {6, 7}};
    std::vector<float3h> pts;
    pts.reserve(32);
    for (const auto& e : edge) {
        float3h a = fr[e[0]], b = fr[e[1]];
        float ylo = std::fmin(a.y, b.y), yhi = std::fmax(a.y, b.y);
        for (float d : {D, -D}) {
            float3h hit;
            if (ylo <= d && d <= yhi && intersect_y_plane(hit, a, b, d)) pts.push_back(hit);
        }
    }
    for (const auto& c : fr)
        if (-D <= c.y && c.y <= D) pts.push_back(c);
    if (dbg) dbg->n_intersections = (int)pts.size();
    if (pts.empty()) return WT_NO_GRID;                  // :894

    // :905-917 — projector: camera copy lifted out of the displaceable slab (mirrored if underwater)
    CameraState projector = cam;
    float cam_y = cam.pos.y;
    bool underwater = cam_y < 0.0f;
    float min_y = D + 1.0f;
    if (cam_y < min_y) projector.lift(underwater ? min_y - 2.0f * cam_y : min_y - cam_y);

    // :920-939 — horizon-safe forward, aimpoint 1 on the base plane
    float3h fwd = cam.forward;
    if (std::fabs(fwd.y) < 0.001f) fwd.y = (fwd.y >= 0.0f ? 1.0f : -1.0f) * 0.001f;
    bool looking_down = fwd.y < 0.0f;
    float3h ray = (looking_down != underwater) ? fwd : reflect3(fwd, float3h{0.0f, 1.0f, 0.0f});
    float3h aim1{0.0f, 0.0f, 0.0f};
    intersect_y_plane(aim1, cam.pos, add(cam.pos, ray), 0.0f);
    // :943-954 — aimpoint 2 and the blend
    float3h aim2 = add(cam.pos, scale(cam.forward, 1.0f));
    aim2.y = 0.0f;
    float3h aim = lerp3(aim2, aim1, std::fabs(fwd.y));
    // :957-963 — aim the projector
    float3h pf = unit(sub(aim, projector.pos));
    float3h pf_flat = unit(float3h{pf.x, 0.0f, pf.z});
    float pitch = -to_degrees(std::acos(dot3(pf, pf_flat)));
    projector.set_rotation(projector.yaw, pitch);
    if (dbg) {
        dbg->aimpoint_final[0] = aim.x; dbg->aimpoint_final[1] = aim.y; dbg->aimpoint_final[2] = aim.z;
        dbg->projector_pitch_deg = projector.pitch;
        dbg->projector_y = projector.pos.y;
    }

    // :967-992 — flatten the points, take them to projector NDC, bound them
    mat4h projector_vp = matmul(projector.proj, projector.view);
    float x_min = 0, x_max = 0, y_min = 0, y_max = 0;
    for (size_t i = 0; i < pts.size(); ++i) {
        float4h ndc = divide_w(transform(projector_vp, float4h{pts[i].x, 0.0f, pts[i].z, 1.0f}));
        if (i == 0) { x_min = x_max = ndc.x; y_min = y_max = ndc.y; continue; }
        // the reference's scan is an if / else-if per axis
        if (ndc.x < x_min) x_min = ndc.x; else if (ndc.x > x_max) x_max = ndc.x;
        if (ndc.y < y_min) y_min = ndc.y; else if (ndc.y > y_max) y_max = ndc.y;
    }
    if (dbg) { dbg->range[0] = x_min; dbg->range[1] = x_max; dbg->range[2] = y_min; dbg->range[3] = y_max; }

    // :1002-1008 — range matrix; note camera PROJECTION with projector VIEW
    mat4h range = {};
    range.m[0][0] = x_max - x_min;
    range.m[1][1] = y_max - y_min;
    range.m[2][2] = 1.0f;
    range.m[3][0] = x_min; range.m[3][1] = y_min; range.m[3][3] = 1.0f;
    mat4h to_world = matmul(invert(matmul(cam.proj, projector.view)), range);

    // :1012-1036 — unit-square corners: near/far points -> world -> infinite line ∩ y = 0
    static const float uv[4][2] = {{0.0f, 0.0f}, {0.0f, 1.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}};   // BL, TL, BR, TR
    for (int i = 0; i < 4; ++i) {
        float4h n = divide_w(transform(to_world, float4h{uv[i][0], uv[i][1], -1.0f, 1.0f}));
        float4h f = divide_w(transform(to_world, float4h{uv[i][0], uv[i][1], 1.0f, 1.0f}));
        float3h hit{0.0f, 0.0f, 0.0f};
        if (!intersect_y_plane(hit, float3h{n.x, n.y, n.z}, float3h{f.x, f.y, f.z}, 0.0f)) return WT_ERR_DEGENERATE;
        corners[i][0] = hit.x; corners[i][1] = hit.y; corners[i][2] = hit.z; corners[i][3] = 1.0f;
    }
    return WT_OK;
}

}  // namespace wt

using namespace wt;

extern "C" {

void wt_default_params(wt_params* p) {   // render-engine.h:224-238
    p->heightmapDisplacementScale = 1.0f;
    p->heightmapSampleScale = 0.02f;
    p->verticalBounceWaveAmplitude = 0.1f;
    p->verticalBounceWavePhase = 0.0f;
    p->waveAnimationTimeInSeconds = 0.0f;
    p->timeOfDayInHours = 9.0f;
    p->sunShininess = 50.0f;
    p->sunStrength = 1.0f;
}

void wt_default_camera(wt_camera* c, float aspect) {   // render-engine.cpp:52, camera.cpp:40-45
    c->fov_deg = 72.0f; c->aspect = aspect; c->z_near = 0.1f; c->z_far = 100.0f;
    c->pos[0] = 0.0f; c->pos[1] = 4.0f; c->pos[2] = 70.0f;
    c->yaw_deg = -90.0f; c->pitch_deg = 0.0f;
}

void wt_default_waves(wt_wave w[4]) {   // render-engine.cpp:46-49
    float inv_len = 1.0f / std::sqrt(1.0f * 1.0f + 1.0f * 1.0f);   // glm::normalize(vec2{1,1})
    w[0] = {0.06f, 1.0f, 2.0f, 1.0f, 1.0f, 0.0f};
    w[1] = {0.1f, 1.0f, 0.2f, 0.0f, 1.0f * inv_len, 1.0f * inv_len};
    w[2] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f};
    w[3] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f};
}

int wt_camera_matrices(const wt_camera* cam, float view[16], float proj[16], float view_proj[16], float view_only_yaw[16], float forward[3]) {
    if (!cam) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_camera_matrices: cam is NULL");
    CameraState c(*cam);
    if (view) store(view, c.view);
    if (proj) store(proj, c.proj);
    if (view_proj) store(view_proj, matmul(c.proj, c.view));
    if (view_only_yaw) {   // render-engine.cpp:279-281
        CameraState y = c;
        y.set_rotation(y.yaw, 0.0f);
        store(view_only_yaw, y.view);
    }
    if (forward) { forward[0] = c.forward.x; forward[1] = c.forward.y; forward[2] = c.forward.z; }
    return WT_OK;
}

float wt_total_amplitude(const wt_wave* waves, unsigned count) {   // render-engine.h:156-157
    float total = 0.0f;
    for (unsigned i = 0; i < count; ++i) total += waves[i].amplitude_A;
    return total;
}

float wt_steepness_Q_i(const wt_wave* w, unsigned count) {   // render-engine.cpp:1073
    float wa = w->frequency_w * w->amplitude_A;
    return wa != 0.0f ? w->steepness_Q / (wa * count) : 0.0f;
}

void wt_sun_position(float hours, float out[3]) {   // render-engine.cpp:289-291
    float days = hours / 24.0f;
    float theta = days * 6.283185307179586476925286766559f - 1.5707963267948966192313216916398f;
    out[0] = std::cos(theta);
    out[1] = std::sin(theta);
    out[2] = 0.0f;
}

float wt_bounce_displacement(float amplitude, float phase) {   // render-engine.cpp:299-300
    float shift = phase * 6.283185307179586476925286766559f;
    return amplitude * std::sin(shift);
}

int wt_project_grid(const wt_camera* cam, float displaceable_amplitude, float corners[4][4]) {
    if (!cam || !corners) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_project_grid: NULL argument");
    int st = project_grid_impl(*cam, displaceable_amplitude, corners, nullptr);
    if (st == WT_ERR_DEGENERATE) set_thread_error(st, "wt_project_grid: projector ray parallel to the base plane");
    return st;
}

void wt_default_animation(wt_animation* out) {   // render-engine.h:218-219,226-227
    if (!out) return;
    out->animationSpeedTimeOfDayInSecondsPerHour = 1.0f;
    out->animationSpeedVerticalBounceWavePhasePeriodInSeconds = 3.0f;
    out->isAnimatingTimeOfDay = 0;
    out->isAnimatingWaves = 1;
}

// One UI frame of the reference's animation state machine: program.cpp:220-225 (time of day) and :452-464
// (wave time, bounce phase).  fp32, incremental accumulation exactly as the reference does it.
int wt_animation_advance(wt_params* p, const wt_animation* anim, float delta_time_s) {
    if (!p || !anim) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_animation_advance: NULL argument");
    if (anim->isAnimatingTimeOfDay && anim->animationSpeedTimeOfDayInSecondsPerHour > 0.0f) {
        const float dh = (1.0f / anim->animationSpeedTimeOfDayInSecondsPerHour) * delta_time_s;
        p->timeOfDayInHours += dh;
        p->timeOfDayInHours = floor_mod(p->timeOfDayInHours, 24.0f);
    }
    if (anim->isAnimatingWaves) {
        p->waveAnimationTimeInSeconds += delta_time_s;
        if (p->waveAnimationTimeInSeconds < 0.0f) p->waveAnimationTimeInSeconds = 0.0f;
        if (anim->animationSpeedVerticalBounceWavePhasePeriodInSeconds > 0.0f) {
            const float dp = (1.0f / anim->animationSpeedVerticalBounceWavePhasePeriodInSeconds) * delta_time_s;
            p->verticalBounceWavePhase += dp;
            p->verticalBounceWavePhase = floor_mod(p->verticalBounceWavePhase, 1.0f);
        }
    }
    return WT_OK;
}

int wt_build_frame(const wt_camera* cam, const wt_params* p, float total_wave_amplitude, wt_frame* out) {
    if (!cam || !p || !out) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_build_frame: NULL argument");
    std::memset(out, 0, sizeof(*out));
    wt_camera_matrices(cam, nullptr, nullptr, out->viewProjection, out->viewMatOnlyYaw, nullptr);
    out->cameraPosition[0] = cam->pos[0]; out->cameraPosition[1] = cam->pos[1]; out->cameraPosition[2] = cam->pos[2];
    out->waveAnimationTimeInSeconds = p->waveAnimationTimeInSeconds;
    out->verticalBounceWaveDisplacement = wt_bounce_displacement(p->verticalBounceWaveAmplitude, p->verticalBounceWavePhase);
    out->heightmapDisplacementScale = p->heightmapDisplacementScale;
    out->heightmapSampleScale = p->heightmapSampleScale;
    out->sunShininess = p->sunShininess;
    out->sunStrength = p->sunStrength;
    wt_sun_position(p->timeOfDayInHours, out->sunPosition);
    // render-engine.cpp:803
    float D = total_wave_amplitude + p->heightmapDisplacementScale + p->verticalBounceWaveAmplitude;
    float c[4][4];
    int st = wt_project_grid(cam, D, c);
    if (st != WT_OK) return st;
    std::memcpy(out->bottomLeftGridPointInWorld, c[0], 16);
    std::memcpy(out->topLeftGridPointInWorld, c[1], 16);
    std::memcpy(out->bottomRightGridPointInWorld, c[2], 16);
    std::memcpy(out->topRightGridPointInWorld, c[3], 16);
    return WT_OK;
}

// ------------------------------------------------------------------------------------------------
// Parameter surface: the reference's sliders and their "force-clamp" code
// ------------------------------------------------------------------------------------------------
static inline float clamp01(float v) { return std::fmin(std::fmax(v, 0.0f), 1.0f); }   // glm::clamp(v, 0, 1)
static inline float floor0(float v) { return v < 0.0f ? 0.0f : v; }                     // `if (v < 0) v = 0`: NaN passes through like in the reference

int wt_wave_from_wavelength_speed(float amplitude, float wavelength, float steepness, float dir_x, float dir_z, float speed, int convention,
                                  wt_wave* out) {
    if (!out) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_wave_from_wavelength_speed: out is NULL");
    if (!(wavelength > 0.0f) || !(amplitude >= 0.0f) || !(speed >= 0.0f) || !(steepness >= 0.0f && steepness <= 1.0f))
        return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_wave_from_wavelength_speed: need wavelength > 0, amplitude >= 0, speed >= 0, steepness in [0,1]");
    if (convention != WT_WAVENUMBER_GPU_GEMS && convention != WT_WAVENUMBER_PHYSICAL)
        return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_wave_from_wavelength_speed: unknown convention");
    const float len = std::sqrt(dir_x * dir_x + dir_z * dir_z);
    if (!(len > 0.0f)) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_wave_from_wavelength_speed: direction is the zero vector");
    // render-engine.h:139-140: "w = 2/L (roughly)", "phi = S x w" (GPU Gems ch.1 eq. 1-2); the physical wavenumber is 2 pi / L
    const float w = (convention == WT_WAVENUMBER_GPU_GEMS ? 2.0f : 6.283185307179586476925286766559f) / wavelength;
    out->amplitude_A = amplitude;
    out->frequency_w = w;
    out->phaseConstant_phi = speed * w;
    out->steepness_Q = steepness;
    out->dir_x = dir_x / len;     // glm::normalize divides by the length
    out->dir_z = dir_z / len;
    return WT_OK;
}

int wt_wave_clamp_ui(wt_wave* w) {   // program.cpp:376-401
    if (!w) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_wave_clamp_ui: NULL argument");
    w->amplitude_A = floor0(w->amplitude_A);               // :376-379 (slider 0..1, only the lower end is forced)
    w->frequency_w = floor0(w->frequency_w);               // :380-383
    w->phaseConstant_phi = floor0(w->phaseConstant_phi);   // :384-387 (slider 0..10)
    w->steepness_Q = clamp01(w->steepness_Q);              // :388-391
    float x = std::fmin(std::fmax(w->dir_x, -1.0f), 1.0f), z = std::fmin(std::fmax(w->dir_z, -1.0f), 1.0f);   // :392-394
    const float len = std::sqrt(x * x + z * z);
    if (len == 0.0f) { x = 0.0f; z = 1.0f; }               // :397
    else { x /= len; z /= len; }                           // :398
    w->dir_x = x; w->dir_z = z;
    return WT_OK;
}

int wt_params_clamp_ui(wt_params* p) {   // program.cpp:201-204, 242-250, 410-437
    if (!p) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_params_clamp_ui: NULL argument");
    p->timeOfDayInHours = std::fmin(std::fmax(p->timeOfDayInHours, 0.0f), 24.0f);   // :201-204
    p->sunShininess = floor0(p->sunShininess);                                      // :242-245 (slider 0..200)
    p->sunStrength = clamp01(p->sunStrength);                                       // :247-250
    p->heightmapSampleScale = floor0(p->heightmapSampleScale);                      // :410-413
    p->heightmapDisplacementScale = floor0(p->heightmapDisplacementScale);          // :415-418
    p->verticalBounceWaveAmplitude = floor0(p->verticalBounceWaveAmplitude);        // :420-423
    p->verticalBounceWavePhase = clamp01(p->verticalBounceWavePhase);               // :425-428
    p->waveAnimationTimeInSeconds = floor0(p->waveAnimationTimeInSeconds);          // :433-436 (slider 0..3600)
    return WT_OK;
}

int wt_animation_clamp_ui(wt_animation* a) {   // program.cpp:212-215, 444-447
    if (!a) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_animation_clamp_ui: NULL argument");
    a->animationSpeedTimeOfDayInSecondsPerHour = floor0(a->animationSpeedTimeOfDayInSecondsPerHour);
    a->animationSpeedVerticalBounceWavePhasePeriodInSeconds = floor0(a->animationSpeedVerticalBounceWavePhasePeriodInSeconds);
    a->isAnimatingTimeOfDay = a->isAnimatingTimeOfDay ? 1 : 0;
    a->isAnimatingWaves = a->isAnimatingWaves ? 1 : 0;
    return WT_OK;
}

int wt_sky_params_clamp_ui(wt_sky_params* s) {   // program.cpp:201-204, 227-250, 254-257
    if (!s) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_sky_params_clamp_ui: NULL argument");
    s->timeOfDayInHours = std::fmin(std::fmax(s->timeOfDayInHours, 0.0f), 24.0f);
    s->cloudProportion = clamp01(s->cloudProportion);          // :227-230
    s->overcastStrength = clamp01(s->overcastStrength);        // :232-235
    s->sunHorizonDarkness = clamp01(s->sunHorizonDarkness);    // :237-240
    s->sunShininess = floor0(s->sunShininess);                 // :242-245
    s->sunStrength = clamp01(s->sunStrength);                  // :247-250
    for (int i = 0; i < 4; ++i) s->fogColourFarAtNoon[i] = clamp01(s->fogColourFarAtNoon[i]);   // :254-257
    s->drawStars = s->drawStars ? 1 : 0; s->drawSkysphere = s->drawSkysphere ? 1 : 0; s->drawClouds = s->drawClouds ? 1 : 0;
    return WT_OK;
}

int wt_compose_params_clamp_ui(wt_compose_params* c, int edited_fog_near) {   // program.cpp:254-272, 466-479
    if (!c) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_compose_params_clamp_ui: NULL argument");
    for (int i = 0; i < 4; ++i) c->fogColourFarAtNoon[i] = clamp01(c->fogColourFarAtNoon[i]);
    if (edited_fog_near) {   // :266-271 — the slider that moved wins, the other bound is shifted
        c->fogDepthRadiusNear = clamp01(c->fogDepthRadiusNear);
        c->fogDepthRadiusFar = std::fmin(std::fmax(c->fogDepthRadiusFar, c->fogDepthRadiusNear), 1.0f);
    } else {                 // :259-264
        c->fogDepthRadiusFar = clamp01(c->fogDepthRadiusFar);
        c->fogDepthRadiusNear = std::fmin(std::fmax(c->fogDepthRadiusNear, 0.0f), c->fogDepthRadiusFar);
    }
    c->tintDeltaDepthThreshold = clamp01(c->tintDeltaDepthThreshold);             // :466-469
    c->waterClarity = clamp01(c->waterClarity);                                   // :471-474
    c->softEdgesDeltaDepthThreshold = clamp01(c->softEdgesDeltaDepthThreshold);   // :476-479
    return WT_OK;
}

void wt_ui_ranges(wt_ui_range out[WT_UI_RANGE_COUNT]) {   // the SliderFloat bounds of program.cpp:201-479, in UI order
    static const wt_ui_range R[WT_UI_RANGE_COUNT] = {
        {"timeOfDayInHours", 0.0f, 24.0f, 1, 1},                                          // :201
        {"animationSpeedTimeOfDayInSecondsPerHour", 0.0f, 60.0f, 1, 0},                   // :212
        {"cloudProportion", 0.0f, 1.0f, 1, 1}, {"overcastStrength", 0.0f, 1.0f, 1, 1},    // :227, :232
        {"sunHorizonDarkness", 0.0f, 1.0f, 1, 1},                                         // :237
        {"sunShininess", 0.0f, 200.0f, 1, 0}, {"sunStrength", 0.0f, 1.0f, 1, 1},          // :242, :247
        {"fogDepthRadiusFar", 0.0f, 1.0f, 1, 1}, {"fogDepthRadiusNear", 0.0f, 1.0f, 1, 1},// :259, :266
        {"amplitude_A", 0.0f, 1.0f, 1, 0}, {"frequency_w", 0.0f, 1.0f, 1, 0},             // :376, :380
        {"phaseConstant_phi", 0.0f, 10.0f, 1, 0}, {"steepness_Q", 0.0f, 1.0f, 1, 1},      // :384, :388
        {"xzDirection_D", -1.0f, 1.0f, 1, 1},                                             // :392
        {"heightmapSampleScale", 0.0f, 1.0f, 1, 0}, {"heightmapDisplacementScale", 0.0f, 1.0f, 1, 0},   // :410, :415
        {"verticalBounceWaveAmplitude", 0.0f, 1.0f, 1, 0}, {"verticalBounceWavePhase", 0.0f, 1.0f, 1, 1},// :420, :425
        {"waveAnimationTimeInSeconds", 0.0f, 3600.0f, 1, 0},                              // :433
        {"animationSpeedVerticalBounceWavePhasePeriodInSeconds", 0.0f, 60.0f, 1, 0},      // :444
        {"tintDeltaDepthThreshold", 0.0f, 1.0f, 1, 1}, {"waterClarity", 0.0f, 1.0f, 1, 1},// :466, :471
        {"softEdgesDeltaDepthThreshold", 0.0f, 1.0f, 1, 1},                               // :476
    };
    if (out) std::memcpy(out, R, sizeof(R));
}

// ------------------------------------------------------------------------------------------------
// Validity domain of WT_FLAG_FAST_PHASE
// ------------------------------------------------------------------------------------------------
// The reference rounds the phase of every wave at every vertex to fp32:
//     k = fl(fl(w * fl(fl(D.x x) + fl(D.z z))) + fl(phi t))                       (water-grid.vert:78)
// so k carries a deterministic error of up to ulp(k)/2 — 1.5e-5 rad at |k| ~ 300 (a short wave at the far side of the
// default grid), 2e-3 rad at phi t = 36000 (the top of the reference's time slider).  The bit-faithful mode reproduces
// those roundings bit for bit.  The recurrence mode evaluates a CLEANER phase (pre-folded w D, the time offset reduced mod
// 2 pi in double, an affine step per row): it does not share the reference's rounding noise, so it sits ~ulp(k) x A_i
// away from the reference per wave, and its finite-difference normals by that noise over twice the vertex spacing.  This
// estimator adds the variances of the roundings the two paths do not share, per wave, and scales by the spacing of the
// grid; the constants are calibrated against measured deviations (profiles/parity_margins.json).
static inline float ulp_of(float x) {
    x = std::fabs(x);
    if (!(x > 0.0f)) return 0.0f;
    int e;
    std::frexp(x, &e);                       // x = m 2^e, m in [0.5, 1)
    return std::ldexp(1.0f, e - 24);
}

// does fl(fl(i du) -/+ du) differ from fl((i -/+ 1) du) for any index of an L-grid (water-grid.vert:123-125,164-167)?  Cached per L.
static bool stencil_arguments_differ(unsigned L) {
    static thread_local unsigned last_L = 0;
    static thread_local bool last = true;
    if (L == last_L) return last;
    const float du = 1.0f / (float)(L - 1);
    bool any = false;
    for (unsigned i = 0; i < L && !any; ++i) {
        const float u = (float)i * du;
        const float minus = u - du, plus = u + du;
        any = minus != (float)((int)i - 1) * du || plus != (float)(i + 1) * du;
    }
    last_L = L; last = any;
    return any;
}

float wt_stencil_flag_rate(unsigned L) {
    if (L < 2) return 0.0f;
    const float du = 1.0f / (float)(L - 1);
    unsigned long long sides = 0;
    for (unsigned i = 0; i < L; ++i) {
        const float u = (float)i * du;
        const float minus = u - du, plus = u + du;
        sides += (minus != (float)((int)i - 1) * du) + (plus != (float)(i + 1) * du);
    }
    return (float)((double)sides / (2.0 * (double)L));
}

int wt_fast_phase_report(const wt_wave* waves, unsigned count, const wt_frame* f, unsigned L, wt_phase_report* out) {
    if ((count && !waves) || !f || L < 2) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_fast_phase_report: bad argument");
    const float* c[4] = {f->bottomLeftGridPointInWorld, f->bottomRightGridPointInWorld, f->topLeftGridPointInWorld, f->topRightGridPointInWorld};
    double rho = 0.0, coord = 0.0;           // largest |xz| and largest |x| or |z| of the grid (the bilinear hull of its corners)
    for (int i = 0; i < 4; ++i) {
        rho = std::fmax(rho, std::hypot((double)c[i][0], (double)c[i][2]));
        coord = std::fmax(coord, std::fmax(std::fabs((double)c[i][0]), std::fabs((double)c[i][2])));
    }
    // vertex spacing along u and along v at the four corners; the noise-to-angle gain is largest where both are small
    auto dist = [&](int a, int b) { return std::hypot((double)c[a][0] - c[b][0], (double)c[a][2] - c[b][2]) / (double)(L - 1); };
    const double hu[2] = {dist(0, 1), dist(2, 3)}, hv[2] = {dist(0, 2), dist(1, 3)};   // bottom / top edge, left / right edge
    double gain = 0.0, hmin = 1e300;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const double u = std::fmax(hu[a], 1e-30), v = std::fmax(hv[b], 1e-30);
            gain = std::fmax(gain, std::sqrt(1.0 / (u * u) + 1.0 / (v * v)));
            hmin = std::fmin(hmin, std::fmin(u, v));
        }
    const double t = f->waveAnimationTimeInSeconds;
    const float urho = ulp_of((float)rho);
    const double uc = ulp_of((float)coord);
    double var_y = 0.0, var_h = 0.0, total = 0.0, kmax = 0.0, flips = 0.0;
    for (unsigned i = 0; i < count; ++i) {
        const wt_wave& w = waves[i];
        const double a = (double)w.frequency_w * rho, K = a + (double)w.phaseConstant_phi * t;
        kmax = std::fmax(kmax, K);
        // roundings of the reference's k the recurrence mode does not reproduce: the final add (ulp(K)), w * d (ulp(a)),
        // d itself (two products and a sum at magnitude rho, times w); and the mode's own: two FMAs at magnitude ~a + pi and
        // the rounding of the pre-folded w D (systematic, relative 2^-24).  Uniform rounding error: variance ulp^2 / 12.
        const double uK = ulp_of((float)K), ua = ulp_of((float)a), ud = (double)w.frequency_w * urho, um = ulp_of((float)(a + 3.2));
        const double var = (uK * uK + ua * ua + 3.0 * ud * ud + 2.0 * um * um) / 12.0 + (a * 5.96e-8) * (a * 5.96e-8) / 3.0;
        const double Qi = wt_steepness_Q_i(&w, count);
        const double ah = (double)w.amplitude_A * Qi;          // horizontal amplitude A Q_i |D|
        var_y += (double)w.amplitude_A * w.amplitude_A * var;
        var_h += ah * ah * var;
        // The shader adds every wave's term to a running fp32 sum that starts at the grid coordinate (magnitude up to
        // `coord`): a term that differs by a fraction of ulp(coord) flips a rounding of that sum now and then
        if (uc > 0.0) flips += std::fmin(1.0, 2.0 * ah * std::sqrt(var) / uc);
        total += w.amplitude_A;
    }
    wt_phase_report r;
    std::memset(&r, 0, sizeof(r));
    // Extremes over a frame in units of the modelled sigmas; the constants are the upper envelope of the deviations
    // measured over grids 1024 .. 16384, 8 .. 256 waves, t = 0 .. 3600 s and the edge cameras (see the header).
    //   position: the phase noise itself + the flipped roundings of the running sums, quantised in ulp(coord);
    //   normal:   the height noise and those flips, each over the vertex spacing (the flips tilt the finite-difference
    //             tangents of a sloped surface: at the start of an animation they are the larger part).
    const double sig_n = std::sqrt(var_y) * gain, flip_n = uc * gain;
    r.est_position_error = (float)(WT_PHASE_K_POSITION * std::sqrt(std::fmax(var_y, var_h)) + (1.0 + 2.5 * std::sqrt(flips)) * uc);
    r.est_normal_error_rad = (float)std::sqrt(WT_PHASE_K_NORMAL * WT_PHASE_K_NORMAL * sig_n * sig_n + WT_PHASE_K_FLIP * WT_PHASE_K_FLIP * flip_n * flip_n);
    r.position_tolerance = (float)(1e-4 * total);
    r.normal_tolerance_rad = 1e-3f;
    r.min_vertex_spacing = (float)hmin;
    r.max_phase = (float)kmax;
    r.valid = (count == 0) || (r.est_position_error * WT_PHASE_MARGIN <= r.position_tolerance && r.est_normal_error_rad * WT_PHASE_MARGIN <= r.normal_tolerance_rad);
    // The bit-faithful kernel reproduces k and the running sums, but (like the recurrence kernel) it evaluates every displaced
    // position ONCE and shares it between the stencils of its four neighbours, where the shader re-evaluates the neighbours at
    // uv +- du (water-grid.vert:164-182) — an argument that differs from the neighbour's own uv by an ulp of u now and then
    // (SURVEY.md H4).  That difference is harmless while k is fine-grained; once ulp(k) is large it flips the rounding of k
    // in the shader's neighbour evaluation, moving that neighbour by A_i ulp(k_i): the shader's normal then carries noise no
    // shared-stencil kernel can reproduce.  Bound: the largest A_i ulp(k_i) of the frame over the vertex spacing (two such
    // flips do coincide somewhere on a 16 M-vertex grid).  Only the five-evaluation kernels (WT_FLAG_STENCIL_ORDER,
    // WT_FLAG_REFERENCE_ORDER: exactly the shader's arguments) are free of it.
    double a_ulp = 0.0;
    for (unsigned i = 0; i < count; ++i) {
        const double K = (double)waves[i].frequency_w * rho + (double)waves[i].phaseConstant_phi * t;
        a_ulp = std::fmax(a_ulp, (double)waves[i].amplitude_A * ulp_of((float)K));
    }
    r.est_shared_stencil_normal_error_rad = (float)(WT_PHASE_K_STENCIL * a_ulp * gain);
    // ... unless no stencil argument of this grid size differs from its neighbour's own at all (du a power of two: 513, 1025, 2049 ...):
    // the bit-faithful kernel then evaluates exactly what the shader does, at any animation time
    if (!stencil_arguments_differ(L)) r.est_shared_stencil_normal_error_rad = 0.0f;
    // what ONE ulp of the fp32 coordinates does to a finite-difference normal: the floor of every evaluation whose sin / cos are
    // not bit-identical to the reference's libm (one flipped rounding of the running sum moves a neighbour by ulp(coord))
    r.coordinate_ulp_normal_error_rad = (float)(0.5 * uc * gain);
    r.shared_stencil_valid = (count == 0) || (r.est_shared_stencil_normal_error_rad * WT_PHASE_MARGIN_STENCIL <= r.normal_tolerance_rad);
    r.recommended = r.valid ? WT_PHASE_RECURRENCE : (r.shared_stencil_valid ? WT_PHASE_BIT_FAITHFUL : WT_PHASE_STENCIL_ORDER);
    if (out) *out = r;
    return r.valid ? 1 : 0;
}

}  // extern "C"
