// CPU prologue of the water path: camera matrices, sun/bounce scalars and the projected-grid
// builder.  The reference runs exactly this on the CPU too (4 points per frame) — it is not a
// "CPU fallback" of the kernel.  fp32 throughout, operation order as in the reference + GLM so that
// the four corner points handed to the kernel are the ones the reference would upload.
//
// Replaces: Camera (src/camera.cpp:47-146), RenderEngine::render's per-frame scalars
// (src/render-engine.cpp:278-300) and its projected-grid block (src/render-engine.cpp:797-1042),
// geometry::Plane/Line + utils::linePlaneIntersection (src/render-engine.h:53-124).
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

}  // extern "C"
