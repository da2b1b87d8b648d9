// TEST INFRASTRUCTURE — not product code (see oracle/README.md).
//
// Reference-side half of the oracle for the projected-grid prologue.  This file is compiled
// TOGETHER WITH the reference's own, unmodified sources where they lie:
//     /root/reference/src/camera.cpp           (wave_tool::Camera)
//     /root/reference/src/render-engine.h      (geometry::Line/Plane/GerstnerWave,
//                                               utils::linePlaneIntersection)
//     /root/reference/deps/glm                 (GLM 0.9.9.8, all matrix/vector arithmetic)
// The grid construction itself is inline in RenderEngine::render()
// (/root/reference/src/render-engine.cpp:797-1042), a function that also issues OpenGL calls
// and therefore cannot be linked headless; its arithmetic is re-stated here step by step, driving
// the reference's own classes, so that every number still comes out of the reference's Camera,
// its line/plane helper and GLM.  Each step cites the reference line it follows.
#include <array>
#include <cstring>
#include <vector>

#include <glm/glm.hpp>
#include <glm/gtc/constants.hpp>

#include "render-engine.h"   // reference header (types only; no RenderEngine object is created)

using wave_tool::Camera;
namespace geo = wave_tool::geometry;
namespace utl = wave_tool::utils;

extern "C" {

struct wtref_camera {   // Camera ctor arguments, render-engine.cpp:52 / camera.h:49
    float fov_deg, aspect, z_near, z_far;
    float pos[3];
    float yaw_deg, pitch_deg;
};

struct wtref_grid_out {
    int has_grid;                 // 0 => render-engine.cpp:894 skipped the water this frame
    int n_intersections;          // intersectionPoints.size() at :894
    float corners[4][4];          // [0]=bottomLeft [1]=topLeft [2]=bottomRight [3]=topRight (:1012-1015)
    float view[16], proj[16], view_proj[16], view_only_yaw[16];   // :278-285, column-major
    float cam_forward[3];
    float aimpoint_1[3], aimpoint_2[3], aimpoint_final[3];         // :929-954
    float projector_pitch_deg, projector_y;                        // after :963
    float range[4];               // x_min, x_max, y_min, y_max (:982-992)
};

// Builds the reference Camera through its own constructor (camera.cpp:47-54): yaw and pitch are
// stored raw (no mod/clamp — that only happens in rotate()/setRotation(), camera.cpp:80-98) and the
// basis is derived by updateBasisVectors() (camera.cpp:130-138).  RenderEngine uses the defaults
// yaw = -90, pitch = 0 (render-engine.cpp:52, camera.cpp:40-45).
static Camera make_camera(const wtref_camera& c) {
    return Camera(c.fov_deg, c.aspect, c.z_near, c.z_far, glm::vec3(c.pos[0], c.pos[1], c.pos[2]),
                  Camera::DEFAULT_FORWARD, Camera::DEFAULT_UP, c.yaw_deg, c.pitch_deg);
}

static void put(float* dst, const glm::mat4& m) { std::memcpy(dst, &m[0][0], 64); }
static void put3(float* dst, const glm::vec3& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

// sum of amplitudes exactly as the GerstnerWave ctor accumulates it (render-engine.h:156-157):
// totalAmplitude += amplitude_A, in construction order, starting from 0.0f
float wtref_total_amplitude(const float* amplitudes, unsigned n) {
    float total = 0.0f;
    for (unsigned i = 0; i < n; ++i) total += amplitudes[i];
    return total;
}

// shader-side steepness, render-engine.cpp:1073
float wtref_steepness_Q_i(float Q, float w, float A, unsigned count) {
    return (w * A) != 0.0f ? Q / (w * A * count) : 0.0f;
}

// sun direction, render-engine.cpp:289-291
void wtref_sun_position(float time_of_day_hours, float out3[3]) {
    float const days{time_of_day_hours / 24.0f};
    float const theta{days * glm::two_pi<float>() - glm::half_pi<float>()};
    out3[0] = glm::cos(theta);
    out3[1] = glm::sin(theta);
    out3[2] = 0.0f;
}

// vertical bounce, render-engine.cpp:299-300
float wtref_bounce_displacement(float amplitude, float phase) {
    float const shift{phase * glm::two_pi<float>()};
    return amplitude * glm::sin(shift);
}

// render-engine.cpp:278-286 + :797-1036.  displaceable_amplitude = TotalAmplitude() +
// heightmapDisplacementScale + verticalBounceWaveAmplitude (:803), formed by the caller.
int wtref_project_grid(const wtref_camera* cin, float displaceable_amplitude, wtref_grid_out* o) {
    std::memset(o, 0, sizeof(*o));
    Camera camera = make_camera(*cin);

    // :278-286
    glm::mat4 const view = camera.getViewMat();
    Camera yawOnly{camera};
    yawOnly.setRotation(yawOnly.getYaw(), 0.0f);
    glm::mat4 const viewYawOnly{yawOnly.getViewMat()};
    glm::mat4 const projection = camera.getProjectionMat();
    glm::mat4 const viewProjection = projection * view;
    glm::mat4 const inverseViewProjection = glm::inverse(viewProjection);
    put(o->view, view); put(o->proj, projection); put(o->view_proj, viewProjection); put(o->view_only_yaw, viewYawOnly);
    put3(o->cam_forward, camera.getForward());

    float const D = displaceable_amplitude;
    geo::Plane const upper{0.0f, 1.0f, 0.0f, D};       // :807
    geo::Plane const base{0.0f, 1.0f, 0.0f, 0.0f};     // :808
    geo::Plane const lower{0.0f, 1.0f, 0.0f, -D};      // :809

    // :815-832  NDC cube corners, index = 4*right + 2*top + far, xy padded by 1.2
    std::array<glm::vec4, 8> corner;
    for (int i = 0; i < 8; ++i) {
        corner[i] = glm::vec4{(i & 4) ? 1.0f : -1.0f, (i & 2) ? 1.0f : -1.0f, (i & 1) ? 1.0f : -1.0f, 1.0f};
        corner[i].x *= 1.2f;
        corner[i].y *= 1.2f;
    }
    // :835-838
    for (auto& c : corner) {
        glm::vec4 const t{inverseViewProjection * c};
        c = t / t.w;
    }
    // :842-853 edge list
    static const unsigned edges[24] = {0,1, 0,2, 0,4, 1,3, 1,5, 2,3, 2,6, 3,7, 4,5, 4,6, 5,7, 6,7};

    std::vector<glm::vec4> hits;
    for (unsigned e = 0; e < 12; ++e) {                 // :860-883
        geo::Line const line{corner[edges[2 * e]], corner[edges[2 * e + 1]]};
        float const lo = glm::min(line.p0.y, line.p1.y), hi = glm::max(line.p0.y, line.p1.y);
        for (geo::Plane const* pl : {&upper, &lower}) {
            if (lo <= pl->d && pl->d <= hi) {
                glm::vec3 p;
                if (utl::linePlaneIntersection(p, line, *pl)) hits.push_back(glm::vec4{p, 1.0f});
            }
        }
    }
    for (auto const& c : corner)                        // :887-891
        if (lower.d <= c.y && c.y <= upper.d) hits.push_back(c);

    o->n_intersections = (int)hits.size();
    if (hits.empty()) return 1;                         // :894 -> no water this frame
    o->has_grid = 1;

    // :905-917 projector = camera copy lifted out of the displaceable slab
    Camera projector{camera};
    float const camY{camera.getPosition().y};
    bool const underwater{camY < 0.0f};
    float const minY{D + 1.0f};
    if (camY < minY) {
        if (underwater) projector.translate(glm::vec3{0.0f, minY - 2.0f * camY, 0.0f});
        else projector.translate(glm::vec3{0.0f, minY - camY, 0.0f});
    }

    // :920-926 horizon-safe forward
    glm::vec3 fwdSafe{camera.getForward()};
    if (glm::abs(fwdSafe.y) < 0.001f) fwdSafe.y = (fwdSafe.y >= 0.0f ? 1.0f : -1.0f) * 0.001f;

    // :929-939 aimpoint 1
    glm::vec3 aim1{0.0f};
    bool const lookingDown{fwdSafe.y < 0.0f};
    if (lookingDown != underwater) {
        utl::linePlaneIntersection(aim1, geo::Line{camera.getPosition(), camera.getPosition() + fwdSafe}, base);
    } else {
        glm::vec3 const mirrored{glm::reflect(fwdSafe, base.getNormalVec())};
        utl::linePlaneIntersection(aim1, geo::Line{camera.getPosition(), camera.getPosition() + mirrored}, base);
    }
    // :943-946 aimpoint 2
    glm::vec3 aim2{camera.getPosition() + 1.0f * camera.getForward()};
    aim2.y = 0.0f;
    // :951-954
    float const a{glm::abs(fwdSafe.y)};
    glm::vec3 const aim{glm::mix(aim2, aim1, a)};
    put3(o->aimpoint_1, aim1); put3(o->aimpoint_2, aim2); put3(o->aimpoint_final, aim);

    // :957-963 aim the projector
    glm::vec3 const pf{glm::normalize(aim - projector.getPosition())};
    glm::vec3 const pfXZ{glm::normalize(glm::vec3{pf.x, 0.0f, pf.z})};
    float const pitchDeg{-glm::degrees(glm::acos(glm::dot(pf, pfXZ)))};
    projector.setRotation(projector.getYaw(), pitchDeg);
    o->projector_pitch_deg = projector.getPitch();
    o->projector_y = projector.getPosition().y;

    // :967-979 flatten, then to projector NDC
    for (auto& h : hits) h.y = 0.0f;
    glm::mat4 const projectorVP{projector.getProjectionMat() * projector.getViewMat()};
    for (auto& h : hits) {
        glm::vec4 const t{projectorVP * h};
        h = t / t.w;
    }
    // :982-992 (note the else-if structure of the reference's min/max scan)
    float x_min = hits[0].x, x_max = hits[0].x, y_min = hits[0].y, y_max = hits[0].y;
    for (size_t i = 1; i < hits.size(); ++i) {
        if (hits[i].x < x_min) x_min = hits[i].x; else if (hits[i].x > x_max) x_max = hits[i].x;
        if (hits[i].y < y_min) y_min = hits[i].y; else if (hits[i].y > y_max) y_max = hits[i].y;
    }
    o->range[0] = x_min; o->range[1] = x_max; o->range[2] = y_min; o->range[3] = y_max;

    // :1002-1008
    glm::mat4 const rangeMat{x_max - x_min, 0.0f, 0.0f, 0.0f,
                             0.0f, y_max - y_min, 0.0f, 0.0f,
                             0.0f, 0.0f, 1.0f, 0.0f,
                             x_min, y_min, 0.0f, 1.0f};
    glm::mat4 const projectorMat{glm::inverse(projection * projector.getViewMat()) * rangeMat};

    // :1012-1036 four corners: near/far points through projectorMat, line ∩ base plane
    static const float uvs[4][2] = {{0.0f, 0.0f}, {0.0f, 1.0f}, {1.0f, 0.0f}, {1.0f, 1.0f}};
    for (int i = 0; i < 4; ++i) {
        glm::vec4 p0{uvs[i][0], uvs[i][1], -1.0f, 1.0f};
        glm::vec4 p1{p0};
        p1.z = 1.0f;
        p0 = projectorMat * p0;
        p0 /= p0.w;
        p1 = projectorMat * p1;
        p1 /= p1.w;
        glm::vec3 hit{0.0f};
        bool const ok = utl::linePlaneIntersection(hit, geo::Line{p0, p1}, base);
        if (!ok) return -2;   // the reference asserts here (:1034); Release builds read garbage
        o->corners[i][0] = hit.x; o->corners[i][1] = hit.y; o->corners[i][2] = hit.z; o->corners[i][3] = 1.0f;
    }
    return 0;
}

}  // extern "C"
