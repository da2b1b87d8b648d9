// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call this file.  The product
// (wave_tool_b200/) never links it and has no CPU fallback.
//
// CPU restatement ("port") of the per-vertex ocean-surface path of TexelBox/wave-tool, in plain
// C++17 with no GLM, templated on the scalar type: T = float is the parity oracle (same fp32
// operation order as the reference + GLM 0.9.9.8, built with -ffp-contract=off), T = double is
// the "truth" twin used for error budgeting.
//
// PINNING: the reference ships no tests or golden vectors for this path.  This
// restatement is pinned instead against the reference ITSELF run in the build container:
// oracle/_ref/libwt_ref_{proj,vert,frag}.so, compiled by oracle/Makefile from the reference's own
// camera.cpp, render-engine.h, GLM and the water-grid.vert / water-grid.frag shader TEXTS (as C++), see
// tests/test_oracle_vs_ref.py (bit-exact) and the committed vectors under tests/golden/.
//
// Each function cites the reference lines it follows (paths relative to /root/reference).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "gl_sampler.h"

namespace {

// ----------------------------------------------------------------------------------------------
// Minimal vector/matrix algebra.  The ORDER of floating-point operations restates GLM 0.9.9.8
// (deps/glm/glm, the reference's only math dependency) so that T=float reproduces it bit for bit.
// ----------------------------------------------------------------------------------------------
template <typename T> struct V3 { T x, y, z; };
template <typename T> struct V4 { T x, y, z, w; };
template <typename T> struct M4 { V4<T> c[4]; };   // column-major, c[j] = column j (glm::mat4[j])

template <typename T> V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> V3<T> operator-(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> V3<T> operator*(V3<T> a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename T> V3<T> operator*(T s, V3<T> a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T> V4<T> operator+(V4<T> a, V4<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
template <typename T> V4<T> operator-(V4<T> a, V4<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
template <typename T> V4<T> operator*(V4<T> a, V4<T> b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
template <typename T> V4<T> operator*(V4<T> a, T s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
template <typename T> V4<T> operator/(V4<T> a, T s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
template <typename T> V3<T> xyz(V4<T> a) { return {a.x, a.y, a.z}; }

// glm/detail/func_geometric.inl:48-55 — tmp = a*b; tmp.x + tmp.y + tmp.z
template <typename T> T dot(V3<T> a, V3<T> b) { T tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
// glm/detail/func_geometric.inl:68-79
template <typename T> V3<T> cross(V3<T> x, V3<T> y) {
    return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
// glm/detail/func_geometric.inl:82-90 with inversesqrt = 1/sqrt (func_exponential.inl:44-50)
template <typename T> V3<T> normalize(V3<T> v) { T inv = (T)1 / std::sqrt(dot(v, v)); return v * inv; }
// glm/detail/func_geometric.inl:104-110 — I - N * dot(N, I) * 2
template <typename T> V3<T> reflect(V3<T> I, V3<T> N) { T d = dot(N, I); return I - (N * d) * (T)2; }
// glm/detail/func_common.inl:526-538 — x * (1 - a) + y * a
template <typename T> T mix(T x, T y, T a) { return x * ((T)1 - a) + y * a; }
template <typename T> V3<T> mix(V3<T> x, V3<T> y, T a) { return {mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)}; }
template <typename T> V4<T> mix(V4<T> x, V4<T> y, T a) { return {mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a), mix(x.w, y.w, a)}; }
template <typename T> T clampT(T x, T lo, T hi) { return std::min(std::max(x, lo), hi); }
template <typename T> T radians(T d) { return d * (T)0.01745329251994329576923690768489L; }
template <typename T> T degrees(T r) { return r * (T)57.295779513082320876798154814105L; }
template <typename T> T gmod(T x, T y) { return x - y * std::floor(x / y); }   // glm::mod

// glm/detail/type_mat4x4.inl:536-572 — (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
template <typename T> V4<T> mul(const M4<T>& m, V4<T> v) {
    V4<T> a0 = m.c[0] * v.x + m.c[1] * v.y;
    V4<T> a1 = m.c[2] * v.z + m.c[3] * v.w;
    return a0 + a1;
}
// glm/detail/type_mat4x4.inl:630-648 — ((A0*b0 + A1*b1) + A2*b2) + A3*b3 per result column
template <typename T> M4<T> mul(const M4<T>& a, const M4<T>& b) {
    M4<T> r;
    for (int j = 0; j < 4; ++j) r.c[j] = a.c[0] * b.c[j].x + a.c[1] * b.c[j].y + a.c[2] * b.c[j].z + a.c[3] * b.c[j].w;
    return r;
}
template <typename T> T& at(M4<T>& m, int col, int row) { return (&m.c[col].x)[row]; }
template <typename T> T at(const M4<T>& m, int col, int row) { return (&m.c[col].x)[row]; }

// glm/detail/func_matrix.inl:294-352 — cofactor inverse, 2x2 sub-determinants shared between columns
template <typename T> M4<T> inverse(const M4<T>& m) {
    auto e = [&](int c, int r) { return at(m, c, r); };
    // sub[k][j]: 2x2 minors built from rows (ra, rb) of columns (2,3), (1,3), (1,2)
    auto minor3 = [&](int ra, int rb, T out[3]) {
        out[0] = e(2, ra) * e(3, rb) - e(3, ra) * e(2, rb);
        out[1] = e(1, ra) * e(3, rb) - e(3, ra) * e(1, rb);
        out[2] = e(1, ra) * e(2, rb) - e(2, ra) * e(1, rb);
    };
    T c0[3], c1[3], c2[3], c3[3], c4[3], c5[3];
    minor3(2, 3, c0);   // Coef00,02,03
    minor3(1, 3, c1);   // Coef04,06,07
    minor3(1, 2, c2);   // Coef08,10,11
    minor3(0, 3, c3);   // Coef12,14,15
    minor3(0, 2, c4);   // Coef16,18,19
    minor3(0, 1, c5);   // Coef20,22,23
    auto fac = [](const T c[3]) { return V4<T>{c[0], c[0], c[1], c[2]}; };
    V4<T> F0 = fac(c0), F1 = fac(c1), F2 = fac(c2), F3 = fac(c3), F4 = fac(c4), F5 = fac(c5);
    V4<T> v0{e(1, 0), e(0, 0), e(0, 0), e(0, 0)};
    V4<T> v1{e(1, 1), e(0, 1), e(0, 1), e(0, 1)};
    V4<T> v2{e(1, 2), e(0, 2), e(0, 2), e(0, 2)};
    V4<T> v3{e(1, 3), e(0, 3), e(0, 3), e(0, 3)};
    V4<T> i0 = v1 * F0 - v2 * F1 + v3 * F2;
    V4<T> i1 = v0 * F0 - v2 * F3 + v3 * F4;
    V4<T> i2 = v0 * F1 - v1 * F3 + v3 * F5;
    V4<T> i3 = v0 * F2 - v1 * F4 + v2 * F5;
    V4<T> sa{(T)1, (T)-1, (T)1, (T)-1}, sb{(T)-1, (T)1, (T)-1, (T)1};
    M4<T> inv{{i0 * sa, i1 * sb, i2 * sa, i3 * sb}};
    V4<T> row0{inv.c[0].x, inv.c[1].x, inv.c[2].x, inv.c[3].x};
    V4<T> d0 = m.c[0] * row0;
    T det = (d0.x + d0.y) + (d0.z + d0.w);
    T ood = (T)1 / det;
    for (int j = 0; j < 4; ++j) inv.c[j] = inv.c[j] * ood;
    return inv;
}

// glm/ext/matrix_clip_space.inl:249-262 (perspectiveRH_NO; glm::perspective's default)
template <typename T> M4<T> perspective(T fovy, T aspect, T zn, T zf) {
    T th = std::tan(fovy / (T)2);
    M4<T> r;
    std::memset(&r, 0, sizeof(r));
    at(r, 0, 0) = (T)1 / (aspect * th);
    at(r, 1, 1) = (T)1 / th;
    at(r, 2, 2) = -(zf + zn) / (zf - zn);
    at(r, 2, 3) = -(T)1;
    at(r, 3, 2) = -((T)2 * zf * zn) / (zf - zn);
    return r;
}
// glm/ext/matrix_transform.inl:99-118 (lookAtRH)
template <typename T> M4<T> look_at(V3<T> eye, V3<T> center, V3<T> up) {
    V3<T> f = normalize(center - eye);
    V3<T> s = normalize(cross(f, up));
    V3<T> u = cross(s, f);
    M4<T> r;
    std::memset(&r, 0, sizeof(r));
    at(r, 0, 0) = s.x; at(r, 1, 0) = s.y; at(r, 2, 0) = s.z;
    at(r, 0, 1) = u.x; at(r, 1, 1) = u.y; at(r, 2, 1) = u.z;
    at(r, 0, 2) = -f.x; at(r, 1, 2) = -f.y; at(r, 2, 2) = -f.z;
    at(r, 3, 0) = -dot(s, eye); at(r, 3, 1) = -dot(u, eye); at(r, 3, 2) = dot(f, eye);
    at(r, 3, 3) = (T)1;
    return r;
}

// ----------------------------------------------------------------------------------------------
// Camera — src/camera.cpp:47-146
// ----------------------------------------------------------------------------------------------
template <typename T> struct Camera {
    T fov, aspect, zn, zf, yaw, pitch;
    V3<T> pos, forward, right, up, up_init;
    M4<T> view, proj;
    // camera.cpp:47-54 (ctor: yaw/pitch stored raw)
    Camera(T fov_, T aspect_, T zn_, T zf_, V3<T> pos_, T yaw_, T pitch_)
        : fov(fov_), aspect(aspect_), zn(zn_), zf(zf_), yaw(yaw_), pitch(pitch_), pos(pos_) {
        up_init = {(T)0, (T)1, (T)0};
        update_basis();
        proj = perspective(radians(fov), aspect, zn, zf);   // camera.cpp:140-142
    }
    void update_basis() {   // camera.cpp:130-138
        T cy = std::cos(radians(yaw)), sy = std::sin(radians(yaw));
        T cp = std::cos(radians(pitch)), sp = std::sin(radians(pitch));
        forward = normalize(V3<T>{cy * cp, sp, sy * cp});
        right = normalize(cross(forward, up_init));
        up = normalize(cross(right, forward));
        update_view();
    }
    void update_view() { view = look_at(pos, pos + forward, up); }   // camera.cpp:144-146
    void set_rotation(T yaw_deg, T pitch_deg) {                        // camera.cpp:93-98
        yaw = gmod(yaw_deg, (T)360);
        pitch = clampT(pitch_deg, (T)-89, (T)89);
        update_basis();
    }
    void translate(V3<T> d) { pos = pos + d; update_view(); }         // camera.cpp:100-104
};

// ----------------------------------------------------------------------------------------------
// utils::linePlaneIntersection — src/render-engine.h:111-124 (infinite line; false iff n.dir == 0)
// ----------------------------------------------------------------------------------------------
template <typename T> bool line_plane(V3<T>& out, V3<T> p0, V3<T> p1, V4<T> plane /*a,b,c,d*/) {
    V3<T> n{plane.x, plane.y, plane.z};
    V3<T> dir = normalize(p1 - p0);
    T denom = dot(n, dir);
    if ((T)0 == denom) return false;
    T t = (plane.w - dot(n, p0)) / denom;
    out = p0 + t * dir;
    return true;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// C interface (ctypes).  f32 = oracle, f64 = truth twin.  Layouts are shared with tests/.
// ----------------------------------------------------------------------------------------------
extern "C" {

struct wto_camera { float fov_deg, aspect, z_near, z_far, pos[3], yaw_deg, pitch_deg; };

struct wto_grid_f32 {
    int has_grid, n_intersections;
    float corners[4][4];   // [0]=bottomLeft [1]=topLeft [2]=bottomRight [3]=topRight
    float view[16], proj[16], view_proj[16], view_only_yaw[16];
    float cam_forward[3], aimpoint_1[3], aimpoint_2[3], aimpoint_final[3];
    float projector_pitch_deg, projector_y;
    float range[4];
};
struct wto_grid_f64 {
    int has_grid, n_intersections;
    double corners[4][4];
    double view[16], proj[16], view_proj[16], view_only_yaw[16];
    double cam_forward[3], aimpoint_1[3], aimpoint_2[3], aimpoint_final[3];
    double projector_pitch_deg, projector_y;
    double range[4];
};

}  // extern "C"

namespace {

template <typename T, typename OUT>
int project_grid(const wto_camera& ci, T D, OUT* o) {
    std::memset(o, 0, sizeof(*o));
    auto put16 = [](auto* dst, const M4<T>& m) { for (int j = 0; j < 4; ++j) for (int r = 0; r < 4; ++r) dst[4 * j + r] = at(m, j, r); };
    auto put3 = [](auto* dst, V3<T> v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };

    Camera<T> cam((T)ci.fov_deg, (T)ci.aspect, (T)ci.z_near, (T)ci.z_far, V3<T>{(T)ci.pos[0], (T)ci.pos[1], (T)ci.pos[2]},
                  (T)ci.yaw_deg, (T)ci.pitch_deg);
    // src/render-engine.cpp:278-286
    Camera<T> yaw_only = cam;
    yaw_only.set_rotation(yaw_only.yaw, (T)0);
    M4<T> vp = mul(cam.proj, cam.view);
    M4<T> ivp = inverse(vp);
    put16(o->view, cam.view); put16(o->proj, cam.proj); put16(o->view_proj, vp); put16(o->view_only_yaw, yaw_only.view);
    put3(o->cam_forward, cam.forward);

    // :807-809 planes y = +D, 0, -D
    V4<T> upper{(T)0, (T)1, (T)0, D}, base{(T)0, (T)1, (T)0, (T)0}, lower{(T)0, (T)1, (T)0, -D};

    // :815-838 padded NDC cube -> world
    V4<T> fc[8];
    for (int i = 0; i < 8; ++i) {
        V4<T> ndc{(i & 4) ? (T)1 : (T)-1, (i & 2) ? (T)1 : (T)-1, (i & 1) ? (T)1 : (T)-1, (T)1};
        ndc.x *= (T)1.2f;   // SAFETY_PADDING_SCALAR is the float literal 1.2f (:828)
        ndc.y *= (T)1.2f;
        V4<T> w = mul(ivp, ndc);
        fc[i] = w / w.w;
    }
    // :842-853
    static const int E[12][2] = {{0, 1}, {0, 2}, {0, 4}, {1, 3}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 6}, {5, 7}, {6, 7}};
    std::vector<V4<T>> pts;
    for (auto& e : E) {   // :860-883
        V3<T> p0 = xyz(fc[e[0]]), p1 = xyz(fc[e[1]]);
        T lo = std::min(p0.y, p1.y), hi = std::max(p0.y, p1.y);
        V3<T> hit;
        if (lo <= upper.w && upper.w <= hi && line_plane(hit, p0, p1, upper)) pts.push_back({hit.x, hit.y, hit.z, (T)1});
        if (lo <= lower.w && lower.w <= hi && line_plane(hit, p0, p1, lower)) pts.push_back({hit.x, hit.y, hit.z, (T)1});
    }
    for (auto& c : fc)    // :887-891
        if (lower.w <= c.y && c.y <= upper.w) pts.push_back(c);
    o->n_intersections = (int)pts.size();
    if (pts.empty()) return 1;   // :894
    o->has_grid = 1;

    // :905-917
    Camera<T> projector = cam;
    T cam_y = cam.pos.y;
    bool underwater = cam_y < (T)0;
    T min_y = D + (T)1;
    if (cam_y < min_y) {
        if (underwater) projector.translate({(T)0, min_y - (T)2 * cam_y, (T)0});
        else projector.translate({(T)0, min_y - cam_y, (T)0});
    }
    // :920-926
    V3<T> fs = cam.forward;
    const T EPS = (T)0.001f;
    if (std::fabs(fs.y) < EPS) fs.y = (fs.y >= (T)0 ? (T)1 : (T)-1) * EPS;
    // :929-939
    V3<T> aim1{(T)0, (T)0, (T)0};
    bool looking_down = fs.y < (T)0;
    if (looking_down != underwater) {
        line_plane(aim1, cam.pos, cam.pos + fs, base);
    } else {
        V3<T> mirrored = reflect(fs, V3<T>{(T)0, (T)1, (T)0});
        line_plane(aim1, cam.pos, cam.pos + mirrored, base);
    }
    // :943-954
    V3<T> aim2 = cam.pos + (T)1 * cam.forward;
    aim2.y = (T)0;
    T a = std::fabs(fs.y);
    V3<T> aim = mix(aim2, aim1, a);
    put3(o->aimpoint_1, aim1); put3(o->aimpoint_2, aim2); put3(o->aimpoint_final, aim);
    // :957-963
    V3<T> pf = normalize(aim - projector.pos);
    V3<T> pf_xz = normalize(V3<T>{pf.x, (T)0, pf.z});
    T pitch_deg = -degrees(std::acos(dot(pf, pf_xz)));
    projector.set_rotation(projector.yaw, pitch_deg);
    o->projector_pitch_deg = projector.pitch;
    o->projector_y = projector.pos.y;
    // :967-979
    M4<T> pvp = mul(projector.proj, projector.view);
    for (auto& p : pts) {
        p.y = (T)0;
        V4<T> c = mul(pvp, p);
        p = c / c.w;
    }
    // :982-992 (else-if scan, as written in the reference)
    T x_min = pts[0].x, x_max = pts[0].x, y_min = pts[0].y, y_max = pts[0].y;
    for (size_t i = 1; i < pts.size(); ++i) {
        if (pts[i].x < x_min) x_min = pts[i].x; else if (pts[i].x > x_max) x_max = pts[i].x;
        if (pts[i].y < y_min) y_min = pts[i].y; else if (pts[i].y > y_max) y_max = pts[i].y;
    }
    o->range[0] = x_min; o->range[1] = x_max; o->range[2] = y_min; o->range[3] = y_max;
    // :1002-1008
    M4<T> range;
    std::memset(&range, 0, sizeof(range));
    at(range, 0, 0) = x_max - x_min; at(range, 1, 1) = y_max - y_min; at(range, 2, 2) = (T)1;
    at(range, 3, 0) = x_min; at(range, 3, 1) = y_min; at(range, 3, 3) = (T)1;
    M4<T> pm = mul(inverse(mul(cam.proj, projector.view)), range);   // camera projection × projector view
    // :1012-1036
    static const T UV[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
    for (int i = 0; i < 4; ++i) {
        V4<T> p0 = mul(pm, V4<T>{UV[i][0], UV[i][1], (T)-1, (T)1});
        p0 = p0 / p0.w;
        V4<T> p1 = mul(pm, V4<T>{UV[i][0], UV[i][1], (T)1, (T)1});
        p1 = p1 / p1.w;
        V3<T> hit{(T)0, (T)0, (T)0};
        if (!line_plane(hit, xyz(p0), xyz(p1), base)) return -2;
        o->corners[i][0] = hit.x; o->corners[i][1] = hit.y; o->corners[i][2] = hit.z; o->corners[i][3] = (T)1;
    }
    return 0;
}

}  // namespace

extern "C" {

// Uniform block of the water program — the subset water-grid.vert:37-64 and the in-scope lines of
// water-grid.frag consume; upload sites src/render-engine.cpp:1057-1110.  All fp32, as uploaded.
struct wto_uniforms {
    float bottomLeft[4], bottomRight[4], topLeft[4], topRight[4];
    float cameraPosition[3];
    unsigned gerstnerWaveCount;
    unsigned gridLength;
    float heightmapDisplacementScale, heightmapSampleScale;
    float verticalBounceWaveDisplacement;
    float viewMatOnlyYaw[16];
    float viewProjection[16];
    float waveAnimationTimeInSeconds;
    float sunPosition[3];
    float sunShininess, sunStrength;
};
struct wto_wave { float A, w, phi, Q_i, Dx, Dz; };   // per-wave uniforms, render-engine.cpp:1076-1080

struct wto_textures {
    const uint8_t* heightmap_rgba; int hm_w, hm_h;   // GL row order
    const uint8_t* sky_faces[6]; int sky_n;           // may be all-NULL when shading is not requested
};

// Uniforms of the screen-space layers of water-grid.frag (:34-55) and its fog block; upload sites
// src/render-engine.cpp:1060-1063,1087-1088,1096,1100,1107,1109-1110.
struct wto_compose_uniforms {
    float fogColourFarAtCurrentTime[4];
    float fogDepthRadiusFar, fogDepthRadiusNear;
    float softEdgesDeltaDepthThreshold, tintDeltaDepthThreshold, waterClarity;
    float zNear, zFar;
    float viewportWidthHeight[2];
    float cameraPosition[3];
    float sunPosition[3];
};
// Inputs of the skybox synthesis (render-engine.cpp:306-442): the RenderEngine fields and per-frame values the
// three sky programs and the fog quad receive, and the three source textures.
struct wto_sky_uniforms {
    float sunPosition[3];            // render-engine.cpp:289-291
    float sunHorizonDarkness, sunShininess, sunStrength;      // :368-371
    float oneMinusCloudProportion, overcastStrength;          // :386-387, :297
    float fogColourFarAtCurrentTime[4];                       // :295, :424
    int draw_stars, draw_skysphere, draw_clouds;              // the m_isVisible / nullptr tests of :333,:352,:381
};
struct wto_sky_textures {
    const uint8_t* gradient_rgba; int gradient_len;           // 1-D sky gradient (program.cpp:644)
    const uint8_t* stars[6]; int stars_n;
    const uint8_t* clouds[6]; int clouds_n;
};

struct wto_screen_textures {
    const float* depth;            // w x h, GL row order, window-space depth in [0,1]
    const uint8_t* reflections;    // w x h RGBA8
    const uint8_t* refractions;    // w x h RGBA8
    int w, h;
};

}  // extern "C"

namespace {

int g_time_offset_fp32 = 0;   // wto_set_time_offset_fp32: see Surface::gerstner (only changes T = double)

template <typename T> struct Surface {
    const wto_uniforms& u;
    const wto_wave* waves;
    wt_gl::Image2D hm;
    V4<T> BL, BR, TL, TR;
    T t, bounce, hm_disp, hm_scale;

    Surface(const wto_uniforms& u_, const wto_wave* w_, const wto_textures& tx) : u(u_), waves(w_) {
        hm = {tx.heightmap_rgba, tx.hm_w, tx.hm_h};
        auto v4 = [](const float* p) { return V4<T>{(T)p[0], (T)p[1], (T)p[2], (T)p[3]}; };
        BL = v4(u.bottomLeft); BR = v4(u.bottomRight); TL = v4(u.topLeft); TR = v4(u.topRight);
        t = (T)u.waveAnimationTimeInSeconds; bounce = (T)u.verticalBounceWaveDisplacement;
        hm_disp = (T)u.heightmapDisplacementScale; hm_scale = (T)u.heightmapSampleScale;
    }
    // water-grid.vert:87-91
    V4<T> grid(T uu, T vv) const { return mix(mix(BL, BR, uu), mix(TL, TR, uu), vv); }
    // water-grid.vert:74-85 — argument is the UNdisplaced (x,z); sin for height, cos for horizontal
    V3<T> gerstner(T x, T z) const {
        V3<T> P{x, (T)0, z};
        for (unsigned i = 0; i < u.gerstnerWaveCount; ++i) {
            T A = (T)waves[i].A, w = (T)waves[i].w, phi = (T)waves[i].phi, Q = (T)waves[i].Q_i;
            T Dx = (T)waves[i].Dx, Dz = (T)waves[i].Dz;
            T dxz = Dx * x + Dz * z;                 // dot(vec2, vec2): tmp.x + tmp.y
            // T = double, "clean phase" variant: the time offset as the shader's fp32 product fl(phi t), everything else
            // exact — the phase a smooth (recurrence) evaluation aims at; used to budget its distance from the reference
            T pt = g_time_offset_fp32 ? (T)((float)waves[i].phi * (float)u.waveAnimationTimeInSeconds) : phi * t;
            T k = w * dxz + pt;
            T ck = std::cos(k), sk = std::sin(k);
            if (g_time_offset_fp32 == 2) {
                // "smooth phase" variant (meaningful for T = float): the phase and its sin / cos evaluated in double from the
                // same fp32 inputs (time offset = the shader's fl(phi t)), everything else — products, the running fp32 sum in
                // wave order — exactly as the shader does it.  Models an evaluation that does not reproduce the reference's
                // rounding of k (the recurrence kernel); its distance from T = float is the price of that.
                const double kd = (double)waves[i].w * ((double)waves[i].Dx * (double)x + (double)waves[i].Dz * (double)z) +
                                  (double)((float)waves[i].phi * (float)u.waveAnimationTimeInSeconds);
                ck = (T)std::cos(kd); sk = (T)std::sin(kd);
            }
            T c = Q * ck;
            V3<T> d{Dx * c, sk, Dz * c};
            P = P + A * d;
        }
        return P;
    }
    // water-grid.vert:93-104 (sample at the DISPLACED position, t-coordinate uses -z)
    void sample_hm(V3<T> P, T texel[4]) const {
        wt_gl::texture2d_mirrored_linear<T>(hm, hm_scale * P.x, hm_scale * (-P.z), texel);
    }
    // water-grid.vert:106-111
    T hm_displacement(const T texel[4]) const { return hm_disp * ((T)2 * (texel[0] - (T)0.5)); }
    // the position pipeline of water-grid.vert:135-147 (also used for the 4 neighbours, :164-182)
    V3<T> displaced(T uu, T vv, T texel_out[4]) const {
        V4<T> g = grid(uu, vv);
        V3<T> P = gerstner(g.x, g.z);
        T tx[4];
        sample_hm(P, tx);
        P.y += hm_displacement(tx);
        P.y += bounce;
        if (texel_out) std::memcpy(texel_out, tx, sizeof(tx));
        return P;
    }
};

// water-grid.frag:71-73,95-111 evaluated at a vertex with the un-interpolated varyings
// (viewVecRaw, normal).  shade4 = (sky.rgb, sun specular), *fresnel = fresnel_f_theta.
template <typename T>
void shade_vertex(const wto_uniforms& u, const wt_gl::CubeRGBA8* sky, V3<T> view_raw, V3<T> n, T* shade4, T* fresnel) {
    T len = std::sqrt(dot(view_raw, view_raw));                         // length(), frag:71
    V3<T> vv3{view_raw.x / len, view_raw.y / len, view_raw.z / len};     // frag:73 (vec3 / float)
    V3<T> R = reflect(V3<T>{-vv3.x, -vv3.y, -vv3.z}, n);                // frag:95
    T sky_c[4] = {0, 0, 0, 0};
    if (sky && sky->face[0]) wt_gl::texture_cube_clamp_linear<T>(*sky, R.x, R.y, R.z, sky_c);   // frag:96
    V3<T> sun{(T)u.sunPosition[0], (T)u.sunPosition[1], (T)u.sunPosition[2]};
    T sd = clampT(dot(R, sun), (T)0, (T)1);
    T sun_term = (T)u.sunStrength * (T)1 * std::pow(sd, (T)u.sunShininess);   // frag:99-100
    T cos_theta = dot(n, vv3);                                          // frag:110
    *fresnel = (T)0.02f + ((T)1 - (T)0.02f) * std::pow((T)1 - cos_theta, (T)5);   // frag:107-111
    shade4[0] = sky_c[0]; shade4[1] = sky_c[1]; shade4[2] = sky_c[2]; shade4[3] = sun_term;
}

// ----------------------------------------------------------------------------------------------
// Skybox synthesis: the colour each texel of the 6 x n x n RGBA8 skybox cubemap receives from
// render-engine.cpp:306-442 — clear, then stars, sky sphere, clouds and the fog quad blended in this order with
// glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA) (:303) into an RGBA8 colour attachment.
// ----------------------------------------------------------------------------------------------
// skysphere.frag:48-76
template <typename T>
void skysphere_colour(const wto_sky_uniforms& u, const wto_sky_textures& tx, V3<T> normalVec, T out[4]) {
    auto mixT = [](T x, T y, T a) { return x * ((T)1 - a) + y * a; };
    V3<T> L = normalize(V3<T>{(T)u.sunPosition[0], (T)u.sunPosition[1], (T)u.sunPosition[2]});
    V3<T> N = normalize(normalVec);
    T a1 = (T)0.25f * (L.y + (T)1), a2 = (T)1 - a1;
    T peak[4], hor[4];
    wt_gl::texture1d_clamp_linear<T>(tx.gradient_rgba, tx.gradient_len, a1, peak);
    wt_gl::texture1d_clamp_linear<T>(tx.gradient_rgba, tx.gradient_len, a2, hor);
    T a3 = dot(N, L);
    T spec = a3 > (T)0 ? (T)u.sunStrength * (T)1 * std::pow(a3, (T)u.sunShininess) : (T)0;
    for (int c = 0; c < 3; ++c) {
        T horizon = mixT(hor[c], (T)0, (T)u.sunHorizonDarkness);
        T diffuse = a3 >= (T)0 ? mixT(horizon, peak[c], a3) : mixT(horizon, (T)0, std::fabs(a3) * ((T)1 - std::fabs(L.y)));
        out[c] = diffuse + spec;
    }
    const T THRESHOLD = (T)0.1f;
    T alpha1 = L.y < THRESHOLD ? (L.y + (T)1) / (THRESHOLD + (T)1) : (T)1;
    T alpha2 = (a3 + (T)1) / (T)2;
    out[3] = std::max(alpha1, alpha2);
}
// skybox-clouds.frag:46-67
template <typename T>
void clouds_colour(const wto_sky_uniforms& u, const wto_sky_textures& tx, V3<T> str, T out[4]) {
    auto mixT = [](T x, T y, T a) { return x * ((T)1 - a) + y * a; };
    wt_gl::CubeRGBA8 cube{{tx.clouds[0], tx.clouds[1], tx.clouds[2], tx.clouds[3], tx.clouds[4], tx.clouds[5]}, tx.clouds_n};
    T raw[4];
    wt_gl::texture_cube_clamp_linear<T>(cube, str.x, str.y, str.z, raw);
    T lin = (T)0.2126f * raw[0] + (T)0.7152f * raw[1] + (T)0.0722f * raw[2];
    T srgb = lin <= (T)0.0031308f ? (T)12.92f * lin : (T)1.055f * std::pow(lin, (T)1 / (T)2.4f) - (T)0.055f;
    T intensity = clampT((T)u.sunPosition[1], (T)0.2f, (T)1);
    T omp = (T)u.oneMinusCloudProportion;
    T col[4] = {intensity, intensity, intensity, srgb > omp ? (srgb - omp) / ((T)1 - omp) : (T)0};
    T a = clampT((T)u.sunPosition[1], (T)0, (T)1) * (T)u.overcastStrength;
    if (srgb > omp) {
        for (int c = 0; c < 3; ++c) out[c] = mixT(col[c], srgb, a);
        out[3] = mixT(col[3], (T)1, a);
    } else {
        out[0] = out[1] = out[2] = (T)0.5f;
        out[3] = a;
    }
}
// One layer blended into an RGBA8 destination: GL 4.1 section 4.1.7 — source, destination and factors clamped to [0,1]
// (fixed-point colour buffer), Cd' = Cs * As + Cd * (1 - As) on all four channels, result converted to unorm8
// (section 2.1.5: round to nearest).
template <typename T>
void blend_rgba8(uint8_t dst[4], const T src_in[4]) {
    T src[4];
    for (int c = 0; c < 4; ++c) src[c] = clampT(src_in[c], (T)0, (T)1);
    for (int c = 0; c < 4; ++c) {
        T d = (T)dst[c] / (T)255;
        T v = clampT(src[c] * src[3] + d * ((T)1 - src[3]), (T)0, (T)1);
        dst[c] = (uint8_t)std::floor(v * (T)255 + (T)0.5);
    }
}
// The direction the rasteriser hands the fragment at the centre of texel (i, j) of face f: the faces are drawn with a
// 90-degree frustum and the look-at bases of render-engine.h:198-203, which make the framebuffer pixel (i, j) the
// cube-map texel (s, t) = ((i + 1/2)/n, (j + 1/2)/n) of that face (GL 4.1 cube-map face table, inverted).
template <typename T>
V3<T> sky_texel_direction(int face, int i, int j, int n) {
    T sc = ((T)2 * ((T)i + (T)0.5)) / (T)n - (T)1, tc = ((T)2 * ((T)j + (T)0.5)) / (T)n - (T)1;
    switch (face) {
        case 0: return V3<T>{(T)1, -tc, -sc};
        case 1: return V3<T>{(T)-1, -tc, sc};
        case 2: return V3<T>{sc, (T)1, tc};
        case 3: return V3<T>{sc, (T)-1, -tc};
        case 4: return V3<T>{sc, -tc, (T)1};
        default: return V3<T>{-sc, -tc, (T)-1};
    }
}
template <typename T>
void sky_texel(const wto_sky_uniforms& u, const wto_sky_textures& tx, int face, int i, int j, int n, uint8_t out[4]) {
    V3<T> d = sky_texel_direction<T>(face, i, j, n);
    out[0] = out[1] = out[2] = out[3] = 255;                      // glClearColor(1,1,1,1) (render-engine.cpp:694), glClear :330
    T src[4];
    if (u.draw_stars) {                                            // skybox-stars.frag:41
        wt_gl::CubeRGBA8 cube{{tx.stars[0], tx.stars[1], tx.stars[2], tx.stars[3], tx.stars[4], tx.stars[5]}, tx.stars_n};
        wt_gl::texture_cube_clamp_linear<T>(cube, d.x, d.y, d.z, src);
        src[3] = (T)1;
        blend_rgba8<T>(out, src);
    }
    if (u.draw_skysphere) { skysphere_colour<T>(u, tx, d, src); blend_rgba8<T>(out, src); }
    if (u.draw_clouds) { clouds_colour<T>(u, tx, d, src); blend_rgba8<T>(out, src); }
    for (int c = 0; c < 4; ++c) src[c] = (T)u.fogColourFarAtCurrentTime[c];   // screen-space-quad.frag:45, solidColour
    blend_rgba8<T>(out, src);
}

// water-grid.frag:66-69,76-89,113-162 evaluated at a vertex: everything of main() that is not the three terms
// shade_vertex covers.  The fragment inputs a rasteriser would interpolate are taken un-interpolated at the
// vertex: gl_FragCoord from gl_Position through the viewport transform (GL 4.1 §2.14.1: xw = (xd + 1) w/2,
// zw = (zd + 1)/2 with the default depth range), the varyings as written by the vertex stage.
template <typename T>
void compose_vertex(const wto_compose_uniforms& u, const wto_screen_textures& st, V3<T> P, T fresnel_f_theta, const T shade4[4], V4<T> clip,
                    T ndc0x, T ndc0y, V3<T> yawn, T colour[4]) {
    auto mixT = [](T x, T y, T a) { return x * ((T)1 - a) + y * a; };
    auto linearize = [&](T depth) { return ((T)u.zNear * depth) / ((T)u.zFar - depth * ((T)u.zFar - (T)u.zNear)); };   // :66-69
    const T vw = (T)u.viewportWidthHeight[0], vh = (T)u.viewportWidthHeight[1];
    V3<T> cam{(T)u.cameraPosition[0], (T)u.cameraPosition[1], (T)u.cameraPosition[2]};
    V3<T> view_raw = cam - P;
    T len = std::sqrt(dot(view_raw, view_raw));                          // :71
    T vdc = clampT(len / (T)u.zFar, (T)0, (T)1);                         // :72
    // gl_FragCoord
    T ndx = clip.x / clip.w, ndy = clip.y / clip.w, ndz = clip.z / clip.w;
    T fx = (ndx + (T)1) * (vw / (T)2), fy = (ndy + (T)1) * (vh / (T)2), fz = (ndz + (T)1) / (T)2;
    T uvx = fx / vw, uvy = fy / vh;                                       // :79
    T h0x = (ndc0x + (T)1) / (T)2, h0y = (ndc0y + (T)1) / (T)2;           // :81
    T back = linearize(wt_gl::texture2d_clamp_nearest_r32f<T>(st.depth, st.w, st.h, uvx, uvy));   // :84
    T hack = back > (T)0.999f ? (T)0 : (T)1;                              // :86
    T frag = linearize(fz);                                               // :87
    T dclamped = clampT(back - frag, (T)0, (T)1);                         // :89-91
    // :118-133 local reflections / refractions
    T ds = (T)0.1f * ((T)1 - vdc);
    T rlx = clampT(mixT(uvx, h0x, ds) + ds * yawn.x, (T)0, (T)1), rly = clampT(mixT(uvy, h0y, ds) + ds * (-yawn.z), (T)0, (T)1);
    T rrx = clampT(mixT(uvx, h0x, ds) + ds * (-yawn.x), (T)0, (T)1), rry = clampT(mixT(uvy, h0y, ds) + ds * yawn.z, (T)0, (T)1);
    T refl[4], refr[4];
    wt_gl::texture2d_clamp_linear<T>(wt_gl::Image2D{st.reflections, st.w, st.h}, rlx, rly, refl);
    wt_gl::texture2d_clamp_linear<T>(wt_gl::Image2D{st.refractions, st.w, st.h}, rrx, rry, refr);
    // :136-141 tint
    const T DEEP[3] = {(T)0.0f, (T)0.341f, (T)0.482f}, SHALLOW[3] = {(T)0.3f, (T)0.941f, (T)0.903f};
    T tint_f = (T)u.tintDeltaDepthThreshold == (T)0 ? (T)0 : hack * ((T)1 - clampT(dclamped / (T)u.tintDeltaDepthThreshold, (T)0, (T)1));
    T sun_y = clampT((T)u.sunPosition[1], (T)0, (T)1);
    T out[3];
    for (int c = 0; c < 3; ++c) {
        T tint = sun_y * mixT(DEEP[c], SHALLOW[c], tint_f);
        T transmission = mixT(tint, refr[c], (T)u.waterClarity * refr[3] * ((T)1 - dclamped));   // :143
        T reflection = mixT(shade4[c] + shade4[3], refl[c], refl[3]);                             // :145 (SUN_BASE_COLOUR = 1)
        out[c] = mixT(transmission, reflection, fresnel_f_theta);                                  // :151
    }
    // :154-156 soft edges
    T hardness = (T)u.softEdgesDeltaDepthThreshold == (T)0 ? (T)1
                 : (T)1 - (hack * ((T)1 - clampT(dclamped / (T)u.softEdgesDeltaDepthThreshold, (T)0, (T)1)));
    // :160-162 fog
    T fa = (T)u.fogColourFarAtCurrentTime[3];
    T fog_a = vdc >= (T)u.fogDepthRadiusFar ? fa : vdc <= (T)u.fogDepthRadiusNear ? (T)0
              : fa * ((vdc - (T)u.fogDepthRadiusNear) / ((T)u.fogDepthRadiusFar - (T)u.fogDepthRadiusNear));
    for (int c = 0; c < 3; ++c) colour[c] = mixT(out[c], (T)u.fogColourFarAtCurrentTime[c], fog_a);
    colour[3] = hardness;
}

// Per-vertex outputs.  pos4 = (x,y,z,1) world position (water-grid.vert `position`),
// nrm4 = (normal.xyz, fresnel or 0), shade4 = (sky.rgb, sun), extra16 = clip(4) ndc0(2) viewVecRaw(3)
// normalVecInViewSpaceOnlyYaw(3) heightmap_colour(4).
template <typename T>
void eval_vertex(const Surface<T>& S, const M4<T>& VP, const M4<T>& VY, const wt_gl::CubeRGBA8* sky, long long col, long long row,
                 T* pos4, T* nrm4, T* shade4, T* extra16) {
    const wto_uniforms& u = S.u;
    // water-grid.vert:123-125.  Exact-integer (col,row) semantics: identical to the shader's float
    // mod()/integer division for gridLength <= 4096 and well-defined beyond.
    T du = (T)1 / (T)(u.gridLength - 1);
    T dv = du;
    T uu = (T)col * du, vv = (T)row * dv;

    T texel[4];
    V3<T> P = S.displaced(uu, vv, texel);                                   // :135-147
    V4<T> clip = mul(VP, V4<T>{P.x, P.y, P.z, (T)1});                        // :150
    V4<T> clip0 = mul(VP, V4<T>{P.x, (T)0, P.z, (T)1});                      // :153
    T ndc0x = clip0.x / clip0.w, ndc0y = clip0.y / clip0.w;                  // :154
    V3<T> cam{(T)u.cameraPosition[0], (T)u.cameraPosition[1], (T)u.cameraPosition[2]};
    V3<T> view_raw = cam - P;                                               // :158
    // :164-182 neighbours at uv -/+ (du,0), uv -/+ (0,dv), full pipeline each
    V3<T> Pmu = S.displaced(uu - du, vv, nullptr);
    V3<T> Ppu = S.displaced(uu + du, vv, nullptr);
    V3<T> Pmv = S.displaced(uu, vv - dv, nullptr);
    V3<T> Ppv = S.displaced(uu, vv + dv, nullptr);
    V3<T> n = normalize(cross(Ppu - Pmu, Ppv - Pmv));                       // :187
    V3<T> view_n = normalize(view_raw);
    if (dot(n, view_n) < (T)0) n = n * (T)-1;                               // :190
    V4<T> ny4 = mul(VY, V4<T>{n.x, n.y, n.z, (T)0});                         // :193
    V3<T> ny = normalize(xyz(ny4));

    pos4[0] = P.x; pos4[1] = P.y; pos4[2] = P.z; pos4[3] = (T)1;
    nrm4[0] = n.x; nrm4[1] = n.y; nrm4[2] = n.z; nrm4[3] = (T)0;
    if (extra16) {
        extra16[0] = clip.x; extra16[1] = clip.y; extra16[2] = clip.z; extra16[3] = clip.w;
        extra16[4] = ndc0x; extra16[5] = ndc0y;
        extra16[6] = view_raw.x; extra16[7] = view_raw.y; extra16[8] = view_raw.z;
        extra16[9] = ny.x; extra16[10] = ny.y; extra16[11] = ny.z;
        for (int c = 0; c < 4; ++c) extra16[12 + c] = texel[c];
    }
    if (shade4) shade_vertex<T>(u, sky, view_raw, n, shade4, nrm4 + 3);
}

template <typename T>
int eval_vertices(const wto_uniforms* u, const wto_wave* waves, const wto_textures* tx, const long long* ids, long long n,
                  long long row0, long long row1, long long col0, long long col1, T* pos4, T* nrm4, T* shade4, T* extra16) {
    if (u->gridLength < 2) return -1;
    Surface<T> S(*u, waves, *tx);
    M4<T> VP, VY;
    for (int j = 0; j < 4; ++j) for (int r = 0; r < 4; ++r) { at(VP, j, r) = (T)u->viewProjection[4 * j + r]; at(VY, j, r) = (T)u->viewMatOnlyYaw[4 * j + r]; }
    wt_gl::CubeRGBA8 cube{{tx->sky_faces[0], tx->sky_faces[1], tx->sky_faces[2], tx->sky_faces[3], tx->sky_faces[4], tx->sky_faces[5]}, tx->sky_n};
    const long long L = u->gridLength;
    const long long ncol = col1 - col0;
    const long long total = ids ? n : (row1 - row0) * ncol;
#pragma omp parallel for schedule(dynamic, 256)
    for (long long k = 0; k < total; ++k) {
        long long col, row;
        if (ids) { col = ids[k] % L; row = ids[k] / L; }
        else { row = row0 + k / ncol; col = col0 + k % ncol; }
        eval_vertex<T>(S, VP, VY, &cube, col, row, pos4 + 4 * k, nrm4 + 4 * k, shade4 ? shade4 + 4 * k : nullptr,
                       extra16 ? extra16 + 16 * k : nullptr);
    }
    return 0;
}

}  // namespace

extern "C" {

int wto_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void wto_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int wto_project_grid_f32(const wto_camera* cam, float displaceable_amplitude, wto_grid_f32* out) {
    return project_grid<float>(*cam, displaceable_amplitude, out);
}
int wto_project_grid_f64(const wto_camera* cam, double displaceable_amplitude, wto_grid_f64* out) {
    return project_grid<double>(*cam, displaceable_amplitude, out);
}

// host-side scalar rules of the boundary
float wto_total_amplitude(const float* A, unsigned n) {   // render-engine.h:156-157 (running fp32 sum)
    float s = 0.0f;
    for (unsigned i = 0; i < n; ++i) s += A[i];
    return s;
}
float wto_steepness_Q_i(float Q, float w, float A, unsigned count) {   // render-engine.cpp:1073
    return (w * A) != 0.0f ? Q / (w * A * count) : 0.0f;
}
void wto_sun_position(float hours, float out3[3]) {   // render-engine.cpp:289-291
    float days = hours / 24.0f;
    float theta = days * 6.283185307179586476925286766559f - 1.5707963267948966192313216916398f;
    out3[0] = std::cos(theta); out3[1] = std::sin(theta); out3[2] = 0.0f;
}
float wto_bounce_displacement(float amplitude, float phase) {   // render-engine.cpp:299-300
    float shift = phase * 6.283185307179586476925286766559f;
    return amplitude * std::sin(shift);
}

// ids != NULL: evaluate the n listed vertex ids (id = row * gridLength + col);
// ids == NULL: evaluate the region rows [row0,row1) x cols [col0,col1), row-major.
int wto_eval_f32(const wto_uniforms* u, const wto_wave* waves, const wto_textures* tx, const long long* ids, long long n,
                 long long row0, long long row1, long long col0, long long col1, float* pos4, float* nrm4, float* shade4, float* extra16) {
    return eval_vertices<float>(u, waves, tx, ids, n, row0, row1, col0, col1, pos4, nrm4, shade4, extra16);
}
void wto_set_time_offset_fp32(int on) { g_time_offset_fp32 = on; }
int wto_eval_f64(const wto_uniforms* u, const wto_wave* waves, const wto_textures* tx, const long long* ids, long long n,
                 long long row0, long long row1, long long col0, long long col1, double* pos4, double* nrm4, double* shade4, double* extra16) {
    return eval_vertices<double>(u, waves, tx, ids, n, row0, row1, col0, col1, pos4, nrm4, shade4, extra16);
}

// The in-scope fragment terms alone, evaluated at caller-supplied vertices: pos4/nrm4 are (n,4) arrays
// (e.g. a kernel's own outputs); writes shade4 (n,4) and fresnel (n).  Lets a test check the shading
// FUNCTION tightly, independent of how errors of the normal propagate through pow(., shininess).
// Skybox synthesis: faces_out = 6 * n * n RGBA8 texels, GL face order.
int wto_sky_synthesize(const wto_sky_uniforms* u, const wto_sky_textures* tx, int n, uint8_t* faces_out) {
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < 6LL * n * n; ++k) {
        const int face = (int)(k / ((long long)n * n)), j = (int)((k / n) % n), i = (int)(k % n);
        sky_texel<float>(*u, *tx, face, i, j, n, faces_out + 4 * k);
    }
    return 0;
}
// the layer functions alone, on caller-supplied directions (pinned against the compiled shader texts)
int wto_skysphere_colour_f32(const wto_sky_uniforms* u, const wto_sky_textures* tx, long long n, const float* dir3, float* colour4) {
    for (long long k = 0; k < n; ++k) skysphere_colour<float>(*u, *tx, V3<float>{dir3[3 * k], dir3[3 * k + 1], dir3[3 * k + 2]}, colour4 + 4 * k);
    return 0;
}
int wto_clouds_colour_f32(const wto_sky_uniforms* u, const wto_sky_textures* tx, long long n, const float* dir3, float* colour4) {
    for (long long k = 0; k < n; ++k) clouds_colour<float>(*u, *tx, V3<float>{dir3[3 * k], dir3[3 * k + 1], dir3[3 * k + 2]}, colour4 + 4 * k);
    return 0;
}

// The screen-space layers + fog of water-grid.frag at caller-supplied vertices (all (n,4) arrays: pos4, nrm4 =
// (normal, fresnel_f_theta), shade4 = (sky.rgb, sun), clip4 = gl_Position, vary_a = (ndc0.xy, ., .), vary_b = (yaw normal, .)).
int wto_compose_f32(const wto_compose_uniforms* u, const wto_screen_textures* st, long long n, const float* pos4, const float* nrm4,
                    const float* shade4, const float* clip4, const float* vary_a, const float* vary_b, float* colour4) {
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < n; ++k) {
        compose_vertex<float>(*u, *st, V3<float>{pos4[4 * k], pos4[4 * k + 1], pos4[4 * k + 2]}, nrm4[4 * k + 3], shade4 + 4 * k,
                              V4<float>{clip4[4 * k], clip4[4 * k + 1], clip4[4 * k + 2], clip4[4 * k + 3]}, vary_a[4 * k], vary_a[4 * k + 1],
                              V3<float>{vary_b[4 * k], vary_b[4 * k + 1], vary_b[4 * k + 2]}, colour4 + 4 * k);
    }
    return 0;
}

int wto_shade_f32(const wto_uniforms* u, const wto_textures* tx, long long n, const float* pos4, const float* nrm4, float* shade4, float* fresnel) {
    wt_gl::CubeRGBA8 cube{{tx->sky_faces[0], tx->sky_faces[1], tx->sky_faces[2], tx->sky_faces[3], tx->sky_faces[4], tx->sky_faces[5]}, tx->sky_n};
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < n; ++k) {
        V3<float> P{pos4[4 * k], pos4[4 * k + 1], pos4[4 * k + 2]};
        V3<float> cam{u->cameraPosition[0], u->cameraPosition[1], u->cameraPosition[2]};
        V3<float> nn{nrm4[4 * k], nrm4[4 * k + 1], nrm4[4 * k + 2]};
        shade_vertex<float>(*u, &cube, cam - P, nn, shade4 + 4 * k, fresnel + k);
    }
    return 0;
}

}  // extern "C"
