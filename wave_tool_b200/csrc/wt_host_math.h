// Host-side fp32 vector/matrix helpers for the CPU prologue of the water path.
//
// The reference does all host math with GLM 0.9.9.8 (deps/glm).  To hand the kernel the very same
// corner points and matrices the reference would upload, every helper here performs its fp32
// operations in the order GLM's scalar code does (cited per function); the translation unit is
// built with -ffp-contract=off so no operation is fused.
#pragma once
#include <cmath>

namespace wt {

struct float3h { float x, y, z; };
struct float4h { float x, y, z, w; };
struct mat4h { float m[4][4]; };   // m[col][row], column-major like glm::mat4

inline float3h add(float3h a, float3h b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3h sub(float3h a, float3h b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3h scale(float3h a, float s) { return {a.x * s, a.y * s, a.z * s}; }

// glm/detail/func_geometric.inl:48-55
inline float dot3(float3h a, float3h b) {
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return px + py + pz;
}
// glm/detail/func_geometric.inl:68-79
inline float3h cross3(float3h a, float3h b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
// glm/detail/func_geometric.inl:82-90; inversesqrt = 1 / sqrt (func_exponential.inl:44-50)
inline float3h unit(float3h v) { return scale(v, 1.0f / std::sqrt(dot3(v, v))); }
// glm/detail/func_common.inl:526-538
inline float lerp1(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float3h lerp3(float3h x, float3h y, float a) { return {lerp1(x.x, y.x, a), lerp1(x.y, y.y, a), lerp1(x.z, y.z, a)}; }
// glm/detail/func_geometric.inl:104-110
inline float3h reflect3(float3h i, float3h n) { float d = dot3(n, i); return sub(i, scale(scale(n, d), 2.0f)); }

inline float to_radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float to_degrees(float rad) { return rad * 57.295779513082320876798154814105f; }
inline float floor_mod(float x, float y) { return x - y * std::floor(x / y); }   // glm::mod
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }

// glm/detail/type_mat4x4.inl:536-572: (c0*v0 + c1*v1) + (c2*v2 + c3*v3)
inline float4h transform(const mat4h& a, float4h v) {
    float4h r;
    float* o = &r.x;
    for (int row = 0; row < 4; ++row) {
        float lo = a.m[0][row] * v.x + a.m[1][row] * v.y;
        float hi = a.m[2][row] * v.z + a.m[3][row] * v.w;
        o[row] = lo + hi;
    }
    return r;
}
// glm/detail/type_mat4x4.inl:630-648: ((A0*b0 + A1*b1) + A2*b2) + A3*b3
inline mat4h matmul(const mat4h& a, const mat4h& b) {
    mat4h r;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row)
            r.m[col][row] = ((a.m[0][row] * b.m[col][0] + a.m[1][row] * b.m[col][1]) + a.m[2][row] * b.m[col][2]) + a.m[3][row] * b.m[col][3];
    return r;
}
inline float4h divide_w(float4h v) { return {v.x / v.w, v.y / v.w, v.z / v.w, v.w / v.w}; }

// glm/detail/func_matrix.inl:294-352 (cofactor expansion sharing the eighteen 2x2 minors)
inline mat4h invert(const mat4h& a) {
    const float (*m)[4] = a.m;
    // minors of row pairs; index: [pair][k], k = column pair (2,3), (1,3), (1,2)
    const int pairs[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
    float s[6][3];
    for (int p = 0; p < 6; ++p) {
        int ra = pairs[p][0], rb = pairs[p][1];
        s[p][0] = m[2][ra] * m[3][rb] - m[3][ra] * m[2][rb];
        s[p][1] = m[1][ra] * m[3][rb] - m[3][ra] * m[1][rb];
        s[p][2] = m[1][ra] * m[2][rb] - m[2][ra] * m[1][rb];
    }
    // Fac_p = (s[p][0], s[p][0], s[p][1], s[p][2]);  Vec_r = (m[1][r], m[0][r], m[0][r], m[0][r])
    auto F = [&](int p, int lane) { return lane < 2 ? s[p][0] : s[p][lane - 1]; };
    auto V = [&](int r, int lane) { return lane == 0 ? m[1][r] : m[0][r]; };
    mat4h inv;
    for (int lane = 0; lane < 4; ++lane) {
        float i0 = V(1, lane) * F(0, lane) - V(2, lane) * F(1, lane) + V(3, lane) * F(2, lane);
        float i1 = V(0, lane) * F(0, lane) - V(2, lane) * F(3, lane) + V(3, lane) * F(4, lane);
        float i2 = V(0, lane) * F(1, lane) - V(1, lane) * F(3, lane) + V(3, lane) * F(5, lane);
        float i3 = V(0, lane) * F(2, lane) - V(1, lane) * F(4, lane) + V(2, lane) * F(5, lane);
        float sa = (lane & 1) ? -1.0f : 1.0f;
        inv.m[0][lane] = i0 * sa;
        inv.m[1][lane] = i1 * -sa;
        inv.m[2][lane] = i2 * sa;
        inv.m[3][lane] = i3 * -sa;
    }
    float d0 = m[0][0] * inv.m[0][0], d1 = m[0][1] * inv.m[1][0], d2 = m[0][2] * inv.m[2][0], d3 = m[0][3] * inv.m[3][0];
    float det = (d0 + d1) + (d2 + d3);
    float ood = 1.0f / det;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) inv.m[c][r] = inv.m[c][r] * ood;
    return inv;
}

// glm/ext/matrix_clip_space.inl:249-262 (perspectiveRH_NO)
inline mat4h perspective_rh_no(float fovy_rad, float aspect, float zn, float zf) {
    float th = std::tan(fovy_rad / 2.0f);
    mat4h r = {};
    r.m[0][0] = 1.0f / (aspect * th);
    r.m[1][1] = 1.0f / th;
    r.m[2][2] = -(zf + zn) / (zf - zn);
    r.m[2][3] = -1.0f;
    r.m[3][2] = -(2.0f * zf * zn) / (zf - zn);
    return r;
}
// glm/ext/matrix_transform.inl:99-118 (lookAtRH)
inline mat4h look_at_rh(float3h eye, float3h center, float3h up) {
    float3h f = unit(sub(center, eye));
    float3h s = unit(cross3(f, up));
    float3h u = cross3(s, f);
    mat4h r = {};
    r.m[0][0] = s.x; r.m[1][0] = s.y; r.m[2][0] = s.z;
    r.m[0][1] = u.x; r.m[1][1] = u.y; r.m[2][1] = u.z;
    r.m[0][2] = -f.x; r.m[1][2] = -f.y; r.m[2][2] = -f.z;
    r.m[3][0] = -dot3(s, eye); r.m[3][1] = -dot3(u, eye); r.m[3][2] = dot3(f, eye);
    r.m[3][3] = 1.0f;
    return r;
}

}  // namespace wt
