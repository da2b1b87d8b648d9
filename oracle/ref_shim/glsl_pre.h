// TEST INFRASTRUCTURE — not product code (see oracle/README.md).
//
// GLSL-as-C++ shim: lets g++ compile the reference's own shader TEXT
// (/root/reference/assets/shaders/water-grid.vert, streamed from where it lies by
// oracle/Makefile with only its "#version" line dropped) against the reference's vendored
// GLM 0.9.9.8.  Nothing from the shader is copied into this repository.
//
// GLSL storage qualifiers map onto C++ as:  uniform -> plain namespace-scope variable (shared,
// read-only while vertices run), out -> thread_local variable (one per OpenMP thread),
// in (only used on function parameters in water-grid.vert) -> nothing.
#pragma once
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../gl_sampler.h"

#ifndef WT_REF_MAX_WAVES
#define WT_REF_MAX_WAVES 4   // the reference's MAX_COUNT (render-engine.h:132, water-grid.vert:35)
#endif

namespace wt_ref_glsl {
using namespace glm;
typedef unsigned int uint;

struct sampler2D {
    wt_gl::Image2D img;
};

// texture(sampler2D, vec2): GL 4.1 LINEAR / MIRRORED_REPEAT / LOD 0 (oracle/gl_sampler.h)
inline vec4 texture(const sampler2D& s, const vec2& st) {
    float c[4];
    wt_gl::texture2d_mirrored_linear<float>(s.img, st.x, st.y, c);
    return vec4(c[0], c[1], c[2], c[3]);
}

// GLSL implicit conversions int/uint -> float that C++ overload resolution does not perform:
// water-grid.vert:125 calls mod(gl_VertexID, gridLength)  (genType mod(genType, genType) on floats)
inline float mod(int x, uint y) { return glm::mod(static_cast<float>(x), static_cast<float>(y)); }

// swizzle proxies (position.xz, (a - b).xyz) do not deduce through GLM's templates
template <typename T, qualifier Q, int E0, int E1, int E2>
inline vec3 normalize(const glm::detail::_swizzle<3, T, Q, E0, E1, E2, -1>& v) { return glm::normalize(vec3(v())); }
template <typename T, qualifier Q, int A0, int A1, int A2, int B0, int B1, int B2>
inline vec3 cross(const glm::detail::_swizzle<3, T, Q, A0, A1, A2, -1>& a,
                  const glm::detail::_swizzle<3, T, Q, B0, B1, B2, -1>& b) { return glm::cross(vec3(a()), vec3(b())); }
template <typename T, qualifier Q, int E0, int E1>
inline vec2 operator/(const glm::detail::_swizzle<2, T, Q, E0, E1, -1, -2>& a, float b) { return vec2(a()) / b; }

// built-ins
thread_local int gl_VertexID;
thread_local vec4 gl_Position;

#define uniform
#define in
#define out thread_local
#define main wt_shader_main
