// TEST INFRASTRUCTURE — not product code (see oracle/README.md).
//
// GLSL-as-C++ shim for the reference's FRAGMENT shader text
// (/root/reference/assets/shaders/water-grid.frag, streamed from where it lies by oracle/Makefile with only
// its "#version" line dropped), compiled against the reference's vendored GLM 0.9.9.8.  Nothing from the
// shader is copied into this repository.  uniform / in -> plain namespace-scope variables (the compiled
// shader is run serially, one "fragment" at a time), out -> plain variable.
#pragma once
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../gl_sampler.h"

namespace wt_ref_glsl_frag {
using namespace glm;
typedef unsigned int uint;

// The three 2-D samplers of the shader differ only in the sampler STATE the reference configures
// (render-engine.cpp:128-136 depth: GL_NEAREST float; :143-151,:172-180 colour: GL_LINEAR RGBA8; all CLAMP_TO_EDGE).
struct sampler2D {
    const float* depth = nullptr;      // non-NULL: single-channel float texture, NEAREST
    wt_gl::Image2D img{nullptr, 0, 0};  // else RGBA8, LINEAR
    int w = 0, h = 0;
};
struct samplerCube {
    wt_gl::CubeRGBA8 cube;
};

inline vec4 texture(const sampler2D& s, const vec2& st) {
    if (s.depth) {
        float d = wt_gl::texture2d_clamp_nearest_r32f<float>(s.depth, s.w, s.h, st.x, st.y);
        return vec4(d, d, d, 1.0f);   // GL: a DEPTH_COMPONENT texture samples as (d, d, d, 1)
    }
    float c[4];
    wt_gl::texture2d_clamp_linear<float>(s.img, st.x, st.y, c);
    return vec4(c[0], c[1], c[2], c[3]);
}
inline vec4 texture(const samplerCube& s, const vec3& r) {
    float c[4];
    wt_gl::texture_cube_clamp_linear<float>(s.cube, r.x, r.y, r.z, c);
    return vec4(c[0], c[1], c[2], c[3]);
}

// GLSL implicit conversions / swizzle proxies that C++ overload resolution does not perform
inline float pow(float x, int y) { return glm::pow(x, static_cast<float>(y)); }

// mix(vec3, vec4.rgb, float) (water-grid.frag:142,144): the swizzle proxy does not deduce through GLM's templates
template <typename T, qualifier Q, int E0, int E1, int E2>
inline vec3 mix(const vec3& x, const glm::detail::_swizzle<3, T, Q, E0, E1, E2, -1>& y, float a) { return glm::mix(x, vec3(y()), a); }

// built-ins
vec4 gl_FragCoord;

#define uniform
#define in
#define out
#define main wt_shader_main
