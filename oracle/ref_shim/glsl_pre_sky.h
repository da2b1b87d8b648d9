// TEST INFRASTRUCTURE — not product code (see oracle/README.md).
//
// GLSL-as-C++ shim for the reference's three sky FRAGMENT shader texts (skysphere.frag, skybox-clouds.frag,
// skybox-stars.frag under /root/reference/assets/shaders, streamed from where they lie by oracle/Makefile with
// only their "#version" lines dropped; each lands in its own namespace).  Nothing from the shaders is copied
// into this repository.  uniform / in / out -> plain namespace-scope variables; run serially.
#pragma once
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../gl_sampler.h"

namespace wt_ref_glsl_sky {
using namespace glm;
typedef unsigned int uint;

// sampler state the reference configures: 1-D gradient RGBA8 GL_LINEAR / CLAMP_TO_EDGE (texture.cpp:40-49);
// star and cloud cubemaps RGBA8 GL_LINEAR / CLAMP_TO_EDGE (render-engine.cpp:1353-1357)
struct sampler1D {
    const uint8_t* rgba = nullptr;
    int length = 0;
};
struct samplerCube {
    wt_gl::CubeRGBA8 cube;
};
inline vec4 texture(const sampler1D& s, float coord) {
    float c[4];
    wt_gl::texture1d_clamp_linear<float>(s.rgba, s.length, coord, c);
    return vec4(c[0], c[1], c[2], c[3]);
}
inline vec4 texture(const samplerCube& s, const vec3& r) {
    float c[4];
    wt_gl::texture_cube_clamp_linear<float>(s.cube, r.x, r.y, r.z, c);
    return vec4(c[0], c[1], c[2], c[3]);
}
// mix(texture(...).rgb, vec3, float) (skysphere.frag:56): the swizzle proxy does not deduce through GLM's templates
template <typename T, qualifier Q, int E0, int E1, int E2>
inline vec3 mix(const glm::detail::_swizzle<3, T, Q, E0, E1, E2, -1>& x, const vec3& y, float a) { return glm::mix(vec3(x()), y, a); }
}  // namespace wt_ref_glsl_sky

#define uniform
#define in
#define out
#define main wt_shader_main
