// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build or call anything under oracle/.
//
// CPU restatement of the OpenGL 4.1 core fixed-function texture sampling the reference's
// shaders delegate to the GL driver.  The reference never implements this itself; it only
// configures it:
//   * 2-D heightmap: RGBA8, GL_MIRRORED_REPEAT on S and T, GL_LINEAR min+mag (=> LOD 0 only)
//       /root/reference/src/texture.cpp:54-72, image flipped vertically at load
//       /root/reference/src/render-engine.cpp:1308-1322
//   * sky cubemap: RGBA8, GL_LINEAR, GL_CLAMP_TO_EDGE, seamless filtering never enabled
//       /root/reference/src/render-engine.cpp:90-110
//   * screen-space inputs of the water fragment shader: depth texture GL_DEPTH_COMPONENT/GL_FLOAT,
//     GL_NEAREST, GL_CLAMP_TO_EDGE (/root/reference/src/render-engine.cpp:128-136); local reflection /
//     refraction colour textures RGBA8, GL_LINEAR, GL_CLAMP_TO_EDGE (:143-151, :172-180)
// Algorithm restated from the OpenGL 4.1 core specification, section 3.8 ("Texturing"):
//   - "Texture Minification", LINEAR:  i0 = wrap(floor(u - 1/2)), i1 = wrap(floor(u - 1/2) + 1),
//     alpha = frac(u - 1/2); tau = (1-a)(1-b) t00 + a(1-b) t10 + (1-a) b t01 + a b t11
//   - "Wrapping", MIRRORED_REPEAT:  (size-1) - mirror((coord mod 2 size) - size),
//     mirror(a) = a >= 0 ? a : -(1 + a);  CLAMP_TO_EDGE: clamp(coord, 0, size-1)
//   - "Cube Map Texture Selection" (major-axis table, s = (sc/|ma| + 1)/2, t = (tc/|ma| + 1)/2)
//   - unsigned normalised fixed point -> float:  f = c / (2^8 - 1)
// Real GL implementations quantise the filter weights (typically 8 fractional bits); this
// restatement keeps exact fp32 (or fp64) weights, which is the limit those implementations
// approximate.  The arithmetic ORDER below is the specification of the oracle: weights are
// formed first, the four weighted texels are summed left to right.
#pragma once
#include <cmath>
#include <cstdint>

namespace wt_gl {

struct Image2D {          // tightly packed RGBA8, row 0 = GL row 0 (i.e. already flipped like stbi does)
    const uint8_t* rgba;
    int w, h;
};

struct CubeRGBA8 {        // 6 square faces in GL order +x,-x,+y,-y,+z,-z
    const uint8_t* face[6];
    int n;
};

inline int wrap_mirrored_repeat(long long i, int size) {
    long long m = i % (2LL * size);
    if (m < 0) m += 2LL * size;           // floored modulo
    long long a = m - size;
    long long mir = a >= 0 ? a : -(1 + a);
    return (int)((size - 1) - mir);
}

inline int wrap_clamp_to_edge(long long i, int size) {
    return (int)(i < 0 ? 0 : (i > size - 1 ? size - 1 : i));
}

template <typename T>
inline T unorm8(uint8_t c) { return (T)c / (T)255; }

// GL_LINEAR fetch of all four channels at LOD 0.  s,t are the (unwrapped) normalised coords.
template <typename T, int (*WRAP)(long long, int)>
inline void sample_linear_rgba(const uint8_t* rgba, int w, int h, T s, T t, T out[4]) {
    T u = s * (T)w;
    T v = t * (T)h;
    T um = u - (T)0.5;
    T vm = v - (T)0.5;
    T fu = std::floor(um);
    T fv = std::floor(vm);
    T a = um - fu;
    T b = vm - fv;
    long long iu = (long long)fu, iv = (long long)fv;
    int i0 = WRAP(iu, w), i1 = WRAP(iu + 1, w);
    int j0 = WRAP(iv, h), j1 = WRAP(iv + 1, h);
    T w00 = ((T)1 - a) * ((T)1 - b);
    T w10 = a * ((T)1 - b);
    T w01 = ((T)1 - a) * b;
    T w11 = a * b;
    const uint8_t* p00 = rgba + 4 * ((size_t)j0 * w + i0);
    const uint8_t* p10 = rgba + 4 * ((size_t)j0 * w + i1);
    const uint8_t* p01 = rgba + 4 * ((size_t)j1 * w + i0);
    const uint8_t* p11 = rgba + 4 * ((size_t)j1 * w + i1);
    for (int c = 0; c < 4; ++c) {
        out[c] = w00 * unorm8<T>(p00[c]) + w10 * unorm8<T>(p10[c]) + w01 * unorm8<T>(p01[c]) +
                 w11 * unorm8<T>(p11[c]);
    }
}

// texture(sampler2D, st) with the reference's heightmap sampler state.
template <typename T>
inline void texture2d_mirrored_linear(const Image2D& img, T s, T t, T out[4]) {
    sample_linear_rgba<T, wrap_mirrored_repeat>(img.rgba, img.w, img.h, s, t, out);
}

// texture(sampler1D, s): RGBA8, GL_LINEAR, GL_CLAMP_TO_EDGE, LOD 0 — the sky-gradient texture
// (/root/reference/src/texture.cpp:40-49).  Same LINEAR rule as the 2-D case, one dimension.
template <typename T>
inline void texture1d_clamp_linear(const uint8_t* rgba, int length, T s, T out[4]) {
    T um = s * (T)length - (T)0.5;
    T fu = std::floor(um);
    T a = um - fu;
    long long iu = (long long)fu;
    int i0 = wrap_clamp_to_edge(iu, length), i1 = wrap_clamp_to_edge(iu + 1, length);
    for (int c = 0; c < 4; ++c) out[c] = ((T)1 - a) * unorm8<T>(rgba[4 * i0 + c]) + a * unorm8<T>(rgba[4 * i1 + c]);
}

// texture(sampler2D, st) with the sampler state of the local reflection / refraction textures.
template <typename T>
inline void texture2d_clamp_linear(const Image2D& img, T s, T t, T out[4]) {
    sample_linear_rgba<T, wrap_clamp_to_edge>(img.rgba, img.w, img.h, s, t, out);
}

// texture(depthTexture2D, st).x: GL_NEAREST ("Texture Minification", NEAREST: i = wrap(floor(u))), CLAMP_TO_EDGE.
template <typename T>
inline T texture2d_clamp_nearest_r32f(const float* texels, int w, int h, T s, T t) {
    long long i = (long long)std::floor(s * (T)w), j = (long long)std::floor(t * (T)h);
    return (T)texels[(size_t)wrap_clamp_to_edge(j, h) * w + wrap_clamp_to_edge(i, w)];
}

// texture(samplerCube, r) with the reference's sky cubemap sampler state (non-seamless).
// Major axis selection: the component of largest magnitude; ties resolved x, then y, then z
// (the spec leaves ties implementation-defined; this order is the oracle's definition).
template <typename T>
inline void texture_cube_clamp_linear(const CubeRGBA8& cube, T rx, T ry, T rz, T out[4]) {
    T ax = std::fabs(rx), ay = std::fabs(ry), az = std::fabs(rz);
    int face;
    T sc, tc, ma;
    if (ax >= ay && ax >= az) {
        ma = ax;
        if (rx >= 0) { face = 0; sc = -rz; tc = -ry; } else { face = 1; sc = rz; tc = -ry; }
    } else if (ay >= az) {
        ma = ay;
        if (ry >= 0) { face = 2; sc = rx; tc = rz; } else { face = 3; sc = rx; tc = -rz; }
    } else {
        ma = az;
        if (rz >= 0) { face = 4; sc = rx; tc = -ry; } else { face = 5; sc = -rx; tc = -ry; }
    }
    T s = (T)0.5 * (sc / ma + (T)1);
    T t = (T)0.5 * (tc / ma + (T)1);
    sample_linear_rgba<T, wrap_clamp_to_edge>(cube.face[face], cube.n, cube.n, s, t, out);
}

}  // namespace wt_gl
