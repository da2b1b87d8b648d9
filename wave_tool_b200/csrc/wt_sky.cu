// Skybox synthesis — replaces the cubemap render pass of src/render-engine.cpp:306-442: for each
// of the 6 x n x n texels the clear colour, then the star cubemap (skybox-stars.frag), the sky sphere
// (skysphere.frag), the cloud cubemap (skybox-clouds.frag) and the fog quad (screen-space-quad.frag) blended in that
// order with glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA) into RGBA8.  One thread per texel; the result is
// written straight into the context's skybox (linear faces + the padded gather atlas the surface kernel samples).
// HBM-bound: 4 B written per texel, the two source cubemaps are read with 2-D locality.
#include <cmath>
#include <string>

#include "wt_ctx.h"

namespace wt {

struct SkyArgs {
    uchar4* out;                 // 6 * n * n texels, GL face order
    int n;
    float inv_n_exact;           // 1 / n when n is a power of two (then x / n == x * (1 / n) exactly), else 0
    const uchar4* stars; int stars_n;     // NULL: layer not drawn
    const uchar4* clouds; int clouds_n;
    int draw_sphere;
    float Lx, Ly, Lz;            // normalize(sunPosition)
    float sun_y;                 // sunPosition.y (un-normalised, clouds shader)
    float peak[3], horizon[3];   // the two gradient samples of skysphere.frag:55-56 (per-frame constants, host)
    float shininess, strength, alpha1;
    float omp, overcast;         // oneMinusCloudProportion, overcastStrength
    float fog[4];
    float inv_gamma;             // fl(1 / 2.4) as the shader's constant expression folds it (skybox-clouds.frag:54)
    float inv_one_minus_omp;     // 1 / (1 - omp) for the approximate path
    unsigned long long* unsure_count;   // diagnostic (may be NULL): layers that were redone on the exact path
    int fast;                    // 1: approximate path with guard bands + exact fallback per texel; 0: exact path only (WT_SKY_EXACT build / debugging)
    cudaSurfaceObject_t atlas;   // the padded gather atlas the surface kernel samples: written by the same kernel (0: not written)
    int fog_identity;            // fog alpha == 0: the fog quad leaves every byte as it is
    unsigned fog_lut[256];       // the fog quad's blend per destination byte d: byte c of entry d = the byte channel c holds afterwards
};

__device__ __forceinline__ float mixs(float x, float y, float a) { return __fadd_rn(__fmul_rn(x, __fsub_rn(1.0f, a)), __fmul_rn(y, a)); }
__device__ __forceinline__ float sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// texture(samplerCube, r): GL 4.1 face selection, per-face CLAMP_TO_EDGE, GL_LINEAR with exact fp32 weights (oracle order)
__device__ __forceinline__ float4 cube_rgba8(const uchar4* faces, int n, float rx, float ry, float rz) {
    const float ax = fabsf(rx), ay = fabsf(ry), az = fabsf(rz);
    int face;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (rx >= 0.0f) { face = 0; sc = -rz; tc = -ry; } else { face = 1; sc = rz; tc = -ry; } }
    else if (ay >= az) { ma = ay; if (ry >= 0.0f) { face = 2; sc = rx; tc = rz; } else { face = 3; sc = rx; tc = -rz; } }
    else { ma = az; if (rz >= 0.0f) { face = 4; sc = rx; tc = -ry; } else { face = 5; sc = -rx; tc = -ry; } }
    // the directions of texel centres have a major component of exactly +-1: x / 1 == x, no division needed
    const float qs = ma == 1.0f ? sc : __fdiv_rn(sc, ma), qt = ma == 1.0f ? tc : __fdiv_rn(tc, ma);
    const float s = __fmul_rn(0.5f, __fadd_rn(qs, 1.0f)), t = __fmul_rn(0.5f, __fadd_rn(qt, 1.0f));
    const float um = __fsub_rn(__fmul_rn(s, (float)n), 0.5f), vm = __fsub_rn(__fmul_rn(t, (float)n), 0.5f);
    int iu, iv;
    const float fu = floor_small(um, iu), fv = floor_small(vm, iv);
    const float al = __fsub_rn(um, fu), be = __fsub_rn(vm, fv);
    const int i0 = min(max(iu, 0), n - 1), i1 = min(max(iu + 1, 0), n - 1), j0 = min(max(iv, 0), n - 1), j1 = min(max(iv + 1, 0), n - 1);
    const unsigned* f = reinterpret_cast<const unsigned*>(faces) + (size_t)face * n * n;
    const unsigned t00 = __ldg(f + (size_t)j0 * n + i0), t10 = __ldg(f + (size_t)j0 * n + i1), t01 = __ldg(f + (size_t)j1 * n + i0), t11 = __ldg(f + (size_t)j1 * n + i1);
    const float oma = __fsub_rn(1.0f, al), omb = __fsub_rn(1.0f, be);
    const float w00 = __fmul_rn(oma, omb), w10 = __fmul_rn(al, omb), w01 = __fmul_rn(oma, be), w11 = __fmul_rn(al, be);
    auto ch = [&](unsigned k) {
        float acc = __fmul_rn(w00, unorm8f(byte_f(t00, k)));
        acc = __fadd_rn(acc, __fmul_rn(w10, unorm8f(byte_f(t10, k))));
        acc = __fadd_rn(acc, __fmul_rn(w01, unorm8f(byte_f(t01, k))));
        return __fadd_rn(acc, __fmul_rn(w11, unorm8f(byte_f(t11, k))));
    };
    return make_float4(ch(0), ch(1), ch(2), 1.0f);
}

// one layer into the RGBA8 destination: operands clamped to [0,1], Cd' = Cs As + Cd (1 - As) on all four channels,
// converted back to unorm8 with round-to-nearest (GL 4.1 sections 4.1.7, 2.1.5)
// dst holds the attachment's bytes as integer-valued floats.
// FAST instantiation: the source colour / alpha carry absolute error bounds (err_rgb, err_a: they came out of MUFU
// approximations of pow); `unsure` is raised when a channel's value before the final floor lies so close to an integer that
// the bound cannot exclude a different byte.  A texel that raises it is recomputed on the exact path.
template <bool FAST>
__device__ __forceinline__ void blend(float dst[4], float r, float g, float b, float a, float err_rgb, float err_a, bool& unsure) {
    const float src[4] = {sat(r), sat(g), sat(b), sat(a)};
    // As == 0: Cs * 0 + Cd * 1 = Cd = c / 255, which converts back to the same byte c: nothing to do
    // (a shortcut, not a discontinuity: blending with As = 0 leaves the same bytes — so an approximate As near 0 needs no guard)
    if (src[3] == 0.0f && (!FAST || err_a == 0.0f)) return;
    const float oma = __fsub_rn(1.0f, src[3]);
    const bool opaque = src[3] == 1.0f && (!FAST || err_a == 0.0f);   // Cs * 1 + Cd * 0 == Cs exactly (Cd is finite): the destination is not read
    float edge = 0.0f;                     // FAST: the largest distance of a guarded channel from the integer it rounds to
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float v;
        if (opaque) {
            v = src[c];
        } else {
            const float d = unorm8f(dst[c]);   // what the blender reads back from the RGBA8 attachment: c / 255
            // in [0, 1 + 2^-23]: the conversion's clamp to [0,1] cannot change the byte (255 (1 + 2^-23) + 1/2 still floors to 255)
            v = __fadd_rn(__fmul_rn(src[c], src[3]), __fmul_rn(d, oma));
        }
        const float s = __fmul_rn(v, 255.0f);
        int unused;
        const float q = floor_small(__fadd_rn(s, 0.5f), unused);
        // the alpha channel depends on As alone: exact, and left out of the guard, when As is
        if (FAST && (c < 3 || err_a != 0.0f)) edge = fmaxf(edge, fabsf(__fsub_rn(s, q)));
        dst[c] = q;
    }
    if (FAST) {
        // q = round-half-up(s): a different byte needs s to move across q +- 1/2.  d(out)/d(src_c) = As <= 1, d(out)/d(As) =
        // src_c - d in [-1, 1] (the alpha channel: 2 As - d in [-1, 2]); 255 byte units per unit, plus the roundings of the blend
        // itself (a few ulp of a value <= 1 and of s + 1/2: 1e-4 byte units is ample)
        const float e = 255.0f * fmaxf(err_rgb + err_a, 2.0f * err_a) + 1e-4f;
        unsure = unsure || ((err_rgb != 0.0f || err_a != 0.0f) && edge > 0.5f - e);
    }
}

// texture(samplerCube, d) for the direction through the centre of texel k of a cubemap of the SAME power-of-two size as the
// target (the reference's own configuration: 2048^2 star / cloud assets, CUBEMAP_LENGTH 2048 — program.cpp:616,656,
// render-engine.h:191): face selection returns the target's face (the major component is exactly +-1, the others are
// smaller), s n - 1/2 is exactly the texel's own index, so the bilinear weights are (1, 0, 0, 0) and 1 v + 0 + 0 + 0 == v.
__device__ __forceinline__ float4 cube_rgba8_aligned(const uchar4* faces, size_t k) {
    const unsigned t = __ldg(reinterpret_cast<const unsigned*>(faces) + k);
    return make_float4(unorm8f(byte_f(t, 0)), unorm8f(byte_f(t, 1)), unorm8f(byte_f(t, 2)), 1.0f);
}

// All layers of one texel (direction d) into dst (integer-valued floats 0..255), the fog quad excepted (fog_layer below).
// The two pow() of the sky shaders are first evaluated with MUFU lg2 / ex2 and the division by (1 - omp) as a multiplication
// by its reciprocal; the blend carries bounds on the error this introduces and tells when a comparison or a byte conversion
// could come out differently.  Only then the layer is redone exactly (libdevice powf, IEEE division) from the bytes the
// previous layers left — every layer starts from exact bytes, so the fallback is local to the layer.  With a.fast == 0 the
// exact forms are used throughout.
template <bool ALIGNED>
__device__ __forceinline__ void sky_texel(const SkyArgs& a, size_t k, float dx, float dy, float dz, float dst[4], unsigned& n_exact) {
    dst[0] = dst[1] = dst[2] = dst[3] = 255.0f;                          // glClearColor(1,1,1,1), render-engine.cpp:694,330
    bool unused = false;
    if (a.stars) {                                                       // skybox-stars.frag:41
        const float4 s = ALIGNED ? cube_rgba8_aligned(a.stars, k) : cube_rgba8(a.stars, a.stars_n, dx, dy, dz);
        blend<false>(dst, s.x, s.y, s.z, 1.0f, 0.0f, 0.0f, unused);
    }
    if (a.draw_sphere) {                                                 // skysphere.frag:48-76
        const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
        const float nx = __fmul_rn(dx, inv), ny = __fmul_rn(dy, inv), nz = __fmul_rn(dz, inv);
        const float a3 = __fadd_rn(__fadd_rn(__fmul_rn(nx, a.Lx), __fmul_rn(ny, a.Ly)), __fmul_rn(nz, a.Lz));
        const float night = __fmul_rn(fabsf(a3), __fsub_rn(1.0f, fabsf(a.Ly)));
        float diffuse[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) diffuse[c] = a3 >= 0.0f ? mixs(a.horizon[c], a.peak[c], a3) : mixs(a.horizon[c], 0.0f, night);
        const float alpha = fmaxf(a.alpha1, __fmul_rn(__fadd_rn(a3, 1.0f), 0.5f));   // x / 2.0 == x * 0.5 exactly
        const float s1 = __fmul_rn(a.strength, 1.0f);
        float spec = 0.0f, err = 0.0f;
        bool exact = !a.fast || a3 <= 0.0f || a.shininess == 0.0f;
        if (a3 > 0.0f) {
            if (exact) {
                spec = __fmul_rn(s1, powf(a3, a.shininess));
            } else {
                // pow(a3, shininess) = 2^(shininess log2 a3): lg2.approx is within 2^-22 absolute on [0.5, 2] and 2 ulp elsewhere, ex2.approx within 2 ulp;
                // the product's own rounding adds 2^-24 |shininess lg|
                const float lg = lg2_approx(a3);
                spec = __fmul_rn(s1, ex2_approx(__fmul_rn(a.shininess, lg)));
                err = __fmul_rn(spec, __fmaf_rn(a.shininess, 2.5e-7f * (1.0f + fabsf(lg)), 6e-7f)) + 1e-30f;
            }
        }
        const float keep[4] = {dst[0], dst[1], dst[2], dst[3]};
        bool unsure = false;
        blend<true>(dst, __fadd_rn(diffuse[0], spec), __fadd_rn(diffuse[1], spec), __fadd_rn(diffuse[2], spec), alpha, err, 0.0f, unsure);
        if (unsure) {                                                    // a byte within the error bound of a boundary: this layer again, exactly
            spec = __fmul_rn(s1, powf(a3, a.shininess));
            dst[0] = keep[0]; dst[1] = keep[1]; dst[2] = keep[2]; dst[3] = keep[3];
            blend<false>(dst, __fadd_rn(diffuse[0], spec), __fadd_rn(diffuse[1], spec), __fadd_rn(diffuse[2], spec), alpha, 0.0f, 0.0f, unused);
            ++n_exact;
        }
    }
    if (a.clouds) {                                                      // skybox-clouds.frag:46-67
        const float4 raw = ALIGNED ? cube_rgba8_aligned(a.clouds, k) : cube_rgba8(a.clouds, a.clouds_n, dx, dy, dz);
        const float lin = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, raw.x), __fmul_rn(0.7152f, raw.y)), __fmul_rn(0.0722f, raw.z));
        const float intensity = fminf(fmaxf(a.sun_y, 0.2f), 1.0f);
        const float oc = __fmul_rn(sat(a.sun_y), a.overcast);
        const float keep[4] = {dst[0], dst[1], dst[2], dst[3]};
        // the layer for a given srgb value: exact (IEEE division) or with the reciprocal form and the error bounds es / eca
        auto layer = [&](float srgb, float es, bool approx, bool& unsure) {
            if (approx && fabsf(__fsub_rn(srgb, a.omp)) <= es) unsure = true;     // the branch itself hangs on the approximation
            if (srgb > a.omp) {
                float ca, eca = 0.0f;
                if (approx) {
                    ca = __fmul_rn(__fsub_rn(srgb, a.omp), a.inv_one_minus_omp);
                    eca = __fmaf_rn(es, a.inv_one_minus_omp, 1.3e-7f * fabsf(ca) + 1e-30f);   // the approximation + the rounding of the reciprocal form
                } else {
                    ca = __fdiv_rn(__fsub_rn(srgb, a.omp), __fsub_rn(1.0f, a.omp));
                }
                const float v = mixs(intensity, srgb, oc);
                if (approx) blend<true>(dst, v, v, v, mixs(ca, 1.0f, oc), __fmul_rn(es, oc), __fmul_rn(eca, __fsub_rn(1.0f, oc)), unsure);
                else blend<false>(dst, v, v, v, mixs(ca, 1.0f, oc), 0.0f, 0.0f, unsure);
            } else {
                blend<false>(dst, 0.5f, 0.5f, 0.5f, oc, 0.0f, 0.0f, unsure);
            }
        };
        if (lin <= 0.0031308f) {
            bool u = false;
            layer(__fmul_rn(12.92f, lin), 0.0f, a.fast != 0, u);          // srgb itself is exact here; only the reciprocal form is approximate
            if (u) { dst[0] = keep[0]; dst[1] = keep[1]; dst[2] = keep[2]; dst[3] = keep[3]; layer(__fmul_rn(12.92f, lin), 0.0f, false, unused); ++n_exact; }
        } else if (a.fast) {
            const float lg = lg2_approx(lin);
            const float p = ex2_approx(__fmul_rn(a.inv_gamma, lg));
            const float srgb = __fsub_rn(__fmul_rn(1.055f, p), 0.055f);
            // exponent error inv_gamma |lg| 2^-22 (+ its rounding), ex2 2^-22 relative, two more roundings; p <= 1
            const float es = __fmul_rn(p, __fmaf_rn(fabsf(lg), 1.0e-7f, 4.5e-7f)) + 7e-8f;
            bool u = false;
            layer(srgb, es, true, u);
            if (u) {
                dst[0] = keep[0]; dst[1] = keep[1]; dst[2] = keep[2]; dst[3] = keep[3];
                layer(__fsub_rn(__fmul_rn(1.055f, powf(lin, a.inv_gamma)), 0.055f), 0.0f, false, unused);
                ++n_exact;
            }
        } else {
            layer(__fsub_rn(__fmul_rn(1.055f, powf(lin, a.inv_gamma)), 0.055f), 0.0f, false, unused);
        }
    }
}

// The fog quad (screen-space-quad.frag:45) blends a per-frame constant colour and alpha over the attachment: its result is a
// function of the destination byte alone, tabulated once per frame by the host (fog_table below, same operations as blend()).
// Returns the packed texel.
__device__ __forceinline__ unsigned fog_layer(const unsigned* lut, unsigned rgba) {
    const unsigned r = lut[rgba & 255u], g = lut[(rgba >> 8) & 255u], b = lut[(rgba >> 16) & 255u], al = lut[rgba >> 24];
    return __byte_perm(__byte_perm(r, g, 0x0050), __byte_perm(b, al, 0x0072), 0x5410);
}

// pack the four integer-valued floats (0..255) into RGBA8: 2^23 + c has c in its low mantissa byte
__device__ __forceinline__ unsigned pack_rgba8(const float dst[4]) {
    const unsigned b0 = __float_as_uint(__fadd_rn(dst[0], 8388608.0f)), b1 = __float_as_uint(__fadd_rn(dst[1], 8388608.0f));
    const unsigned b2 = __float_as_uint(__fadd_rn(dst[2], 8388608.0f)), b3 = __float_as_uint(__fadd_rn(dst[3], 8388608.0f));
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256) sky_synth_kernel(const __grid_constant__ SkyArgs a) {
    // grid = (ceil(n / 256), n, 6): the face and the row come from the block index (no integer divisions, and the face
    // switch below is uniform per block)
    const int face = blockIdx.z, j = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ unsigned lut[256];
    if (!a.fog_identity) {
        lut[threadIdx.x] = a.fog_lut[threadIdx.x];
        __syncthreads();
    }
    if (i < a.n) {
        const unsigned k = ((unsigned)face * a.n + j) * a.n + i;     // 6 n^2 <= 6 * 2^28 for n <= 16384
        // direction through the centre of texel (i, j) of the face (inverse of the GL cube-map face table)
        // (2 (i + 1/2)) / n - 1; for a power-of-two n the division is a multiplication by the exact reciprocal
        const float ni = __fmul_rn(2.0f, __fadd_rn((float)i, 0.5f)), nj = __fmul_rn(2.0f, __fadd_rn((float)j, 0.5f));
        const float sc = __fsub_rn(a.inv_n_exact != 0.0f ? __fmul_rn(ni, a.inv_n_exact) : __fdiv_rn(ni, (float)a.n), 1.0f);
        const float tc = __fsub_rn(a.inv_n_exact != 0.0f ? __fmul_rn(nj, a.inv_n_exact) : __fdiv_rn(nj, (float)a.n), 1.0f);
        float dx, dy, dz;
        switch (face) {
            case 0: dx = 1.0f; dy = -tc; dz = -sc; break;
            case 1: dx = -1.0f; dy = -tc; dz = sc; break;
            case 2: dx = sc; dy = 1.0f; dz = tc; break;
            case 3: dx = sc; dy = -1.0f; dz = -tc; break;
            case 4: dx = sc; dy = -tc; dz = 1.0f; break;
            default: dx = -sc; dy = -tc; dz = -1.0f; break;
        }
        float dst[4];
        unsigned n_exact = 0;
        sky_texel<ALIGNED>(a, k, dx, dy, dz, dst, n_exact);
        unsigned rgba = pack_rgba8(dst);
        if (!a.fog_identity) rgba = fog_layer(lut, rgba);
        if (a.unsure_count && n_exact) atomicAdd(a.unsure_count, (unsigned long long)n_exact);
        reinterpret_cast<unsigned*>(a.out)[k] = rgba;
        if (a.atlas) {
            // the gather atlas: tile column = face / 2, tile row = face % 2, 1-texel border replicating the edge
            const int tile = a.n + 2, x0 = (face / 2) * tile, y0 = (face % 2) * tile;
            const uchar4 v = make_uchar4(rgba & 255u, (rgba >> 8) & 255u, (rgba >> 16) & 255u, rgba >> 24);
            auto put_row = [&](int y) {
                surf2Dwrite(v, a.atlas, (x0 + i + 1) * 4, y);
                if (i == 0) surf2Dwrite(v, a.atlas, x0 * 4, y);
                if (i == a.n - 1) surf2Dwrite(v, a.atlas, (x0 + a.n + 1) * 4, y);
            };
            put_row(y0 + j + 1);
            if (j == 0) put_row(y0);                                                    // uniform per block (j = blockIdx.y)
            if (j == a.n - 1) put_row(y0 + a.n + 1);
        }
    }
}

// linear faces -> the padded gather atlas (tile column = face / 2, tile row = face % 2, 1-texel border replicating the edge)
__global__ void __launch_bounds__(256) sky_atlas_kernel(const uchar4* faces, int n, cudaSurfaceObject_t surf) {
    // grid = (ceil(tile / 256), tile, 6): one block row per atlas row of one face tile
    const int tile = n + 2;
    const int face = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < tile) {
        const int i = min(max(x - 1, 0), n - 1), j = min(max(y - 1, 0), n - 1);
        surf2Dwrite(faces[((size_t)face * n + j) * n + i], surf, ((face / 2) * tile + x) * (int)sizeof(uchar4), (face % 2) * tile + y);
    }
}

// (re)allocates the atlas array, its texture object and the surface object the kernels write it through
static cudaError_t allocate_sky_atlas(wt_ctx* ctx, int n, cudaArray_t* arr, cudaTextureObject_t* tex) {
    cudaError_t e;
    const int tile = n + 2, W = 3 * tile, H = 2 * tile;
    {
        if (*tex) { cudaDestroyTextureObject(*tex); *tex = 0; }
        if (*arr) { cudaFreeArray(*arr); *arr = nullptr; }
        cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
        if ((e = cudaMallocArray(arr, &fmt, W, H, cudaArrayTextureGather | cudaArraySurfaceLoadStore)) != cudaSuccess) return e;
        cudaResourceDesc res;
        std::memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;
        res.res.array.array = *arr;
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 0;
        if ((e = cudaCreateTextureObject(tex, &res, &td, nullptr)) != cudaSuccess) return e;
        if (ctx->sky_surf) { cudaDestroySurfaceObject(ctx->sky_surf); ctx->sky_surf = 0; }
        if ((e = cudaCreateSurfaceObject(&ctx->sky_surf, &res)) != cudaSuccess) return e;   // kept for the life of the array
    }
    return cudaSuccess;
}

cudaError_t rebuild_sky_atlas(wt_ctx* ctx, const uchar4* d_faces, int n, cudaArray_t* arr, cudaTextureObject_t* tex, bool allocate, cudaStream_t stream) {
    cudaError_t e;
    const int tile = n + 2;
    if (allocate && (e = allocate_sky_atlas(ctx, n, arr, tex)) != cudaSuccess) return e;
    sky_atlas_kernel<<<dim3((tile + 255) / 256, tile, 6), 256, 0, stream>>>(d_faces, n, ctx->sky_surf);
    ctx->launches++;
    return cudaGetLastError();
}

// texture(sampler1D, s).rgb of the sky gradient on the host (two samples per frame): GL_LINEAR, CLAMP_TO_EDGE, fp32
static void gradient_sample(const std::vector<uint8_t>& g, float s, float out[3]) {
    const int len = (int)(g.size() / 4);
    const float um = s * (float)len - 0.5f;
    const float fu = std::floor(um);
    const float a = um - fu;
    const long long iu = (long long)fu;
    const int i0 = (int)std::min<long long>(std::max<long long>(iu, 0), len - 1), i1 = (int)std::min<long long>(std::max<long long>(iu + 1, 0), len - 1);
    for (int c = 0; c < 3; ++c) out[c] = (1.0f - a) * ((float)g[4 * i0 + c] / 255.0f) + a * ((float)g[4 * i1 + c] / 255.0f);
}

// The fog quad's blend for every destination byte (same operations, in the same order, as the device blend<false>():
// plain fp32, no contraction — x86-64 host code has no fused multiply-add unless asked for).
static void fog_table(const float fog[4], unsigned lut[256]) {
    auto sat = [](float x) { return std::fmin(std::fmax(x, 0.0f), 1.0f); };
    const float src[4] = {sat(fog[0]), sat(fog[1]), sat(fog[2]), sat(fog[3])};
    const volatile float oma = 1.0f - src[3];
    for (int d = 0; d < 256; ++d) {
        unsigned e = 0;
        for (int c = 0; c < 4; ++c) {
            const volatile float dn = (float)d / 255.0f;
            const volatile float t0 = src[c] * src[3], t1 = dn * oma;
            const volatile float v = sat(t0 + t1);
            const volatile float t2 = v * 255.0f;
            const float q = std::floor(t2 + 0.5f);
            e |= (unsigned)q << (8 * c);
        }
        lut[d] = e;
    }
}

}  // namespace wt

using namespace wt;

extern "C" {

void wt_default_sky_params(wt_sky_params* out) {   // render-engine.h:220-221,228,230-233
    if (!out) return;
    out->timeOfDayInHours = 9.0f;
    out->cloudProportion = 0.6f;
    out->overcastStrength = 0.0f;
    out->sunHorizonDarkness = 0.25f;
    out->sunShininess = 50.0f;
    out->sunStrength = 1.0f;
    out->fogColourFarAtNoon[0] = out->fogColourFarAtNoon[1] = out->fogColourFarAtNoon[2] = 1.0f;
    out->fogColourFarAtNoon[3] = 0.1f;
    out->drawStars = out->drawSkysphere = out->drawClouds = 1;
}

int wt_set_sky_layers(wt_ctx* ctx, const uint8_t* gradient_rgba, int gradient_len, const uint8_t* const stars[6], int stars_n,
                      const uint8_t* const clouds[6], int clouds_n) {
    if (!ctx) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_set_sky_layers: ctx is NULL");
    auto fail = [&](int code, const char* msg) { ctx->err = msg; set_thread_error(code, msg); return code; };
    if (!gradient_rgba || gradient_len <= 0) return fail(WT_ERR_INVALID_ARGUMENT, "wt_set_sky_layers: the sky gradient is required");
    if ((stars && (stars_n <= 0 || stars_n > 16384)) || (clouds && (clouds_n <= 0 || clouds_n > 16384))) return fail(WT_ERR_INVALID_ARGUMENT, "wt_set_sky_layers: bad cubemap size");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    ctx->sky_gradient.assign(gradient_rgba, gradient_rgba + 4 * (size_t)gradient_len);
    cudaFree(ctx->d_stars); cudaFree(ctx->d_clouds);
    ctx->d_stars = ctx->d_clouds = nullptr;
    ctx->stars_n = ctx->clouds_n = 0;
    cudaError_t e = cudaSuccess;
    auto upload = [&](const uint8_t* const faces[6], int n, uchar4** dst) {
        const size_t fb = (size_t)n * n * 4;
        if ((e = cudaMalloc(dst, 6 * fb)) != cudaSuccess) return;
        for (int f = 0; f < 6 && e == cudaSuccess; ++f) {
            if (!faces[f]) { e = cudaErrorInvalidValue; return; }
            e = cudaMemcpy((uint8_t*)*dst + f * fb, faces[f], fb, cudaMemcpyHostToDevice);
        }
    };
    if (stars) { upload(stars, stars_n, &ctx->d_stars); if (e == cudaSuccess) ctx->stars_n = stars_n; }
    if (e == cudaSuccess && clouds) { upload(clouds, clouds_n, &ctx->d_clouds); if (e == cudaSuccess) ctx->clouds_n = clouds_n; }
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(WT_ERR_CUDA, cudaGetErrorString(e));
    return WT_OK;
}

int wt_synthesize_skybox(wt_ctx* ctx, const wt_sky_params* p, int n, void* stream_) {
    if (!ctx) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_synthesize_skybox: ctx is NULL");
    auto fail = [&](int code, const std::string& msg) { ctx->err = msg; set_thread_error(code, msg.c_str()); return code; };
    if (!p || n <= 0 || n > 16384) return fail(WT_ERR_INVALID_ARGUMENT, "wt_synthesize_skybox: bad argument");
    if (ctx->sky_gradient.empty()) return fail(WT_ERR_NOT_READY, "wt_set_sky_layers has not been called");
    cudaStream_t stream = (cudaStream_t)stream_;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaSuccess;
    const bool fresh = ctx->sky_n != n || !ctx->d_sky || !ctx->sky_surface_ok;
    if (fresh) {
        cudaFree(ctx->d_sky);
        ctx->d_sky = nullptr;
        e = cudaMalloc(&ctx->d_sky, 6 * (size_t)n * n * sizeof(uchar4));
    }
    SkyArgs a;
    std::memset(&a, 0, sizeof(a));
    if (e == cudaSuccess) {
        a.out = ctx->d_sky; a.n = n;
        a.inv_n_exact = (n & (n - 1)) == 0 ? 1.0f / (float)n : 0.0f;
        a.stars = p->drawStars ? ctx->d_stars : nullptr; a.stars_n = ctx->stars_n;
        a.clouds = p->drawClouds ? ctx->d_clouds : nullptr; a.clouds_n = ctx->clouds_n;
        a.draw_sphere = p->drawSkysphere;
        float sun[3];
        wt_sun_position(p->timeOfDayInHours, sun);                       // render-engine.cpp:289-291
        const float inv = 1.0f / std::sqrt(sun[0] * sun[0] + sun[1] * sun[1] + sun[2] * sun[2]);   // glm::normalize = v * inversesqrt(dot)
        a.Lx = sun[0] * inv; a.Ly = sun[1] * inv; a.Lz = sun[2] * inv;
        a.sun_y = sun[1];
        const float a1 = 0.25f * (a.Ly + 1.0f), a2 = 1.0f - a1;         // skysphere.frag:53-54
        float hor[3];
        gradient_sample(ctx->sky_gradient, a1, a.peak);
        gradient_sample(ctx->sky_gradient, a2, hor);
        for (int c = 0; c < 3; ++c) a.horizon[c] = hor[c] * (1.0f - p->sunHorizonDarkness) + 0.0f * p->sunHorizonDarkness;   // :56
        a.shininess = p->sunShininess; a.strength = p->sunStrength;
        a.alpha1 = a.Ly < 0.1f ? (a.Ly + 1.0f) / (0.1f + 1.0f) : 1.0f;    // :70
        // by day (L.y >= 0.1) the sky sphere is drawn with alpha = max(1, .) = 1: Cs * 1 + Cd * 0 is exactly Cs, so whatever
        // the star layer left in the attachment cannot reach the result — skip sampling it (same bytes, a third less work)
        if (a.draw_sphere && a.alpha1 == 1.0f) a.stars = nullptr;
        a.omp = 1.0f - p->cloudProportion;                               // render-engine.cpp:297
        a.inv_gamma = 1.0f / 2.4f;
        a.inv_one_minus_omp = 1.0f / (1.0f - a.omp);                     // only read where srgb > omp, i.e. omp < 1
#ifdef WT_SKY_EXACT
        a.fast = 0;
#else
        a.fast = std::isfinite(a.inv_one_minus_omp) ? 1 : 0;
#endif
        a.overcast = p->overcastStrength;
        const float sy = std::fmin(std::fmax(sun[1], 0.0f), 1.0f);
        for (int c = 0; c < 3; ++c) a.fog[c] = sy * p->fogColourFarAtNoon[c];   // :295
        a.fog[3] = p->fogColourFarAtNoon[3];
        if (!ctx->d_sky_unsure) e = cudaMalloc(&ctx->d_sky_unsure, sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_sky_unsure, 0, sizeof(unsigned long long), stream);
        a.unsure_count = e == cudaSuccess ? ctx->d_sky_unsure : nullptr;
        a.fog_identity = !(std::fmin(std::fmax(a.fog[3], 0.0f), 1.0f) > 0.0f);   // As == 0 (or NaN): Cd comes back as the same byte
        if (!a.fog_identity) fog_table(a.fog, a.fog_lut);
        if (e == cudaSuccess && fresh) e = allocate_sky_atlas(ctx, n, &ctx->sky_array, &ctx->sky_tex);
        a.atlas = ctx->sky_surf;
        if (e == cudaSuccess) {
            // star / cloud cubemaps of the target's own power-of-two size are read texel for texel (cube_rgba8_aligned)
            const bool aligned = a.inv_n_exact != 0.0f && (!a.stars || a.stars_n == n) && (!a.clouds || a.clouds_n == n);
            const dim3 grid((n + 255) / 256, n, 6);
            if (aligned) sky_synth_kernel<true><<<grid, 256, 0, stream>>>(a);
            else sky_synth_kernel<false><<<grid, 256, 0, stream>>>(a);
            ctx->launches++;
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) { ctx->sky_n = n; ctx->sky_surface_ok = true; }
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(WT_ERR_CUDA, cudaGetErrorString(e));
    return WT_OK;
}

int wt_debug_fog_table(const float fog_rgba[4], uint32_t out[256]) {
    if (!fog_rgba || !out) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_debug_fog_table: NULL argument");
    unsigned lut[256];
    fog_table(fog_rgba, lut);
    for (int d = 0; d < 256; ++d) out[d] = lut[d];
    return WT_OK;
}

int wt_debug_sky_exact_texels(wt_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_debug_sky_exact_texels: NULL argument");
    *out = 0;
    if (!ctx->d_sky_unsure) return WT_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    unsigned long long v = 0;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(&v, ctx->d_sky_unsure, sizeof(v), cudaMemcpyDeviceToHost);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return set_thread_error(WT_ERR_CUDA, ctx->err.c_str()); }
    *out = v;
    return WT_OK;
}

int wt_get_skybox_rgba8(wt_ctx* ctx, uint8_t* out, int* n_out) {
    if (!ctx) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_get_skybox_rgba8: ctx is NULL");
    if (n_out) *n_out = ctx->sky_n;
    if (!out) return WT_OK;
    if (!ctx->d_sky) { ctx->err = "no skybox has been set or synthesized"; return set_thread_error(WT_ERR_NOT_READY, ctx->err.c_str()); }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, ctx->d_sky, 6 * (size_t)ctx->sky_n * ctx->sky_n * 4, cudaMemcpyDeviceToHost);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return set_thread_error(WT_ERR_CUDA, ctx->err.c_str()); }
    return WT_OK;
}

}  // extern "C"
