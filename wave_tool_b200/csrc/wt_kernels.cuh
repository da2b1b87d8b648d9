// Device code of the water-surface path (sm_100a).  See DESIGN.md for the derivation.
//
//   surface_warp_kernel   — the production kernel.  One WARP owns a strip of 62 output columns
//                           (+2 halo columns; lane l holds columns l and l+32 of the strip as the two
//                           halves of Blackwell packed-fp32 registers: FFMA2 / FMUL2) and marches
//                           over rows, R rows per step.  No CTA barrier: the 4-neighbour stencil is
//                           served from a warp-private shared-memory ring.  Fuses: bilinear grid
//                           placement (water-grid.vert:87-91), the Gerstner sum (:74-85), heightmap
//                           detail (:93-111), bounce (:147), finite-difference normal + view-facing
//                           flip (:164-190) and, optionally, the vertex-evaluable fragment terms
//                           (water-grid.frag:71-73,95-111).  Two phase modes:
//                             MODE 0  bit-faithful fp32 phase + one MUFU sin/cos pair per vertex-wave
//                             MODE 1  one MUFU seed per wave per R rows, the other rows by complex
//                                     rotation (the phase is affine in the row index along a column)
//                           With VARY also gl_Position, the height-0 NDC position, the heightmap sample and
//                           the yaw-only view-space normal (water-grid.vert:141,150-154,193).
//                           Work items (strip x row chunk) are static (one per CTA) or pulled from an
//                           atomic queue by a persistent grid (wt_launch.cuh).
//   surface_reference_order_kernel (wt_aux_kernels.cuh) — straight transliteration of the shader (5 full
//                           evaluations per vertex, libdevice sincosf, IEEE sqrt/div); used to cross-check
//                           the production kernel at full grid sizes and as the "naive GPU port" timing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wavetool_b200.h"

// build-time switches used by the helpers below (the kernel's own knobs are defined next to it)
#ifndef WT_HM_MAGIC_FLOOR
#define WT_HM_MAGIC_FLOOR 0     // 1: bit-faithful kernel floors the heightmap coordinates on the FP32 / integer pipes instead of the XU pipe (FRND, F2I).
                                // Measured 1.2 % SLOWER at cfg3 (0.703 vs 0.695 ms): the XU pipe is not what the per-vertex stage waits for
#endif
#ifndef WT_HM_SCALAR_BLEND
#define WT_HM_SCALAR_BLEND 1    // heightmap footprint blended per column in scalar form (no register moves to pair the two columns' texels up)
#endif

namespace wt {

// One wave, pre-folded on the host (wt_set_waves).  MODE 0 reads (Dx,Dz,w); MODE 1 is handed a
// table whose (Dx,Dz) already hold w*Dx, w*Dz.  cX = A*Q_i*Dx, cZ = A*Q_i*Dz.  For the production
// kernel the launcher overwrites `phi` with the per-frame phase offset fl(phi * t), which the shader
// recomputes per vertex per wave (water-grid.vert:78) — MODE 1 gets it reduced mod 2pi; the
// reference-order kernel keeps the raw phi (and Q_i in cX) and multiplies in the kernel.
struct __align__(16) WaveK {
    float Dx, Dz, w, phi;
    float cX, A, cZ, pad;
};

#ifndef WT_PAD_TABLE
#define WT_PAD_TABLE 1          // 1: one spare (zero) entry behind the tiers up to 256 waves, so the one-wave-ahead fetch of the wave loop needs no index clamp
#endif
template <int CAP>
struct WaveTable {
    static constexpr int PAD = (WT_PAD_TABLE && CAP <= 256) ? 1 : 0;   // the 1000-wave tier sits at the kernel-parameter size limit
    WaveK w[CAP + PAD];
};

struct EvalArgs;
#ifndef WT_BATCH_MAX_FRAMES
#define WT_BATCH_MAX_FRAMES 16
#endif
// frames per grouped launch: 16 x (680 + 256) bytes of kernel parameters (the limit is 32 KB).  Measured on 1024^2 x 8 waves:
// 8 frames per launch 8.48 us / frame, 16 frames 7.87 (9.84 -> 9.02 with digests)
constexpr int BATCH_MAX_FRAMES = WT_BATCH_MAX_FRAMES;
struct NoBatch { int unused; };
constexpr int STRIP_COLS = 62;   // output columns per warp: 64 held columns minus the two halo columns

struct EvalArgs {
    // grid corners, x and z only (y is discarded by water-grid.vert:138 `position.xz`)
    float blx, blz, brx, brz, tlx, tlz, trx, trz;
    float t, bounce, hm_disp, hm_scale;
    float camx, camy, camz, du;
    float sunx, suny, sunz, sun_shininess;
    float sun_strength;
    float one;          // 1.0f the compiler cannot see: x*one + y is an exact add ptxas cannot contract
    int n_waves;
    // heightmap, reference-order kernel: R channel as exact unorm floats (c/255.0f), row-major, GL row order
    const float* hm_r;
    int hm_w, hm_h, hm_pow2;
    int hm_magic_floor;   // production kernel, bit-faithful mode: every heightmap sample coordinate of this frame is far inside +-2^22 texels (hm_fetch2)
    // heightmap, production kernel: GL_MIRRORED_REPEAT unrolled into a (2w x 2h) table of bilinear
    // footprints (t00, t10, t01, t11) so that one 16-byte load serves a sample
    const float4* hm_quads;
    // sky cubemap: 6 faces of n*n RGBA8, GL face order (never NULL for the production kernel: a 1x1 black
    // cubemap stands in when no skybox was set)
    const uchar4* sky;
    int sky_n;
    // production kernel: the same texels as a 2-D texture atlas (block-linear CUDA array, point/gather access):
    // face f occupies the (n+2)^2 tile at column f/2 (the major axis), row f%2 (its sign), its edge texels replicated into the 1-texel
    // border, so a 2x2 gather anywhere inside a face IS the per-face CLAMP_TO_EDGE bilinear footprint
    unsigned long long sky_tex;
    // outputs (element 0 = (row_begin, col 0))
    float4* pos;
    float4* nrm;
    float4* shade;
    wt_checksum* cks;   // NULL unless WT_FLAG_CHECKSUM: where the finished digest goes (device memory or mapped host memory)
    // production kernel: a pre-armed accumulator record + a count of finished warps (wt_ctx ring); the last warp to
    // finish decodes the record into *cks and re-arms both, so a digest costs no extra launch
    wt_checksum* cks_acc;
    int* cks_done;
    // WT_FLAG_VARYINGS: the remaining vertex-stage outputs of water-grid.vert (:150-154, :141, :193)
    float4* clip;       // gl_Position = viewProjection * position
    float4* vary_a;     // (xyPositionNDCSpaceHeight0.x, .y, heightmap_colour.r, 0)
    float4* vary_b;     // (normalVecInViewSpaceOnlyYaw.xyz, 0)
    float vp[16];       // viewProjection, column-major
    float vy[12];       // upper-left 3 rows of viewMatOnlyYaw's first three columns: vy[3*c + r]
    int L, row_begin, row_end;
    int n_strips;       // warps across a row
    // Work items = (strip, chunk of rows).  The chunks come from the CANONICAL layout of the grid (wt_ctx.h: a function of
    // grid_len and the kernel instantiation only), restricted to the chunks that intersect [row_begin, row_end): n_seg
    // segments of equally sized chunks; chunk q of the launch lies in segment s with seg_prefix[s] <= q < seg_prefix[s+1]
    // and covers rows [seg_row0[s] + (q - seg_prefix[s]) * seg_rows[s], ... + seg_rows[s]) clipped at seg_end[s].  A chunk's
    // marching steps always start at its canonical first row - 1 (+ a multiple of R), whatever part of it the call
    // writes, so a vertex's rounding never depends on the shard or the device.
    int n_seg;
    int seg_prefix[17], seg_row0[16], seg_rows[16], seg_end[16];
    int dbg;            // experiment knob (WT_DBG): bit 0 = skip heights/epilogue
    // dynamic work distribution (NULL = static: item = blockIdx.x): queue[0] = next item, queue[1] = warps done
    int* queue;
    int n_items;
    // batched launch (BATCH instantiations: several small frames in ONE persistent launch): item = frame * items_per_frame + the
    // frame's own work item; the frames' own EvalArgs and phase offsets travel in a third kernel parameter (BatchTable)
    int* batch_done;             // per frame of the group: finished work items (the last one decodes the frame's digest record and re-arms the counter)
    int items_per_frame;
    unsigned long long* trace;   // experiment knob (WT_TRACE): per work item {smid, item, t_start, t_end} in globaltimer ns
};

// the per-frame arguments of a grouped launch, passed by value as a kernel parameter (constant bank, indexed by the frame)
struct BatchTable {
    EvalArgs fa[BATCH_MAX_FRAMES];
    float phi[BATCH_MAX_FRAMES][64];   // fl(phi t) per wave (reduced mod 2 pi for the recurrence kernel)
};
template <bool BATCH> struct BatchParam { typedef NoBatch type; };
template <> struct BatchParam<true> { typedef BatchTable type; };
template <bool BATCH, class BT>
__device__ __forceinline__ const EvalArgs& frame_args(const EvalArgs& a, const BT& bt, int frame) {
    if constexpr (BATCH) return bt.fa[frame];
    else return a;
}
template <bool BATCH, class BT>
__device__ __forceinline__ float frame_phi(const BT& bt, int frame, int wave, float dflt) {
    if constexpr (BATCH) return bt.phi[frame][wave];
    else return dflt;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }   // never contracted
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

// GLSL/GLM mix(x, y, a) = x*(1-a) + y*a with every operation rounded (water-grid.vert:87-91)
__device__ __forceinline__ float mix_rn(float x, float y, float a) {
    return add_rn(mul_rn(x, sub_rn(1.0f, a)), mul_rn(y, a));
}

// sin and cos of a phase of any magnitude met in practice (|k| < 2^22 * 2pi):
// Cody-Waite reduction k - n*2pi with the product formed exactly inside the FMA (2pi split into
// fp32 hi + lo), then the MUFU approximations on [-pi, pi] where their absolute error is ~2^-21.
__device__ __forceinline__ void sincos_mufu(float k, float& s, float& c) {
    const float INV_2PI = 0.15915494309189535f;
    const float MAGIC = 12582912.0f;             // 1.5 * 2^23: adding it rounds to the nearest integer
    const float TWO_PI_HI = 6.2831854820251465f; // fl32(2pi)
    const float TWO_PI_LO = -1.7484556000744883e-07f;   // 2pi - TWO_PI_HI
    float n = __fmaf_rn(k, INV_2PI, MAGIC) - MAGIC;
    float r = __fmaf_rn(n, -TWO_PI_HI, k);
    r = __fmaf_rn(n, -TWO_PI_LO, r);
    s = __sinf(r);
    c = __cosf(r);
}

// exact c / 255.0f for c in [0,255]: reciprocal multiply + one FMA correction step (verified
// exhaustively in tests/test_host_math.py against IEEE division)
__device__ __forceinline__ float unorm8(unsigned c) {
    const float R = 1.0f / 255.0f;
    float f = (float)c;
    float q = f * R;
    float rem = __fmaf_rn(-q, 255.0f, f);
    return __fmaf_rn(rem, R, q);
}

// Byte and floor conversions kept off the XU pipe (I2F / F2I / FRND issue at a quarter of the FP32 rate):
//  * byte k of a word -> float: PRMT builds the bit pattern of 2^23 + c, one add removes the 2^23 (exact);
//  * floor(x) for |x| < 2^22: x + 1.5 * 2^23 is positive, so adding with round-toward-zero floors it (exact).
__device__ __forceinline__ float byte_f(unsigned word, unsigned k) { return __fadd_rn(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | k)), -8388608.0f); }
__device__ __forceinline__ float unorm8f(float f) {   // exact c / 255.0f for integer-valued f in [0,255] (see unorm8)
    const float R = 1.0f / 255.0f;
    const float q = __fmul_rn(f, R);
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, f), R, q);
}
__device__ __forceinline__ float floor_small(float x, int& as_int) {
    const float MAGIC = 12582912.0f;
    const float t = __fadd_rz(x, MAGIC);
    as_int = __float_as_int(t) - 0x4B400000;
    return __fsub_rn(t, MAGIC);
}
// x / c for a divisor c that is the same for every vertex of a launch, given rc = fl(1 / c): the quotient estimate x rc is
// corrected with its exact FMA residual — correctly rounded like IEEE division (Markstein 1990; what __fdiv_rn's own fast
// path does after computing the reciprocal), three FP32-pipe operations instead of a reciprocal on the XU pipe plus ~8.
// Needs finite, normal x / c (the callers' operands are screen coordinates, depths and thresholds in [0, 1e5]).
__device__ __forceinline__ float div_by_uniform(float x, float c, float rc) {
    const float q = __fmul_rn(x, rc);
    return __fmaf_rn(__fmaf_rn(-q, c, x), rc, q);
}

// GL_MIRRORED_REPEAT on a texel index (GL 4.1 core §3.8.9): (size-1) - mirror((i mod 2 size) - size)
__device__ __forceinline__ int wrap_mirror(int i, int size, int pow2) {
    int two = 2 * size;
    int m;
    if (pow2) {
        m = i & (two - 1);
    } else {
        m = i % two;
        if (m < 0) m += two;
    }
    int a = m - size;
    int mir = a >= 0 ? a : ~a;   // -(1 + a)
    return (size - 1) - mir;
}

// texture(heightmap, st).r — GL_LINEAR at LOD 0 with exact fp32 weights; the operation order is the
// oracle's (oracle/gl_sampler.h): weights first, four weighted texels summed left to right.
__device__ __forceinline__ float heightmap_r(const EvalArgs& a, float s, float t) {
    float um = sub_rn(mul_rn(s, (float)a.hm_w), 0.5f);
    float vm = sub_rn(mul_rn(t, (float)a.hm_h), 0.5f);
    float fu = floorf(um), fv = floorf(vm);
    float al = sub_rn(um, fu), be = sub_rn(vm, fv);
    int iu = (int)fu, iv = (int)fv;
    int i0 = wrap_mirror(iu, a.hm_w, a.hm_pow2), i1 = wrap_mirror(iu + 1, a.hm_w, a.hm_pow2);
    int j0 = wrap_mirror(iv, a.hm_h, a.hm_pow2), j1 = wrap_mirror(iv + 1, a.hm_h, a.hm_pow2);
    const float* r0 = a.hm_r + (size_t)j0 * a.hm_w;
    const float* r1 = a.hm_r + (size_t)j1 * a.hm_w;
    float t00 = __ldg(r0 + i0), t10 = __ldg(r0 + i1), t01 = __ldg(r1 + i0), t11 = __ldg(r1 + i1);
    float oma = sub_rn(1.0f, al), omb = sub_rn(1.0f, be);
    float w00 = mul_rn(oma, omb), w10 = mul_rn(al, omb), w01 = mul_rn(oma, be), w11 = mul_rn(al, be);
    float acc = mul_rn(w00, t00);
    acc = add_rn(acc, mul_rn(w10, t10));
    acc = add_rn(acc, mul_rn(w01, t01));
    acc = add_rn(acc, mul_rn(w11, t11));
    return acc;
}

// water-grid.vert:141-147 — heightmap bump at the DISPLACED position, then the bounce
__device__ __forceinline__ float finish_height(const EvalArgs& a, float px, float py, float pz) {
    float r = heightmap_r(a, mul_rn(a.hm_scale, px), mul_rn(a.hm_scale, -pz));
    float bump = mul_rn(a.hm_disp, mul_rn(2.0f, sub_rn(r, 0.5f)));
    return add_rn(add_rn(py, bump), a.bounce);
}

// texture(skybox, R).rgb — GL 4.1 cube map face selection, per-face CLAMP_TO_EDGE bilinear
__device__ __forceinline__ float3 sky_rgb(const EvalArgs& a, float rx, float ry, float rz) {
    float ax = fabsf(rx), ay = fabsf(ry), az = fabsf(rz);
    int face;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) {
        ma = ax;
        if (rx >= 0.0f) { face = 0; sc = -rz; tc = -ry; } else { face = 1; sc = rz; tc = -ry; }
    } else if (ay >= az) {
        ma = ay;
        if (ry >= 0.0f) { face = 2; sc = rx; tc = rz; } else { face = 3; sc = rx; tc = -rz; }
    } else {
        ma = az;
        if (rz >= 0.0f) { face = 4; sc = rx; tc = -ry; } else { face = 5; sc = -rx; tc = -ry; }
    }
    float s = 0.5f * (__fdiv_rn(sc, ma) + 1.0f);
    float t = 0.5f * (__fdiv_rn(tc, ma) + 1.0f);
    const int n = a.sky_n;
    float um = sub_rn(mul_rn(s, (float)n), 0.5f), vm = sub_rn(mul_rn(t, (float)n), 0.5f);
    float fu = floorf(um), fv = floorf(vm);
    float al = sub_rn(um, fu), be = sub_rn(vm, fv);
    int iu = (int)fu, iv = (int)fv;
    int i0 = min(max(iu, 0), n - 1), i1 = min(max(iu + 1, 0), n - 1);
    int j0 = min(max(iv, 0), n - 1), j1 = min(max(iv + 1, 0), n - 1);
    const uchar4* f = a.sky + (size_t)face * n * n;
    uchar4 t00 = __ldg(f + (size_t)j0 * n + i0), t10 = __ldg(f + (size_t)j0 * n + i1);
    uchar4 t01 = __ldg(f + (size_t)j1 * n + i0), t11 = __ldg(f + (size_t)j1 * n + i1);
    float oma = sub_rn(1.0f, al), omb = sub_rn(1.0f, be);
    float w00 = mul_rn(oma, omb), w10 = mul_rn(al, omb), w01 = mul_rn(oma, be), w11 = mul_rn(al, be);
    float3 o;
    o.x = add_rn(add_rn(add_rn(mul_rn(w00, unorm8(t00.x)), mul_rn(w10, unorm8(t10.x))), mul_rn(w01, unorm8(t01.x))), mul_rn(w11, unorm8(t11.x)));
    o.y = add_rn(add_rn(add_rn(mul_rn(w00, unorm8(t00.y)), mul_rn(w10, unorm8(t10.y))), mul_rn(w01, unorm8(t01.y))), mul_rn(w11, unorm8(t11.y)));
    o.z = add_rn(add_rn(add_rn(mul_rn(w00, unorm8(t00.z)), mul_rn(w10, unorm8(t10.z))), mul_rn(w01, unorm8(t01.z))), mul_rn(w11, unorm8(t11.z)));
    return o;
}

struct P3 { float x, y, z; };

// order-preserving float <-> uint key for atomicMin/atomicMax
__device__ __forceinline__ unsigned float_key(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_float(unsigned k) {
    unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

struct ChecksumAcc {
    unsigned long long pos = 0, nrm = 0, shd = 0, count = 0;
    float ymin = 3.402823466e+38f, ymax = -3.402823466e+38f;   // +-FLT_MAX; converted to ordered keys in flush()
    // sums and y range of one vertex; the vertex COUNT is added per vertex by add(), per work item by the production kernel
    __device__ __forceinline__ void add_sums(const float4& p, const float4& n) {
        pos += (unsigned long long)__float_as_uint(p.x) + __float_as_uint(p.y) + __float_as_uint(p.z);
        nrm += (unsigned long long)__float_as_uint(n.x) + __float_as_uint(n.y) + __float_as_uint(n.z) + __float_as_uint(n.w);
        ymin = fminf(ymin, p.y);
        ymax = fmaxf(ymax, p.y);
    }
    __device__ __forceinline__ void add(const float4& p, const float4& n) {
        add_sums(p, n);
        count += 1;
    }
    // Both vertices of a lane's output row, branch-free: lanes 0 and 31 hold one halo column each (never stored, never counted),
    // and a branch around the per-column sums made every warp diverge twice per row (2-3 % of the kernel's time).  Each column's
    // words are summed first (at most 2^34), masked as a whole, then the two columns go into the running sum together.
    __device__ __forceinline__ void add_pair(const float4& pa, const float4& na, const float4& sa, bool keep_a,
                                             const float4& pb, const float4& nb, const float4& sb, bool keep_b, bool shade) {
        const unsigned long long ma = 0ull - (unsigned long long)keep_a, mb = 0ull - (unsigned long long)keep_b;
        auto w = [](float f) { return (unsigned long long)__float_as_uint(f); };
        pos += ((w(pa.x) + w(pa.y) + w(pa.z)) & ma) + ((w(pb.x) + w(pb.y) + w(pb.z)) & mb);
        nrm += ((w(na.x) + w(na.y) + w(na.z) + w(na.w)) & ma) + ((w(nb.x) + w(nb.y) + w(nb.z) + w(nb.w)) & mb);
        if (shade) shd += ((w(sa.x) + w(sa.y) + w(sa.z) + w(sa.w)) & ma) + ((w(sb.x) + w(sb.y) + w(sb.z) + w(sb.w)) & mb);
        const float ya = keep_a ? pa.y : ymin, yb = keep_b ? pb.y : ymin;     // the running value itself is neutral
        ymin = fminf(ymin, fminf(ya, yb));
        const float za = keep_a ? pa.y : ymax, zb = keep_b ? pb.y : ymax;
        ymax = fmaxf(ymax, fmaxf(za, zb));
    }
    __device__ __forceinline__ void add_shade(const float4& s) {
        shd += (unsigned long long)__float_as_uint(s.x) + __float_as_uint(s.y) + __float_as_uint(s.z) + __float_as_uint(s.w);
    }
    // warp-reduce, then one set of atomics per warp.  While accumulating, the min_y/max_y words of
    // the record hold order-preserving keys; checksum_finalize_kernel decodes them.
    __device__ __forceinline__ void flush(wt_checksum* out) {
        unsigned kmin = float_key(ymin), kmax = float_key(ymax);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pos += __shfl_xor_sync(0xffffffffu, pos, o);
            nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
            shd += __shfl_xor_sync(0xffffffffu, shd, o);
            count += __shfl_xor_sync(0xffffffffu, count, o);
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if ((threadIdx.x & 31) == 0 && count) {
            atomicAdd((unsigned long long*)&out->position_bits, pos);
            atomicAdd((unsigned long long*)&out->normal_bits, nrm);
            atomicAdd((unsigned long long*)&out->shade_bits, shd);
            atomicAdd((unsigned long long*)&out->vertices, count);
            atomicMin((unsigned*)&out->min_y, kmin);
            atomicMax((unsigned*)&out->max_y, kmax);
        }
    }
};

// single-instruction MUFU approximations (flush-to-zero forms: no denormal fix-up code around them)
__device__ __forceinline__ float rsq_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ------------------------------------------------------------------------------------------------
// Packed fp32 (Blackwell FFMA2 / FMUL2): both halves are rounded like the scalar instruction.
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it does not for the scalar forms),
// so wherever the reference's rounding ORDER matters the add is written a*one + b with a run-time
// one == 1.0f (EvalArgs::one): bit-identical to the add, one instruction, and not contractible.
// ------------------------------------------------------------------------------------------------
typedef float2 f2;
__device__ __forceinline__ f2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }                     // only where neither operand is a product
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ f2 add2x(f2 a, f2 b, float one) { return __ffma2_rn(a, bc(one), b); }    // fl(a + b), exact order
__device__ __forceinline__ f2 sub2x(f2 a, f2 b, float one) { return __ffma2_rn(b, bc(-one), a); }   // fl(a - b), exact order

// sin and cos of both halves: the Cody-Waite reduction of sincos_mufu, packed
__device__ __forceinline__ void sincos2(f2 k, f2& s, f2& c) {
    const float INV_2PI = 0.15915494309189535f;
    const float MAGIC = 12582912.0f;
    const float TWO_PI_HI = 6.2831854820251465f;
    const float TWO_PI_LO = -1.7484556000744883e-07f;
#ifdef WT_NO_CW
    f2 r = k; (void)INV_2PI; (void)MAGIC; (void)TWO_PI_HI; (void)TWO_PI_LO;
#else
    f2 n = add2(fma2(k, bc(INV_2PI), bc(MAGIC)), bc(-MAGIC));
    f2 r = fma2(n, bc(-TWO_PI_HI), k);
    r = fma2(n, bc(-TWO_PI_LO), r);
#endif
    s = make_float2(__sinf(r.x), __sinf(r.y));
    c = make_float2(__cosf(r.x), __cosf(r.y));
}

// texture(heightmap, scale*(x,-z)).r for the two columns of a lane, from the footprint table; then
// water-grid.vert:141-147.  Same fp32 operations in the same order as heightmap_r/finish_height.
__device__ __forceinline__ int wrap_period_any(int i, int period) {
    int m = i % period;
    return m < 0 ? m + period : m;
}
// non-power-of-two heightmaps: the four footprint coordinates modulo (2w, 2h); kept out of line so the
// integer-division sequences stay out of the hot instruction stream of the usual power-of-two case
static __device__ __noinline__ int4 wrap_period4_any(int4 v, int w2, int h2) {
    return make_int4(wrap_period_any(v.x, w2), wrap_period_any(v.y, w2), wrap_period_any(v.z, h2), wrap_period_any(v.w, h2));
}
// Split in two so that the table loads of several rows are in flight together: hm_fetch2 computes the
// bilinear weights and issues the two 16-byte footprint loads, hm_finish2 blends and applies :141-147.
struct HmTap { f2 al, be; float4 qa, qb; };
template <bool POW2, bool MAGIC_FLOOR>
__device__ __forceinline__ HmTap hm_fetch2(const EvalArgs& a, f2 px, f2 pz) {
    const float one = a.one;
    f2 s = mul2(bc(a.hm_scale), px);
    f2 t = mul2(bc(-a.hm_scale), pz);
    f2 um = sub2x(mul2(s, bc((float)a.hm_w)), bc(0.5f), one);
    f2 vm = sub2x(mul2(t, bc((float)a.hm_h)), bc(0.5f), one);
    f2 fu, fv;
    int4 ij;
    if (MAGIC_FLOOR && (WT_HM_MAGIC_FLOOR == 2 || a.hm_magic_floor)) {   // (2: experiment, no fallback compiled in)
        // floor() and the float -> int conversion without the XU pipe (FRND and F2I issue at a quarter of the FP32 rate, and
        // the bit-faithful kernel is bound by that pipe): for |x| < 2^22 the sum x + 1.5 * 2^23 is a positive number with a
        // unit in the last place of 1, so adding with round-toward-zero floors x exactly; the integer sits in the mantissa.
        // The launcher clears hm_magic_floor when the sample coordinates of a frame could leave that range.
        const float M = 12582912.0f;
        const f2 tu = __fadd2_rz(um, bc(M)), tv = __fadd2_rz(vm, bc(M));
        fu = add2(tu, bc(-M)); fv = add2(tv, bc(-M));
        ij = make_int4(__float_as_int(tu.x) - 0x4B400000, __float_as_int(tu.y) - 0x4B400000, __float_as_int(tv.x) - 0x4B400000, __float_as_int(tv.y) - 0x4B400000);
    } else {
        fu = make_float2(floorf(um.x), floorf(um.y)); fv = make_float2(floorf(vm.x), floorf(vm.y));
        ij = make_int4((int)fu.x, (int)fu.y, (int)fv.x, (int)fv.y);
    }
    HmTap h;
    h.al = sub2(um, fu); h.be = sub2(vm, fv);
    const int w2 = 2 * a.hm_w, h2 = 2 * a.hm_h;
    if (POW2) {
        ij.x &= w2 - 1; ij.y &= w2 - 1; ij.z &= h2 - 1; ij.w &= h2 - 1;
    } else {
        ij = wrap_period4_any(ij, w2, h2);
    }
    h.qa = __ldg(a.hm_quads + (unsigned)(ij.z * w2 + ij.x));   // (2w)(2h) <= 2^30 entries (wt_set_heightmap_rgba8)
    h.qb = __ldg(a.hm_quads + (unsigned)(ij.w * w2 + ij.y));
    return h;
}
__device__ __forceinline__ f2 hm_finish2(const EvalArgs& a, const HmTap& h, f2 py, f2& texel_r) {
    const float one = a.one;
    f2 oma = sub2(bc(1.0f), h.al), omb = sub2(bc(1.0f), h.be);
    f2 w00 = mul2(oma, omb), w10 = mul2(h.al, omb), w01 = mul2(oma, h.be), w11 = mul2(h.al, h.be);
#if WT_HM_SCALAR_BLEND
    // The two columns' texels arrive in two separate 16-byte loads: blending them as scalars (same operations, same
    // order, same roundings) costs the FMA pipe exactly what the packed form costs and saves the eight register moves
    // that would pair (qa.k, qb.k) up
    float aA = mul_rn(w00.x, h.qa.x), aB = mul_rn(w00.y, h.qb.x);
    aA = add_rn(mul_rn(w10.x, h.qa.y), aA); aB = add_rn(mul_rn(w10.y, h.qb.y), aB);
    aA = add_rn(mul_rn(w01.x, h.qa.z), aA); aB = add_rn(mul_rn(w01.y, h.qb.z), aB);
    aA = add_rn(mul_rn(w11.x, h.qa.w), aA); aB = add_rn(mul_rn(w11.y, h.qb.w), aB);
    f2 acc = make_float2(aA, aB);
    (void)one;
#else
    f2 acc = mul2(w00, make_float2(h.qa.x, h.qb.x));
    acc = add2x(mul2(w10, make_float2(h.qa.y, h.qb.y)), acc, one);
    acc = add2x(mul2(w01, make_float2(h.qa.z, h.qb.z)), acc, one);
    acc = add2x(mul2(w11, make_float2(h.qa.w, h.qb.w)), acc, one);
#endif
    texel_r = acc;   // heightmap_colour.r (:141)
    f2 bump = mul2(bc(a.hm_disp), mul2(bc(2.0f), sub2(acc, bc(0.5f))));
    return add2(add2x(bump, py, one), bc(a.bounce));
}

// texture(skybox, R).rgb for the production kernel: GL 4.1 cube map face selection, per-face
// CLAMP_TO_EDGE bilinear.  (Reciprocal and weights are not bit-ordered like the oracle: the sky colour
// has its own, absolute tolerance.)
// texture(skybox, R).rgb for the production kernel: GL 4.1 cube map face selection, per-face CLAMP_TO_EDGE
// bilinear.  The three channels of the 2x2 footprint come from three texture-gather instructions (tld4.r/g/b)
// on the padded face atlas; the bilinear weights are exact fp32, computed here (the texture unit's own filter
// would quantise them to 8 bits).  (Reciprocal and weights are not bit-ordered like the oracle: the sky
// colour has its own, absolute tolerance.)
struct SkyTap { float4 r, g, b; float al, be; };   // per channel: texels (i0,j1), (i1,j1), (i1,j0), (i0,j0) as c/255
__device__ __forceinline__ float4 tld4(unsigned long long tex, float x, float y, int comp) {
    float4 v;
    if (comp == 0) asm volatile("tld4.r.2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6}];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(tex), "f"(x), "f"(y));
    else if (comp == 1) asm volatile("tld4.g.2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6}];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(tex), "f"(x), "f"(y));
    else asm volatile("tld4.b.2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6}];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(tex), "f"(x), "f"(y));
    return v;
}
__device__ __forceinline__ SkyTap sky_fetch(const EvalArgs& a, float rx, float ry, float rz) {
    // Face selection without branches.  ma = max(|rx|,|ry|,|rz|); ties resolve x before y before z like the
    // GL 4.1 table the oracle follows.
    const float ax = fabsf(rx), ay = fabsf(ry), az = fabsf(rz);
    const bool isx = ax >= ay && ax >= az;
    const bool isy = !isx && ay >= az;
    const float ma = fmaxf(ax, fmaxf(ay, az));
    const float major = isx ? rx : (isy ? ry : rz);
    const bool pos = major >= 0.0f;
    // sc: +x:-rz  -x:+rz  +-y:+rx  +z:+rx  -z:-rx        tc: +y:+rz  -y:-rz  else:-ry
    const float sc_abs = isx ? rz : rx;
    const bool sc_neg = isx ? pos : (isy ? false : !pos);
    const float sc = sc_neg ? -sc_abs : sc_abs;
    const float tc_abs = isy ? rz : ry;
    const bool tc_neg = isy ? !pos : true;
    const float tc = tc_neg ? -tc_abs : tc_abs;
    const int n = a.sky_n;
    const float hn = 0.5f * (float)n;
    const float inv = rcp_approx(ma);
    // um = (0.5 (sc/|ma| + 1)) n - 0.5 in [-0.5, n - 0.5]; floor/frac without the XU pipe: adding 1.5*2^23 rounds
    // to the nearest integer.  Rounding um - 0.5 instead of flooring um only differs on exact integers, where the
    // neighbouring footprint with weight 1 yields the same sample.
    const float MAGIC = 12582912.0f;
    const float um = fmaf(sc * inv, hn, hn - 0.5f), vm = fmaf(tc * inv, hn, hn - 0.5f);
    const float fu = ((um - 0.5f) + MAGIC) - MAGIC, fv = ((vm - 0.5f) + MAGIC) - MAGIC;   // floor(um), floor(vm) in {-1 .. n-1}
    SkyTap t;
    t.al = um - fu; t.be = vm - fv;
    // atlas tile origin of the face (+1 for the border) and the footprint's centre: texels (fu, fu+1) of the face
    // are atlas texels X0+1+fu, X0+2+fu, whose shared edge is at X0 + fu + 2 — the gather coordinate
    const float tile = (float)(n + 2);
    const float X0 = isx ? 0.0f : (isy ? tile : tile + tile);   // tile column = major axis
    const float Y0 = pos ? 0.0f : tile;                          // tile row = sign of the major component
    const float gx = X0 + fu + 2.0f, gy = Y0 + fv + 2.0f;
    t.r = tld4(a.sky_tex, gx, gy, 0);
    t.g = tld4(a.sky_tex, gx, gy, 1);
    t.b = tld4(a.sky_tex, gx, gy, 2);
    return t;
}
__device__ __forceinline__ float3 sky_blend(const SkyTap& t) {
    // gather order: x = (i0, j1), y = (i1, j1), z = (i1, j0), w = (i0, j0)
    const float oma = 1.0f - t.al, omb = 1.0f - t.be;
    const f2 wxy = mul2(make_float2(oma, t.al), bc(t.be));     // (w01, w11)
    const f2 wzw = mul2(make_float2(t.al, oma), bc(omb));      // (w10, w00)
    const f2 sr = fma2(wxy, make_float2(t.r.x, t.r.y), mul2(wzw, make_float2(t.r.z, t.r.w)));
    const f2 sg = fma2(wxy, make_float2(t.g.x, t.g.y), mul2(wzw, make_float2(t.g.z, t.g.w)));
    const f2 sb = fma2(wxy, make_float2(t.b.x, t.b.y), mul2(wzw, make_float2(t.b.z, t.b.w)));
    return make_float3(sr.x + sr.y, sg.x + sg.y, sb.x + sb.y);
}

// One output row in flight between the two halves of the output stage (see surface_warp_kernel)
struct EpiRow {
    f2 nx, ny, nz, fres, sp;
    SkyTap skyA, skyB;
    int row, slot;   // slot: ring row of the vertex itself (its position is re-read when the row is stored)
};

// ------------------------------------------------------------------------------------------------
// Production kernel.
//   One warp = one work item = (strip of STRIP_COLS output columns) x (chunk of rows_per_chunk rows).
//   Lane l holds the strip's columns  l  and  l + 32  (strip column 0 and 63 are the halo columns
//   whose positions feed the stencil but are not written) as the .x/.y halves of packed registers.
//   Per step of R rows:  grid placement -> sum over waves (constants broadcast from the constant
//   bank) -> heightmap + bounce -> final positions into a warp-private shared-memory ring of R+2 rows
//   (laid out so that one 8-byte load yields the packed left / centre / right neighbours of both
//   columns) -> for the R rows whose up/down neighbours are now known: normal, flip, shading,
//   coalesced float4 stores.  Only __syncwarp(); warps never wait for each other.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// dynamic work distribution: lane 0 takes the next item off the queue, the warp shares it
__device__ __forceinline__ int next_item(int* queue, int lane) {
    int v = 0;
    if (lane == 0) v = atomicAdd(queue, 1);
    return __shfl_sync(0xffffffffu, v, 0);
}

// build-time tuning knobs (tools/build_variant.py compiles alternative values side by side)
#ifndef WT_WAVE_UNROLL
#define WT_WAVE_UNROLL 8        // wave-loop unroll factor of MODE 1 (measured: 1 -> 0.447 ms, 4 -> 0.443, 8 -> 0.438, 16 -> 0.506 at cfg3)
#endif
#ifndef WT_WAVE_UNROLL_EXACT
#define WT_WAVE_UNROLL_EXACT 2   // MODE 0 runs at its 80-register cap: 2 -> -0.6 %, 4 and 8 slower
#endif
#ifndef WT_LB_RECUR
#define WT_LB_RECUR 16          // __launch_bounds__ min CTAs (= warps) per SM, MODE 1
#endif
#ifndef WT_LB_RECUR_BIG
#define WT_LB_RECUR_BIG 12      // ... MODE 1 with more than 8 rows per step (the > 64-wave tiers): 168 registers, 3 warps per scheduler
#endif
#ifndef WT_LB_EXACT
#define WT_LB_EXACT 24
#endif
#ifndef WT_HM_BATCH
#define WT_HM_BATCH 4           // rows whose heightmap footprint loads are in flight together
#endif
#ifndef WT_EPI_UNROLL_EXACT
#define WT_EPI_UNROLL_EXACT 1
#endif
#ifndef WT_EPI_UNROLL
#define WT_EPI_UNROLL 2         // rows of the output stage in flight per warp (MODE 1)
#endif
#ifndef WT_SEAMLESS_RING
#define WT_SEAMLESS_RING 0      // 1: no seam entries in the ring; lane 0 / lane 31 fetch their cross-half neighbours through per-lane offsets (4 scalar
                                // loads per plane instead of 2 pair loads, no divergent seam writes)
#endif
#ifndef WT_CHEB
#define WT_CHEB 1               // 1: rows 2.. of a MODE 1 step by the three-term recurrence, 0: by complex rotations (cfg3: 0.405 vs 0.438 ms)
#endif
#ifndef WT_SYM
#define WT_SYM 0                // 1: symmetric-pair form of the MODE 1 wave loop (needs an odd WT_R_RECUR).  EXPERIMENT ONLY: 0.396 ms at
                                // cfg3, closer to the fp64 evaluation than the reference is, but it cannot follow the reference's own
                                // fp32 accumulation order and misses the parity bar where that noise exceeds it (profiles/README.md)
#endif
#define WT_STR(x) #x
#define WT_UNROLL(n) _Pragma(WT_STR(unroll n))
#define WT_UNROLL_IF(c, a, b) _Pragma(WT_STR(unroll ((c) ? (a) : (b))))

constexpr int WARPS_PER_CTA = 1;   // one warp per CTA: every loop bound derives from blockIdx alone, so ptxas keeps the wave constants in uniform registers
constexpr int RING_COLS = 34;   // entries per ring row: [0] = (-, A of lane 31), [1 + l] = (A_l, B_l), [33] = (B of lane 0, -)

template <int R, bool VARY = false>
struct RingLayout {
    static constexpr int ROWS = R + 2;
    static constexpr int PLANE = ROWS * RING_COLS;                   // float2 entries per component
    static constexpr int PLANES = VARY ? 4 : 3;                      // x, y, z (+ the heightmap sample with WT_FLAG_VARYINGS)
    static constexpr int SEAM_SCRATCH = 16;                          // float2 entries behind the planes: one float per lane (see the seam stores)
    static constexpr int BYTES_PER_WARP = (PLANES * PLANE + SEAM_SCRATCH) * (int)sizeof(float2);
    // WT_FLAG_CHECKSUM in the bit-faithful kernel (80 registers, every live value counts): the warp's running digest lives
    // behind the ring — [4][32] 64-bit words (position, normal, shade sums, (min y, max y); one column per lane)
    static constexpr int CKS_BYTES_PER_WARP = 4 * 32 * (int)sizeof(unsigned long long);
};

template <int CAP, int R, int MODE, bool SHADE, bool CKS, bool HMP2, bool VARY, bool BATCH = false>
__global__ void __launch_bounds__(32, MODE == 1 ? (R > 8 ? WT_LB_RECUR_BIG : WT_LB_RECUR) : WT_LB_EXACT)
surface_warp_kernel(const __grid_constant__ EvalArgs a, const __grid_constant__ WaveTable<CAP> wt, const __grid_constant__ typename BatchParam<BATCH>::type bt) {
    extern __shared__ float2 ring_all[];
    using RL = RingLayout<R, VARY>;
    const int lane = threadIdx.x;
    float2* const ring = ring_all;
    const bool dyn = a.queue != nullptr;
    int item = blockIdx.x;
    if (dyn) item = next_item(a.queue, lane);
    // The digest accumulators.  Recurrence kernel (128+ registers): in registers.  Bit-faithful kernel: in shared memory, updated
    // row by row — eight more live registers across the wave loop and the output stage cost it 4.7 % (measured with the sums
    // themselves stubbed out), the read-modify-write of this lane's own column of the record costs less (profiles/README.md).
    // The vertex count stays in a register: it is touched once per work item (measured: 3.6 % with it in a register, 4.8 % in the record).
    constexpr bool CKS_SMEM = CKS && MODE == 0;
    ChecksumAcc acc;
    unsigned long long* const cks_slot = reinterpret_cast<unsigned long long*>(ring_all + RL::PLANES * RL::PLANE + RL::SEAM_SCRATCH) + lane;
    auto cks_reset = [&]() {
        cks_slot[0] = 0; cks_slot[32] = 0; cks_slot[64] = 0;
        *reinterpret_cast<float2*>(cks_slot + 96) = make_float2(3.402823466e+38f, -3.402823466e+38f);
    };
    auto cks_collect = [&]() {
        acc.pos = cks_slot[0]; acc.nrm = cks_slot[32]; acc.shd = cks_slot[64];
        const float2 yr = *reinterpret_cast<const float2*>(cks_slot + 96);
        acc.ymin = yr.x; acc.ymax = yr.y;
    };
    if (CKS_SMEM) cks_reset();
  while (item < a.n_items) {
    unsigned long long t_start = 0;
    if (a.trace) t_start = global_ns();
    int fitem = item, cur_frame = 0;
    if (BATCH) {
        cur_frame = item / a.items_per_frame;
        fitem = item - cur_frame * a.items_per_frame;
    }
    const EvalArgs& A = frame_args<BATCH>(a, bt, cur_frame);
    const int strip = fitem % A.n_strips, chunk = fitem / A.n_strips;
    const int strip0 = strip * STRIP_COLS;
    int seg = 0;
    while (seg + 1 < A.n_seg && chunk >= A.seg_prefix[seg + 1]) ++seg;
    // canonical rows of the chunk [c0, c1); the rows this call writes [r0, r1)
    const int c0 = A.seg_row0[seg] + (chunk - A.seg_prefix[seg]) * A.seg_rows[seg];
    const int c1 = min(c0 + A.seg_rows[seg], A.seg_end[seg]);
    const int r0 = max(c0, A.row_begin), r1 = min(c1, A.row_end);
    const int colA = strip0 - 1 + lane, colB = colA + 32;
    const bool outA = lane >= 1 && colA < A.L;
    const bool outB = lane <= 30 && colB < A.L;
    const float one = A.one;

    int slot_new = 2;   // ring row that receives the first row of the next step; rows -2/-1 of the first step are never read

    // marching steps sit on the chunk's own lattice (first rows c0 - 1 + k R); a call that writes only part of the
    // chunk starts at the step that holds row r0 - 1 — every step is seeded on its own, so skipping steps changes nothing
    for (int row_a = c0 - 1 + ((r0 - c0) / R) * R; row_a <= r1; row_a += R) {
        // water-grid.vert:123-125,87-91: u = col*du is fixed per lane, so the two inner mixes are too.  They
        // are re-derived every step (a dozen instructions) instead of living in 12 registers across the wave
        // loop; the empty asm keeps the compiler from hoisting them back out.
        int col_a = colA;
        asm volatile("" : "+r"(col_a));
        const f2 u = make_float2(mul_rn((float)col_a, A.du), mul_rn((float)(col_a + 32), A.du));
        const f2 omu = sub2(bc(1.0f), u);
        const f2 m1x = add2x(mul2(bc(A.blx), omu), mul2(bc(A.brx), u), one), m1z = add2x(mul2(bc(A.blz), omu), mul2(bc(A.brz), u), one);
        const f2 m2x = add2x(mul2(bc(A.tlx), omu), mul2(bc(A.trx), u), one), m2z = add2x(mul2(bc(A.tlz), omu), mul2(bc(A.trz), u), one);
        // MODE 1: world-space step between consecutive rows of this column (the grid is bilinear, hence
        // affine in the row index along a column), so every wave's phase advances by a constant per row
        const f2 ex = mul2(sub2(m2x, m1x), bc(A.du)), ez = mul2(sub2(m2z, m1z), bc(A.du));
        f2 px[R], py[R], pz[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const float v = mul_rn((float)(row_a + j), A.du);
            const float omv = sub_rn(1.0f, v);
            px[j] = add2x(mul2(m1x, bc(omv)), mul2(m2x, bc(v)), one);
            pz[j] = add2x(mul2(m1z, bc(omv)), mul2(m2z, bc(v)), one);
            py[j] = bc(0.0f);
        }
        // ---- the hot loop: sum of Gerstner waves (water-grid.vert:74-85) ----
        if (MODE == 0) {
            f2 gx[R], gz[R];
#pragma unroll
            for (int j = 0; j < R; ++j) { gx[j] = px[j]; gz[j] = pz[j]; }
            WT_UNROLL(WT_WAVE_UNROLL_EXACT)
            for (int i = 0; i < A.n_waves; ++i) {
                WaveK wk = wt.w[i];   // phi holds fl(phaseConstant_phi * t) for this launch
                wk.phi = frame_phi<BATCH>(bt, cur_frame, i, wk.phi);
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    // frequency_w * dot(D, xz) + phi * t, every operation rounded as GLSL/GLM does
                    const f2 d = add2x(mul2(bc(wk.Dx), gx[j]), mul2(bc(wk.Dz), gz[j]), one);
                    const f2 k = add2x(mul2(bc(wk.w), d), bc(wk.phi), one);
                    f2 s, c;
                    sincos2(k, s, c);
                    px[j] = fma2(bc(wk.cX), c, px[j]);
                    py[j] = fma2(bc(wk.A), s, py[j]);
                    pz[j] = fma2(bc(wk.cZ), c, pz[j]);
                }
            }
        } else if (WT_SYM && (R & 1)) {
            // Symmetric-pair form of the recurrence mode.  The R = 2H+1 rows of a step sit at phases
            // k_c + j*dk, j = -H..H, around the centre row, and
            //   cos(k_c +- j dk) = cos k_c cos(j dk) -+ sin k_c sin(j dk),   sin(k_c +- j dk) = sin k_c cos(j dk) +- cos k_c sin(j dk)
            // so the SUM and the DIFFERENCE of the two rows of a pair each take one FMA per component per wave:
            //   a_j += (c cos k_c) C_j,  b_j += (c sin k_c) S_j,   P[c+j] = a_j - b_j,  P[c-j] = a_j + b_j   (x, z; y with +-)
            // 6 FMAs per wave per row PAIR instead of 2 x (4 rotation + 3 accumulation).  C_j, S_j = cos / sin(j dk) come from
            // one MUFU pair, angle doubling (j even) and one rotation (j odd): depth <= 3 instead of a chain of R-1 rotations.
            constexpr int H = R / 2;
            const f2 gxc = px[H], gzc = pz[H];
            f2 ax[H], bx[H], ay[H], by[H], az[H], bz[H];
#pragma unroll
            for (int j = 0; j < H; ++j) { ax[j] = bx[j] = ay[j] = by[j] = az[j] = bz[j] = bc(0.0f); }
            f2 ccx = bc(0.0f), ccy = bc(0.0f), ccz = bc(0.0f);
            WaveK wk_next = wt.w[0];   // (Dx,Dz) hold w*D, phi holds fl(phi*t) reduced mod 2pi
            WT_UNROLL(WT_WAVE_UNROLL)
            for (int i = 0; i < A.n_waves; ++i) {
                WaveK wk = wk_next;
                wk.phi = frame_phi<BATCH>(bt, cur_frame, i, wk.phi);
                wk_next = wt.w[(WaveTable<CAP>::PAD || i + 1 < A.n_waves) ? i + 1 : i];
                f2 s, c;
                sincos2(fma2(bc(wk.Dx), gxc, fma2(bc(wk.Dz), gzc, bc(wk.phi))), s, c);
                const f2 dk = fma2(bc(wk.Dx), ex, mul2(bc(wk.Dz), ez));
                f2 C[H + 1], S[H + 1];
                S[1] = make_float2(__sinf(dk.x), __sinf(dk.y));
                C[1] = make_float2(__cosf(dk.x), __cosf(dk.y));
                const f2 qc = mul2(bc(wk.cX), c), qs = mul2(bc(wk.cX), s);
                const f2 rc = mul2(bc(wk.cZ), c), rs = mul2(bc(wk.cZ), s);
                const f2 yc = mul2(bc(wk.A), c), ys = mul2(bc(wk.A), s);
                ccx = add2(ccx, qc); ccy = add2(ccy, ys); ccz = add2(ccz, rc);
#pragma unroll
                for (int j = 1; j <= H; ++j) {
                    if (j > 1) {
                        if ((j & 1) == 0) {
                            // e^{i 2t} from e^{i t}: cos 2t = 1 - 2 sin^2 t (exact form for small t), sin 2t = 2 sin t cos t
                            const f2 t2 = add2(S[j / 2], S[j / 2]);
                            C[j] = fma2(make_float2(-t2.x, -t2.y), S[j / 2], bc(1.0f));
                            S[j] = mul2(t2, C[j / 2]);
                        } else {
                            C[j] = fma2(C[j - 1], C[1], mul2(make_float2(-S[j - 1].x, -S[j - 1].y), S[1]));
                            S[j] = fma2(S[j - 1], C[1], mul2(C[j - 1], S[1]));
                        }
                    }
                    ax[j - 1] = fma2(qc, C[j], ax[j - 1]); bx[j - 1] = fma2(qs, S[j], bx[j - 1]);
                    az[j - 1] = fma2(rc, C[j], az[j - 1]); bz[j - 1] = fma2(rs, S[j], bz[j - 1]);
                    ay[j - 1] = fma2(ys, C[j], ay[j - 1]); by[j - 1] = fma2(yc, S[j], by[j - 1]);
                }
            }
            px[H] = add2(px[H], ccx); py[H] = ccy; pz[H] = add2(pz[H], ccz);
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                px[H + j] = add2(px[H + j], sub2(ax[j - 1], bx[j - 1])); px[H - j] = add2(px[H - j], add2(ax[j - 1], bx[j - 1]));
                pz[H + j] = add2(pz[H + j], sub2(az[j - 1], bz[j - 1])); pz[H - j] = add2(pz[H - j], add2(az[j - 1], bz[j - 1]));
                py[H + j] = add2(ay[j - 1], by[j - 1]); py[H - j] = sub2(ay[j - 1], by[j - 1]);
            }
        } else {
            const f2 gx0 = px[0], gz0 = pz[0];
            WaveK wk_next = wt.w[0];   // (Dx,Dz) hold w*D, phi holds fl(phi*t) reduced mod 2pi
            WT_UNROLL(WT_WAVE_UNROLL)
            for (int i = 0; i < A.n_waves; ++i) {
                // the next wave's constants are fetched from the constant bank one trip ahead
                WaveK wk = wk_next;
                wk.phi = frame_phi<BATCH>(bt, cur_frame, i, wk.phi);
                wk_next = wt.w[(WaveTable<CAP>::PAD || i + 1 < A.n_waves) ? i + 1 : i];
                f2 s, c;
                sincos2(fma2(bc(wk.Dx), gx0, fma2(bc(wk.Dz), gz0, bc(wk.phi))), s, c);
                // per-row phase step: small unless the wave is undersampled by the grid; MUFU's own range
                // handling (error ~ ulp of the angle in turns) is ample for it
                const f2 dk = fma2(bc(wk.Dx), ex, mul2(bc(wk.Dz), ez));
                const f2 sd = make_float2(__sinf(dk.x), __sinf(dk.y));
                const f2 cd = make_float2(__cosf(dk.x), __cosf(dk.y));
                px[0] = fma2(bc(wk.cX), c, px[0]);
                py[0] = fma2(bc(wk.A), s, py[0]);
                pz[0] = fma2(bc(wk.cZ), c, pz[0]);
#if WT_CHEB
                // Row 1 by one complex rotation, e^{i(k + dk)} = e^{ik} e^{i dk}; rows 2.. by the three-term recurrence
                //   x_{j+1} = 2 cos(dk) x_j - x_{j-1}        (x = cos or sin of k + j dk)
                // — one FMA per value instead of two operations.  Re-seeded every R rows, so a rounding error e of one
                // step reaches row n as e (n - m) at most (U_{n-m-1}(cos dk) <= n - m): <= 2.2e-6 of a unit wave at row 7, 5.3e-6 at row 11
                // (tests/test_recurrence_math.py).
                f2 cp = c, sp = s;   // row j - 1
                c = fma2(cp, cd, mul2(make_float2(-sp.x, -sp.y), sd));
                s = fma2(sp, cd, mul2(cp, sd));
                const f2 tc = add2(cd, cd);
                px[1] = fma2(bc(wk.cX), c, px[1]);
                py[1] = fma2(bc(wk.A), s, py[1]);
                pz[1] = fma2(bc(wk.cZ), c, pz[1]);
#pragma unroll
                for (int j = 2; j < R; ++j) {
                    const f2 cn = fma2(tc, c, make_float2(-cp.x, -cp.y));
                    const f2 sn = fma2(tc, s, make_float2(-sp.x, -sp.y));
                    cp = c; sp = s; c = cn; s = sn;
                    px[j] = fma2(bc(wk.cX), c, px[j]);
                    py[j] = fma2(bc(wk.A), s, py[j]);
                    pz[j] = fma2(bc(wk.cZ), c, pz[j]);
                }
#else
#pragma unroll
                for (int j = 1; j < R; ++j) {
                    // e^{i(k + dk)} = e^{ik} e^{i dk}
                    const f2 c2 = fma2(c, cd, mul2(make_float2(-s.x, -s.y), sd));
                    s = fma2(s, cd, mul2(c, sd));
                    c = c2;
                    px[j] = fma2(bc(wk.cX), c, px[j]);
                    py[j] = fma2(bc(wk.A), s, py[j]);
                    pz[j] = fma2(bc(wk.cZ), c, pz[j]);
                }
#endif
            }
        }
        if (A.dbg & 1) {
            f2 t = bc(0.f);
#pragma unroll
            for (int j = 0; j < R; ++j) t = add2(t, add2(px[j], add2(py[j], pz[j])));
            if (t.x + t.y == 123.456f) A.pos[0] = make_float4(t.x, t.y, 0.f, 0.f);
            continue;
        }
        // ---- heightmap detail + bounce, final positions into the ring ----
        __syncwarp();   // the previous step's readers are done with the rows being overwritten
        {
            // HM_BATCH rows at a time: all their footprint loads are issued before the first blend
            constexpr int HM_BATCH = R < WT_HM_BATCH ? R : WT_HM_BATCH;
            int slot = slot_new;
            // The seam entries of the ring.  Measured (profiles/README.md): the branch-free form takes 6 % off the small-frame tier with
            // digests (1024^2 x 8 waves) and is neutral at 8192^2 / 16384^2, but costs the shading instantiations 0.7 % (cfg3): they keep the branches.
            constexpr bool SEAM_BRANCHLESS = !SHADE;
            float* const seam_scratch = reinterpret_cast<float*>(ring_all + RL::PLANES * RL::PLANE) + lane;
            const int seam_sa = lane == 31 ? 2 * RL::PLANE : 0, seam_sb = lane == 0 ? 2 * RL::PLANE : 0;   // plane stride of the seam entry / none in the scratch word
#pragma unroll
            for (int j0 = 0; j0 < R; j0 += HM_BATCH) {
                HmTap tap[HM_BATCH];
#pragma unroll
                for (int jj = 0; jj < HM_BATCH; ++jj)
                    if (j0 + jj < R) tap[jj] = hm_fetch2<HMP2, MODE == 0 && WT_HM_MAGIC_FLOOR>(A, px[j0 + jj], pz[j0 + jj]);
#pragma unroll
                for (int jj = 0; jj < HM_BATCH; ++jj) {
                    const int j = j0 + jj;
                    if (j >= R) break;
                    f2 texel_r;
                    py[j] = hm_finish2(A, tap[jj], py[j], texel_r);
                    float2* const row = ring + slot * RING_COLS;
                    if (VARY) row[3 * RL::PLANE + 1 + lane] = texel_r;
                    row[1 + lane] = px[j];
                    row[RL::PLANE + 1 + lane] = py[j];
                    row[2 * RL::PLANE + 1 + lane] = pz[j];
#if !WT_SEAMLESS_RING
                    if (SEAM_BRANCHLESS) {
                        // lane 31 puts its A value into entry 0's .y and lane 0 its B value into entry 33's .x; every other lane
                        // stores into its own scratch word behind the ring (nobody reads it): no divergent region per row
                        float* const rowf = reinterpret_cast<float*>(row);
                        float* const pa = lane == 31 ? rowf + 1 : seam_scratch;
                        float* const pb = lane == 0 ? rowf + 2 * 33 : seam_scratch;
                        pa[0] = px[j].x; pa[seam_sa] = py[j].x; pa[2 * seam_sa] = pz[j].x;
                        pb[0] = px[j].y; pb[seam_sb] = py[j].y; pb[2 * seam_sb] = pz[j].y;
                    } else {
                        if (lane == 31) {
                            row[0] = make_float2(0.f, px[j].x); row[RL::PLANE] = make_float2(0.f, py[j].x); row[2 * RL::PLANE] = make_float2(0.f, pz[j].x);
                        }
                        if (lane == 0) {
                            row[33] = make_float2(px[j].y, 0.f); row[RL::PLANE + 33] = make_float2(py[j].y, 0.f); row[2 * RL::PLANE + 33] = make_float2(pz[j].y, 0.f);
                        }
                    }
#endif
                    slot = slot + 1 == RL::ROWS ? 0 : slot + 1;
                }
            }
        }
        __syncwarp();

        // ---- outputs for rows row_a-1 .. row_a+R-2 (their up/down neighbours are in the ring now) ----
        // Straight-line code, only the stores are predicated: the R rows are independent, so their
        // shared-memory reads, MUFU chains and cubemap gathers overlap instead of queueing row by row.
        // The rows are independent: with two of them per loop trip their ring reads, MUFU chains and cubemap
        // gathers overlap.  Only the stores are predicated.
        int sd_ = slot_new - 2 < 0 ? slot_new - 2 + RL::ROWS : slot_new - 2;   // row_a - 2
#if WT_SEAMLESS_RING
        // float offsets of this lane's left / right neighbours inside a ring row (see front())
        const int oLx = 2 * (lane ? lane : 1), oLy = lane ? 2 * lane + 1 : 2 * 32;
        const int oRx = lane < 31 ? 2 * (lane + 2) : 2 * 1 + 1, oRy = lane < 31 ? 2 * (lane + 2) + 1 : 2 * 32 + 1;
#endif
        // front half of one output row: ring reads, normal, view-facing flip, shading terms, cubemap gathers ISSUED
        auto front = [&](EpiRow& e, const int j) {
            const int sc_ = sd_ + 1 == RL::ROWS ? 0 : sd_ + 1;
            const int su_ = sc_ + 1 == RL::ROWS ? 0 : sc_ + 1;
            const float2* const rc = ring + sc_ * RING_COLS + lane;
            const float2* const rd = ring + sd_ * RING_COLS + lane + 1;
            const float2* const ru = ring + su_ * RING_COLS + lane + 1;
            sd_ = sc_;
            e.row = row_a - 1 + j;
            e.slot = sc_;
            const f2 cx = rc[1], cy = rc[RL::PLANE + 1], cz = rc[2 * RL::PLANE + 1];
            // (right - left), (up - down): water-grid.vert:164-187
#if WT_SEAMLESS_RING
            // entry 1 + l holds (A_l, B_l).  Left neighbours: (A_{l-1}, B_{l-1}) = entry l — except lane 0, whose B column
            // (strip column 32) has A_31 to its left; right neighbours: entry l + 2 — except lane 31, whose A column (31) has
            // B_0 to its right.  Lane 0's A and lane 31's B are halo columns: any finite value serves as their neighbour.
            const float* const rf = reinterpret_cast<const float*>(ring + sc_ * RING_COLS);
            constexpr int PF = 2 * RL::PLANE;   // floats per plane
            const f2 lx = make_float2(rf[oLx], rf[oLy]), rx_ = make_float2(rf[oRx], rf[oRy]);
            const f2 ly = make_float2(rf[PF + oLx], rf[PF + oLy]), ry_ = make_float2(rf[PF + oRx], rf[PF + oRy]);
            const f2 lz = make_float2(rf[2 * PF + oLx], rf[2 * PF + oLy]), rz_ = make_float2(rf[2 * PF + oRx], rf[2 * PF + oRy]);
            const f2 ux = sub2(rx_, lx), uy = sub2(ry_, ly), uz = sub2(rz_, lz);
#else
            const f2 ux = sub2(rc[2], rc[0]), uy = sub2(rc[RL::PLANE + 2], rc[RL::PLANE]), uz = sub2(rc[2 * RL::PLANE + 2], rc[2 * RL::PLANE]);
#endif
            const f2 vx = sub2(ru[0], rd[0]), vy = sub2(ru[RL::PLANE], rd[RL::PLANE]), vz = sub2(ru[2 * RL::PLANE], rd[2 * RL::PLANE]);
            f2 nx = fma2(uy, vz, mul2(make_float2(-vy.x, -vy.y), uz));
            f2 ny = fma2(uz, vx, mul2(make_float2(-vz.x, -vz.y), ux));
            f2 nz = fma2(ux, vy, mul2(make_float2(-vx.x, -vx.y), uy));
            const f2 n2 = fma2(nx, nx, fma2(ny, ny, mul2(nz, nz)));
            f2 inv = make_float2(rsq_approx(n2.x), rsq_approx(n2.y));
            const f2 wx = sub2(bc(A.camx), cx), wy = sub2(bc(A.camy), cy), wz = sub2(bc(A.camz), cz);   // viewVecRaw, :158
            // :190 — the sign of dot(normal, normalize(viewVecRaw)) is the sign of the unnormalised dot
            const f2 facing = fma2(nx, wx, fma2(ny, wy, mul2(nz, wz)));
            inv.x = facing.x < 0.0f ? -inv.x : inv.x;
            inv.y = facing.y < 0.0f ? -inv.y : inv.y;
            nx = mul2(nx, inv); ny = mul2(ny, inv); nz = mul2(nz, inv);
            e.nx = nx; e.ny = ny; e.nz = nz;
            e.fres = bc(0.0f); e.sp = bc(0.0f);
            if (SHADE) {
                const f2 w2 = fma2(wx, wx, fma2(wy, wy, mul2(wz, wz)));
                const f2 winv = make_float2(rsq_approx(w2.x), rsq_approx(w2.y));
                const f2 e_x = mul2(wx, winv), e_y = mul2(wy, winv), e_z = mul2(wz, winv);          // frag:71-73
                const f2 f = fma2(nx, e_x, fma2(ny, e_y, mul2(nz, e_z)));                            // dot(normal, viewVec) >= 0
                const f2 f2x = add2(f, f);
                // R = reflect(-viewVec, normal) = -viewVec + 2 dot(normal, viewVec) normal   (frag:95)
                const f2 rx = fma2(nx, f2x, make_float2(-e_x.x, -e_x.y)), ry = fma2(ny, f2x, make_float2(-e_y.x, -e_y.y)),
                         rz = fma2(nz, f2x, make_float2(-e_z.x, -e_z.y));
                e.skyA = sky_fetch(A, rx.x, ry.x, rz.x);                                          // frag:96
                e.skyB = sky_fetch(A, rx.y, ry.y, rz.y);
                f2 sdot = fma2(rx, bc(A.sunx), fma2(ry, bc(A.suny), mul2(rz, bc(A.sunz))));
                sdot = make_float2(fminf(fmaxf(sdot.x, 0.0f), 1.0f), fminf(fmaxf(sdot.y, 0.0f), 1.0f));
                // pow(sd, shininess) (frag:100) = 2^(shininess log2 sd): 0^y = 0 for y > 0; 0^0 = 1 is patched below
                const f2 lg = mul2(make_float2(lg2_approx(sdot.x), lg2_approx(sdot.y)), bc(A.sun_shininess));
                f2 sp = make_float2(ex2_approx(lg.x), ex2_approx(lg.y));
                if (A.sun_shininess == 0.0f) sp = bc(1.0f);
                e.sp = mul2(sp, bc(A.sun_strength));
                const f2 om = sub2(bc(1.0f), f);                                                     // frag:110-111
                const f2 om2 = mul2(om, om);
                e.fres = fma2(mul2(mul2(om2, om2), om), bc(1.0f - 0.02f), bc(0.02f));
            }
        };
        // back half: texels converted and blended, position re-read from the ring, the three float4 stores
        auto back = [&](const EpiRow& e) {
            const int row = e.row;
            const bool valid = row >= r0 && row < r1;
            const float2* const rc = ring + e.slot * RING_COLS + lane;
            const f2 cx = rc[1], cy = rc[RL::PLANE + 1], cz = rc[2 * RL::PLANE + 1];
            float4 sA = make_float4(0.f, 0.f, 0.f, 0.f), sB = sA;
            if (SHADE) {
                const float3 skyA = sky_blend(e.skyA), skyB = sky_blend(e.skyB);
                sA = make_float4(skyA.x, skyA.y, skyA.z, e.sp.x);
                sB = make_float4(skyB.x, skyB.y, skyB.z, e.sp.y);
            }
            // one element offset per row; the stores themselves are the only predicated instructions
            const size_t o = (size_t)(row - A.row_begin) * A.L + colA;
            const bool stA = outA && valid, stB = outB && valid;
            const float4 opA = make_float4(cx.x, cy.x, cz.x, 1.0f), onA = make_float4(e.nx.x, e.ny.x, e.nz.x, e.fres.x);
            const float4 opB = make_float4(cx.y, cy.y, cz.y, 1.0f), onB = make_float4(e.nx.y, e.ny.y, e.nz.y, e.fres.y);
            float4* const pp = A.pos + o;
            float4* const pn = A.nrm + o;
            if (stA) __stcs(pp, opA);
            if (stB) __stcs(pp + 32, opB);
            if (stA) __stcs(pn, onA);
            if (stB) __stcs(pn + 32, onB);
            if (SHADE) {
                float4* const ps = A.shade + o;
                if (stA) __stcs(ps, sA);
                if (stB) __stcs(ps + 32, sB);
            }
            if (VARY) {
                // gl_Position = viewProjection * position (:150); the same with y = 0, .xy / .w (:153-154)
                const float* const m = A.vp;
                f2 cl[4], c0[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    c0[r] = fma2(bc(m[8 + r]), cz, fma2(bc(m[r]), cx, bc(m[12 + r])));
                    cl[r] = fma2(bc(m[4 + r]), cy, c0[r]);
                }
                const f2 iw = make_float2(__fdividef(1.0f, c0[3].x), __fdividef(1.0f, c0[3].y));
                const f2 ndx = mul2(c0[0], iw), ndy = mul2(c0[1], iw);
                // normalVecInViewSpaceOnlyYaw = normalize((viewMatOnlyYaw * vec4(normal, 0)).xyz) (:193)
                const float* const y = A.vy;
                f2 yn[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) yn[r] = fma2(bc(y[6 + r]), e.nz, fma2(bc(y[3 + r]), e.ny, mul2(bc(y[r]), e.nx)));
                const f2 y2 = fma2(yn[0], yn[0], fma2(yn[1], yn[1], mul2(yn[2], yn[2])));
                const f2 yi = make_float2(rsq_approx(y2.x), rsq_approx(y2.y));
                const f2 hr = rc[3 * RL::PLANE + 1];
                float4* const pc = A.clip + o;
                float4* const pa = A.vary_a + o;
                float4* const pb = A.vary_b + o;
                if (stA) __stcs(pc, make_float4(cl[0].x, cl[1].x, cl[2].x, cl[3].x));
                if (stB) __stcs(pc + 32, make_float4(cl[0].y, cl[1].y, cl[2].y, cl[3].y));
                if (stA) __stcs(pa, make_float4(ndx.x, ndy.x, hr.x, 0.0f));
                if (stB) __stcs(pa + 32, make_float4(ndx.y, ndy.y, hr.y, 0.0f));
                if (stA) __stcs(pb, make_float4(yn[0].x * yi.x, yn[1].x * yi.x, yn[2].x * yi.x, 0.0f));
                if (stB) __stcs(pb + 32, make_float4(yn[0].y * yi.y, yn[1].y * yi.y, yn[2].y * yi.y, 0.0f));
            }
            if (CKS_SMEM) {
                auto w = [](float f) { return (unsigned long long)__float_as_uint(f); };
                float2* const yr = reinterpret_cast<float2*>(cks_slot + 96);
                if (stA) {
                    cks_slot[0] += w(opA.x) + w(opA.y) + w(opA.z);
                    cks_slot[32] += w(onA.x) + w(onA.y) + w(onA.z) + w(onA.w);
                    if (SHADE) cks_slot[64] += w(sA.x) + w(sA.y) + w(sA.z) + w(sA.w);
                    *yr = make_float2(fminf(yr->x, opA.y), fmaxf(yr->y, opA.y));
                }
                if (stB) {
                    cks_slot[0] += w(opB.x) + w(opB.y) + w(opB.z);
                    cks_slot[32] += w(onB.x) + w(onB.y) + w(onB.z) + w(onB.w);
                    if (SHADE) cks_slot[64] += w(sB.x) + w(sB.y) + w(sB.z) + w(sB.w);
                    *yr = make_float2(fminf(yr->x, opB.y), fmaxf(yr->y, opB.y));
                }
            } else if (CKS && valid) {
                // `valid` is uniform and the halo lanes are masked, not branched around: no divergence (5.7 % -> 3.9 % of the recurrence kernel)
                acc.add_pair(opA, onA, sA, stA, opB, onB, sB, stB, SHADE);
            }
        };
        {
            // MODE 0 runs at 80 registers (24 warps per SM): one row at a time there, two rows in flight otherwise
            WT_UNROLL_IF(MODE == 0, WT_EPI_UNROLL_EXACT, WT_EPI_UNROLL)
            for (int j = 0; j < R; ++j) {
                EpiRow e;
                front(e, j);
                back(e);
            }
        }
        slot_new = slot_new + R >= RL::ROWS ? slot_new + R - RL::ROWS : slot_new + R;
    }
    // every row of [r0, r1) was stored exactly once for each of this lane's in-range columns
    if (CKS) {
        const unsigned long long nv = (unsigned long long)max(r1 - r0, 0) * ((outA ? 1 : 0) + (outB ? 1 : 0));
        acc.count += nv;
    }
    if (BATCH && CKS_SMEM) { cks_collect(); cks_reset(); }
    if (BATCH && CKS) {
        // one digest per FRAME: the frame's own record (armed by the launcher) takes this item's sums; the item that completes the
        // frame decodes the record in place
        acc.flush(A.cks);
        acc = ChecksumAcc();
        if (lane == 0) {
            __threadfence();
            if (atomicAdd(a.batch_done + cur_frame, 1) == a.items_per_frame - 1) {
                __threadfence();
                wt_checksum* const r = A.cks;
                const float lo = key_float(atomicAdd((unsigned*)&r->min_y, 0u)), hi = key_float(atomicAdd((unsigned*)&r->max_y, 0u));
                r->min_y = lo; r->max_y = hi;
                a.batch_done[cur_frame] = 0;
            }
        }
    }
    if (a.trace && lane == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long* t = a.trace + 4 * (size_t)item;
        t[0] = smid; t[1] = blockIdx.x; t[2] = t_start; t[3] = global_ns();
    }
    if (!dyn) break;
    __syncwarp();   // every lane is done with the ring before the next item overwrites it
    item = next_item(a.queue, lane);
  }
    if (dyn && lane == 0) {
        // the last warp to run dry re-arms the queue for the next launch that uses it
        if (atomicAdd(a.queue + 1, 1) == (int)gridDim.x - 1) { a.queue[1] = 0; __threadfence(); a.queue[0] = 0; }
    }
    if (CKS && !BATCH) {
        if (CKS_SMEM) cks_collect();
        acc.flush(a.cks_acc);
        if (lane == 0) {
            __threadfence();
            if (atomicAdd(a.cks_done, 1) == (int)gridDim.x - 1) {
                __threadfence();
                wt_checksum* const r = a.cks_acc;
                wt_checksum out;
                out.position_bits = atomicAdd((unsigned long long*)&r->position_bits, 0ULL);
                out.normal_bits = atomicAdd((unsigned long long*)&r->normal_bits, 0ULL);
                out.shade_bits = atomicAdd((unsigned long long*)&r->shade_bits, 0ULL);
                out.vertices = atomicAdd((unsigned long long*)&r->vertices, 0ULL);
                out.min_y = key_float(atomicAdd((unsigned*)&r->min_y, 0u));
                out.max_y = key_float(atomicAdd((unsigned*)&r->max_y, 0u));
                *a.cks = out;
                // re-arm for the next launch that is handed this record
                r->position_bits = 0; r->normal_bits = 0; r->shade_bits = 0; r->vertices = 0;
                *(unsigned*)&r->min_y = 0xffffffffu;
                *(unsigned*)&r->max_y = 0u;
                __threadfence();
                *a.cks_done = 0;
                __threadfence_system();   // *cks may live in mapped host memory (wt_step_host)
            }
        }
    }
}

}  // namespace wt
