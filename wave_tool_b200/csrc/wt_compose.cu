// The screen-space layers of the water fragment shader evaluated per vertex: depth delta,
// local reflection / refraction lookups, tint, Fresnel mix, soft edges and fog — water-grid.frag:66-69,76-89,113-162,
// i.e. everything of main() that the fused surface kernel does not already produce.  A second, streaming pass over
// the vertex stage's outputs (position, normal + fresnel, shade, gl_Position and the two varyings records): its
// texture inputs come from other render passes of the host application (render-engine.cpp:128-188 FBO attachments).
// The fragment inputs a rasteriser would interpolate are taken un-interpolated at the vertex; gl_FragCoord follows
// from gl_Position by the viewport transform.
#include <string>

#include "wt_ctx.h"

namespace wt {

struct ComposeArgs {
    const float4 *pos, *nrm, *shade, *clip, *va, *vb;
    float4* colour;
    const float* depth;        // w x h, GL row order, window-space depth
    const uchar4* refl;        // w x h RGBA8
    const uchar4* refr;
    int w, h;
    long long n;
    float fog[4], fog_far, fog_near, soft_edges, tint_threshold, clarity, z_near, z_far, vw, vh;
    float camx, camy, camz, sun_y;
    // reciprocals of the per-launch divisors (div_by_uniform) and the hoisted viewport halves
    float r_zfar, r_vw, r_vh, r_tint, r_soft, r_fog, half_vw, half_vh;
    // clip == NULL: gl_Position, the height-0 NDC position and the yaw-only view-space normal are recomputed here from
    // position and normal (the same FMA chains as the surface kernel's WT_FLAG_VARYINGS) instead of being read back
    float vp[16], vy[12];
};

__device__ __forceinline__ float mix_f(float x, float y, float a) { return __fadd_rn(__fmul_rn(x, __fsub_rn(1.0f, a)), __fmul_rn(y, a)); }
__device__ __forceinline__ float clamp_f(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// texture(sampler2D, st) of an RGBA8 texture: GL_LINEAR, GL_CLAMP_TO_EDGE, exact fp32 weights, the oracle's operation order.
// lut[c] = c / 255 exactly (IEEE division, built per block by unorm8f): a table look-up replaces the six-instruction
// byte -> float -> / 255 sequence, 32 times per vertex — the conversions were 29 % of this pass's instructions.
__device__ __forceinline__ float4 sample_rgba8_clamp_linear(const uchar4* tex, const float* lut, int w, int h, float s, float t) {
    // s, t in [0, 1] (the callers clamp them): um, vm in [-0.5, w - 0.5], so floor(um) is in [-1, w - 1] — each of the four
    // footprint indices needs one clamp only.  floor and the float -> int conversion by the magic-number add (off the XU pipe).
    const float um = __fsub_rn(__fmul_rn(s, (float)w), 0.5f), vm = __fsub_rn(__fmul_rn(t, (float)h), 0.5f);
    int iu, iv;
    const float fu = floor_small(um, iu), fv = floor_small(vm, iv);
    const float a = __fsub_rn(um, fu), b = __fsub_rn(vm, fv);
    const unsigned i0 = (unsigned)max(iu, 0), i1 = (unsigned)min(iu + 1, w - 1), j0 = (unsigned)max(iv, 0), j1 = (unsigned)min(iv + 1, h - 1);
    const unsigned* const tx = reinterpret_cast<const unsigned*>(tex);
    const unsigned r0 = j0 * (unsigned)w, r1 = j1 * (unsigned)w;        // w, h <= 16384: the texel index fits 32 bits
    const unsigned t00 = __ldg(tx + (r0 + i0)), t10 = __ldg(tx + (r0 + i1));
    const unsigned t01 = __ldg(tx + (r1 + i0)), t11 = __ldg(tx + (r1 + i1));
    const float oma = __fsub_rn(1.0f, a), omb = __fsub_rn(1.0f, b);
    const float w00 = __fmul_rn(oma, omb), w10 = __fmul_rn(a, omb), w01 = __fmul_rn(oma, b), w11 = __fmul_rn(a, b);
    auto ch = [&](unsigned k) {
        float acc = __fmul_rn(w00, lut[(t00 >> (8 * k)) & 255u]);
        acc = __fadd_rn(acc, __fmul_rn(w10, lut[(t10 >> (8 * k)) & 255u]));
        acc = __fadd_rn(acc, __fmul_rn(w01, lut[(t01 >> (8 * k)) & 255u]));
        return __fadd_rn(acc, __fmul_rn(w11, lut[(t11 >> (8 * k)) & 255u]));
    };
    return make_float4(ch(0), ch(1), ch(2), ch(3));
}

__global__ void __launch_bounds__(256) compose_kernel(const ComposeArgs a) {
    __shared__ float lut[256];
    lut[threadIdx.x] = unorm8f((float)threadIdx.x);     // launched with 256 threads per block
    __syncthreads();
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < a.n; k += (long long)gridDim.x * blockDim.x) {
        const float4 P = __ldcs(a.pos + k), N = __ldcs(a.nrm + k), S = __ldcs(a.shade + k);
        float4 C, A, B;
        if (a.clip) {
            C = __ldcs(a.clip + k); A = __ldcs(a.va + k); B = __ldcs(a.vb + k);
        } else {
            const float* const m = a.vp;
            float c0[4], cl[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                c0[r] = fmaf(m[8 + r], P.z, fmaf(m[r], P.x, m[12 + r]));
                cl[r] = fmaf(m[4 + r], P.y, c0[r]);
            }
            C = make_float4(cl[0], cl[1], cl[2], cl[3]);
            const float iw = __fdividef(1.0f, c0[3]);
            A = make_float4(c0[0] * iw, c0[1] * iw, 0.0f, 0.0f);
            const float* const y = a.vy;
            float yn[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) yn[r] = fmaf(y[6 + r], N.z, fmaf(y[3 + r], N.y, y[r] * N.x));
            const float yi = rsqrtf(fmaf(yn[0], yn[0], fmaf(yn[1], yn[1], yn[2] * yn[2])));
            B = make_float4(yn[0] * yi, yn[1] * yi, yn[2] * yi, 0.0f);
        }
        auto linearize = [&](float d) { return __fdiv_rn(__fmul_rn(a.z_near, d), __fsub_rn(a.z_far, __fmul_rn(d, __fsub_rn(a.z_far, a.z_near)))); };   // frag:66-69
        const float wx = __fsub_rn(a.camx, P.x), wy = __fsub_rn(a.camy, P.y), wz = __fsub_rn(a.camz, P.z);
        const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(wx, wx), __fmul_rn(wy, wy)), __fmul_rn(wz, wz)));   // frag:71
        const float vdc = clamp_f(div_by_uniform(len, a.z_far, a.r_zfar), 0.0f, 1.0f);                                   // frag:72
        // gl_FragCoord: viewport transform of gl_Position (GL 4.1 section 2.14.1, default depth range)
        const float ndx = __fdiv_rn(C.x, C.w), ndy = __fdiv_rn(C.y, C.w), ndz = __fdiv_rn(C.z, C.w);
        const float fx = __fmul_rn(__fadd_rn(ndx, 1.0f), a.half_vw), fy = __fmul_rn(__fadd_rn(ndy, 1.0f), a.half_vh);   // x / 2 == x * 0.5 exactly
        const float fz = __fmul_rn(__fadd_rn(ndz, 1.0f), 0.5f);
        const float uvx = div_by_uniform(fx, a.vw, a.r_vw), uvy = div_by_uniform(fy, a.vh, a.r_vh);                      // frag:79
        const float h0x = __fmul_rn(__fadd_rn(A.x, 1.0f), 0.5f), h0y = __fmul_rn(__fadd_rn(A.y, 1.0f), 0.5f);             // frag:81
        // texture(depthTexture2D, uv).x: GL_NEAREST, CLAMP_TO_EDGE (render-engine.cpp:128-136)
        // (screen coordinates of vertices the frustum-padded grid places far off screen are clamped before the conversion)
        int di, dj;
        floor_small(clamp_f(__fmul_rn(uvx, (float)a.w), -1.0f, (float)a.w), di);
        floor_small(clamp_f(__fmul_rn(uvy, (float)a.h), -1.0f, (float)a.h), dj);
        di = min(max(di, 0), a.w - 1); dj = min(max(dj, 0), a.h - 1);
        const float back = linearize(__ldg(a.depth + ((unsigned)dj * (unsigned)a.w + (unsigned)di)));                                            // frag:84
        const float hack = back > 0.999f ? 0.0f : 1.0f;                                                                 // frag:86
        const float dcl = clamp_f(__fsub_rn(back, linearize(fz)), 0.0f, 1.0f);                                           // frag:87-91
        // frag:118-133 local reflections / refractions
        const float ds = __fmul_rn(0.1f, __fsub_rn(1.0f, vdc));
        const float mx = mix_f(uvx, h0x, ds), my = mix_f(uvy, h0y, ds);
        const float rlx = clamp_f(__fadd_rn(mx, __fmul_rn(ds, B.x)), 0.0f, 1.0f), rly = clamp_f(__fadd_rn(my, __fmul_rn(ds, -B.z)), 0.0f, 1.0f);
        const float rrx = clamp_f(__fadd_rn(mx, __fmul_rn(ds, -B.x)), 0.0f, 1.0f), rry = clamp_f(__fadd_rn(my, __fmul_rn(ds, B.z)), 0.0f, 1.0f);
        const float4 refl = sample_rgba8_clamp_linear(a.refl, lut, a.w, a.h, rlx, rly);
        const float4 refr = sample_rgba8_clamp_linear(a.refr, lut, a.w, a.h, rrx, rry);
        // frag:136-151 tint, transmission, reflection, Fresnel mix
        const float tint_f = a.tint_threshold == 0.0f ? 0.0f : __fmul_rn(hack, __fsub_rn(1.0f, clamp_f(div_by_uniform(dcl, a.tint_threshold, a.r_tint), 0.0f, 1.0f)));
        const float t_w = __fmul_rn(__fmul_rn(a.clarity, refr.w), __fsub_rn(1.0f, dcl));
        auto channel = [&](float deep, float shallow, float refr_c, float sky_c, float refl_c) {
            const float tint = __fmul_rn(a.sun_y, mix_f(deep, shallow, tint_f));
            const float transmission = mix_f(tint, refr_c, t_w);
            const float reflection = mix_f(__fadd_rn(sky_c, S.w), refl_c, refl.w);
            return mix_f(transmission, reflection, N.w);
        };
        const float cr = channel(0.0f, 0.3f, refr.x, S.x, refl.x), cg = channel(0.341f, 0.941f, refr.y, S.y, refl.y), cb = channel(0.482f, 0.903f, refr.z, S.z, refl.z);
        // frag:154-156 soft edges
        const float hard = a.soft_edges == 0.0f ? 1.0f : __fsub_rn(1.0f, __fmul_rn(hack, __fsub_rn(1.0f, clamp_f(div_by_uniform(dcl, a.soft_edges, a.r_soft), 0.0f, 1.0f))));
        // frag:160-162 fog
        const float fog_a = vdc >= a.fog_far ? a.fog[3] : vdc <= a.fog_near ? 0.0f
                            : __fmul_rn(a.fog[3], div_by_uniform(__fsub_rn(vdc, a.fog_near), __fsub_rn(a.fog_far, a.fog_near), a.r_fog));
        __stcs(a.colour + k, make_float4(mix_f(cr, a.fog[0], fog_a), mix_f(cg, a.fog[1], fog_a), mix_f(cb, a.fog[2], fog_a), hard));
    }
}

}  // namespace wt

using namespace wt;

extern "C" {

void wt_default_compose_params(wt_compose_params* out, const wt_camera* cam, unsigned viewport_w, unsigned viewport_h) {
    if (!out) return;
    out->fogColourFarAtNoon[0] = out->fogColourFarAtNoon[1] = out->fogColourFarAtNoon[2] = 1.0f;   // render-engine.h:221
    out->fogColourFarAtNoon[3] = 0.1f;
    out->fogDepthRadiusFar = 1.0f;                 // :222
    out->fogDepthRadiusNear = 0.0f;                // :223
    out->softEdgesDeltaDepthThreshold = 0.05f;     // :229
    out->tintDeltaDepthThreshold = 0.1f;           // :234
    out->waterClarity = 0.3f;                      // :235
    out->zNear = cam ? cam->z_near : 0.1f;         // render-engine.cpp:1109-1110
    out->zFar = cam ? cam->z_far : 100.0f;
    out->viewportWidth = (float)viewport_w;        // render-engine.cpp:1107
    out->viewportHeight = (float)viewport_h;
}

int wt_set_screen_textures(wt_ctx* ctx, const float* depth, const uint8_t* refl, const uint8_t* refr, int w, int h) {
    if (!ctx) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_set_screen_textures: ctx is NULL");
    auto fail = [&](int code, const char* msg) { ctx->err = msg; set_thread_error(code, msg); return code; };
    if (!depth || !refl || !refr || w <= 0 || h <= 0 || w > 16384 || h > 16384) return fail(WT_ERR_INVALID_ARGUMENT, "wt_set_screen_textures: bad textures");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_depth); cudaFree(ctx->d_refl); cudaFree(ctx->d_refr);
    ctx->d_depth = nullptr; ctx->d_refl = ctx->d_refr = nullptr;
    const size_t n = (size_t)w * h;
    cudaError_t e = cudaMalloc(&ctx->d_depth, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_refl, n * sizeof(uchar4));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_refr, n * sizeof(uchar4));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_depth, depth, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_refl, refl, n * sizeof(uchar4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_refr, refr, n * sizeof(uchar4), cudaMemcpyHostToDevice);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(WT_ERR_CUDA, cudaGetErrorString(e));
    ctx->screen_w = w; ctx->screen_h = h;
    return WT_OK;
}

int wt_compose_frame(wt_ctx* ctx, const wt_frame* frame, const wt_compose_params* p, unsigned L, unsigned r0, unsigned r1,
                     const wt_outputs* in, void* d_colour, void* stream_) {
    if (!ctx) return set_thread_error(WT_ERR_INVALID_ARGUMENT, "wt_compose_frame: ctx is NULL");
    auto fail = [&](int code, const std::string& msg) { ctx->err = msg; set_thread_error(code, msg.c_str()); return code; };
    if (!frame || !p || !in || !d_colour) return fail(WT_ERR_INVALID_ARGUMENT, "wt_compose_frame: NULL argument");
    if (L < 2 || r0 >= r1 || r1 > L) return fail(WT_ERR_INVALID_ARGUMENT, "need grid_len >= 2 and 0 <= row_begin < row_end <= grid_len");
    if (!in->position || !in->normal || !in->shade)
        return fail(WT_ERR_INVALID_ARGUMENT, "wt_compose_frame needs the position, normal and shade outputs of an evaluation with WT_FLAG_SHADE");
    const bool have_varyings = in->clip_position && in->varyings_a && in->varyings_b;
    if (!have_varyings && (in->clip_position || in->varyings_a || in->varyings_b))
        return fail(WT_ERR_INVALID_ARGUMENT, "wt_compose_frame: pass all three varyings buffers (WT_FLAG_VARYINGS) or none (they are then recomputed)");
    if (!ctx->d_depth) return fail(WT_ERR_NOT_READY, "wt_set_screen_textures has not been called");
    ComposeArgs a;
    a.pos = (const float4*)in->position; a.nrm = (const float4*)in->normal; a.shade = (const float4*)in->shade;
    a.clip = have_varyings ? (const float4*)in->clip_position : nullptr;
    a.va = (const float4*)in->varyings_a; a.vb = (const float4*)in->varyings_b;
    std::memcpy(a.vp, frame->viewProjection, sizeof(a.vp));
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) a.vy[3 * c + r] = frame->viewMatOnlyYaw[4 * c + r];
    a.colour = (float4*)d_colour;
    a.depth = ctx->d_depth; a.refl = ctx->d_refl; a.refr = ctx->d_refr; a.w = ctx->screen_w; a.h = ctx->screen_h;
    a.n = (long long)(r1 - r0) * L;
    const float sy = std::fmin(std::fmax(frame->sunPosition[1], 0.0f), 1.0f);
    for (int c = 0; c < 3; ++c) a.fog[c] = sy * p->fogColourFarAtNoon[c];   // fogColourFarAtCurrentTime, render-engine.cpp:295
    a.fog[3] = p->fogColourFarAtNoon[3];
    a.fog_far = p->fogDepthRadiusFar; a.fog_near = p->fogDepthRadiusNear; a.soft_edges = p->softEdgesDeltaDepthThreshold;
    a.tint_threshold = p->tintDeltaDepthThreshold; a.clarity = p->waterClarity; a.z_near = p->zNear; a.z_far = p->zFar;
    a.vw = p->viewportWidth; a.vh = p->viewportHeight;
    a.camx = frame->cameraPosition[0]; a.camy = frame->cameraPosition[1]; a.camz = frame->cameraPosition[2];
    a.sun_y = sy;
    a.r_zfar = 1.0f / a.z_far; a.r_vw = 1.0f / a.vw; a.r_vh = 1.0f / a.vh;
    a.r_tint = a.tint_threshold != 0.0f ? 1.0f / a.tint_threshold : 0.0f;
    a.r_soft = a.soft_edges != 0.0f ? 1.0f / a.soft_edges : 0.0f;
    a.r_fog = 1.0f / (a.fog_far - a.fog_near);      // only read where fog_near < vdc < fog_far
    a.half_vw = a.vw / 2.0f; a.half_vh = a.vh / 2.0f;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    const int blocks = (int)std::min<long long>((a.n + 255) / 256, (long long)ctx->sm_count * 16);
    compose_kernel<<<std::max(blocks, 1), 256, 0, (cudaStream_t)stream_>>>(a);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(WT_ERR_CUDA, cudaGetErrorString(e));
    return WT_OK;
}

}  // extern "C"
