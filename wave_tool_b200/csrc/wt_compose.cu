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
    cudaTextureObject_t refl;  // w x h RGBA8 arrays: point sampling, clamp to edge, normalised-float reads (exactly c / 255), gather-able
    cudaTextureObject_t refr;
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
// The 2 x 2 footprint comes from the texture unit, one gather per channel: it clamps the footprint to the edge and converts
// unorm8 to exactly fl(c / 255) (all 256 values checked on the device, tools/probes/probe_unorm.cu) — the byte -> float
// conversions and the four loads with their index arithmetic were 40 % of this pass's instructions.
__device__ __forceinline__ float4 sample_rgba8_clamp_linear(cudaTextureObject_t tex, int w, int h, float s, float t) {
    // s, t in [0, 1] (the callers clamp them): um, vm in [-0.5, w - 0.5]; floor by the magic-number add (off the XU pipe)
    const float um = __fsub_rn(__fmul_rn(s, (float)w), 0.5f), vm = __fsub_rn(__fmul_rn(t, (float)h), 0.5f);
    int iu, iv;
    const float fu = floor_small(um, iu), fv = floor_small(vm, iv);
    const float a = __fsub_rn(um, fu), b = __fsub_rn(vm, fv);
    // texels (fu, fu + 1) x (fv, fv + 1): their shared corner (fu + 1, fv + 1) is the gather coordinate (unnormalised, exact in fp32)
    const float gx = __fadd_rn(fu, 1.0f), gy = __fadd_rn(fv, 1.0f);
    const float oma = __fsub_rn(1.0f, a), omb = __fsub_rn(1.0f, b);
    const float w00 = __fmul_rn(oma, omb), w10 = __fmul_rn(a, omb), w01 = __fmul_rn(oma, b), w11 = __fmul_rn(a, b);
    auto ch = [&](int k) {
        // gather order: x = (i0, j1), y = (i1, j1), z = (i1, j0), w = (i0, j0)
        const float4 g = k == 0 ? tex2Dgather<float4>(tex, gx, gy, 0) : k == 1 ? tex2Dgather<float4>(tex, gx, gy, 1)
                       : k == 2 ? tex2Dgather<float4>(tex, gx, gy, 2) : tex2Dgather<float4>(tex, gx, gy, 3);
        float acc = __fmul_rn(w00, g.w);
        acc = __fadd_rn(acc, __fmul_rn(w10, g.z));
        acc = __fadd_rn(acc, __fmul_rn(w01, g.x));
        return __fadd_rn(acc, __fmul_rn(w11, g.y));
    };
    return make_float4(ch(0), ch(1), ch(2), ch(3));
}

#ifndef WT_COMPOSE_MIN_BLOCKS
#define WT_COMPOSE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, WT_COMPOSE_MIN_BLOCKS) compose_kernel(const ComposeArgs a) {
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
        const float4 refl = sample_rgba8_clamp_linear(a.refl, a.w, a.h, rlx, rly);
        const float4 refr = sample_rgba8_clamp_linear(a.refr, a.w, a.h, rrx, rry);
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
    // build into temporaries: the context keeps its previous textures unless every step succeeds
    float* d_depth = nullptr;
    cudaArray_t arr[2] = {nullptr, nullptr};
    cudaTextureObject_t tex[2] = {0, 0};
    const size_t n = (size_t)w * h;
    cudaError_t e = cudaMalloc(&d_depth, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_depth, depth, n * sizeof(float), cudaMemcpyHostToDevice);
    const uint8_t* const src[2] = {refl, refr};
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
        e = cudaMallocArray(&arr[k], &fmt, (size_t)w, (size_t)h, cudaArrayTextureGather);
        if (e == cudaSuccess) e = cudaMemcpy2DToArray(arr[k], 0, 0, src[k], (size_t)w * 4, (size_t)w * 4, (size_t)h, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) break;
        cudaResourceDesc res;
        std::memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;
        res.res.array.array = arr[k];
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;   // GL_CLAMP_TO_EDGE (render-engine.cpp:138-188)
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 0;
        e = cudaCreateTextureObject(&tex[k], &res, &td, nullptr);
    }
    if (e != cudaSuccess) {
        for (int k = 0; k < 2; ++k) { if (tex[k]) cudaDestroyTextureObject(tex[k]); if (arr[k]) cudaFreeArray(arr[k]); }
        cudaFree(d_depth);
        if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
        return fail(WT_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(ctx->d_depth);
    if (ctx->refl_tex) cudaDestroyTextureObject(ctx->refl_tex);
    if (ctx->refr_tex) cudaDestroyTextureObject(ctx->refr_tex);
    if (ctx->refl_arr) cudaFreeArray(ctx->refl_arr);
    if (ctx->refr_arr) cudaFreeArray(ctx->refr_arr);
    ctx->d_depth = d_depth;
    ctx->refl_arr = arr[0]; ctx->refr_arr = arr[1];
    ctx->refl_tex = tex[0]; ctx->refr_tex = tex[1];
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
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
    a.depth = ctx->d_depth; a.refl = ctx->refl_tex; a.refr = ctx->refr_tex; a.w = ctx->screen_w; a.h = ctx->screen_h;
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
