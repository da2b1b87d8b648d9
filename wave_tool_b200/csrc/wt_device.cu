// Device context and the evaluation entry points of the C ABI (include/wavetool_b200.h).
//
// Replaces the host side of the reference's water draw call: uniform upload + glDrawElements
// (src/render-engine.cpp:1045-1122).  There is NO CPU fallback: every entry point here needs a CUDA
// device and fails with WT_ERR_NO_DEVICE / WT_ERR_CUDA otherwise.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "wt_ctx.h"
#include "wt_aux_kernels.cuh"

namespace wt {

static thread_local std::string g_thread_error;

int set_thread_error(int code, const char* msg) {
    g_thread_error = msg ? msg : "";
    return code;
}

}  // namespace wt

using namespace wt;

namespace wt {
Knobs g_knobs;

// Work items are warps: n_strips strips of STRIP_COLS columns x chunks of rows.  Every item marches its rows + 2
// (the +-1-row stencil halo is recomputed, never exchanged), in steps of R.  Cost model for grids with few items:
// the SMs hold `slots` warps at a time; a partially filled round costs as much as a full one; smaller chunks
// balance better but pay the 2 halo rows more often.
Plan make_plan(int slots, int L, int rows, int R) {
    Plan best;
    double best_cost = 1e300;
    const int n_strips = (L + STRIP_COLS - 1) / STRIP_COLS;
    auto consider = [&](int rpc) {
        rpc = ((rpc + 2 + R - 1) / R) * R - 2;   // rpc + 2 a multiple of R: no marching step is partly wasted
        if (rpc < 1) rpc = R - 2 > 0 ? R - 2 : 1;
        const int n_chunks = (rows + rpc - 1) / rpc;
        const long long items = (long long)n_strips * n_chunks;
        const long long full_rounds = (items + slots - 1) / slots;
        const double cost = (double)full_rounds * (rpc + 2) * (1.0 + 0.02 * full_rounds);   // mild preference for fewer rounds
        if (cost < best_cost) {
            best_cost = cost;
            best.n_strips = n_strips; best.rows_per_chunk = rpc; best.n_chunks = n_chunks;
            best.ctas = (int)((items + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
        }
    };
    for (int rounds = 1; rounds <= 6; ++rounds) {
        const long long want = std::max<long long>(1, (long long)rounds * slots / n_strips);
        consider((int)((rows + want - 1) / want));
    }
    // ... and chunks that do not fill the machine once but pay the 2 halo rows less often (a 1024^2 grid: 14-row items in one
    // round instead of 6-row items in two — the same time for a single frame, a quarter less work in a grouped launch)
    for (int steps = 2; steps <= 4; ++steps) consider(steps * R - 2);
    return best;
}

Layout canonical_layout(int L, int R, int mode) {
    Layout lay;
    lay.n_strips = (L + STRIP_COLS - 1) / STRIP_COLS;
    // resident warps of a B200 for this instantiation (__launch_bounds__ of surface_warp_kernel) — a constant of the
    // layout, NOT the occupancy of the device in use
    const int slots = NOMINAL_SMS * (mode == 1 ? (R > 8 ? WT_LB_RECUR_BIG : WT_LB_RECUR) : WT_LB_EXACT);
    auto round_rows = [R](int want) { int r = ((want + 2 + R - 1) / R) * R - 2; return r < 1 ? R : r; };   // rows + 2 a multiple of R
    int big = mode == 1 ? DYN_ROWS_RECUR : DYN_ROWS_EXACT, small = mode == 1 ? DYN_SMALL_RECUR : DYN_SMALL_EXACT, tail = DYN_TAIL_PERCENT;
    const bool forced = g_knobs.dyn >= 0;           // development builds only (WT_DEV_KNOBS)
    if (forced) big = g_knobs.dyn;
    if (g_knobs.dyn_small >= 0) small = g_knobs.dyn_small;
    if (g_knobs.dyn_tail >= 0) tail = g_knobs.dyn_tail;
    if (big > 0) {
        lay.rows_big = round_rows(big);
        lay.rows_small = (small > 0 && small < big && tail > 0) ? round_rows(small) : 0;
        if (!lay.rows_small && !forced) {
            // One level of equal items (the bit-faithful kernel): the launch takes ceil(items / resident warps) rounds of one item
            // each, an item costs its rows + the 2 halo rows — pick the item height that minimises the product (measured at 4096^2,
            // profiles/README.md: 26 rows 0.685 ms = 3 rounds of 28, 30 rows 0.697 = 3 x 32, 38 rows 0.740 = 3 x 40, 42 rows 0.697 = 2 x 44)
            long long best = -1;
            for (int m = 4; m <= 24; ++m) {
                const int r = m * R - 2;
                const long long items = (long long)lay.n_strips * ((L + r - 1) / r);
                const long long cost = ((items + slots - 1) / slots) * (r + 2);
                if (items >= 2LL * slots && (best < 0 || cost < best)) { best = cost; lay.rows_big = r; }
            }
        }
        lay.tail_pct = lay.rows_small ? tail : 0;
        // bands only matter for a two-level layout (they give every row shard of a box its own share of small items); a single
        // level of equal items is balanced as it is, and measured 2 % faster in one band (profiles/README.md)
        lay.bands = (lay.rows_small && L >= LAYOUT_BANDS * 4 * lay.rows_big) ? LAYOUT_BANDS : 1;
        long long chunks = 0;
        for (int b = 0; b < lay.bands; ++b) {
            int lo, hi;
            lay.band_range(L, b, lo, hi);
            const int nb = lay.n_big(hi - lo);
            chunks += nb + (lay.rows_small ? (hi - lo - nb * lay.rows_big + lay.rows_small - 1) / lay.rows_small : 0);
        }
        if (forced || chunks * lay.n_strips >= 2LL * slots) return lay;
    }
    // few items: one level of equal chunks sized to fill the machine
    const Plan p = make_plan(slots, L, L, R);
    lay.bands = 1; lay.rows_big = p.rows_per_chunk; lay.rows_small = 0; lay.tail_pct = 0;
    return lay;
}

void layout_segments(const Layout& lay, int L, int row_begin, int row_end, EvalArgs& a) {
    a.n_strips = lay.n_strips;
    a.n_seg = 0;
    a.seg_prefix[0] = 0;
    auto push = [&](int row0, int rows, int end) {   // chunks of `rows` rows from row0 up to `end`, those meeting [row_begin, row_end)
        if (rows <= 0 || row0 >= end) return;
        const int lo = std::max(row_begin, row0), hi = std::min(row_end, end);
        if (lo >= hi) return;
        const int c_lo = (lo - row0) / rows, c_hi = (hi - 1 - row0) / rows;
        const int s = a.n_seg++;
        a.seg_row0[s] = row0 + c_lo * rows; a.seg_rows[s] = rows; a.seg_end[s] = end;
        a.seg_prefix[s + 1] = a.seg_prefix[s] + (c_hi - c_lo + 1);
    };
    for (int pass = 0; pass < 2; ++pass) {           // big items of every band first, the small ones after them
        for (int b = 0; b < lay.bands; ++b) {
            int lo, hi;
            lay.band_range(L, b, lo, hi);
            const int split = lay.rows_small ? lo + lay.n_big(hi - lo) * lay.rows_big : hi;
            if (pass == 0) push(lo, lay.rows_big, split);
            else push(split, lay.rows_small, hi);
        }
    }
    a.n_items = a.n_strips * a.seg_prefix[a.n_seg];
}

}  // namespace wt



namespace {



struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int fail(wt_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    set_thread_error(code, msg.c_str());
    return code;
}

#define WT_CUDA(ctx, expr)                                                                              \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) return fail(ctx, WT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// Builds the (3 (n+2)) x (2 (n+2)) RGBA8 atlas of a cubemap (tile column = face / 2, tile row = face % 2, every
// face surrounded by a 1-texel border replicating its edge: per-face GL_CLAMP_TO_EDGE) in a CUDA array created
// for texture gather, and a texture object over it (point access, unnormalised coordinates, unorm8 -> float).
cudaError_t make_sky_atlas(const uint8_t* const faces[6], int n, cudaArray_t* arr_out, cudaTextureObject_t* tex_out) {
    const int tile = n + 2, W = 3 * tile, H = 2 * tile;
    std::vector<uchar4> atlas((size_t)W * H);
    for (int f = 0; f < 6; ++f) {
        const uchar4* src = reinterpret_cast<const uchar4*>(faces[f]);
        const int x0 = (f / 2) * tile, y0 = (f % 2) * tile;
        for (int y = 0; y < tile; ++y) {
            const int sy = std::min(std::max(y - 1, 0), n - 1);
            uchar4* dst = atlas.data() + (size_t)(y0 + y) * W + x0;
            const uchar4* row = src + (size_t)sy * n;
            dst[0] = row[0];
            std::memcpy(dst + 1, row, (size_t)n * sizeof(uchar4));
            dst[tile - 1] = row[n - 1];
        }
    }
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    cudaError_t e = cudaMallocArray(arr_out, &fmt, W, H, cudaArrayTextureGather);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy2DToArray(*arr_out, 0, 0, atlas.data(), (size_t)W * sizeof(uchar4), (size_t)W * sizeof(uchar4), H, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    cudaResourceDesc res;
    std::memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = *arr_out;
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    return cudaCreateTextureObject(tex_out, &res, &td, nullptr);
}

template <int CAP>
void fill_table(WaveTable<CAP>& t, const std::vector<WaveK>& src) {
    std::memset(&t, 0, sizeof(t));
    std::memcpy(t.w, src.data(), src.size() * sizeof(WaveK));
}

template <int CAP, bool SHADE>
cudaError_t launch_reference_order(wt_ctx* ctx, EvalArgs& a, cudaStream_t stream) {
    static thread_local WaveTable<CAP> table;
    fill_table(table, ctx->tab_raw);
    const long long n = (long long)(a.row_end - a.row_begin) * a.L;
    int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    surface_reference_order_kernel<CAP, SHADE><<<std::max(blocks, 1), 256, 0, stream>>>(a, table);
    ctx->launches++;
    return cudaGetLastError();
}

template <int CAP, bool SHADE>
cudaError_t launch_stencil_order(wt_ctx* ctx, EvalArgs& a, cudaStream_t stream) {
    static thread_local WaveTable<CAP> table;
    fill_table(table, ctx->tab_exact);
    const long long pairs = ((long long)(a.row_end - a.row_begin) * a.L + 1) / 2;
    int blocks = (int)std::min<long long>((pairs + 127) / 128, (long long)ctx->sm_count * 32);
    surface_stencil_order_kernel<CAP, SHADE><<<std::max(blocks, 1), 128, 0, stream>>>(a, table);
    ctx->launches++;
    return cudaGetLastError();
}

cudaError_t launch_stencil(wt_ctx* ctx, EvalArgs& a, bool shade, cudaStream_t stream) {
    switch (wave_capacity(ctx->n_waves)) {
        case 64: return shade ? launch_stencil_order<64, true>(ctx, a, stream) : launch_stencil_order<64, false>(ctx, a, stream);
        case 256: return shade ? launch_stencil_order<256, true>(ctx, a, stream) : launch_stencil_order<256, false>(ctx, a, stream);
        default: return shade ? launch_stencil_order<1000, true>(ctx, a, stream) : launch_stencil_order<1000, false>(ctx, a, stream);
    }
}

// The index table of the stencil fix-up: i is flagged where the shader's neighbour argument fl(fl(i du) -/+ du) is not the
// neighbour's own fl((i -/+ 1) du) (water-grid.vert:123-125,164-167).  Same table for rows and columns; cached per grid size.
cudaError_t ensure_fixup_tables(wt_ctx* ctx, int L, float du) {
    if (ctx->fix_L == L && ctx->d_fix_flags) return cudaSuccess;
    std::vector<unsigned char> flags((size_t)L);
    std::vector<int> idx;
    long long sides = 0;
    for (int i = 0; i < L; ++i) {
        const float u = (float)i * du, lo = (float)(i - 1) * du, hi = (float)(i + 1) * du;
        const float minus = u - du, plus = u + du;
        flags[i] = (unsigned char)((minus != lo ? 1 : 0) | (plus != hi ? 2 : 0));
        if (flags[i]) idx.push_back(i);
        sides += (flags[i] & 1) + (flags[i] >> 1);
    }
    std::vector<int> perm;
    perm.reserve((size_t)L);
    for (int i = 0; i < L; ++i) if (!flags[i]) perm.push_back(i);
    perm.insert(perm.end(), idx.begin(), idx.end());
    auto drop = [&]() {
        cudaFree(ctx->d_fix_flags); cudaFree(ctx->d_fix_idx); cudaFree(ctx->d_fix_perm);
        ctx->d_fix_flags = nullptr; ctx->d_fix_idx = ctx->d_fix_perm = nullptr; ctx->fix_L = 0; ctx->fix_n = 0;
    };
    drop();
    cudaError_t e = cudaMalloc(&ctx->d_fix_flags, (size_t)L);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_fix_idx, std::max<size_t>(idx.size(), 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_fix_perm, (size_t)L * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_fix_flags, flags.data(), (size_t)L, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !idx.empty()) e = cudaMemcpy(ctx->d_fix_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_fix_perm, perm.data(), (size_t)L * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { drop(); return e; }
    ctx->fix_L = L; ctx->fix_n = (int)idx.size();
    ctx->fix_side_rate = (float)sides / (2.0f * (float)L);
    return cudaSuccess;
}

// Which form of WT_FLAG_STENCIL_ORDER runs on this grid (tables ensured).  The fix-up form costs about one production
// evaluation plus 1.6 (scalar code) per re-evaluated neighbour, 4 x side_rate of them per vertex; the one-kernel form 4.5.
// The rate is a property of du = 1 / (L - 1): 0.003 at 4096, 0.04 at 1024, 0.19 at 8192, 0.11 at 16384, ~0.5 for an odd size like 700.
bool stencil_runs_full(const wt_ctx* ctx, const EvalArgs& a) {
    if (a.clip) return false;               // WT_FLAG_VARYINGS: only the production kernel writes them
    if (ctx->stencil_form == 1) return true;
    if (ctx->stencil_form == 2) return false;
    return ctx->fix_side_rate > 0.3f;
}

template <int CAP, bool SHADE>
cudaError_t launch_stencil_fixup(wt_ctx* ctx, EvalArgs& a, cudaStream_t stream) {
    if (ctx->fix_n == 0) return cudaSuccess;
    static thread_local WaveTable<CAP> table;
    fill_table(table, ctx->tab_exact);
    const long long jobs = (long long)ctx->fix_n * (2LL * a.L - ctx->fix_n);
    const int blocks = (int)std::min<long long>((jobs + 63) / 64, (long long)ctx->sm_count * 24 * 4);   // one warp per CTA, 4 tiles of 64 jobs or more each
    const FixupArgs f = {ctx->d_fix_flags, ctx->d_fix_idx, ctx->d_fix_perm, ctx->fix_n};
    stencil_fixup_kernel<CAP, SHADE><<<std::max(blocks, 1), 32, 0, stream>>>(a, table, f);
    ctx->launches++;
    return cudaGetLastError();
}

// WT_FLAG_STENCIL_ORDER: the bit-faithful production kernel, then the fix-up pass over the rows and columns whose stencil
// arguments differ from their neighbours' own (see stencil_fixup_kernel).
cudaError_t launch_stencil_hybrid(wt_ctx* ctx, EvalArgs& a, bool shade, unsigned flags, cudaStream_t stream) {
    cudaError_t e = ensure_fixup_tables(ctx, a.L, a.du);
    if (e != cudaSuccess) return e;
    const unsigned pflags = flags & ~(WT_FLAG_STENCIL_ORDER | WT_FLAG_FAST_PHASE | WT_FLAG_AUTO_PHASE);
    e = shade ? launch_production_exact_shade(ctx, a, pflags, stream) : launch_production_exact(ctx, a, pflags, stream);
    if (e != cudaSuccess) return e;
    switch (wave_capacity(ctx->n_waves)) {
        case 64: return shade ? launch_stencil_fixup<64, true>(ctx, a, stream) : launch_stencil_fixup<64, false>(ctx, a, stream);
        case 256: return shade ? launch_stencil_fixup<256, true>(ctx, a, stream) : launch_stencil_fixup<256, false>(ctx, a, stream);
        default: return shade ? launch_stencil_fixup<1000, true>(ctx, a, stream) : launch_stencil_fixup<1000, false>(ctx, a, stream);
    }
}

cudaError_t launch_reference(wt_ctx* ctx, EvalArgs& a, bool shade, cudaStream_t stream) {
    switch (wave_capacity(ctx->n_waves)) {
        case 64: return shade ? launch_reference_order<64, true>(ctx, a, stream) : launch_reference_order<64, false>(ctx, a, stream);
        case 256: return shade ? launch_reference_order<256, true>(ctx, a, stream) : launch_reference_order<256, false>(ctx, a, stream);
        default: return shade ? launch_reference_order<1000, true>(ctx, a, stream) : launch_reference_order<1000, false>(ctx, a, stream);
    }
}

constexpr unsigned FIVE_EVAL_FLAGS = WT_FLAG_REFERENCE_ORDER | WT_FLAG_STENCIL_ORDER;   // the shader's own stencil arguments: more than one launch per frame, never grouped; digest in a device record

int validate_eval(wt_ctx* ctx, const wt_frame* frame, unsigned L, unsigned r0, unsigned r1, void* d_pos, void* d_nrm, void* d_shade,
                  void* d_cks, unsigned flags) {
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (!frame) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "frame is NULL");
    if (L < 2) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "grid_len must be >= 2 (water-grid.vert:53)");
    if (L > 32768) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "grid_len must be <= 32768");
    if (r0 >= r1 || r1 > L) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "need 0 <= row_begin < row_end <= grid_len");
    if (!d_pos || !d_nrm) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "position/normal output buffers are NULL");
    if ((flags & WT_FLAG_SHADE) && !d_shade) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "WT_FLAG_SHADE needs a shade buffer");
    if ((flags & WT_FLAG_CHECKSUM) && !d_cks) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "WT_FLAG_CHECKSUM needs a checksum buffer");
    if (!ctx->waves_set) return fail(ctx, WT_ERR_NOT_READY, "wt_set_waves has not been called");
    if (!ctx->d_hm_r || !ctx->d_hm_quads) return fail(ctx, WT_ERR_NOT_READY, "wt_set_heightmap_rgba8 has not been called (the reference draws no water without a heightmap, program.cpp:721)");
    return WT_OK;
}

EvalArgs make_args(const wt_ctx* ctx, const wt_frame& f, unsigned L, unsigned r0, unsigned r1, void* d_pos, void* d_nrm, void* d_shade,
                   void* d_cks, unsigned flags) {
    EvalArgs a;
    std::memset(&a, 0, sizeof(a));
    a.blx = f.bottomLeftGridPointInWorld[0]; a.blz = f.bottomLeftGridPointInWorld[2];
    a.brx = f.bottomRightGridPointInWorld[0]; a.brz = f.bottomRightGridPointInWorld[2];
    a.tlx = f.topLeftGridPointInWorld[0]; a.tlz = f.topLeftGridPointInWorld[2];
    a.trx = f.topRightGridPointInWorld[0]; a.trz = f.topRightGridPointInWorld[2];
    a.t = f.waveAnimationTimeInSeconds;
    a.bounce = f.verticalBounceWaveDisplacement;
    a.hm_disp = f.heightmapDisplacementScale;
    a.hm_scale = f.heightmapSampleScale;
    a.camx = f.cameraPosition[0]; a.camy = f.cameraPosition[1]; a.camz = f.cameraPosition[2];
    a.du = 1.0f / (float)(L - 1);   // water-grid.vert:123 (IEEE division on the host)
    a.sunx = f.sunPosition[0]; a.suny = f.sunPosition[1]; a.sunz = f.sunPosition[2];
    a.sun_shininess = f.sunShininess;
    a.sun_strength = f.sunStrength;
    a.n_waves = (int)ctx->n_waves;
    a.one = 1.0f;
    a.hm_r = ctx->d_hm_r; a.hm_quads = ctx->d_hm_quads; a.hm_w = ctx->hm_w; a.hm_h = ctx->hm_h;
    a.hm_pow2 = ((ctx->hm_w & (ctx->hm_w - 1)) == 0) && ((ctx->hm_h & (ctx->hm_h - 1)) == 0);
    {
        // largest heightmap sample coordinate this frame can produce, in texels: the grid is the bilinear hull of its corners,
        // a vertex moves horizontally by at most sum_i A_i Q_i |D_i| (ctx->max_horizontal, wt_set_waves)
        float reach = 0.0f;
        for (const float* c : {f.bottomLeftGridPointInWorld, f.bottomRightGridPointInWorld, f.topLeftGridPointInWorld, f.topRightGridPointInWorld})
            reach = std::fmax(reach, std::fmax(std::fabs(c[0]), std::fabs(c[2])));
        const float texels = (reach + ctx->max_horizontal + 1.0f) * std::fabs(f.heightmapSampleScale) * (float)std::max(ctx->hm_w, ctx->hm_h) + 2.0f;
        a.hm_magic_floor = texels < 2097152.0f;   // 2^21: half the range in which the magic-number floor is exact
    }
    a.sky = ctx->d_sky; a.sky_n = ctx->sky_n; a.sky_tex = ctx->sky_tex;
    if (!a.sky) { a.sky = ctx->d_sky_black; a.sky_n = 1; a.sky_tex = ctx->sky_tex_black; }
    a.pos = (float4*)d_pos; a.nrm = (float4*)d_nrm; a.shade = (float4*)d_shade;
    a.cks = (flags & WT_FLAG_CHECKSUM) ? (wt_checksum*)d_cks : nullptr;
    std::memcpy(a.vp, f.viewProjection, sizeof(a.vp));
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) a.vy[3 * c + r] = f.viewMatOnlyYaw[4 * c + r];
    a.L = (int)L; a.row_begin = (int)r0; a.row_end = (int)r1;
    return a;
}

// WT_FLAG_AUTO_PHASE for one frame: the fastest kernel whose deviation from the reference wt_fast_phase_report bounds inside
// the tolerances (this frame's corners and time): recurrence, else bit-faithful, else the stencil-order form.
unsigned resolve_phase(const wt_ctx* ctx, const EvalArgs& a, unsigned flags) {
    if (!(flags & WT_FLAG_AUTO_PHASE)) return flags;
    flags &= ~WT_FLAG_AUTO_PHASE;
    if (flags & FIVE_EVAL_FLAGS) return flags;
    wt_frame f;
    std::memset(&f, 0, sizeof(f));
    f.bottomLeftGridPointInWorld[0] = a.blx; f.bottomLeftGridPointInWorld[2] = a.blz;
    f.bottomRightGridPointInWorld[0] = a.brx; f.bottomRightGridPointInWorld[2] = a.brz;
    f.topLeftGridPointInWorld[0] = a.tlx; f.topLeftGridPointInWorld[2] = a.tlz;
    f.topRightGridPointInWorld[0] = a.trx; f.topRightGridPointInWorld[2] = a.trz;
    f.waveAnimationTimeInSeconds = a.t;
    wt_phase_report rep;
    if (wt_fast_phase_report(ctx->waves.data(), (unsigned)ctx->waves.size(), &f, (unsigned)a.L, &rep) < 0) return flags & ~WT_FLAG_FAST_PHASE;
    flags &= ~WT_FLAG_FAST_PHASE;
    if (rep.recommended == WT_PHASE_RECURRENCE) flags |= WT_FLAG_FAST_PHASE;
    else if (rep.recommended == WT_PHASE_STENCIL_ORDER) flags |= WT_FLAG_STENCIL_ORDER;
    return flags;
}

cudaError_t launch_eval(wt_ctx* ctx, EvalArgs& a, unsigned flags, cudaStream_t stream) {
    const bool shade = flags & WT_FLAG_SHADE;
    ctx->last_phase_mode = (flags & WT_FLAG_REFERENCE_ORDER) ? WT_PHASE_REFERENCE_ORDER : (flags & WT_FLAG_STENCIL_ORDER) ? WT_PHASE_STENCIL_ORDER
                           : (flags & WT_FLAG_FAST_PHASE) ? WT_PHASE_RECURRENCE : WT_PHASE_BIT_FAITHFUL;
    ctx->mode_launches[ctx->last_phase_mode]++;
    if (flags & WT_FLAG_REFERENCE_ORDER) return launch_reference(ctx, a, shade, stream);
    if (flags & WT_FLAG_STENCIL_ORDER) {
        const cudaError_t e = ensure_fixup_tables(ctx, a.L, a.du);
        if (e != cudaSuccess) return e;
        return stencil_runs_full(ctx, a) ? launch_stencil(ctx, a, shade, stream) : launch_stencil_hybrid(ctx, a, shade, flags, stream);
    }
    if (flags & WT_FLAG_FAST_PHASE)
        return shade ? launch_production_recur_shade(ctx, a, flags, stream) : launch_production_recur(ctx, a, flags, stream);
    return shade ? launch_production_exact_shade(ctx, a, flags, stream) : launch_production_exact(ctx, a, flags, stream);
}

// One frame: resolves WT_FLAG_AUTO_PHASE, launches the kernel and — for the reference-order kernel, which only accumulates
// into the digest record — arms the record before and decodes it after (the production kernels do both themselves).
cudaError_t launch_frame(wt_ctx* ctx, EvalArgs& a, unsigned flags, cudaStream_t stream) {
    flags = resolve_phase(ctx, a, flags);
    // (WT_FLAG_STENCIL_ORDER is the production kernel + the fix-up pass, which corrects the finished record in place)
    bool accumulates = (flags & WT_FLAG_REFERENCE_ORDER) != 0;
    if (flags & WT_FLAG_STENCIL_ORDER) {
        const cudaError_t e = ensure_fixup_tables(ctx, a.L, a.du);
        if (e != cudaSuccess) return e;
        accumulates = stencil_runs_full(ctx, a);
    }
    const bool ref_cks = a.cks && accumulates;
    if (ref_cks) { checksum_init_kernel<<<1, 32, 0, stream>>>(a.cks, 1); ctx->launches++; }
    const cudaError_t e = launch_eval(ctx, a, flags, stream);
    if (e != cudaSuccess) return e;
    if (ref_cks) { checksum_finalize_kernel<<<1, 32, 0, stream>>>(a.cks, 1); ctx->launches++; }
    return cudaGetLastError();
}

}  // namespace

extern "C" {

const char* wt_version(void) { return "wavetool_b200 0.1.0 (sm_100a)"; }

const char* wt_last_error(const wt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_thread_error.c_str(); }

int wt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int wt_create(wt_ctx** out, int device) {
    if (!out) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_create: out is NULL");
    *out = nullptr;
    int n = wt_device_count();
    if (n <= 0) return fail(nullptr, WT_ERR_NO_DEVICE, "wt_create: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_create: device index out of range");
    DeviceGuard g(device);
    cudaDeviceProp prop;
    WT_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(nullptr, WT_ERR_NO_DEVICE, "wt_create: device is not sm_100 (kernels are built for sm_100a only)");
    wt_ctx* ctx = new wt_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // every allocation of the context; on any failure the half-built context is destroyed again
    auto init = [&]() -> cudaError_t {
        cudaError_t e;
        if ((e = cudaMalloc(&ctx->s_cks, sizeof(wt_checksum))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ctx->d_sky_black, 6 * sizeof(uchar4))) != cudaSuccess) return e;
        if ((e = cudaMemset(ctx->d_sky_black, 0, 6 * sizeof(uchar4))) != cudaSuccess) return e;
        const uint8_t black[4] = {0, 0, 0, 0};
        const uint8_t* const faces[6] = {black, black, black, black, black, black};
        if ((e = make_sky_atlas(faces, 1, &ctx->sky_array_black, &ctx->sky_tex_black)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ctx->d_cks_acc, N_QUEUES * sizeof(wt_checksum))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ctx->d_cks_done, N_QUEUES * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMemset(ctx->d_cks_done, 0, N_QUEUES * sizeof(int))) != cudaSuccess) return e;
        checksum_init_kernel<<<(N_QUEUES + 127) / 128, 128>>>(ctx->d_cks_acc, (int)N_QUEUES);   // armed: zero sums, min/max keys at their identities
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ctx->d_queues, 2 * N_QUEUES * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMemset(ctx->d_queues, 0, 2 * N_QUEUES * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ctx->d_batch_done, BATCH_MAX_FRAMES * N_QUEUES * sizeof(int))) != cudaSuccess) return e;
        return cudaMemset(ctx->d_batch_done, 0, BATCH_MAX_FRAMES * N_QUEUES * sizeof(int));
    };
    const cudaError_t e = init();
    if (e != cudaSuccess) {
        wt_destroy(ctx);
        return fail(nullptr, WT_ERR_CUDA, std::string("wt_create: ") + cudaGetErrorString(e));
    }
    *out = ctx;
    return WT_OK;
}

void wt_destroy(wt_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaFree(ctx->d_hm_r); cudaFree(ctx->d_hm_quads);
    cudaFree(ctx->d_sky); cudaFree(ctx->d_sky_black);
    if (ctx->sky_surf) cudaDestroySurfaceObject(ctx->sky_surf);
    if (ctx->sky_tex) cudaDestroyTextureObject(ctx->sky_tex);
    if (ctx->sky_tex_black) cudaDestroyTextureObject(ctx->sky_tex_black);
    if (ctx->sky_array) cudaFreeArray(ctx->sky_array);
    if (ctx->sky_array_black) cudaFreeArray(ctx->sky_array_black);
    cudaFree(ctx->s_pos); cudaFree(ctx->s_nrm); cudaFree(ctx->s_shade); cudaFree(ctx->s_cks);
    cudaFree(ctx->d_queues); cudaFree(ctx->d_trace); cudaFree(ctx->d_batch_done);
    cudaFree(ctx->d_cks_acc); cudaFree(ctx->d_cks_done); cudaFree(ctx->d_step_cks);
    cudaFree(ctx->d_fix_flags); cudaFree(ctx->d_fix_idx); cudaFree(ctx->d_fix_perm);
    cudaFree(ctx->d_depth);
    if (ctx->refl_tex) cudaDestroyTextureObject(ctx->refl_tex);
    if (ctx->refr_tex) cudaDestroyTextureObject(ctx->refr_tex);
    if (ctx->refl_arr) cudaFreeArray(ctx->refl_arr);
    if (ctx->refr_arr) cudaFreeArray(ctx->refr_arr);
    cudaFree(ctx->d_stars); cudaFree(ctx->d_clouds); cudaFree(ctx->d_sky_unsure);
    if (ctx->h_cks_pinned) cudaFreeHost(ctx->h_cks_pinned);
    for (unsigned i = 0; i < STEP_RING; ++i) if (ctx->step_event[i]) cudaEventDestroy(ctx->step_event[i]);
    for (int i = 0; i < 4; ++i) {
        if (ctx->aux_stream[i]) cudaStreamDestroy(ctx->aux_stream[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_step_fork) cudaEventDestroy(ctx->ev_step_fork);
    delete ctx;
}

int wt_set_waves(wt_ctx* ctx, const wt_wave* waves, unsigned count) {
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_set_waves: ctx is NULL");
    if (count > WT_MAX_WAVES) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_set_waves: count exceeds WT_MAX_WAVES");
    if (count && !waves) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_set_waves: waves is NULL");
    for (unsigned i = 0; i < count; ++i) {   // the reference's constructor asserts, render-engine.h:147-154
        const wt_wave& w = waves[i];
        float len = std::sqrt(w.dir_x * w.dir_x + w.dir_z * w.dir_z);
        if (!(w.amplitude_A >= 0.0f) || !(w.frequency_w >= 0.0f) || !(w.phaseConstant_phi >= 0.0f) || !(w.steepness_Q >= 0.0f && w.steepness_Q <= 1.0f) ||
            !(len >= 1.0f - 0.001f && len <= 1.0f + 0.001f))
            return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_set_waves: wave " + std::to_string(i) + " violates the GerstnerWave constraints (render-engine.h:147-154)");
    }
    ctx->tab_exact.assign(count, WaveK{});
    ctx->tab_fast.assign(count, WaveK{});
    ctx->tab_raw.assign(count, WaveK{});
    for (unsigned i = 0; i < count; ++i) {
        const wt_wave& w = waves[i];
        const float Qi = wt_steepness_Q_i(&w, count);   // render-engine.cpp:1073
        // horizontal coefficient A*Q_i*D folded in double and rounded once (the shader rounds three
        // times: A * (D * (Q_i * cos)); the difference is <= 2 ulp of a term bounded by Q/(w*count))
        const float cX = (float)((double)w.amplitude_A * (double)Qi * (double)w.dir_x);
        const float cZ = (float)((double)w.amplitude_A * (double)Qi * (double)w.dir_z);
        ctx->tab_exact[i] = {w.dir_x, w.dir_z, w.frequency_w, w.phaseConstant_phi, cX, w.amplitude_A, cZ, 0.f};
        ctx->tab_fast[i] = {(float)((double)w.frequency_w * w.dir_x), (float)((double)w.frequency_w * w.dir_z), w.frequency_w, w.phaseConstant_phi, cX, w.amplitude_A, cZ, 0.f};
        ctx->tab_raw[i] = {w.dir_x, w.dir_z, w.frequency_w, w.phaseConstant_phi, Qi, w.amplitude_A, 0.f, 0.f};
    }
    ctx->max_horizontal = 0.0f;
    for (unsigned i = 0; i < count; ++i) ctx->max_horizontal += std::fabs(ctx->tab_exact[i].cX) + std::fabs(ctx->tab_exact[i].cZ);
    ctx->waves.assign(waves, waves + count);
    ctx->n_waves = count;
    ctx->waves_set = true;
    return WT_OK;
}

int wt_fast_phase_valid(const wt_ctx* ctx, const wt_frame* frame, unsigned L, wt_phase_report* report) {
    if (!ctx || !frame) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_fast_phase_valid: NULL argument");
    if (!ctx->waves_set) return fail(nullptr, WT_ERR_NOT_READY, "wt_fast_phase_valid: wt_set_waves has not been called");
    return wt_fast_phase_report(ctx->waves.data(), (unsigned)ctx->waves.size(), frame, L, report);
}

int wt_last_phase_mode(const wt_ctx* ctx) { return ctx ? ctx->last_phase_mode : -1; }

void wt_phase_mode_counts(const wt_ctx* ctx, uint64_t out[4]) {
    if (!out) return;
    for (int k = 0; k < 4; ++k) out[k] = ctx ? ctx->mode_launches[k] : 0;
}

int wt_set_heightmap_rgba8(wt_ctx* ctx, const uint8_t* rgba, int w, int h) {
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_set_heightmap_rgba8: ctx is NULL");
    // the footprint table takes 64 bytes per texel: 4096 x 4096 texels = 1 GiB is the most it is allowed to grow to
    if (!rgba || w <= 0 || h <= 0 || w > 8192 || h > 8192 || (long long)w * h > (1LL << 24))
        return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_set_heightmap_rgba8: bad image (need 1 <= width, height <= 8192 and width * height <= 2^24)");
    DeviceGuard g(ctx->device);
    // water-grid.vert:106-111 reads only .r; stage it as the exact unorm value c/255.0f
    std::vector<float> r((size_t)w * h);
    for (size_t i = 0; i < r.size(); ++i) r[i] = (float)rgba[4 * i] / 255.0f;
    // build into temporaries; the context changes only after every step has succeeded
    float* d_r = nullptr;
    float4* d_quads = nullptr;
    auto build = [&]() -> cudaError_t {
        cudaError_t e;
        if ((e = cudaMalloc(&d_r, r.size() * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMemcpy(d_r, r.data(), r.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&d_quads, 4 * r.size() * sizeof(float4))) != cudaSuccess) return e;
        build_heightmap_quads_kernel<<<ctx->sm_count * 4, 256>>>(d_r, w, h, d_quads);
        ctx->launches++;
        return cudaDeviceSynchronize();
    };
    const cudaError_t e = build();
    if (e != cudaSuccess) {
        cudaFree(d_r); cudaFree(d_quads);
        cudaGetLastError();
        return fail(ctx, WT_ERR_CUDA, std::string("wt_set_heightmap_rgba8: ") + cudaGetErrorString(e) + " (the previous heightmap, if any, is still in place)");
    }
    cudaFree(ctx->d_hm_r); cudaFree(ctx->d_hm_quads);
    ctx->d_hm_r = d_r; ctx->d_hm_quads = d_quads;
    ctx->hm_w = w; ctx->hm_h = h;
    return WT_OK;
}

int wt_set_skybox_rgba8(wt_ctx* ctx, const uint8_t* const faces[6], int n) {
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_set_skybox_rgba8: ctx is NULL");
    if (!faces || n <= 0 || n > 16384) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_set_skybox_rgba8: bad cubemap");
    for (int f = 0; f < 6; ++f) if (!faces[f]) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_set_skybox_rgba8: face pointer is NULL");
    DeviceGuard g(ctx->device);
    // build into temporaries; the context changes only after every step has succeeded
    const size_t face_bytes = (size_t)n * n * 4;
    uchar4* d_sky = nullptr;
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    auto build = [&]() -> cudaError_t {
        cudaError_t e;
        if ((e = cudaMalloc(&d_sky, 6 * face_bytes)) != cudaSuccess) return e;
        for (int f = 0; f < 6; ++f)
            if ((e = cudaMemcpy((uint8_t*)d_sky + f * face_bytes, faces[f], face_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        return make_sky_atlas(faces, n, &arr, &tex);
    };
    const cudaError_t e = build();
    if (e != cudaSuccess) {
        cudaFree(d_sky);
        if (tex) cudaDestroyTextureObject(tex);
        if (arr) cudaFreeArray(arr);
        cudaGetLastError();
        return fail(ctx, WT_ERR_CUDA, std::string("wt_set_skybox_rgba8: ") + cudaGetErrorString(e) + " (the previous skybox, if any, is still in place)");
    }
    cudaFree(ctx->d_sky);
    if (ctx->sky_surf) { cudaDestroySurfaceObject(ctx->sky_surf); ctx->sky_surf = 0; }
    if (ctx->sky_tex) cudaDestroyTextureObject(ctx->sky_tex);
    if (ctx->sky_array) cudaFreeArray(ctx->sky_array);
    ctx->d_sky = d_sky; ctx->sky_array = arr; ctx->sky_tex = tex;
    ctx->sky_n = n;
    ctx->sky_surface_ok = false;   // this array was not created for surface writes: a later synthesis re-creates it
    return WT_OK;
}

int wt_eval_frame(wt_ctx* ctx, const wt_frame* frame, unsigned L, unsigned r0, unsigned r1, void* d_pos, void* d_nrm, void* d_shade,
                  void* d_cks, unsigned flags, void* stream_) {
    int st = validate_eval(ctx, frame, L, r0, r1, d_pos, d_nrm, d_shade, d_cks, flags);
    if (st != WT_OK) return st;
    DeviceGuard g(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    EvalArgs a = make_args(ctx, *frame, L, r0, r1, d_pos, d_nrm, d_shade, d_cks, flags);
    WT_CUDA(ctx, launch_frame(ctx, a, flags, stream));
    return WT_OK;
}

int wt_eval_frame_ex(wt_ctx* ctx, const wt_frame* frame, unsigned L, unsigned r0, unsigned r1, const wt_outputs* out, unsigned flags, void* stream_) {
    if (!out) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_eval_frame_ex: outputs is NULL");
    int st = validate_eval(ctx, frame, L, r0, r1, out->position, out->normal, out->shade, out->checksum, flags);
    if (st != WT_OK) return st;
    const bool vary = flags & WT_FLAG_VARYINGS;
    if (vary && (!out->clip_position || !out->varyings_a || !out->varyings_b))
        return fail(ctx, WT_ERR_INVALID_ARGUMENT, "WT_FLAG_VARYINGS needs the clip_position, varyings_a and varyings_b buffers");
    if (vary && (flags & WT_FLAG_REFERENCE_ORDER))
        return fail(ctx, WT_ERR_INVALID_ARGUMENT, "WT_FLAG_VARYINGS is not available with WT_FLAG_REFERENCE_ORDER");
    DeviceGuard g(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    EvalArgs a = make_args(ctx, *frame, L, r0, r1, out->position, out->normal, out->shade, out->checksum, flags);
    if (vary) { a.clip = (float4*)out->clip_position; a.vary_a = (float4*)out->varyings_a; a.vary_b = (float4*)out->varyings_b; }
    WT_CUDA(ctx, launch_frame(ctx, a, flags, stream));
    return WT_OK;
}

// water mesh tessellation, src/program.cpp:686-717: two counter-clockwise triangles per grid cell, cells in row-major
// order from the bottom-left, vertex id = row * grid_len + col
int wt_grid_indices(unsigned L, uint32_t* out, uint64_t capacity, uint64_t* count) {
    if (L < 2 || L > 32768) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_grid_indices: grid_len must be in [2, 32768]");
    const uint64_t n = 6ull * (L - 1) * (L - 1);
    if (count) *count = n;
    if (!out) return WT_OK;
    if (capacity < n) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_grid_indices: capacity too small");
    uint64_t k = 0;
    for (unsigned row = 0; row + 1 < L; ++row) {
        for (unsigned col = 0; col + 1 < L; ++col) {
            const uint32_t bl = row * L + col, br = bl + 1, tl = bl + L, tr = tl + 1;
            out[k++] = bl; out[k++] = tr; out[k++] = tl;
            out[k++] = bl; out[k++] = br; out[k++] = tr;
        }
    }
    return WT_OK;
}

}  // extern "C"

namespace {
// Batches of SMALL frames (a 1024^2 frame is ~1300 warp items: half the machine, a few microseconds): consecutive
// frames go to alternating internal streams so that their kernels overlap and one frame's tail is filled by the
// next frame's head.  Frames that share a ring slot stay on one stream (k divides ring_depth), so every slot is
// still written in frame order.  Fork/join with events: the caller's stream ordering is preserved.
struct FanOut {
    wt_ctx* ctx;
    cudaStream_t user;
    int k = 1;
    FanOut(wt_ctx* c, cudaStream_t u, unsigned ring_depth, unsigned n_frames, unsigned L, unsigned rows) : ctx(c), user(u) {
        const unsigned long long verts = (unsigned long long)L * rows;
        // large frames too (measured at 4096^2 x 64 waves + shading: 0.684 -> 0.659 ms per frame with two streams): a frame's
        // persistent grid runs dry over its last round of work items, the next frame's CTAs take the SMs it frees
        int want = verts <= (1ull << 20) ? 4 : 2;
        if (g_knobs.fanout >= 1) want = std::min(g_knobs.fanout, 4);
        while (want > 1 && (ring_depth % want != 0 || n_frames < (unsigned)want)) --want;
        k = want;
    }
    cudaError_t begin() {
        if (k == 1) return cudaSuccess;
        cudaError_t e;
        if (!ctx->ev_fork && (e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return e;
        for (int i = 0; i < k; ++i) {
            if (!ctx->aux_stream[i] && (e = cudaStreamCreateWithFlags(&ctx->aux_stream[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
            if (!ctx->ev_join[i] && (e = cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        if ((e = cudaEventRecord(ctx->ev_fork, user)) != cudaSuccess) return e;
        for (int i = 0; i < k; ++i) if ((e = cudaStreamWaitEvent(ctx->aux_stream[i], ctx->ev_fork, 0)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    cudaStream_t stream_for(unsigned frame) const { return k == 1 ? user : ctx->aux_stream[frame % k]; }
    cudaError_t end() {
        if (k == 1) return cudaSuccess;
        cudaError_t e;
        for (int i = 0; i < k; ++i) {
            if ((e = cudaEventRecord(ctx->ev_join[i], ctx->aux_stream[i])) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(user, ctx->ev_join[i], 0)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
};

// Groups consecutive SMALL frames of a batch / an animation into single persistent launches (wt_launch.cuh: launch_warp_batch).
// A frame joins a group when it resolves to one of the two production kernels without shading terms or extra varyings; a group
// holds frames of ONE kernel that write different ring slots (at most ring_depth of them).  Everything else is launched on its
// own, after the pending group, so the stream order of the frames is kept.
struct FrameBatcher {
    wt_ctx* ctx;
    cudaStream_t stream;
    unsigned ring_depth;
    bool enabled;
    // Two lanes: the groups of the lower and of the upper half of the ring go to two internal streams, so that the tail of one
    // group (a persistent grid running dry) overlaps the head of the next.  A group never straddles the halves and a ring slot is
    // always written from the same lane: every slot is still written in frame order.  half == 0: one lane, the caller's stream.
    unsigned half = 0;
    FanOut* fan = nullptr;
    std::vector<EvalArgs> group;
    unsigned group_flags = 0, group_slot0 = 0;
    FrameBatcher(wt_ctx* c, cudaStream_t s, unsigned ring, unsigned n_frames, unsigned L, unsigned rows, unsigned flags)
        : ctx(c), stream(s), ring_depth(ring),
          enabled(ring >= 2 && n_frames >= 2 && c->n_waves <= 64 && (unsigned long long)L * rows <= (1ull << 22) && g_knobs.fanout != 1 &&
                  !(flags & (FIVE_EVAL_FLAGS | WT_FLAG_SHADE | WT_FLAG_VARYINGS))) {
        if (enabled && ring >= 4 && ring % 2 == 0 && n_frames > ring / 2) half = ring / 2;
    }
    int lanes() const { return half ? 2 : 1; }
    cudaStream_t lane_stream(unsigned slot) const { return half && fan ? fan->stream_for(slot / half) : stream; }
    cudaError_t flush() {
        if (group.empty()) return cudaSuccess;
        cudaError_t e;
        const cudaStream_t lane = lane_stream(group_slot0);
        if (group.size() == 1) {
            e = launch_frame(ctx, group[0], group_flags, lane);
        } else {
            const int mode = (group_flags & WT_FLAG_FAST_PHASE) ? WT_PHASE_RECURRENCE : WT_PHASE_BIT_FAITHFUL;
            if (group[0].cks) {   // the frames' digest records are consecutive: arm them (zero sums, min / max keys at their identities)
                checksum_init_kernel<<<1, 32, 0, lane>>>(group[0].cks, (int)group.size());
                ctx->launches++;
            }
            ctx->last_phase_mode = mode;
            ctx->mode_launches[mode] += group.size();
            ctx->batched_frames += group.size();
            e = mode == WT_PHASE_RECURRENCE ? launch_production_batch_recur(ctx, group, lane) : launch_production_batch_exact(ctx, group, lane);
        }
        group.clear();
        return e;
    }
    // `a` with flags as the caller passed them (WT_FLAG_AUTO_PHASE unresolved); slot: the frame's ring slot; other_stream: where a
    // frame that cannot be grouped goes when nothing is grouped at all
    cudaError_t add(EvalArgs& a, unsigned flags, unsigned slot, cudaStream_t other_stream) {
        const unsigned rflags = resolve_phase(ctx, a, flags);
        const bool groupable = enabled && !(rflags & (FIVE_EVAL_FLAGS | WT_FLAG_SHADE | WT_FLAG_VARYINGS));
        cudaError_t e;
        if (!groupable) {
            if ((e = flush()) != cudaSuccess) return e;
            return launch_frame(ctx, a, rflags, enabled ? lane_stream(slot) : other_stream);
        }
        const size_t cap = std::min<size_t>(half ? half : ring_depth, (size_t)BATCH_MAX_FRAMES);
        if (!group.empty() && (group.size() >= cap || (half && slot % half == 0) || ((rflags ^ group_flags) & WT_FLAG_FAST_PHASE) ||
                               (group[0].cks && a.cks != group[0].cks + group.size())))
            if ((e = flush()) != cudaSuccess) return e;
        if (group.empty()) group_slot0 = slot;
        group.push_back(a);
        group_flags = rflags;
        return cudaSuccess;
    }
};
}  // namespace

extern "C" {

int wt_eval_batch(wt_ctx* ctx, const wt_frame* frames, unsigned n_frames, unsigned L, unsigned r0, unsigned r1, void* d_pos_ring,
                  void* d_nrm_ring, void* d_shade_ring, unsigned ring_depth, void* d_checksums, unsigned flags, void* stream_) {
    if (n_frames == 0) return WT_OK;
    if (ring_depth == 0) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_eval_batch: ring_depth must be >= 1");
    int st = validate_eval(ctx, frames, L, r0, r1, d_pos_ring, d_nrm_ring, d_shade_ring, d_checksums, flags);
    if (st != WT_OK) return st;
    DeviceGuard g(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool cks = flags & WT_FLAG_CHECKSUM;
    const size_t slot = (size_t)(r1 - r0) * L;
    FrameBatcher batcher(ctx, stream, ring_depth, n_frames, L, r1 - r0, flags);
    FanOut fan(ctx, stream, batcher.enabled ? batcher.lanes() : ring_depth, n_frames, L, r1 - r0);   // grouped frames: one stream per lane
    batcher.fan = &fan;
    if (fan.k != batcher.lanes()) batcher.half = 0;
    WT_CUDA(ctx, fan.begin());
    for (unsigned f = 0; f <= n_frames; ++f) {
        cudaError_t e;
        if (f < n_frames) {
            const size_t off = (size_t)(f % ring_depth) * slot;
            EvalArgs a = make_args(ctx, frames[f], L, r0, r1, (float4*)d_pos_ring + off, (float4*)d_nrm_ring + off,
                                   d_shade_ring ? (float4*)d_shade_ring + off : nullptr, cks ? (wt_checksum*)d_checksums + f : nullptr, flags);
            e = batcher.add(a, flags, f % ring_depth, fan.stream_for(f));
        } else {
            e = batcher.flush();
        }
        if (e != cudaSuccess) {   // the internal streams were forked off the caller's: join them before reporting
            fan.end();
            return fail(ctx, WT_ERR_CUDA, std::string("wt_eval_batch: ") + cudaGetErrorString(e));
        }
    }
    WT_CUDA(ctx, fan.end());
    WT_CUDA(ctx, cudaGetLastError());
    return WT_OK;
}

int wt_run_animation(wt_ctx* ctx, const wt_camera* cam, wt_params* params, const wt_animation* anim, float total_amp, float dt,
                     unsigned stride, unsigned n_frames, unsigned L, unsigned r0, unsigned r1, void* d_pos_ring, void* d_nrm_ring,
                     void* d_shade_ring, unsigned ring_depth, void* d_checksums, unsigned flags, void* stream_, unsigned* n_evaluated) {
    if (n_evaluated) *n_evaluated = 0;
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_run_animation: ctx is NULL");
    if (!cam || !params || !anim) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_run_animation: NULL argument");
    if (ring_depth == 0 || stride == 0) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_run_animation: ring_depth and frame_stride must be >= 1");
    if (n_frames == 0) return WT_OK;
    wt_frame probe;
    std::memset(&probe, 0, sizeof(probe));
    int st = validate_eval(ctx, &probe, L, r0, r1, d_pos_ring, d_nrm_ring, d_shade_ring, d_checksums, flags);
    if (st != WT_OK) return st;
    DeviceGuard g(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool cks = flags & WT_FLAG_CHECKSUM;
    const size_t slot = (size_t)(r1 - r0) * L;
    unsigned drawn = 0;
    FrameBatcher batcher(ctx, stream, ring_depth, n_frames, L, r1 - r0, flags);
    FanOut fan(ctx, stream, batcher.enabled ? batcher.lanes() : ring_depth, n_frames, L, r1 - r0);   // grouped frames: one stream per lane
    batcher.fan = &fan;
    if (fan.k != batcher.lanes()) batcher.half = 0;
    WT_CUDA(ctx, fan.begin());
    // after begin() the internal streams have been forked off the caller's: every exit path joins them again
    auto bail = [&](int code, const std::string& msg) { fan.end(); return fail(ctx, code, msg); };
    for (unsigned k = 0; k < n_frames; ++k) {
        wt_frame frame;
        st = wt_build_frame(cam, params, total_amp, &frame);
        if (st < 0) return bail(st, std::string("wt_run_animation: ") + wt_last_error(nullptr));
        if (st == WT_OK) {
            const size_t off = (size_t)(k % ring_depth) * slot;
            EvalArgs a = make_args(ctx, frame, L, r0, r1, (float4*)d_pos_ring + off, (float4*)d_nrm_ring + off,
                                   d_shade_ring ? (float4*)d_shade_ring + off : nullptr, cks ? (wt_checksum*)d_checksums + k : nullptr, flags);
            const cudaError_t e = batcher.add(a, flags, k % ring_depth, fan.stream_for(k));
            if (e != cudaSuccess) return bail(WT_ERR_CUDA, std::string("wt_run_animation: ") + cudaGetErrorString(e));
            ++drawn;
        } else if (cks) {   // nothing drawn: zero digest (after the pending group, which may hold the neighbouring records)
            if (batcher.flush() != cudaSuccess || cudaMemsetAsync((wt_checksum*)d_checksums + k, 0, sizeof(wt_checksum), batcher.enabled ? stream : fan.stream_for(k)) != cudaSuccess)
                return bail(WT_ERR_CUDA, "wt_run_animation: zeroing a digest failed");
        }
        for (unsigned s_ = 0; s_ < stride; ++s_) wt_animation_advance(params, anim, dt);
    }
    {
        const cudaError_t e = batcher.flush();
        if (e != cudaSuccess) return bail(WT_ERR_CUDA, std::string("wt_run_animation: ") + cudaGetErrorString(e));
    }
    WT_CUDA(ctx, fan.end());
    WT_CUDA(ctx, cudaGetLastError());
    if (n_evaluated) *n_evaluated = drawn;
    return WT_OK;
}

// One whole frame from HOST inputs to a HOST result: camera + parameters -> wt_build_frame (the reference's CPU
// prologue) -> fused kernel into the caller's device buffers, whose last warp writes the frame's digest straight into
// a mapped host record.  wt_step_submit enqueues a frame and returns a ticket, wt_step_wait blocks until that frame's
// digest has arrived: with up to STEP_RING frames in flight the host prepares frame k+1 while the GPU evaluates
// frame k, as a render loop does.  wt_step_host = submit + wait.
int wt_step_submit(wt_ctx* ctx, const wt_camera* cam, const wt_params* params, float total_amp, unsigned L, unsigned r0, unsigned r1,
                   void* d_pos, void* d_nrm, void* d_shade, unsigned flags, void* stream_, unsigned* ticket) {
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_step_submit: ctx is NULL");
    if (!cam || !params || !ticket) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_step_submit: NULL argument");
    wt_frame frame;
    int st = wt_build_frame(cam, params, total_amp, &frame);
    if (st < 0) return fail(ctx, st, std::string("wt_step_submit: ") + wt_last_error(nullptr));
    DeviceGuard g(ctx->device);
    if (!ctx->h_cks_pinned) {
        // the events first: the mapped records are only published once everything a ticket needs exists
        for (unsigned i = 0; i < STEP_RING; ++i)
            if (!ctx->step_event[i]) WT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->step_event[i], cudaEventDisableTiming));
        wt_checksum* h = nullptr;
        WT_CUDA(ctx, cudaHostAlloc(&h, STEP_RING * sizeof(wt_checksum), cudaHostAllocMapped));
        ctx->h_cks_pinned = h;
    }
    // a ticket is only handed out for a frame that really was enqueued (or that legitimately draws nothing): on an
    // evaluation error nothing is consumed and no slot is left waiting for an event that was never recorded
    const unsigned t = ctx->step_next;
    const unsigned slot = t % STEP_RING;
    const cudaStream_t user = (cudaStream_t)stream_;
    // Two lanes: consecutive steps run on two internal streams forked off the caller's, so that the tail of one frame's persistent
    // grid overlaps the head of the next (measured at 4096^2 x 64 waves + shading: 0.684 -> 0.659 ms per frame).  The step sees
    // everything enqueued on the caller's stream before it; its results are ordered by wt_step_wait (the ticket's event).  A step
    // that writes the position buffer of a pending step of the other lane waits for that step first.
    const int lane = (int)(t & 1u);
    if (!ctx->aux_stream[lane]) WT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream[lane], cudaStreamNonBlocking));
    if (!ctx->ev_step_fork) WT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_step_fork, cudaEventDisableTiming));
    cudaStream_t stream = ctx->aux_stream[lane];
    void* const lane_stream = (void*)stream;
    if (st == WT_OK) {
        WT_CUDA(ctx, cudaEventRecord(ctx->ev_step_fork, user));
        WT_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_step_fork, 0));
        for (unsigned back = 1; back < STEP_RING && back <= t; ++back) {
            const unsigned ps = (t - back) % STEP_RING;
            if ((back & 1u) && ctx->step_buffer[ps] == d_pos) WT_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->step_event[ps], 0));
        }
        int ev = validate_eval(ctx, &frame, L, r0, r1, d_pos, d_nrm, d_shade, ctx->h_cks_pinned + slot, flags | WT_FLAG_CHECKSUM);
        if (ev != WT_OK) return ev;
        EvalArgs probe = make_args(ctx, frame, L, r0, r1, d_pos, d_nrm, d_shade, nullptr, flags);
        const unsigned rflags = resolve_phase(ctx, probe, flags) | WT_FLAG_CHECKSUM;
        if (rflags & FIVE_EVAL_FLAGS) {
            // the five-evaluation kernels accumulate its digest with atomics: into a device record, copied out afterwards
            if (!ctx->d_step_cks) WT_CUDA(ctx, cudaMalloc(&ctx->d_step_cks, STEP_RING * sizeof(wt_checksum)));
            ev = wt_eval_frame(ctx, &frame, L, r0, r1, d_pos, d_nrm, d_shade, ctx->d_step_cks + slot, rflags, lane_stream);
            if (ev != WT_OK) return ev;
            WT_CUDA(ctx, cudaMemcpyAsync(ctx->h_cks_pinned + slot, ctx->d_step_cks + slot, sizeof(wt_checksum), cudaMemcpyDeviceToHost, stream));
        } else {
            // the production kernels' last warp writes the finished digest straight into the mapped host record: no copy to enqueue
            ev = wt_eval_frame(ctx, &frame, L, r0, r1, d_pos, d_nrm, d_shade, ctx->h_cks_pinned + slot, rflags, lane_stream);
            if (ev != WT_OK) return ev;
        }
    }
    ctx->step_buffer[slot] = st == WT_OK ? d_pos : nullptr;
    WT_CUDA(ctx, cudaEventRecord(ctx->step_event[slot], stream));
    ctx->step_status[slot] = st;
    ctx->step_next = t + 1;
    *ticket = t;
    return st;
}

int wt_step_wait(wt_ctx* ctx, unsigned ticket, wt_checksum* h_out) {
    if (!ctx || !h_out) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_step_wait: NULL argument");
    if (!ctx->h_cks_pinned || ticket >= ctx->step_next || ctx->step_next - ticket > STEP_RING)
        return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_step_wait: unknown ticket (at most 4 frames may be in flight)");
    const unsigned slot = ticket % STEP_RING;
    DeviceGuard g(ctx->device);
    WT_CUDA(ctx, cudaEventSynchronize(ctx->step_event[slot]));
    if (ctx->step_status[slot] == WT_NO_GRID) { std::memset(h_out, 0, sizeof(*h_out)); return WT_NO_GRID; }
    *h_out = ctx->h_cks_pinned[slot];
    return WT_OK;
}

int wt_step_host(wt_ctx* ctx, const wt_camera* cam, const wt_params* params, float total_amp, unsigned L, unsigned r0, unsigned r1,
                 void* d_pos, void* d_nrm, void* d_shade, unsigned flags, wt_checksum* h_out, void* stream_) {
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_step_host: ctx is NULL");
    if (!h_out) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_step_host: NULL argument");
    unsigned ticket = 0;
    int st = wt_step_submit(ctx, cam, params, total_amp, L, r0, r1, d_pos, d_nrm, d_shade, flags, stream_, &ticket);
    if (st < 0) return st;
    return wt_step_wait(ctx, ticket, h_out);
}

int wt_eval_frame_host(wt_ctx* ctx, const wt_frame* frame, unsigned L, unsigned r0, unsigned r1, float* h_pos, float* h_nrm, float* h_shade,
                       wt_checksum* h_cks, unsigned flags) {
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (L < 2 || r0 >= r1 || r1 > L) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "need grid_len >= 2 and 0 <= row_begin < row_end <= grid_len");
    DeviceGuard g(ctx->device);
    const size_t n = (size_t)(r1 - r0) * L;
    if (n > ctx->s_elems) {
        cudaFree(ctx->s_pos); cudaFree(ctx->s_nrm); cudaFree(ctx->s_shade);
        ctx->s_pos = ctx->s_nrm = ctx->s_shade = nullptr;
        ctx->s_elems = 0;
        WT_CUDA(ctx, cudaMalloc(&ctx->s_pos, n * sizeof(float4)));
        WT_CUDA(ctx, cudaMalloc(&ctx->s_nrm, n * sizeof(float4)));
        WT_CUDA(ctx, cudaMalloc(&ctx->s_shade, n * sizeof(float4)));
        ctx->s_elems = n;
    }
    if (h_cks) flags |= WT_FLAG_CHECKSUM; else flags &= ~WT_FLAG_CHECKSUM;
    int st = wt_eval_frame(ctx, frame, L, r0, r1, ctx->s_pos, ctx->s_nrm, ctx->s_shade, ctx->s_cks, flags, nullptr);
    if (st != WT_OK) return st;
    if (h_pos) WT_CUDA(ctx, cudaMemcpy(h_pos, ctx->s_pos, n * sizeof(float4), cudaMemcpyDeviceToHost));
    if (h_nrm) WT_CUDA(ctx, cudaMemcpy(h_nrm, ctx->s_nrm, n * sizeof(float4), cudaMemcpyDeviceToHost));
    if (h_shade && (flags & WT_FLAG_SHADE)) WT_CUDA(ctx, cudaMemcpy(h_shade, ctx->s_shade, n * sizeof(float4), cudaMemcpyDeviceToHost));
    if (h_cks) WT_CUDA(ctx, cudaMemcpy(h_cks, ctx->s_cks, sizeof(wt_checksum), cudaMemcpyDeviceToHost));
    WT_CUDA(ctx, cudaDeviceSynchronize());
    return WT_OK;
}

uint64_t wt_launch_count(const wt_ctx* ctx) { return ctx ? ctx->launches : 0; }

uint64_t wt_batched_frame_count(const wt_ctx* ctx) { return ctx ? ctx->batched_frames : 0; }

void wt_debug_reload_env(void) { g_knobs.load(); }

int wt_debug_full_stencil(wt_ctx* ctx, int on) {
    if (!ctx) return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_debug_full_stencil: ctx is NULL");
    if (on < 0 || on > 2) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_debug_full_stencil: 0 (automatic), 1 (five evaluations per vertex) or 2 (fix-up form)");
    ctx->stencil_form = on;
    return WT_OK;
}

int wt_debug_layout(unsigned L, int rows_per_step, int phase_mode, unsigned row_begin, unsigned row_end, int* out, unsigned capacity) {
    if (L < 2 || L > 32768 || rows_per_step < 1 || row_begin >= row_end || row_end > L || !out)
        return fail(nullptr, WT_ERR_INVALID_ARGUMENT, "wt_debug_layout: bad argument");
    EvalArgs a;
    std::memset(&a, 0, sizeof(a));
    layout_segments(canonical_layout((int)L, rows_per_step, phase_mode), (int)L, (int)row_begin, (int)row_end, a);
    const int n = a.seg_prefix[a.n_seg];
    if ((unsigned)n > capacity) return n;
    int k = 0;
    for (int s = 0; s < a.n_seg; ++s)
        for (int q = a.seg_prefix[s]; q < a.seg_prefix[s + 1]; ++q, ++k) {
            const int c0 = a.seg_row0[s] + (q - a.seg_prefix[s]) * a.seg_rows[s];
            out[2 * k] = c0;
            out[2 * k + 1] = std::min(c0 + a.seg_rows[s], a.seg_end[s]);
        }
    return n;
}

int wt_debug_read_trace(wt_ctx* ctx, uint64_t* out, unsigned max_items) {
    if (!ctx || !out) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_debug_read_trace: NULL argument");
    DeviceGuard g(ctx->device);
    const size_t n = std::min<size_t>(ctx->trace_items, max_items);
    WT_CUDA(ctx, cudaDeviceSynchronize());
    if (n) WT_CUDA(ctx, cudaMemcpy(out, ctx->d_trace, n * 32, cudaMemcpyDeviceToHost));
    return (int)n;
}

int wt_measure_peak(wt_ctx* ctx, int kind, double* per_second) {
    if (!ctx || !per_second) return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_measure_peak: NULL argument");
    DeviceGuard g(ctx->device);
    cudaEvent_t e0, e1;
    WT_CUDA(ctx, cudaEventCreate(&e0));
    WT_CUDA(ctx, cudaEventCreate(&e1));
    float* sink = nullptr;
    WT_CUDA(ctx, cudaMalloc(&sink, 256));
    double best = 0.0;
    const int blocks = ctx->sm_count * 8, threads = 256;
    if (kind == 0 || kind == 1 || kind == 3) {
        const int iters = 4096;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (kind == 0) peak_mufu_kernel<<<blocks, threads>>>(sink, iters);
            else if (kind == 1) peak_fma_kernel<<<blocks, threads>>>(sink, iters);
            else peak_fma2_kernel<<<blocks, threads>>>(sink, iters);
            cudaEventRecord(e1);
            WT_CUDA(ctx, cudaEventSynchronize(e1));
            ctx->launches++;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            // kind 0: sin+cos PAIRS per second (2 MUFU ops each); kind 1: FLOP/s (2 per FMA)
            double work = (double)blocks * threads * 8.0 * iters * (kind == 0 ? 1.0 : kind == 1 ? 2.0 : 4.0);
            if (rep > 0) best = std::max(best, work / (ms * 1e-3));
        }
    } else if (kind == 2) {
        const size_t n = (size_t)1 << 26;   // 64 Mi float4 = 1 GiB
        float4* buf = nullptr;
        WT_CUDA(ctx, cudaMalloc(&buf, n * sizeof(float4)));
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            peak_store_kernel<<<ctx->sm_count * 16, 256>>>(buf, n);
            cudaEventRecord(e1);
            WT_CUDA(ctx, cudaEventSynchronize(e1));
            ctx->launches++;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0) best = std::max(best, (double)n * sizeof(float4) / (ms * 1e-3));
        }
        cudaFree(buf);
    } else {
        cudaFree(sink);
        return fail(ctx, WT_ERR_INVALID_ARGUMENT, "wt_measure_peak: kind must be 0, 1, 2 or 3");
    }
    cudaFree(sink);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *per_second = best;
    return WT_OK;
}

}  // extern "C"
