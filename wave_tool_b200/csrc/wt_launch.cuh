// Launch code of surface_warp_kernel, included by the four wt_launch_*.cu translation units.
#pragma once
#include <algorithm>
#include <cmath>

#include "wt_ctx.h"

namespace wt {

template <int CAP>
void fill_table(WaveTable<CAP>& t, const std::vector<WaveK>& src) {
    std::memset(&t, 0, sizeof(t));
    std::memcpy(t.w, src.data(), src.size() * sizeof(WaveK));
}

template <int CAP, int MODE, bool SHADE, bool CKS, bool HMP2, bool VARY>
cudaError_t launch_warp_vary(wt_ctx* ctx, EvalArgs& a, const std::vector<WaveK>& tab, cudaStream_t stream) {
    constexpr int R = MODE == 1 ? (CAP > 64 ? R_RECUR_BIG : R_RECUR) : R_EXACT;
    auto kernel = surface_warp_kernel<CAP, R, MODE, SHADE, CKS, HMP2, VARY>;
    const int smem = WARPS_PER_CTA * (RingLayout<R, VARY>::BYTES_PER_WARP + (CKS && MODE == 0 ? RingLayout<R, VARY>::CKS_BYTES_PER_WARP : 0)) + g_knobs.extra_smem;
    // per instantiation: occupancy and the segments of the last (grid, row range) seen
    static thread_local struct { int L = -1, r0 = -1, r1 = -1, dev = -1, slots = 0; unsigned gen = 0; EvalArgs seg; } cache;
    if (cache.L != a.L || cache.r0 != a.row_begin || cache.r1 != a.row_end || cache.dev != ctx->device || cache.gen != g_knobs.generation) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * WARPS_PER_CTA, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        cache.slots = ctx->sm_count * per_sm * WARPS_PER_CTA;
        layout_segments(canonical_layout(a.L, R, MODE), a.L, a.row_begin, a.row_end, cache.seg);
        cache.L = a.L; cache.r0 = a.row_begin; cache.r1 = a.row_end; cache.dev = ctx->device; cache.gen = g_knobs.generation;
    }
    a.n_strips = cache.seg.n_strips; a.n_seg = cache.seg.n_seg; a.n_items = cache.seg.n_items;
    std::memcpy(a.seg_prefix, cache.seg.seg_prefix, sizeof(a.seg_prefix));
    std::memcpy(a.seg_row0, cache.seg.seg_row0, sizeof(a.seg_row0));
    std::memcpy(a.seg_rows, cache.seg.seg_rows, sizeof(a.seg_rows));
    std::memcpy(a.seg_end, cache.seg.seg_end, sizeof(a.seg_end));
    // Work distribution.  Few items (fewer than two per resident warp): one item per CTA.  Otherwise a persistent grid
    // pulls items off an atomic queue — the warp schedulers favour some resident warps heavily, so equal static shares
    // finish far apart (profiles/README.md); with queued items every warp stays busy until the queue runs dry.  Which
    // warp takes which item never changes a result.
    int slots = cache.slots;
    if (g_knobs.dyn_warps > 0) slots = std::min(slots, ctx->sm_count * g_knobs.dyn_warps);
    int ctas = a.n_items;
    a.queue = nullptr;
    if ((long long)a.n_items >= 2LL * slots && g_knobs.dyn != 0) {
        ctas = slots;
        a.queue = ctx->d_queues + 2 * (ctx->queue_next++ % N_QUEUES);
    }
    a.trace = nullptr;
    {
        if (g_knobs.trace) {
            if (ctx->trace_cap < (size_t)a.n_items) {
                cudaFree(ctx->d_trace);
                ctx->d_trace = nullptr;
                if (cudaMalloc(&ctx->d_trace, (size_t)a.n_items * 32) == cudaSuccess) ctx->trace_cap = a.n_items;
            }
            a.trace = ctx->d_trace;
            ctx->trace_items = a.trace ? a.n_items : 0;
        }
    }
    static thread_local WaveTable<CAP> table;
    fill_table(table, tab);
    for (size_t i = 0; i < tab.size(); ++i) {
        const float c = tab[i].phi * a.t;   // fl(phi * t), water-grid.vert:78
        // MODE 1 adds the offset inside an FMA chain: hand it over reduced mod 2pi (in double) so it costs no precision
        table.w[i].phi = MODE == 1 ? (float)std::remainder((double)c, 6.283185307179586476925) : c;
    }
    a.dbg = g_knobs.dbg;
    if (CKS) {
        const unsigned slot = ctx->cks_next++ % N_QUEUES;
        a.cks_acc = ctx->d_cks_acc + slot;
        a.cks_done = ctx->d_cks_done + slot;
    }
    kernel<<<ctas, 32 * WARPS_PER_CTA, smem, stream>>>(a, table, NoBatch{0});
    ctx->launches++;
    return cudaGetLastError();
}

template <int CAP, int MODE, bool SHADE, bool CKS, bool HMP2>
cudaError_t launch_warp_hm(wt_ctx* ctx, EvalArgs& a, const std::vector<WaveK>& tab, cudaStream_t stream) {
    return a.clip ? launch_warp_vary<CAP, MODE, SHADE, CKS, HMP2, true>(ctx, a, tab, stream) : launch_warp_vary<CAP, MODE, SHADE, CKS, HMP2, false>(ctx, a, tab, stream);
}

template <int CAP, int MODE, bool SHADE, bool CKS>
cudaError_t launch_warp_cks(wt_ctx* ctx, EvalArgs& a, const std::vector<WaveK>& tab, cudaStream_t stream) {
    // non-power-of-two heightmaps take the instantiation with the out-of-line modulo
    return a.hm_pow2 ? launch_warp_hm<CAP, MODE, SHADE, CKS, true>(ctx, a, tab, stream) : launch_warp_hm<CAP, MODE, SHADE, CKS, false>(ctx, a, tab, stream);
}

template <int CAP, int MODE, bool SHADE>
cudaError_t launch_warp(wt_ctx* ctx, EvalArgs& a, const std::vector<WaveK>& tab, cudaStream_t stream) {
    return a.cks ? launch_warp_cks<CAP, MODE, SHADE, true>(ctx, a, tab, stream) : launch_warp_cks<CAP, MODE, SHADE, false>(ctx, a, tab, stream);
}


// Several SMALL frames in ONE persistent launch (wt_eval_batch / wt_run_animation; <= 64 waves, no shading terms, no extra
// varyings): the work items of all frames go through one queue, so the SMs stay busy across frame boundaries and the host
// pays one launch per group instead of one per frame.  The frames' own arguments (corners, time offsets, outputs, digest
// record) travel as a third kernel parameter (BatchTable, constant bank, indexed by the frame of the work item).
// The frames of a group must write different ring slots (the caller groups them so) and share grid_len / row range.
template <int MODE, bool CKS, bool HMP2>
cudaError_t launch_warp_batch(wt_ctx* ctx, std::vector<EvalArgs>& fa, const std::vector<WaveK>& tab, cudaStream_t stream) {
    constexpr int CAP = 64;
    constexpr int R = MODE == 1 ? R_RECUR : R_EXACT;
    auto kernel = surface_warp_kernel<CAP, R, MODE, false, CKS, HMP2, false, true>;
    const int F = (int)fa.size();
    if (F < 1 || F > BATCH_MAX_FRAMES) return cudaErrorInvalidValue;
    const int smem = WARPS_PER_CTA * (RingLayout<R, false>::BYTES_PER_WARP + (CKS && MODE == 0 ? RingLayout<R, false>::CKS_BYTES_PER_WARP : 0));
    static thread_local struct { int dev = -1, slots = 0; } cache;
    if (cache.dev != ctx->device) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * WARPS_PER_CTA, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        cache.slots = ctx->sm_count * per_sm * WARPS_PER_CTA;
        cache.dev = ctx->device;
    }
    EvalArgs& a0 = fa[0];
    layout_segments(canonical_layout(a0.L, R, MODE), a0.L, a0.row_begin, a0.row_end, a0);
    static thread_local BatchTable bt;
    for (int f = 0; f < F; ++f) {
        EvalArgs& a = fa[f];
        a.n_strips = a0.n_strips; a.n_seg = a0.n_seg; a.n_items = a0.n_items;
        std::memcpy(a.seg_prefix, a0.seg_prefix, sizeof(a.seg_prefix));
        std::memcpy(a.seg_row0, a0.seg_row0, sizeof(a.seg_row0));
        std::memcpy(a.seg_rows, a0.seg_rows, sizeof(a.seg_rows));
        std::memcpy(a.seg_end, a0.seg_end, sizeof(a.seg_end));
        a.queue = nullptr; a.trace = nullptr; a.dbg = 0;
        bt.fa[f] = a;
        for (size_t i = 0; i < tab.size(); ++i) {
            const float c = tab[i].phi * a.t;   // fl(phi * t), water-grid.vert:78
            bt.phi[f][i] = MODE == 1 ? (float)std::remainder((double)c, 6.283185307179586476925) : c;
        }
    }
    EvalArgs top = a0;
    top.batch_done = ctx->d_batch_done + BATCH_MAX_FRAMES * (ctx->batch_next++ % N_QUEUES);   // self re-arming, like the work queues
    top.items_per_frame = a0.n_items;
    top.n_items = a0.n_items * F;
    int ctas = top.n_items;
    top.queue = nullptr;
    if ((long long)top.n_items >= 2LL * cache.slots) {
        ctas = cache.slots;
        top.queue = ctx->d_queues + 2 * (ctx->queue_next++ % N_QUEUES);
    }
    static thread_local WaveTable<CAP> table;
    fill_table(table, tab);
    kernel<<<ctas, 32 * WARPS_PER_CTA, smem, stream>>>(top, table, bt);
    ctx->launches++;
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_production_batch(wt_ctx* ctx, std::vector<EvalArgs>& fa, cudaStream_t stream) {
    const std::vector<WaveK>& tab = MODE == 1 ? ctx->tab_fast : ctx->tab_exact;
    const bool cks = fa[0].cks != nullptr, p2 = fa[0].hm_pow2 != 0;
    if (cks) return p2 ? launch_warp_batch<MODE, true, true>(ctx, fa, tab, stream) : launch_warp_batch<MODE, true, false>(ctx, fa, tab, stream);
    return p2 ? launch_warp_batch<MODE, false, true>(ctx, fa, tab, stream) : launch_warp_batch<MODE, false, false>(ctx, fa, tab, stream);
}

// One (MODE, SHADE) slice of the dispatch: wave-table capacity tier, checksum, heightmap wrap.
template <int MODE, bool SHADE>
cudaError_t launch_production(wt_ctx* ctx, EvalArgs& a, cudaStream_t stream) {
    const std::vector<WaveK>& tab = MODE == 1 ? ctx->tab_fast : ctx->tab_exact;
    switch (wave_capacity(ctx->n_waves)) {
        case 64: return launch_warp<64, MODE, SHADE>(ctx, a, tab, stream);
        case 256: return launch_warp<256, MODE, SHADE>(ctx, a, tab, stream);
        default: return launch_warp<1000, MODE, SHADE>(ctx, a, tab, stream);
    }
}

}  // namespace wt
