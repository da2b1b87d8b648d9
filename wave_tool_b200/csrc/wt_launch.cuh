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
    const int smem = WARPS_PER_CTA * RingLayout<R, VARY>::BYTES_PER_WARP + g_knobs.extra_smem;
    static thread_local struct { int L = -1, rows = -1, dev = -1, slots = 0; Plan p; } cache;   // per instantiation
    const int rows = a.row_end - a.row_begin;
    if (cache.L != a.L || cache.rows != rows || cache.dev != ctx->device) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * WARPS_PER_CTA, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        cache.slots = ctx->sm_count * per_sm * WARPS_PER_CTA;
        cache.p = make_plan(cache.slots, a.L, rows, R);
        cache.L = a.L; cache.rows = rows; cache.dev = ctx->device;
    }
    Plan p = cache.p;
    // Dynamic work distribution: the warp schedulers favour some resident warps heavily, so equal static
    // shares finish far apart (profiles/README.md); with small items handed out through an atomic queue every
    // warp stays busy until the queue runs dry.  Used when there are enough items to go round several times;
    // WT_DYN overrides the rows per item (0 = static plan).
    const bool forced = g_knobs.dyn >= 0;
    const int dyn_rows = forced ? g_knobs.dyn : (MODE == 1 ? DYN_ROWS_RECUR : DYN_ROWS_EXACT);
    a.queue = nullptr;
    if (dyn_rows > 0) {
        auto round_rows = [](int want) { int r = ((want + 2 + R - 1) / R) * R - 2; return r < 1 ? R : r; };   // rows + 2 a multiple of R
        const int rpc = round_rows(dyn_rows);
        // two levels: the last `tail` percent of the rows go out in small items, after all the big ones
        const int small_want = g_knobs.dyn_small >= 0 ? g_knobs.dyn_small : (MODE == 1 ? DYN_SMALL_RECUR : DYN_SMALL_EXACT);
        const int tail_pct = g_knobs.dyn_tail >= 0 ? g_knobs.dyn_tail : DYN_TAIL_PERCENT;
        int n_big = (rows + rpc - 1) / rpc, n_small = 0, rps = 0;
        if (small_want > 0 && small_want < dyn_rows && tail_pct > 0) {
            rps = round_rows(small_want);
            n_big = (int)((long long)rows * (100 - tail_pct) / 100 / rpc);
            n_small = (rows - n_big * rpc + rps - 1) / rps;
        }
        const long long items = (long long)p.n_strips * (n_big + n_small);
        int slots = cache.slots;
        if (g_knobs.dyn_warps > 0) slots = std::min(slots, ctx->sm_count * g_knobs.dyn_warps);
        if (forced || items >= 2LL * slots) {
            p.rows_per_chunk = rpc;
            p.n_chunks = n_big + n_small;
            p.n_big = n_big; p.rows_small = rps;
            p.ctas = (int)std::min<long long>(items, slots);
            a.queue = ctx->d_queues + 2 * (ctx->queue_next++ % N_QUEUES);
        }
    }
    a.n_strips = p.n_strips;
    a.rows_per_chunk = p.rows_per_chunk;
    a.n_chunks = p.n_chunks;
    a.n_big = p.n_big < 0 ? p.n_chunks : p.n_big;
    a.rows_small = p.rows_small > 0 ? p.rows_small : p.rows_per_chunk;
    a.n_items = p.n_strips * p.n_chunks;
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
    kernel<<<p.ctas, 32 * WARPS_PER_CTA, smem, stream>>>(a, table);
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
