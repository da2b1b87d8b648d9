// CUDA-side internals shared by wt_device.cu and the wt_launch_*.cu translation units: the context, the
// development knobs and the launch planning of surface_warp_kernel.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "wt_internal.h"
#include "wt_kernels.cuh"

struct wt_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    // wave tables (host copies; handed to the kernels as __grid_constant__ parameters per launch)
    unsigned n_waves = 0;
    bool waves_set = false;
    std::vector<wt::WaveK> tab_exact, tab_fast, tab_raw;
    std::vector<wt_wave> waves;     // as handed to wt_set_waves (wt_fast_phase_valid)
    float max_horizontal = 0.0f;    // sum_i |A_i Q_i D_i|_1: the furthest a vertex can move sideways
    int last_phase_mode = -1;       // WT_PHASE_*: what the most recent evaluation launch ran
    // stencil fix-up tables of the grid size last used with WT_FLAG_STENCIL_ORDER (they depend on grid_len alone)
    unsigned char* d_fix_flags = nullptr;
    int *d_fix_idx = nullptr, *d_fix_perm = nullptr;
    int fix_L = 0, fix_n = 0;
    float fix_side_rate = 0.0f;     // probability that one of a vertex's four stencil arguments is flagged (set bits / 2 L)
    int stencil_form = 0;           // wt_debug_full_stencil: 0 = whichever is cheaper for the grid, 1 = five evaluations per vertex, 2 = fix-up form
    uint64_t mode_launches[4] = {0, 0, 0, 0};   // evaluation launches per kernel, indexed by WT_PHASE_* (wt_phase_mode_counts)
    // textures
    float* d_hm_r = nullptr;        // R channel, c/255.0f (reference-order kernel)
    float4* d_hm_quads = nullptr;   // mirrored-repeat footprint table (production kernel)
    int hm_w = 0, hm_h = 0;
    uchar4* d_sky = nullptr;        // 6 faces RGBA8 (reference-order kernel)
    uchar4* d_sky_black = nullptr;  // 1x1 black cubemap: shading without a skybox reads sky = 0, no branch in the production kernel
    int sky_n = 0;
    // production kernel: padded face atlas in a block-linear CUDA array + a gather-capable texture object
    cudaArray_t sky_array = nullptr, sky_array_black = nullptr;
    cudaTextureObject_t sky_tex = 0, sky_tex_black = 0;
    // sources of the skybox synthesis (wt_set_sky_layers); sky_surface_ok: sky_array was created for surface writes
    std::vector<uint8_t> sky_gradient;
    uchar4 *d_stars = nullptr, *d_clouds = nullptr;
    int stars_n = 0, clouds_n = 0;
    bool sky_surface_ok = false;
    unsigned long long* d_sky_unsure = nullptr;   // texels of the last synthesis that took the exact path (wt_debug_sky_exact_texels)
    cudaSurfaceObject_t sky_surf = 0;
    // screen-space inputs of the compose pass (wt_set_screen_textures)
    float* d_depth = nullptr;
    // reflection / refraction colour textures as gather-able arrays (texture unit: footprint fetch + exact c / 255)
    cudaArray_t refl_arr = nullptr, refr_arr = nullptr;
    cudaTextureObject_t refl_tex = 0, refr_tex = 0;
    int screen_w = 0, screen_h = 0;
    // scratch for the *_host entry points
    float4 *s_pos = nullptr, *s_nrm = nullptr, *s_shade = nullptr;
    size_t s_elems = 0;
    wt_checksum* s_cks = nullptr;
    wt_checksum* d_step_cks = nullptr;     // device records of wt_step_submit frames that run the reference-order kernel
    wt_checksum* h_cks_pinned = nullptr;   // mapped host records of wt_step_host / wt_step_submit (STEP_RING + 1 of them)
    cudaEvent_t step_event[4] = {nullptr, nullptr, nullptr, nullptr};
    int step_status[4] = {0, 0, 0, 0};
    unsigned step_next = 0;
    const void* step_buffer[4] = {nullptr, nullptr, nullptr, nullptr};   // position buffer each pending step writes (lane hazards, wt_step_submit)
    cudaEvent_t ev_step_fork = nullptr;
    uint64_t launches = 0;
    uint64_t batched_frames = 0;    // frames evaluated inside grouped launches (wt_launch.cuh: launch_warp_batch)
    // dynamic work distribution: a ring of self re-arming {next item, warps done} counters, one per launch in flight
    int* d_queues = nullptr;
    unsigned queue_next = 0;
    int* d_batch_done = nullptr;    // grouped launches: per group BATCH_MAX_FRAMES finished-item counters (self re-arming)
    unsigned batch_next = 0;
    // small frames: internal streams on which consecutive frames of a batch overlap (created on first use)
    cudaStream_t aux_stream[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    // digests: a ring of pre-armed accumulator records + finished-warp counters, one per production launch in flight
    wt_checksum* d_cks_acc = nullptr;
    int* d_cks_done = nullptr;
    unsigned cks_next = 0;
    // WT_TRACE diagnostic: per work item {smid, cta, t_start, t_end} of the most recent production launch
    unsigned long long* d_trace = nullptr;
    size_t trace_cap = 0, trace_items = 0;
};
constexpr unsigned N_QUEUES = 256;
constexpr unsigned STEP_RING = 4;   // frames in flight between wt_step_submit and wt_step_wait

namespace wt {

// Development knobs.  They exist only in builds compiled with -DWT_DEV_KNOBS (tools/build_variant.py); the
// production library ignores the environment: its launch plans are a function of the grid alone.
//   WT_DYN        rows per dynamically queued work item (0 = static plan; unset = built-in policy)
//   WT_DYN_SMALL  rows per item of the tail of the dynamic plan (0 = single level), WT_DYN_TAIL = percent of the rows it covers
//   WT_DYN_WARPS  cap on the resident warps per SM of the dynamic plan
//   WT_TRACE      record per-work-item {SM, CTA, start, end} (wt_debug_read_trace)
//   WT_DBG        bit 0: wave loop only (skips heightmap and output stage; outputs are NOT written)
//   WT_EXTRA_SMEM extra dynamic shared memory per CTA (caps the CTAs per SM)
//   WT_FANOUT     internal streams used for batches of small frames (1 = none; unset = built-in policy)
struct Knobs {
    int dyn = -1, dyn_small = -1, dyn_tail = -1, dyn_warps = 0, trace = 0, dbg = 0, extra_smem = 0, fanout = -1;
    unsigned generation = 0;   // bumped by every load(): cached launch plans are keyed on it
    void load() {
#ifdef WT_DEV_KNOBS
        auto geti = [](const char* name, int dflt) { const char* e = std::getenv(name); return e && *e ? std::atoi(e) : dflt; };
        dyn = geti("WT_DYN", -1); dyn_small = geti("WT_DYN_SMALL", -1); dyn_tail = geti("WT_DYN_TAIL", -1); dyn_warps = geti("WT_DYN_WARPS", 0); trace = geti("WT_TRACE", 0); dbg = geti("WT_DBG", 0);
        extra_smem = geti("WT_EXTRA_SMEM", 0); fanout = geti("WT_FANOUT", -1);
#endif
        ++generation;
    }
    Knobs() { load(); }
};
extern Knobs g_knobs;

// ------------------------------------------------------------------------------------------------
// Launch planning for surface_warp_kernel
// ------------------------------------------------------------------------------------------------
#ifndef WT_R_EXACT
#define WT_R_EXACT 4
#endif
#ifndef WT_R_RECUR
#define WT_R_RECUR 8
#endif
constexpr int R_EXACT = WT_R_EXACT;   // rows per marching step, MODE 0 (keeps the undisplaced grid point of every row live)
constexpr int DYN_ROWS_RECUR = 46;   // rows per dynamically queued work item, MODE 1 (+2 halo rows = 6 marching steps of 8 rows, 4 of 12)
constexpr int DYN_ROWS_EXACT = 30;   // ... MODE 0 (8 marching steps)
constexpr int DYN_SMALL_RECUR = 14, DYN_SMALL_EXACT = 0;   // rows per item of the tail of the launch (0 = single level)
constexpr int DYN_TAIL_PERCENT = 15;   // measured at cfg3 (profiles/README.md): 38/14/25 % 0.4075 ms, 46/14/25 % 0.4026, 46/14/15 % 0.3982
constexpr int R_RECUR = WT_R_RECUR;   // rows per marching step, MODE 1 (one MUFU seed per wave, row 1 by rotation, rows 2.. by the three-term recurrence)
#ifndef WT_R_RECUR_BIG
#define WT_R_RECUR_BIG 12
#endif
constexpr int R_RECUR_BIG = WT_R_RECUR_BIG;   // ... for the wave-table tiers above 64 waves: the wave loop dominates there, and 12 rows per seed at
                                              // 12 warps per SM beat 8 rows at 16 (cfg4: 3.83 vs 4.27 ms; at 64 waves 0.412 vs 0.405 ms)

struct Plan {
    int n_strips = 1, rows_per_chunk = 1, n_chunks = 1, ctas = 1;
};

Plan make_plan(int slots, int L, int rows, int R);

// Canonical work layout of a grid: which rows form one work item.  A function of (grid_len, rows per marching
// step, phase mode) ONLY — not of the device, the occupancy or the row range of the call — because in the
// recurrence mode the rounding of a vertex depends on where its item's marching steps start: with a canonical
// layout every vertex has the same bits whatever the shard, the stream or the GPU it is evaluated on.
//   The rows are cut into `bands` equal bands (the row shards of an 8-GPU box, wave_tool_b200/shard.py:row_shard);
//   each band holds n_big items of rows_big rows followed by items of rows_small rows (the tail that keeps every
//   warp busy until the queue is empty); the last item of a band is clipped at the band's end.
constexpr int NOMINAL_SMS = 148;     // B200; the layout must not depend on the device actually used
#ifndef WT_LAYOUT_BANDS
#define WT_LAYOUT_BANDS 8
#endif
constexpr int LAYOUT_BANDS = WT_LAYOUT_BANDS;
constexpr int MAX_SEGS = 2 * LAYOUT_BANDS;
struct Layout {
    int n_strips = 1, bands = 1, rows_big = 1, rows_small = 0, tail_pct = 0;
    void band_range(int L, int b, int& lo, int& hi) const {
        const int base = L / bands, rem = L % bands;
        lo = b * base + (b < rem ? b : rem);
        hi = lo + base + (b < rem ? 1 : 0);
    }
    int n_big(int band_rows) const {
        if (rows_small <= 0) return (band_rows + rows_big - 1) / rows_big;
        return (int)((long long)band_rows * (100 - tail_pct) / 100 / rows_big);
    }
};
Layout canonical_layout(int L, int R, int mode);
// The work items of rows [row_begin, row_end) under `lay`, as segments of equally sized chunks (big items of all
// bands first, then the small ones): fills the seg_* fields, n_seg and n_items of `a`.
void layout_segments(const Layout& lay, int L, int row_begin, int row_end, EvalArgs& a);

inline int wave_capacity(unsigned n) { return n <= 64 ? 64 : n <= 256 ? 256 : 1000; }

// The production kernel's instantiations are spread over four translation units (MODE x SHADE) so that
// they compile in parallel (wave_tool_b200/build.py); each defines one of these.
cudaError_t launch_production_exact(wt_ctx* ctx, EvalArgs& a, unsigned flags, cudaStream_t stream);
cudaError_t launch_production_exact_shade(wt_ctx* ctx, EvalArgs& a, unsigned flags, cudaStream_t stream);
cudaError_t launch_production_recur(wt_ctx* ctx, EvalArgs& a, unsigned flags, cudaStream_t stream);
cudaError_t launch_production_recur_shade(wt_ctx* ctx, EvalArgs& a, unsigned flags, cudaStream_t stream);
// several small frames in one persistent launch (wt_launch.cuh: launch_warp_batch)
cudaError_t launch_production_batch_exact(wt_ctx* ctx, std::vector<EvalArgs>& frames, cudaStream_t stream);
cudaError_t launch_production_batch_recur(wt_ctx* ctx, std::vector<EvalArgs>& frames, cudaStream_t stream);

}  // namespace wt
