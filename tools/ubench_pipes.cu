// Development micro-benchmark: do the XU (MUFU) and FMA pipes of an sm_100a sub-partition overlap when one
// warp interleaves FFMA2/FFMA with MUFU.SIN/COS?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NF2, int NMUFU, int NF1>
__global__ void mix_kernel(float* sink, int iters) {
    float2 x[8];
    float m[8], y[8];
    const float2 a = make_float2(1.0000001f, 0.9999999f), b = make_float2(1e-9f * threadIdx.x, 1e-9f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = make_float2(0.001f * (threadIdx.x + 1) + 0.1f * j, 0.5f + 0.1f * j); m[j] = 0.01f * j + 0.001f * threadIdx.x; y[j] = m[j]; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < NF2; ++j) x[(r * NF2 + j) & 7] = __ffma2_rn(x[(r * NF2 + j) & 7], a, b);
#pragma unroll
            for (int j = 0; j < NF1; ++j) y[(r * NF1 + j) & 7] = __fmaf_rn(y[(r * NF1 + j) & 7], a.x, b.x);
#pragma unroll
            for (int j = 0; j < NMUFU; ++j) {
                float r_;
                asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(r_) : "f"(m[(r * NMUFU + j) & 7]));
                m[(r * NMUFU + j) & 7] = r_;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j].x + x[j].y + m[j] + y[j];
    if (s == 123.456f) sink[0] = s;
}

template <int NF2, int NMUFU, int NF1>
void run(const char* name, int warps_per_smsp) {
    float* sink;
    cudaMalloc(&sink, 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, blocks = 148 * 4, threads = 32 * warps_per_smsp;   // 4 CTAs per SM, each warps_per_smsp warps
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        mix_kernel<NF2, NMUFU, NF1><<<blocks, threads>>>(sink, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    // cycles per loop round (8 groups) per warp-instruction slot, at 1.965 GHz, per sub-partition
    const double cyc = best * 1e-3 * 1.965e9 / (double)iters / 8.0 / warps_per_smsp;
    printf("%-34s warps/SMSP=%d  %.3f ms  -> %.2f cycles per group per warp (FFMA2 x%d, FFMA x%d, MUFU x%d)\n", name, warps_per_smsp, best, cyc, NF2, NF1, NMUFU);
}

int main() {
    for (int w : {1, 2, 4, 8}) {
        run<8, 0, 0>("FFMA2 only", w);
        run<0, 1, 0>("MUFU only", w);
        run<8, 1, 0>("8 FFMA2 + 1 MUFU", w);
        run<0, 0, 8>("FFMA only", w);
        run<0, 1, 8>("8 FFMA + 1 MUFU", w);
        run<0, 2, 8>("8 FFMA + 2 MUFU", w);
        run<4, 1, 0>("4 FFMA2 + 1 MUFU", w);
    }
    return 0;
}
