// wt_headless — a C++ host program on top of the C ABI (include/wavetool_b200.h), the way a C++ application such as
// the reference (src/program.cpp / src/render-engine.cpp) would drive the library: set the wave table and textures
// once, then let wt_run_animation advance the animation state and evaluate frame after frame on the GPU.
//
//   g++ -std=c++17 -O2 -Iinclude -I/usr/local/cuda/include examples/wt_headless.cpp \
//       -Lwave_tool_b200 -lwavetool_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/wave_tool_b200 -o wt_headless
//   ./wt_headless [grid_len=513] [frames=8] [waves.bin]      (waves.bin: wt_wave records; default: the reference's 4 slots)
//
// Prints one line per frame: the frame's digest (wt_checksum).  No CPU fallback: fails without an sm_100 GPU.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "wavetool_b200.h"

#define CHECK_WT(expr)                                                                       \
    do {                                                                                     \
        int st_ = (expr);                                                                    \
        if (st_ < 0) { std::fprintf(stderr, "%s failed: %s\n", #expr, wt_last_error(ctx)); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const unsigned L = argc > 1 ? (unsigned)std::atoi(argv[1]) : 513;   // GRID_LENGTH, src/program.cpp:688
    const unsigned frames = argc > 2 ? (unsigned)std::atoi(argv[2]) : 8;
    wt_ctx* ctx = nullptr;
    if (wt_create(&ctx, 0) != WT_OK) { std::fprintf(stderr, "wt_create: %s\n", wt_last_error(nullptr)); return 2; }

    std::vector<wt_wave> waves(4);
    wt_default_waves(waves.data());                                     // src/render-engine.cpp:46-49
    if (argc > 3) {
        FILE* f = std::fopen(argv[3], "rb");
        if (!f) { std::perror(argv[3]); return 1; }
        std::fseek(f, 0, SEEK_END);
        const long bytes = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        waves.resize((size_t)bytes / sizeof(wt_wave));
        if (std::fread(waves.data(), sizeof(wt_wave), waves.size(), f) != waves.size()) return 1;
        std::fclose(f);
    }
    CHECK_WT(wt_set_waves(ctx, waves.data(), (unsigned)waves.size()));

    // a small procedural heightmap (the application would hand over the pixels stbi gave it, src/render-engine.cpp:1308-1322)
    const int hw = 64;
    std::vector<uint8_t> hm((size_t)hw * hw * 4);
    for (int y = 0; y < hw; ++y)
        for (int x = 0; x < hw; ++x) {
            const uint8_t v = (uint8_t)(96 + ((x * 7 + y * 13) % 64));
            uint8_t* p = &hm[4 * ((size_t)y * hw + x)];
            p[0] = p[1] = p[2] = v; p[3] = 255;
        }
    CHECK_WT(wt_set_heightmap_rgba8(ctx, hm.data(), hw, hw));

    const size_t n = (size_t)L * L, ring = 2;
    void *d_pos = nullptr, *d_nrm = nullptr, *d_cks = nullptr;
    if (cudaMalloc(&d_pos, ring * n * 16) != cudaSuccess || cudaMalloc(&d_nrm, ring * n * 16) != cudaSuccess ||
        cudaMalloc(&d_cks, frames * sizeof(wt_checksum)) != cudaSuccess) { std::fprintf(stderr, "cudaMalloc failed\n"); return 2; }

    wt_camera cam;
    wt_default_camera(&cam, 1.0f);                                      // src/render-engine.cpp:52
    wt_params params;
    wt_default_params(&params);
    wt_animation anim;
    wt_default_animation(&anim);
    unsigned drawn = 0;
    CHECK_WT(wt_run_animation(ctx, &cam, &params, &anim, wt_total_amplitude(waves.data(), (unsigned)waves.size()), 1.0f / 60.0f, 1, frames, L, 0, L,
                              d_pos, d_nrm, nullptr, (unsigned)ring, d_cks, WT_FLAG_CHECKSUM, nullptr, &drawn));
    std::vector<wt_checksum> cks(frames);
    if (cudaMemcpy(cks.data(), d_cks, frames * sizeof(wt_checksum), cudaMemcpyDeviceToHost) != cudaSuccess) return 2;
    for (unsigned f = 0; f < frames; ++f)
        std::printf("frame %u position_bits %llu normal_bits %llu min_y %.9g max_y %.9g vertices %llu\n", f, (unsigned long long)cks[f].position_bits,
                    (unsigned long long)cks[f].normal_bits, cks[f].min_y, cks[f].max_y, (unsigned long long)cks[f].vertices);
    std::printf("%u of %u frames drawn, t = %.9g s, %llu kernel launches\n", drawn, frames, params.waveAnimationTimeInSeconds,
                (unsigned long long)wt_launch_count(ctx));
    cudaFree(d_pos); cudaFree(d_nrm); cudaFree(d_cks);
    wt_destroy(ctx);
    return 0;
}
