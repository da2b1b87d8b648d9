// TEST INFRASTRUCTURE — not product code (see oracle/README.md).
// Closes the translation unit opened by glsl_pre.h + the streamed water-grid.vert text and
// exposes the compiled reference vertex shader through a small C interface (ctypes-friendly).
#undef uniform
#undef in
#undef out
#undef main

}  // namespace wt_ref_glsl

extern "C" {

// One record per wave, exactly the five uniforms the reference uploads per wave
// (render-engine.cpp:1076-1080): amplitude_A, frequency_w, phaseConstant_phi, steepness_Q_i, D.x, D.y
struct wtref_wave_uniform { float A, w, phi, Q_i, Dx, Dz; };

// Mirrors the uniform uploads of render-engine.cpp:1057-1110 that water-grid.vert consumes.
struct wtref_vert_uniforms {
    float bottomLeft[4], bottomRight[4], topLeft[4], topRight[4];
    float cameraPosition[3];
    unsigned gerstnerWaveCount;
    unsigned gridLength;
    float heightmapDisplacementScale, heightmapSampleScale;
    float verticalBounceWaveDisplacement;
    float viewMatOnlyYaw[16];   // column-major
    float viewProjection[16];   // column-major
    float waveAnimationTimeInSeconds;
};

int wtref_max_waves() { return WT_REF_MAX_WAVES; }

int wtref_vert_set_uniforms(const wtref_vert_uniforms* u, const wtref_wave_uniform* waves,
                            const uint8_t* heightmap_rgba, int hm_w, int hm_h) {
    using namespace wt_ref_glsl;
    if (u->gerstnerWaveCount > MAX_COUNT_OF_GERSTNER_WAVES) return -1;
    std::memcpy(&bottomLeftGridPointInWorld, u->bottomLeft, 16);
    std::memcpy(&bottomRightGridPointInWorld, u->bottomRight, 16);
    std::memcpy(&topLeftGridPointInWorld, u->topLeft, 16);
    std::memcpy(&topRightGridPointInWorld, u->topRight, 16);
    std::memcpy(&cameraPosition, u->cameraPosition, 12);
    gerstnerWaveCount = u->gerstnerWaveCount;
    for (unsigned i = 0; i < u->gerstnerWaveCount; ++i) {
        gerstnerWaves[i].amplitude_A = waves[i].A;
        gerstnerWaves[i].frequency_w = waves[i].w;
        gerstnerWaves[i].phaseConstant_phi = waves[i].phi;
        gerstnerWaves[i].steepness_Q_i = waves[i].Q_i;
        gerstnerWaves[i].xzDirection_D = vec2(waves[i].Dx, waves[i].Dz);
    }
    gridLength = u->gridLength;
    heightmap.img.rgba = heightmap_rgba;
    heightmap.img.w = hm_w;
    heightmap.img.h = hm_h;
    heightmapDisplacementScale = u->heightmapDisplacementScale;
    heightmapSampleScale = u->heightmapSampleScale;
    verticalBounceWaveDisplacement = u->verticalBounceWaveDisplacement;
    std::memcpy(&viewMatOnlyYaw, u->viewMatOnlyYaw, 64);
    std::memcpy(&viewProjection, u->viewProjection, 64);
    waveAnimationTimeInSeconds = u->waveAnimationTimeInSeconds;
    return 0;
}

// Per-vertex record of every vertex-stage output of water-grid.vert:66-70,150.
struct wtref_vert_out {
    float gl_Position[4];
    float heightmap_colour[4];
    float normal[3];
    float normalVecInViewSpaceOnlyYaw[3];
    float viewVecRaw[3];
    float xyPositionNDCSpaceHeight0[2];
};

// Runs the shader's main() for n vertex ids (OpenMP over ids).  ids == NULL means the
// contiguous range [id_begin, id_begin + n).
void wtref_vert_run(const int* ids, int id_begin, long long n, wtref_vert_out* outv) {
    using namespace wt_ref_glsl;
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < n; ++k) {
        gl_VertexID = ids ? ids[k] : (int)(id_begin + k);
        wt_shader_main();
        wtref_vert_out& o = outv[k];
        std::memcpy(o.gl_Position, &gl_Position, 16);
        std::memcpy(o.heightmap_colour, &heightmap_colour, 16);
        std::memcpy(o.normal, &normal, 12);
        std::memcpy(o.normalVecInViewSpaceOnlyYaw, &normalVecInViewSpaceOnlyYaw, 12);
        std::memcpy(o.viewVecRaw, &viewVecRaw, 12);
        std::memcpy(o.xyPositionNDCSpaceHeight0, &xyPositionNDCSpaceHeight0, 8);
    }
}

}  // extern "C"
