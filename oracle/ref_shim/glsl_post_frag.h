// TEST INFRASTRUCTURE — not product code (see oracle/README.md).
// Closes the translation unit opened by glsl_pre_frag.h + the streamed water-grid.frag text and exposes the
// compiled reference fragment shader through a small C interface (ctypes-friendly).  Serial by design: the
// shader's `in` variables are plain globals here.
#undef uniform
#undef in
#undef out
#undef main

}  // namespace wt_ref_glsl_frag

extern "C" {

// Mirrors the uniform uploads of render-engine.cpp:1060-1063,1087-1110 that water-grid.frag consumes.
struct wtref_frag_uniforms {
    float fogColourFarAtCurrentTime[4];
    float fogDepthRadiusFar, fogDepthRadiusNear;
    float softEdgesDeltaDepthThreshold;
    float sunPosition[3];
    float sunShininess, sunStrength;
    float tintDeltaDepthThreshold;
    float viewportWidthHeight[2];
    float waterClarity;
    float zFar, zNear;
};

int wtref_frag_set_uniforms(const wtref_frag_uniforms* u, const float* depth, const uint8_t* reflections_rgba, const uint8_t* refractions_rgba,
                            int w, int h, const uint8_t* const sky_faces[6], int sky_n) {
    using namespace wt_ref_glsl_frag;
    std::memcpy(&fogColourFarAtCurrentTime, u->fogColourFarAtCurrentTime, 16);
    fogDepthRadiusFar = u->fogDepthRadiusFar;
    fogDepthRadiusNear = u->fogDepthRadiusNear;
    softEdgesDeltaDepthThreshold = u->softEdgesDeltaDepthThreshold;
    std::memcpy(&sunPosition, u->sunPosition, 12);
    sunShininess = u->sunShininess;
    sunStrength = u->sunStrength;
    tintDeltaDepthThreshold = u->tintDeltaDepthThreshold;
    std::memcpy(&viewportWidthHeight, u->viewportWidthHeight, 8);
    waterClarity = u->waterClarity;
    zFar = u->zFar;
    zNear = u->zNear;
    depthTexture2D.depth = depth; depthTexture2D.w = w; depthTexture2D.h = h;
    localReflectionsTexture2D.img = wt_gl::Image2D{reflections_rgba, w, h};
    localRefractionsTexture2D.img = wt_gl::Image2D{refractions_rgba, w, h};
    for (int f = 0; f < 6; ++f) skybox.cube.face[f] = sky_faces[f];
    skybox.cube.n = sky_n;
    return 0;
}

// One record per "fragment": the varyings of water-grid.frag:57-61 plus gl_FragCoord.
struct wtref_frag_in {
    float gl_FragCoord[4];
    float heightmap_colour[4];
    float normal[3];
    float normalVecInViewSpaceOnlyYaw[3];
    float viewVecRaw[3];
    float xyPositionNDCSpaceHeight0[2];
};

// Runs the shader's main() for n fragments; colour4 receives the `out vec4 colour` of each.
void wtref_frag_run(const wtref_frag_in* inv, long long n, float* colour4) {
    using namespace wt_ref_glsl_frag;
    for (long long k = 0; k < n; ++k) {
        const wtref_frag_in& f = inv[k];
        std::memcpy(&gl_FragCoord, f.gl_FragCoord, 16);
        std::memcpy(&heightmap_colour, f.heightmap_colour, 16);
        std::memcpy(&normal, f.normal, 12);
        std::memcpy(&normalVecInViewSpaceOnlyYaw, f.normalVecInViewSpaceOnlyYaw, 12);
        std::memcpy(&viewVecRaw, f.viewVecRaw, 12);
        std::memcpy(&xyPositionNDCSpaceHeight0, f.xyPositionNDCSpaceHeight0, 8);
        wt_shader_main();
        std::memcpy(colour4 + 4 * k, &colour, 16);
    }
}

}  // extern "C"
