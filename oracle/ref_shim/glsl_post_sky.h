// TEST INFRASTRUCTURE — not product code (see oracle/README.md).
// Closes the translation unit opened by glsl_pre_sky.h + the three streamed sky fragment shader texts and exposes
// them through a small C interface.  Serial by design.
#undef uniform
#undef in
#undef out
#undef main

extern "C" {

// skysphere.frag: uniforms of render-engine.cpp:367-372, input normalVec -> colour
void wtref_skysphere_run(const uint8_t* gradient_rgba, int gradient_len, float sunHorizonDarkness, const float sunPosition[3], float sunShininess,
                         float sunStrength, const float* normalVec3, long long n, float* colour4) {
    using namespace wt_ref_sky_sphere;
    skysphere.rgba = gradient_rgba; skysphere.length = gradient_len;
    wt_ref_sky_sphere::sunHorizonDarkness = sunHorizonDarkness;
    std::memcpy(&wt_ref_sky_sphere::sunPosition, sunPosition, 12);
    wt_ref_sky_sphere::sunShininess = sunShininess;
    wt_ref_sky_sphere::sunStrength = sunStrength;
    for (long long k = 0; k < n; ++k) {
        std::memcpy(&normalVec, normalVec3 + 3 * k, 12);
        wt_shader_main();
        std::memcpy(colour4 + 4 * k, &colour, 16);
    }
}

// skybox-clouds.frag: uniforms of render-engine.cpp:386-392, input STR -> colour
void wtref_clouds_run(const uint8_t* const faces[6], int face_n, float oneMinusCloudProportion, float overcastStrength, const float sunPosition[3],
                      const float* str3, long long n, float* colour4) {
    using namespace wt_ref_sky_clouds;
    for (int f = 0; f < 6; ++f) skyboxClouds.cube.face[f] = faces[f];
    skyboxClouds.cube.n = face_n;
    wt_ref_sky_clouds::oneMinusCloudProportion = oneMinusCloudProportion;
    wt_ref_sky_clouds::overcastStrength = overcastStrength;
    std::memcpy(&wt_ref_sky_clouds::sunPosition, sunPosition, 12);
    for (long long k = 0; k < n; ++k) {
        std::memcpy(&STR, str3 + 3 * k, 12);
        wt_shader_main();
        std::memcpy(colour4 + 4 * k, &colour, 16);
    }
}

// skybox-stars.frag: input STR -> colour
void wtref_stars_run(const uint8_t* const faces[6], int face_n, const float* str3, long long n, float* colour4) {
    using namespace wt_ref_sky_stars;
    for (int f = 0; f < 6; ++f) skyboxStars.cube.face[f] = faces[f];
    skyboxStars.cube.n = face_n;
    for (long long k = 0; k < n; ++k) {
        std::memcpy(&STR, str3 + 3 * k, 12);
        wt_shader_main();
        std::memcpy(colour4 + 4 * k, &colour, 16);
    }
}

}  // extern "C"
