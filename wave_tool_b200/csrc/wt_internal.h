// Internal declarations shared by the host prologue (wt_host.cpp) and the CUDA side (wt_device.cu).
#pragma once
#include "../../include/wavetool_b200.h"

namespace wt {

// Records a ctx-less error for wt_last_error(NULL) on this thread and returns `code`.
int set_thread_error(int code, const char* msg);

// Intermediate values of the projected-grid builder, exposed to the tests through
// wt_project_grid_debug (not part of the drop-in surface).
struct GridDebug {
    int n_intersections;
    float aimpoint_final[3];
    float projector_pitch_deg, projector_y;
    float range[4];
};

int project_grid_impl(const wt_camera& cam, float displaceable_amplitude, float corners[4][4], GridDebug* dbg);

}  // namespace wt
