// Header-compatible shim of include/rtpose/renderFunctions.h of the reference (:1-18): the three render entry points
// examples/rtpose/rtpose.cpp:277-296 calls, same signatures and argument meaning (device pointers), forwarded to the C ABI of
// libposeengine.so (pe_render_device).  `boxsize` and `centers` are unused by the reference's implementation as well.
#ifndef RENDER_FUNCTIONS_H
#define RENDER_FUNCTIONS_H

#include <vector>

#include "../poseengine.h"

#define RENDER_MAX_PEOPLE 96

inline void render_mpi_parts(float* canvas, int w_canvas, int h_canvas, int w_net, int h_net, float* heatmaps, int /*boxsize*/,
                             float* /*centers*/, float* poses, std::vector<int> num_people, int part) {
    pe_render_device(0, canvas, w_canvas, h_canvas, w_net, h_net, heatmaps, poses, num_people.data(), (int)num_people.size(), part, 0);
}
inline void render_coco_parts(float* canvas, int w_canvas, int h_canvas, int w_net, int h_net, float* heatmaps, int /*boxsize*/,
                              float* /*centers*/, float* poses, std::vector<int> num_people, int part, bool googly_eyes = 0) {
    pe_render_device(1, canvas, w_canvas, h_canvas, w_net, h_net, heatmaps, poses, num_people.data(), (int)num_people.size(), part,
                     googly_eyes ? 1 : 0);
}
inline void render_coco_aff(float* canvas, int w_canvas, int h_canvas, int w_net, int h_net, float* heatmaps, int /*boxsize*/,
                            float* /*centers*/, float* poses, std::vector<int> num_people, int part, int num_parts_accum) {
    pe_render_device(2, canvas, w_canvas, h_canvas, w_net, h_net, heatmaps, poses, num_people.data(), (int)num_people.size(), part,
                     num_parts_accum);
}

#endif  // RENDER_FUNCTIONS_H
