// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Not part of the product.
//
// Prelude that lets the reference's OWN render kernels (src/rtpose/renderFunctions.cu:4-329, 394-975: getColor*,
// cubic_interpolation, render_pose_29parts{,_heatmap}, render_pose_coco_{parts,heatmap,heatmap2,affinity}) compile
// stand-alone with nvcc for sm_100a.  oracle/build_ref.py splices the reference line ranges between this prelude and
// ref_render_launch.cuh in a temp dir; only oracle/_ref/libref_render.so is kept.  No reference source is stored here.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <cstdio>
#include <vector>

#define RENDER_MAX_PEOPLE 96   // include/rtpose/renderFunctions.h:6

namespace caffe {
inline int updiv(int a, int b) { return (a + b - 1) / b; }   // src/caffe/cpm/util/math_functions.cpp:5-7
}  // namespace caffe
