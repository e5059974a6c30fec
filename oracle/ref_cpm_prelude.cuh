// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Not part of the product.
//
// Prelude that lets the reference's OWN CUDA kernels
//   src/caffe/cpm/layers/imresize_layer.cu:8-18,97-155   (cubic_interpolation, imresize_cubic_kernel)
//   src/caffe/cpm/layers/nms_layer.cu:13-113             (nms_register_kernel, writeResultKernel)
// compile stand-alone with nvcc for sm_100a, without Caffe.  oracle/build_ref.py splices the
// reference line ranges between this prelude and ref_cpm_launch.cuh in a temp dir; only the
// resulting oracle/_ref/libref_cpm.so is kept.  No reference source is stored in this repo.
#pragma once
#include <cuda_runtime.h>
#include <thrust/scan.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <cmath>
#include <cstdio>

#define NUMBER_THREADS_PER_BLOCK_1D 16
#define NUMBER_THREADS_PER_BLOCK 256

namespace caffe {
// src/caffe/cpm/util/math_functions.cpp:5-7
inline int updiv(int a, int b) { return (a + b - 1) / b; }
}  // namespace caffe
