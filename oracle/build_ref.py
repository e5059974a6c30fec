#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Build oracle/_ref/ from the reference sources WHERE THEY LIE.

Nothing under /root/reference is copied into the repository: line ranges of the reference's own
files are spliced between our preludes/launchers in a temporary directory, compiled, and only
the resulting shared objects are kept under oracle/_ref/ (git-ignored, travels to the GPU box).

  libref_host.so  g++   src/rtpose/modelDescriptor{,Factory}.cpp (unmodified)
                        examples/rtpose/rtpose.cpp:144-152, 549-751, 808-1076 (ColumnCompare,
                        connectLimbs, connectLimbsCOCO)  src/caffe/util/im2col.cpp:8-56 (im2col_cpu)
                        conv_layer.cpp:27-39, base_conv_layer.cpp:259-271, 277-279, base_conv_layer.hpp:100-105,
                        math_functions.cpp:12-21 (ConvolutionLayer::Forward_cpu down to the cblas_sgemm call, BLAS loaded at run time)
                        second translation unit: rtpose.cpp:239-269 (process_and_pad_image), :474-479 (display scale),
                        :509-511 (per-scale target size), :1395-1414 (JSON writer), :271-300 (render() dispatch,
                        launchers replaced by recording stand-ins), :103-129 + :1551-1592, 1606-1671 (handleKey without its window calls);
                        src/caffe/layers/pooling_layer.cpp:90-105, 151-186 (pooled size, MAX loop),
                        src/caffe/layers/relu_layer.cpp:15-18
  libref_cpm.so   nvcc  src/caffe/cpm/layers/imresize_layer.cu:8-18,97-155
                        src/caffe/cpm/layers/nms_layer.cu:13-113
                        (sm_100a; default -fmad, as the reference Makefile:410 passes no fmad flag)
  libref_render.so nvcc src/rtpose/renderFunctions.cu:4-329, 394-975 (the six render kernels + colour helpers)

Flags: the reference Makefile (:326,405,409) uses `-O3 -march=native -std=c++11`; we drop
-march=native so that the host arithmetic is the portable SSE2 one (no machine-dependent FMA
contraction) - stated in DESIGN.md.

The reference's own build system is not run.  If /root/reference is absent (GPU box) this script
is a no-op and the prebuilt .so files are used.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RTPOSE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def lines(path, ranges):
    src = open(os.path.join(REF, path)).read().split("\n")
    out = []
    for a, b in ranges:
        out.append("#line %d \"%s\"" % (a, path))
        out.extend(src[a - 1:b])
    return "\n".join(out) + "\n"


HOST_WRAPPER = r'''
// ---- wrappers (ours) ----
extern "C" int ref_connect(int model, const float* heatmap, const float* peaks, int max_peaks,
                           int netW, int netH, int dispW, int dispH,
                           int min_cnt, float min_score, float inter_thr, int min_above,
                           float* joints, double* subset_out, int subset_cap, int* subset_rows) {
    NET_RESOLUTION_WIDTH = netW; NET_RESOLUTION_HEIGHT = netH;
    DISPLAY_RESOLUTION_WIDTH = dispW; DISPLAY_RESOLUTION_HEIGHT = dispH;
    global.connect_min_subset_cnt = min_cnt; global.connect_min_subset_score = min_score;
    global.connect_inter_threshold = inter_thr; global.connect_inter_min_above_threshold = min_above;
    std::unique_ptr<ModelDescriptor> md;
    ModelDescriptorFactory::createModelDescriptor(
        model == 0 ? ModelDescriptorFactory::Type::MPI_15 : ModelDescriptorFactory::Type::COCO_18, md);
    std::vector<std::vector<double>> subset;
    std::vector<std::vector<std::vector<double>>> connection;
    int cnt = model == 0 ? connectLimbs(subset, connection, heatmap, peaks, max_peaks, joints, md.get())
                         : connectLimbsCOCO(subset, connection, heatmap, peaks, max_peaks, joints, md.get());
    if (subset_rows) *subset_rows = (int)subset.size();
    if (subset_out) {
        const int w = md->get_number_parts() + 3;
        for (int i = 0; i < (int)subset.size() && i < subset_cap; i++)
            for (int j = 0; j < w; j++) subset_out[i * w + j] = subset[i][j];
    }
    return cnt;
}

extern "C" int ref_model_descriptor(int model, int* num_parts, int* num_limbs, int* limb_seq, int* map_idx,
                                    char* names, int names_cap) {
    std::unique_ptr<ModelDescriptor> md;
    try {
        ModelDescriptorFactory::createModelDescriptor(
            model == 0 ? ModelDescriptorFactory::Type::MPI_15 : ModelDescriptorFactory::Type::COCO_18, md);
    } catch (...) { return -1; }
    *num_parts = md->get_number_parts();
    *num_limbs = md->number_limb_sequence();
    for (int i = 0; i < 2 * *num_limbs; i++) { limb_seq[i] = md->get_limb_sequence()[i]; map_idx[i] = md->get_map_idx()[i]; }
    std::string all;
    const int nmaps = *num_parts + 1 + 2 * *num_limbs;
    for (int i = 0; i < nmaps; i++) { all += md->get_part_name(i); all += "\n"; }
    snprintf(names, names_cap, "%s", all.c_str());
    return 0;
}

extern "C" void ref_im2col(const float* im, int channels, int height, int width, int kh, int kw,
                           int ph, int pw, int sh, int sw, float* col) {
    caffe::im2col_cpu<float>(im, channels, height, width, kh, kw, ph, pw, sh, sw, 1, 1, col);
}
'''


# Second translation unit: small pieces of the per-frame host path and of the Caffe CPU layers the oracle restates, each a line
# range of the reference spliced into a function body of ours whose locals carry the names the reference code uses.
HOST2_PRELUDE = r"""
// TEST INFRASTRUCTURE ONLY (oracle/_ref build): generated by oracle/build_ref.py, never committed.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
using std::min; using std::max; using std::vector;
struct RefNullStream2 { template <typename T> RefNullStream2& operator<<(const T&) { return *this; } };
#define REF_CHECK_OP(a, op, b) if (!((a)op(b))) { fprintf(stderr, "ref CHECK failed: %s %s %s\n", #a, #op, #b); abort(); } else RefNullStream2()
#define CHECK_GE(a, b) REF_CHECK_OP(a, >=, b)
#define CHECK_LE(a, b) REF_CHECK_OP(a, <=, b)
#define CHECK_LT(a, b) REF_CHECK_OP(a, <, b)
namespace cv { struct Mat { int cols, rows; unsigned char* data; }; }   // the three members process_and_pad_image reads
struct RefBlob {                                                        // Blob::num / offset(n, c) for an N x C x H x W blob
    int n_, c_, h_, w_;
    int num() const { return n_; }
    int offset(int n, int c) const { return (n * c_ + c) * h_ * w_; }
};
"""

HOST2_BODY_A = r"""
extern "C" void ref_process_and_pad_image(float* target, unsigned char* data, int ow, int oh, int tw, int th, int normalize) {
    cv::Mat m{ow, oh, data};
    process_and_pad_image(target, m, tw, th, normalize != 0);
}
extern "C" double ref_display_scale(int cols, int rows, int DISPLAY_RESOLUTION_WIDTH, int DISPLAY_RESOLUTION_HEIGHT) {
    struct { int cols, rows; } image_uchar_orig = {cols, rows};
"""
HOST2_BODY_B = r"""
    return scale;
}
extern "C" void ref_scale_target(int NET_RESOLUTION_WIDTH, int NET_RESOLUTION_HEIGHT, float START_SCALE, float SCALE_GAP, int i, int* tw, int* th) {
    int target_width, target_height;
"""
HOST2_BODY_C = r"""
    *tw = target_width; *th = target_height;
}
extern "C" void ref_write_json(const char* fname, const float* joints, int numPeople, int num_parts, double frame_scale) {
    struct { const float* joints; int numPeople; double scale; } frame = {joints, numPeople, frame_scale};
    double scale = 1.0/frame.scale;   // rtpose.cpp:1384
    {
"""
HOST2_BODY_D = r"""
    }
}
// PoolingLayer<float>::Reshape (pooled size) + Forward_cpu, MAX case
extern "C" int ref_maxpool(const float* bottom_data, int num, int channels_, int height_, int width_, int kernel, int stride, int pad,
                           float* top_data, int* pooled_hw) {
    typedef float Dtype;
    const int kernel_h_ = kernel, kernel_w_ = kernel, stride_h_ = stride, stride_w_ = stride, pad_h_ = pad, pad_w_ = pad;
    int pooled_height_, pooled_width_;
"""
HOST2_BODY_E = r"""
    pooled_hw[0] = pooled_height_; pooled_hw[1] = pooled_width_;
    if (!top_data) return 0;
    RefBlob b{num, channels_, height_, width_}, t{num, channels_, pooled_height_, pooled_width_};
    vector<RefBlob*> bottom{&b}, top{&t};
    const int top_count = num * channels_ * pooled_height_ * pooled_width_;
    const bool use_top_mask = false;
    float* top_mask = NULL;
    vector<int> mask_store(top_count, -1);                       // caffe_set(top_count, -1, mask)            :147
    int* mask = mask_store.data();
    for (int i = 0; i < top_count; i++) top_data[i] = -FLT_MAX;  // caffe_set(top_count, Dtype(-FLT_MAX), ..) :149
"""
HOST2_BODY_F = r"""
    return 0;
}
extern "C" void ref_relu(const float* bottom_data, float* top_data, int count, float negative_slope) {
    typedef float Dtype;
"""
HOST2_BODY_G = r"""
}
// render() (rtpose.cpp:271-300): which of the three launchers runs for a --part_to_show value, and with which arguments.  The launchers
// are recording stand-ins with the signatures of include/rtpose/renderFunctions.h; the dispatch itself is the reference's text.
struct RefMD2 { int parts; int get_number_parts() const { return parts; } };
struct RefNetCopy2 { float* canvas; float* joints; std::vector<int> num_people; RefMD2* up_model_descriptor; };
static RefMD2 g_md2;
static std::vector<RefNetCopy2> net_copies(1);
#define LOG(x) RefNullStream2()
static struct RefGlobal2 {   // the members of `struct Global` that are not queues: rtpose.cpp:103-129, spliced
"""
HOST2_BODY_G2 = r"""
} global;
static int DISPLAY_RESOLUTION_WIDTH = 1280, DISPLAY_RESOLUTION_HEIGHT = 720, NET_RESOLUTION_WIDTH = 656, NET_RESOLUTION_HEIGHT = 368;
const int BOX_SIZE = 368;
static double get_wall_time() { return 0; }
#define VLOG(x) RefNullStream2()
static int g_rec[4];
static void render_mpi_parts(float*, int, int, int, int, float*, int, float*, float*, std::vector<int>, int part) {
    g_rec[0] = 0; g_rec[1] = part; g_rec[2] = 0; g_rec[3]++; }
static void render_coco_parts(float*, int, int, int, int, float*, int, float*, float*, std::vector<int>, int part, bool googly_eyes) {
    g_rec[0] = 1; g_rec[1] = part; g_rec[2] = googly_eyes; g_rec[3]++; }
static void render_coco_aff(float*, int, int, int, int, float*, int, float*, float*, std::vector<int>, int part, int num_parts_accum) {
    g_rec[0] = 2; g_rec[1] = part; g_rec[2] = num_parts_accum; g_rec[3]++; }
"""
HOST2_BODY_H = r"""
extern "C" int ref_render_dispatch(int num_parts, int part_to_show, int googly_eyes, int* out3) {
    g_md2.parts = num_parts;
    net_copies[0].up_model_descriptor = &g_md2;
    global.part_to_show = part_to_show; global.uistate.is_googly_eyes = googly_eyes != 0;
    g_rec[3] = 0;
    render(0, NULL);
    for (int i = 0; i < 3; i++) out3[i] = g_rec[i];
    return g_rec[3];   // launches made (0 for a model that is neither 15 nor 18 parts)
}
"""

HOST2_BODY_I = r"""
// state: f[3] = nms_threshold, connect_min_subset_score, connect_inter_threshold; i[7] = connect_inter_min_above_threshold,
// connect_min_subset_cnt, part_to_show, is_googly_eyes, is_video_paused, current_frame, quit_threads - in and out
extern "C" void ref_handle_keys(const int* keys, int n, int has_video, float* f, int* i) {
    FLAGS_video = has_video ? "clip.avi" : "";
    global.nms_threshold = f[0]; global.connect_min_subset_score = f[1]; global.connect_inter_threshold = f[2];
    global.connect_inter_min_above_threshold = i[0]; global.connect_min_subset_cnt = i[1]; global.part_to_show = i[2];
    global.uistate.is_googly_eyes = i[3] != 0; global.uistate.is_video_paused = i[4] != 0; global.uistate.current_frame = i[5];
    global.quit_threads = false;
    for (int k = 0; k < n; k++) handleKey(keys[k]);
    f[0] = global.nms_threshold; f[1] = global.connect_min_subset_score; f[2] = global.connect_inter_threshold;
    i[0] = global.connect_inter_min_above_threshold; i[1] = global.connect_min_subset_cnt; i[2] = global.part_to_show;
    i[3] = global.uistate.is_googly_eyes; i[4] = global.uistate.is_video_paused; i[5] = global.uistate.current_frame; i[6] = global.quit_threads;
}
"""


def host2_tu():
    return (HOST2_PRELUDE
            + lines("examples/rtpose/rtpose.cpp", [(239, 269)])
            + HOST2_BODY_A + lines("examples/rtpose/rtpose.cpp", [(474, 479)])
            + HOST2_BODY_B + lines("examples/rtpose/rtpose.cpp", [(509, 511)])
            + HOST2_BODY_C + lines("examples/rtpose/rtpose.cpp", [(1395, 1414)])
            + HOST2_BODY_D + lines("src/caffe/layers/pooling_layer.cpp", [(90, 105)])
            + HOST2_BODY_E + lines("src/caffe/layers/pooling_layer.cpp", [(151, 186)])
            + HOST2_BODY_F + lines("src/caffe/layers/relu_layer.cpp", [(15, 18)])
            + HOST2_BODY_G + lines("examples/rtpose/rtpose.cpp", [(103, 129)])
            + HOST2_BODY_G2 + lines("examples/rtpose/rtpose.cpp", [(271, 300)])
            + HOST2_BODY_H
            + "static std::string FLAGS_video;\n"
            + lines("examples/rtpose/rtpose.cpp", [(1551, 1592), (1606, 1671)])     # handleKey without the cv:: window block of the 'f' key
            + HOST2_BODY_I)

# Convolution forward of the reference (ConvolutionLayer::Forward_cpu -> forward_cpu_gemm / forward_cpu_bias -> caffe_cpu_gemm ->
# cblas_sgemm) as members of a stand-in class that carries the members those bodies read.  The BLAS is third-party in the reference
# (Makefile:369-386); cblas_sgemm is forwarded to whatever library ref_load_blas() opens - the tests hand it the same OpenBLAS the
# oracle uses, so the two must then agree bit for bit.
CONV_A = r"""
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
typedef void (*ref_sgemm_t)(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
static ref_sgemm_t g_ref_sgemm = 0;
static inline void cblas_sgemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int M, int N, int K, float alpha, const float* A, int lda,
                               const float* B, int ldb, float beta, float* C, int ldc) {
    g_ref_sgemm(o, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
template <typename Dtype>
void caffe_cpu_gemm(const CBLAS_TRANSPOSE TransA, const CBLAS_TRANSPOSE TransB, const int M, const int N, const int K, const Dtype alpha,
                    const Dtype* A, const Dtype* B, const Dtype beta, Dtype* C);
"""
CONV_B = r"""
template <typename T> struct RefBuf {   // the cpu_data() / mutable_cpu_data() face of Blob / SyncedMemory
    std::vector<T> v;
    T* mutable_cpu_data() { return v.data(); }
    const T* cpu_data() const { return v.data(); }
};
template <typename T> struct RefPtr {
    T* p;
    const T* cpu_data() const { return p; }
    T* mutable_cpu_data() { return p; }
};
template <typename Dtype>
struct BaseConvolutionLayer {
    bool is_1x1_, bias_term_, force_nd_im2col_;
    int num_spatial_axes_, group_, conv_out_channels_, conv_in_channels_, conv_out_spatial_dim_, kernel_dim_, weight_offset_, col_offset_,
        output_offset_, num_output_, out_spatial_dim_, num_, bottom_dim_, top_dim_;
    RefBuf<Dtype> col_buffer_, bias_multiplier_;
    RefBuf<int> conv_input_shape_, kernel_shape_, pad_, stride_, dilation_;
    std::vector<RefPtr<Dtype>*> blobs_;
    inline void conv_im2col_cpu(const Dtype* data, Dtype* col_buff) {   // base_conv_layer.hpp:98-105, the 2-D branch
"""
CONV_C = r"""
    }
    void forward_cpu_gemm(const Dtype* input, const Dtype* weights, Dtype* output, bool skip_im2col = false) {
"""
CONV_D = r"""
    }
    void forward_cpu_bias(Dtype* output, const Dtype* bias) {
"""
CONV_E = r"""
    }
};
template <typename Dtype>
struct ConvolutionLayer : public BaseConvolutionLayer<Dtype> {
    void Forward_cpu(const std::vector<RefPtr<Dtype>*>& bottom, const std::vector<RefPtr<Dtype>*>& top) {
"""
CONV_F = r"""
    }
};
"""
CONV_WRAPPER = r"""
#include <dlfcn.h>
extern "C" int ref_load_blas(const char* path) {
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    void* f = dlsym(h, "cblas_sgemm");
    if (!f) f = dlsym(h, "scipy_cblas_sgemm");
    if (!f) return -2;
    caffe::g_ref_sgemm = (caffe::ref_sgemm_t)f;
    return 0;
}
// stride 1, dilation 1, group 1 (every convolution of the pose nets); members set as LayerSetUp / Reshape set them
// (base_conv_layer.cpp:95-254: kernel_dim_ = blobs_[0]->count(1), weight_offset_ = conv_out_channels_ * kernel_dim_ / group_, ...)
extern "C" int ref_conv_forward(const float* in, int n, int cin, int h, int w, const float* weight, const float* bias, int cout, int k, int pad,
                                float* out) {
    if (!caffe::g_ref_sgemm) return -1;
    caffe::ConvolutionLayer<float> L;
    const int oh = h + 2 * pad - k + 1, ow = w + 2 * pad - k + 1;
    L.is_1x1_ = k == 1 && pad == 0; L.bias_term_ = bias != 0; L.force_nd_im2col_ = false; L.num_spatial_axes_ = 2; L.group_ = 1;
    L.conv_out_channels_ = L.num_output_ = cout; L.conv_in_channels_ = cin;
    L.conv_out_spatial_dim_ = L.out_spatial_dim_ = oh * ow; L.kernel_dim_ = cin * k * k;
    L.weight_offset_ = cout * L.kernel_dim_; L.col_offset_ = L.kernel_dim_ * oh * ow; L.output_offset_ = cout * oh * ow;
    L.num_ = n; L.bottom_dim_ = cin * h * w; L.top_dim_ = cout * oh * ow;
    L.col_buffer_.v.resize((size_t)L.kernel_dim_ * oh * ow);
    L.bias_multiplier_.v.assign((size_t)oh * ow, 1.0f);                                  // caffe_set(..., Dtype(1), ...) :251-252
    L.conv_input_shape_.v = {cin, h, w}; L.kernel_shape_.v = {k, k}; L.pad_.v = {pad, pad}; L.stride_.v = {1, 1}; L.dilation_.v = {1, 1};
    caffe::RefPtr<float> wb{const_cast<float*>(weight)}, bb{const_cast<float*>(bias)}, ib{const_cast<float*>(in)}, ob{out};
    L.blobs_ = {&wb, &bb};
    std::vector<caffe::RefPtr<float>*> bottom{&ib}, top{&ob};
    L.Forward_cpu(bottom, top);
    return 0;
}
"""


# warmup()'s model selection (rtpose.cpp:212-229): the thresholds each model starts with
DEFAULTS_A = r"""
#ifndef CHECK
#define CHECK(c) if (!(c)) { fprintf(stderr, "ref CHECK failed: %s\n", #c); abort(); } else RefNullStream()
#endif
extern "C" void ref_model_defaults(int num_parts, float* nms_threshold, int* min_subset_cnt, float* min_subset_score, float* inter_threshold,
                                   int* inter_min_above_threshold) {
    struct RefNetCopy { int nms_num_parts; std::unique_ptr<ModelDescriptor> up_model_descriptor; };
    std::vector<RefNetCopy> net_copies(1);
    const int device_id = 0;
    net_copies[device_id].nms_num_parts = num_parts;
"""
DEFAULTS_B = r"""
    *nms_threshold = global.nms_threshold; *min_subset_cnt = global.connect_min_subset_cnt; *min_subset_score = global.connect_min_subset_score;
    *inter_threshold = global.connect_inter_threshold; *inter_min_above_threshold = global.connect_inter_min_above_threshold;
}
"""


def build_host(tmp):
    tu = ('#include "%s"\n' % os.path.join(HERE, "ref_host_prelude.h")
          + lines("examples/rtpose/rtpose.cpp", [(144, 152), (549, 751), (808, 1076)])
          + "namespace caffe {\n" + lines("src/caffe/util/im2col.cpp", [(8, 56)])
          + CONV_A + lines("src/caffe/util/math_functions.cpp", [(12, 21)])
          + CONV_B + lines("include/caffe/layers/base_conv_layer.hpp", [(100, 105)])
          + CONV_C + lines("src/caffe/layers/base_conv_layer.cpp", [(259, 271)])
          + CONV_D + lines("src/caffe/layers/base_conv_layer.cpp", [(277, 279)])
          + CONV_E + lines("src/caffe/layers/conv_layer.cpp", [(27, 39)])
          + CONV_F + "}\n"
          + HOST_WRAPPER + CONV_WRAPPER
          + DEFAULTS_A + lines("examples/rtpose/rtpose.cpp", [(212, 229)]) + DEFAULTS_B)
    src = os.path.join(tmp, "ref_host_tu.cpp")
    open(src, "w").write(tu)
    src2 = os.path.join(tmp, "ref_host_tu2.cpp")
    open(src2, "w").write(host2_tu())
    out = os.path.join(OUT, "libref_host.so")
    cmd = ["g++", "-O3", "-std=c++11", "-fPIC", "-shared", "-w", "-I" + os.path.join(REF, "include"), src, src2,
           os.path.join(REF, "src/rtpose/modelDescriptor.cpp"),
           os.path.join(REF, "src/rtpose/modelDescriptorFactory.cpp"), "-o", out, "-ldl"]
    subprocess.check_call(cmd)
    return out


def build_cpm(tmp, keep_sass=False):
    tu = ('#include "%s"\n' % os.path.join(HERE, "ref_cpm_prelude.cuh")
          + "namespace caffe {\n"
          + lines("src/caffe/cpm/layers/imresize_layer.cu", [(8, 18), (97, 155)])
          + lines("src/caffe/cpm/layers/nms_layer.cu", [(13, 113)])
          + "}\n"
          + '#include "%s"\n' % os.path.join(HERE, "ref_cpm_launch.cuh"))
    src = os.path.join(tmp, "ref_cpm_tu.cu")
    open(src, "w").write(tu)
    out = os.path.join(OUT, "libref_cpm.so")
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-shared", "-w", src, "-o", out]
    subprocess.check_call(cmd)
    if keep_sass:
        sass = subprocess.check_output(["cuobjdump", "-sass", out]).decode()
        open(os.path.join(OUT, "ref_cpm.sass"), "w").write(sass)
    return out


def build_render(tmp, keep_sass=False):
    tu = ('#include "%s"\n' % os.path.join(HERE, "ref_render_prelude.cuh")
          + lines("src/rtpose/renderFunctions.cu", [(4, 329), (394, 975)])
          + '#include "%s"\n' % os.path.join(HERE, "ref_render_launch.cuh"))
    src = os.path.join(tmp, "ref_render_tu.cu")
    open(src, "w").write(tu)
    out = os.path.join(OUT, "libref_render.so")
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-shared", "-w", src, "-o", out]
    subprocess.check_call(cmd)
    if keep_sass:
        sass = subprocess.check_output(["cuobjdump", "-sass", out]).decode()
        open(os.path.join(OUT, "ref_render.sass"), "w").write(sass)
    return out


def main():
    if not os.path.isdir(REF):
        print("build_ref: %s absent - keeping prebuilt oracle/_ref" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="rtpose_ref_")
    try:
        print("built", build_host(tmp))
        if shutil.which("nvcc"):
            print("built", build_cpm(tmp, keep_sass="--sass" in sys.argv))
            print("built", build_render(tmp, keep_sass="--sass" in sys.argv))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
