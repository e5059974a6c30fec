/* TEST INFRASTRUCTURE ONLY - the CPU oracle.  Nothing in the product path may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * It restates, on the CPU, the reference's algorithm for the north-star hot path
 * (SURVEY.md section 8a):
 *   conv/ReLU/pool/concat  Caffe CPU arithmetic: im2col + cblas_sgemm + bias gemm
 *                          (conv_layer.cpp:25-40, base_conv_layer.cpp:257-279, im2col.cpp:19-55,
 *                           math_functions.cpp:13-21, relu_layer.cpp:9-19, pooling_layer.cpp:90-93,128-187,
 *                           concat_layer.cpp:57-74)
 *   ImResize               the reference's CUDA kernel arithmetic (imresize_layer.cu:8-18,98-155),
 *                          including nvcc's FMA contraction pattern, read from the SASS of the
 *                          reference kernels compiled here (oracle/_ref)
 *   NMS                    the reference's CUDA kernels (nms_layer.cu:14-46,49-113)
 *   connectLimbs / COCO    examples/rtpose/rtpose.cpp:549-751, 808-1076
 *   preprocess             rtpose.cpp:239-269, 508-518 + OpenCV INTER_AREA (third-party, restated; pinned
 *                          by tests/golden fixtures generated with cv2 4.13)
 *   JSON                   rtpose.cpp:1383-1416
 *
 * Parity pinning: connectLimbs*, im2col, the model descriptors, process_and_pad_image + the scale
 * arithmetic of the preprocessing (rtpose.cpp:239-269, 474-479, 509-511), the JSON block (:1395-1414),
 * MAX pooling and ReLU (pooling_layer.cpp:90-105,151-186, relu_layer.cpp:15-18) are checked bit-for-bit
 * against the reference's own code compiled from /root/reference (oracle/_ref/libref_host.so, recipe:
 * oracle/build_ref.py); ImResize and NMS are checked bit-for-bit against the reference's own CUDA
 * kernels (oracle/_ref/libref_cpm.so) on the GPU box.  The convolution is checked bit-for-bit against
 * ConvolutionLayer::Forward_cpu compiled from the reference with its cblas_sgemm bound to the same OpenBLAS
 * this file loads; the BLAS binary itself is third-party and unpinned in the reference (Makefile:369-386),
 * so against another BLAS the bar is upstream Caffe's naive-loop 1e-4.
 */
#ifndef RTPOSE_ORACLE_H
#define RTPOSE_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MODEL_MPI_15 = 0, ORC_MODEL_COCO_18 = 1 };

/* ---- BLAS hook: dlopen a library exporting cblas_sgemm (or scipy_cblas_sgemm). Returns 0 on success. */
int orc_load_blas(const char* path);
int orc_have_blas(void);
void orc_set_threads(int n);

/* ---- Caffe layer arithmetic (NCHW fp32) */
void orc_im2col(const float* im, int channels, int height, int width, int kh, int kw, int ph, int pw,
                int sh, int sw, float* col);
void orc_conv2d(const float* in, int n, int cin, int h, int w, const float* weight, const float* bias, int cout,
                int k, int pad, float* out);
void orc_relu(float* x, size_t count);
void orc_maxpool(const float* in, int n, int c, int h, int w, int k, int stride, int pad, float* out);
int orc_pooled_dim(int in, int k, int stride, int pad);

/* ---- model descriptors (modelDescriptorFactory.cpp:6-28,30-55) */
int orc_model_num_parts(int model);
int orc_model_num_limbs(int model);
int orc_model_num_maps(int model); /* parts + bkg + 2*limbs */
const int* orc_model_limb_seq(int model);
const int* orc_model_map_idx(int model);
const char* orc_model_map_name(int model, int idx);

/* ---- the deploy graph (pose_deploy_linevec.prototxt) */
typedef struct OrcNet OrcNet;
OrcNet* orc_net_create(int model);
OrcNet* orc_net_create_stages(int model, int stages); /* 1, 2, 4: model/mpi/pose_deploy_linevec_{1,2,4}.prototxt */
void orc_net_destroy(OrcNet* net);
int orc_net_num_layers(const OrcNet* net);
/* type: "Convolution","ReLU","Pooling","Concat","ImResize","Nms" ; returns 0 if idx valid */
int orc_net_layer_info(const OrcNet* net, int idx, char* name, char* type, char* bottoms, char* top,
                       int* num_output, int* kernel, int* pad, int* stride);
int orc_net_num_convs(const OrcNet* net);
/* conv idx in prototxt order: name[64], cout, cin, k */
int orc_net_conv_info(const OrcNet* net, int conv_idx, char* name, int* cout, int* cin, int* k);
int orc_net_set_weights(OrcNet* net, const char* conv_name, const float* w, const float* b);
/* input: num x 3 x h x w (planar BGR, normalised); out: num x C x h/8 x w/8 (concat_stage7).  Returns 0. */
int orc_net_forward(OrcNet* net, const float* input, int num, int h, int w, float* out);
/* run and also fetch one intermediate blob by top name (for layer-wise parity); blob_out may be NULL */
int orc_net_forward_blob(OrcNet* net, const float* input, int num, int h, int w, const char* blob, float* blob_out,
                         size_t blob_cap, int* bc, int* bh, int* bw);
double orc_net_flops(int model, int h, int w);

/* ---- ImResize (GPU-kernel semantics): src num x c x h8 x w8 -> dst c x th x tw */
void orc_imresize(const float* src, int num, int channels, int h8, int w8, int th, int tw, float start_scale,
                  float scale_gap, float* dst);
/* one output value, for spot checks of the fused engine */
float orc_imresize_at(const float* src, int num, int channels, int h8, int w8, int th, int tw, float start_scale,
                      float scale_gap, int c, int y, int x);

/* ---- NMS (GPU-kernel semantics): map = `channels` full-res maps (channels > num_parts); peaks:
 * num_parts x (max_peaks+1) x 3 (pre-zeroed by the callee).  peaks[part][0][0] = TOTAL count (unclamped). */
void orc_nms(const float* map, int channels, int height, int width, int num_parts, int max_peaks, float threshold,
             float* peaks);

/* ---- connectLimbs / connectLimbsCOCO */
typedef struct {
    int min_subset_cnt;      /* 3 */
    float min_subset_score;  /* 0.4 */
    float inter_threshold;   /* COCO 0.05, MPI 0.01 */
    int inter_min_above;     /* COCO 9, MPI 8 */
    int clamp_counts;        /* 1: n = min(count, max_peaks)  (documented extension; 0 = reference, UB beyond) */
} OrcConnectParams;
void orc_default_params(int model, float* nms_threshold, OrcConnectParams* p);
/* heatmap: C x netH x netW full-res; joints: up to 96 x num_parts x 3.  Returns number of people.
 * subset_out (optional): rows x (num_parts+3) doubles, creation order. */
int orc_connect(int model, const float* heatmap, const float* peaks, int max_peaks, int net_w, int net_h, int disp_w,
                int disp_h, const OrcConnectParams* p, float* joints, double* subset_out, int subset_cap,
                int* subset_rows);

/* ---- display image (rtpose.cpp:474-487): uniform scale s, top-left anchored, cv::warpAffine INTER_CUBIC, black border */
double orc_display_scale(int cols, int rows, int disp_w, int disp_h);
void orc_warp_affine_cubic_u8c3(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw, double scale);

/* ---- preprocess (rtpose.cpp:508-518, 239-269) */
/* cv::resize(..., INTER_AREA), 8UC3: area decimation when both axes shrink, OpenCV's fixed-point bilinear "area mode"
 * as soon as one axis enlarges.  Returns 0. */
int orc_resize_area_u8c3(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw);
void orc_scale_target(int net_w, int net_h, double start_scale, double scale_gap, int i, int* tw, int* th);
/* display image (disp_h x disp_w x 3 BGR u8) -> num_scales x 3 x net_h x net_w floats */
int orc_preprocess(const uint8_t* disp, int disp_h, int disp_w, int net_h, int net_w, int num_scales,
                   double start_scale, double scale_gap, float* out);

/* ---- renderers: render() rtpose.cpp:271-300 + src/rtpose/renderFunctions.cu (GPU-only in the reference).  canvas: 3 x h x w float
 * planar BGR in/out; heatmaps: full-resolution resized_map (num_maps x h_net x w_net), read when part_to_show > 0. */
void orc_canvas_from_u8(const uint8_t* bgr, int h, int w, float* canvas);
void orc_canvas_to_u8(const float* canvas, int h, int w, uint8_t* bgr);
int orc_render_dispatch(int model, int part_to_show, int googly_eyes, int* out3); /* {launcher, part, googly / num_parts_accum} */
int orc_render(int model, float* canvas, int w_canvas, int h_canvas, int w_net, int h_net, const float* heatmaps,
               const float* poses, int num_people, int part_to_show, int googly_eyes);

/* ---- JSON (rtpose.cpp:1383-1416).  Returns bytes written (excluding NUL), or needed size if > cap. */
int orc_json(const float* joints, int num_people, int num_parts, double frame_scale, char* buf, int cap);

/* ---- whole frame: display image -> joints (+ optional peaks / stride-8 maps) */
int orc_process_frame(OrcNet* net, int model, const uint8_t* disp, int disp_h, int disp_w, int net_h, int net_w,
                      int num_scales, double start_scale, double scale_gap, float nms_threshold,
                      const OrcConnectParams* p, float* joints, float* peaks_out, float* maps8_out);

#ifdef __cplusplus
}
#endif
#endif
