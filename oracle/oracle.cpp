// TEST INFRASTRUCTURE ONLY - the CPU oracle (see oracle.h for scope, citations and pinning).
// Build: oracle/Makefile  (g++ -O2 -mfma -ffp-contract=off: every FMA below is an EXPLICIT fmaf()/fma()
// call placed where nvcc contracts the reference's CUDA expressions; nothing else may be contracted).
#include "oracle.h"

#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

// ------------------------------------------------------------------------------------------------
// BLAS hook.  Caffe calls cblas_sgemm(CblasRowMajor, NoTrans, NoTrans, M, N, K, 1, A, K, B, N, beta, C, N)
// (math_functions.cpp:13-21).  The BLAS itself is third-party and unpinned in the reference
// (Makefile:369-386: ATLAS / MKL / OpenBLAS); we dlopen the OpenBLAS that ships in this image.
// Without one we fall back to a plain triple loop (same contraction order k-inner, fp32).
// ------------------------------------------------------------------------------------------------
typedef void (*sgemm_fn)(int order, int ta, int tb, int m, int n, int k, float alpha, const float* a, int lda,
                         const float* b, int ldb, float beta, float* c, int ldc);
static sgemm_fn g_sgemm = nullptr;
static void (*g_blas_set_threads)(int) = nullptr;

extern "C" int orc_load_blas(const char* path) {
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    void* f = dlsym(h, "cblas_sgemm");
    if (!f) f = dlsym(h, "scipy_cblas_sgemm");
    if (!f) f = dlsym(h, "cblas_sgemm64_");
    if (!f) return -2;
    g_sgemm = (sgemm_fn)f;
    void* t = dlsym(h, "openblas_set_num_threads");
    if (!t) t = dlsym(h, "scipy_openblas_set_num_threads");
    g_blas_set_threads = (void (*)(int))t;
    return 0;
}
extern "C" int orc_have_blas(void) { return g_sgemm != nullptr; }
extern "C" void orc_set_threads(int n) {
    if (g_blas_set_threads) g_blas_set_threads(n);
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
}

// C[MxN] = A[MxK] * B[KxN] + beta*C   (row major)
static void gemm_rm(int M, int N, int K, const float* A, const float* B, float beta, float* C) {
    if (g_sgemm) {
        g_sgemm(101 /*RowMajor*/, 111 /*NoTrans*/, 111, M, N, K, 1.0f, A, K, B, N, beta, C, N);
        return;
    }
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; m++) {
        float* c = C + (size_t)m * N;
        if (beta == 0.f) for (int n = 0; n < N; n++) c[n] = 0.f;
        for (int k = 0; k < K; k++) {
            const float a = A[(size_t)m * K + k];
            const float* b = B + (size_t)k * N;
            for (int n = 0; n < N; n++) c[n] = c[n] + a * b[n];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Caffe layer arithmetic
// ------------------------------------------------------------------------------------------------
// im2col.cpp:19-55: col rows ordered (channel, kernel_row, kernel_col); zero padding; dilation 1.
extern "C" void orc_im2col(const float* im, int channels, int height, int width, int kh, int kw, int ph, int pw,
                           int sh, int sw, float* col) {
    const int oh = (height + 2 * ph - kh) / sh + 1;
    const int ow = (width + 2 * pw - kw) / sw + 1;
#pragma omp parallel for schedule(static)
    for (int row = 0; row < channels * kh * kw; row++) {
        const int c = row / (kh * kw), kr = (row / kw) % kh, kc = row % kw;
        const float* src = im + (size_t)c * height * width;
        float* dst = col + (size_t)row * oh * ow;
        for (int y = 0; y < oh; y++) {
            const int iy = y * sh - ph + kr;
            if (iy < 0 || iy >= height) {
                for (int x = 0; x < ow; x++) dst[y * ow + x] = 0.f;
                continue;
            }
            for (int x = 0; x < ow; x++) {
                const int ix = x * sw - pw + kc;
                dst[y * ow + x] = (ix >= 0 && ix < width) ? src[iy * width + ix] : 0.f;
            }
        }
    }
}

// conv_layer.cpp:25-40 -> base_conv_layer.cpp:257-279: per image  out = W * col  then  out += bias * ones
extern "C" void orc_conv2d(const float* in, int n, int cin, int h, int w, const float* weight, const float* bias,
                           int cout, int k, int pad, float* out) {
    const int oh = h + 2 * pad - k + 1, ow = w + 2 * pad - k + 1;
    const size_t osp = (size_t)oh * ow;
    std::vector<float> col;
    std::vector<float> ones(osp, 1.0f);
    if (k != 1) col.resize((size_t)cin * k * k * osp);
    for (int i = 0; i < n; i++) {
        const float* im = in + (size_t)i * cin * h * w;
        float* o = out + (size_t)i * cout * osp;
        const float* cb = im;  // is_1x1_: col buffer is the input itself (base_conv_layer.cpp:260-266)
        if (k != 1) { orc_im2col(im, cin, h, w, k, k, pad, pad, 1, 1, col.data()); cb = col.data(); }
        gemm_rm(cout, (int)osp, cin * k * k, weight, cb, 0.f, o);
        if (bias) {  // forward_cpu_bias: gemm(M=cout, N=osp, K=1, bias, ones, beta=1)
            if (g_sgemm) gemm_rm(cout, (int)osp, 1, bias, ones.data(), 1.f, o);
            else for (int c = 0; c < cout; c++) for (size_t p = 0; p < osp; p++) o[c * osp + p] = o[c * osp + p] + bias[c] * 1.0f;
        }
    }
}

// relu_layer.cpp:9-19 with negative_slope 0:  max(x,0) + 0*min(x,0)
extern "C" void orc_relu(float* x, size_t count) {
    for (size_t i = 0; i < count; i++) x[i] = std::max(x[i], 0.f) + 0.f * std::min(x[i], 0.f);
}

// pooling_layer.cpp:90-93 (+ clip when padded, :94-106)
extern "C" int orc_pooled_dim(int in, int k, int stride, int pad) {
    int p = (int)ceilf((float)(in + 2 * pad - k) / stride) + 1;
    if (pad && (p - 1) * stride >= in + pad) --p;
    return p;
}
// pooling_layer.cpp:128-187 MAX: start at -FLT_MAX, strict > (first max wins)
extern "C" void orc_maxpool(const float* in, int n, int c, int h, int w, int k, int stride, int pad, float* out) {
    const int ph = orc_pooled_dim(h, k, stride, pad), pw = orc_pooled_dim(w, k, stride, pad);
#pragma omp parallel for schedule(static)
    for (int nc = 0; nc < n * c; nc++) {
        const float* b = in + (size_t)nc * h * w;
        float* t = out + (size_t)nc * ph * pw;
        for (int y = 0; y < ph; y++)
            for (int x = 0; x < pw; x++) {
                int hs = y * stride - pad, ws = x * stride - pad;
                const int he = std::min(hs + k, h), we = std::min(ws + k, w);
                hs = std::max(hs, 0); ws = std::max(ws, 0);
                float m = -3.402823466e+38F;
                for (int yy = hs; yy < he; yy++)
                    for (int xx = ws; xx < we; xx++)
                        if (b[yy * w + xx] > m) m = b[yy * w + xx];
                t[y * pw + x] = m;
            }
    }
}

// ------------------------------------------------------------------------------------------------
// Model descriptors (modelDescriptorFactory.cpp:6-28, 30-55; modelDescriptor.cpp:4-20)
// ------------------------------------------------------------------------------------------------
struct ModelDesc {
    int num_parts, num_limbs;
    std::vector<int> limb_seq, map_idx;
    std::vector<std::string> names;
};
static const ModelDesc& model_desc(int model) {
    static ModelDesc md[2];
    static bool init = false;
    if (!init) {
        const char* mpi[] = {"Head", "Neck", "RShoulder", "RElbow", "RWrist", "LShoulder", "LElbow", "LWrist", "RHip",
                             "RKnee", "RAnkle", "LHip", "LKnee", "LAnkle", "Chest", "Bkg"};
        const char* coco[] = {"Nose", "Neck", "RShoulder", "RElbow", "RWrist", "LShoulder", "LElbow", "LWrist", "RHip", "RKnee",
                              "RAnkle", "LHip", "LKnee", "LAnkle", "REye", "LEye", "REar", "LEar", "Bkg"};
        md[0].num_parts = 15;
        md[0].limb_seq = {0, 1, 1, 2, 2, 3, 3, 4, 1, 5, 5, 6, 6, 7, 1, 14, 14, 11, 11, 12, 12, 13, 14, 8, 8, 9, 9, 10};
        md[0].map_idx = {16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 38, 39, 40, 41, 42, 43, 32, 33, 34, 35, 36, 37};
        md[1].num_parts = 18;
        md[1].limb_seq = {1, 2, 1, 5, 2, 3, 3, 4, 5, 6, 6, 7, 1, 8, 8, 9, 9, 10, 1, 11, 11, 12, 12, 13, 1, 0, 0, 14, 14, 16, 0, 15, 15, 17, 2, 16, 5, 17};
        md[1].map_idx = {31, 32, 39, 40, 33, 34, 35, 36, 41, 42, 43, 44, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 47, 48, 49, 50, 53, 54, 51, 52, 55, 56, 37, 38, 45, 46};
        for (int m = 0; m < 2; m++) {
            ModelDesc& d = md[m];
            d.num_limbs = (int)d.limb_seq.size() / 2;
            d.names.assign(d.num_parts + 1 + 2 * d.num_limbs, "");
            for (int i = 0; i <= d.num_parts; i++) d.names[i] = m == 0 ? mpi[i] : coco[i];
            for (int l = 0; l < d.num_limbs; l++) {  // createPartToName
                const std::string base = d.names[d.limb_seq[2 * l]] + "->" + d.names[d.limb_seq[2 * l + 1]];
                d.names[d.map_idx[2 * l]] = base + "(X)";
                d.names[d.map_idx[2 * l + 1]] = base + "(Y)";
            }
        }
        init = true;
    }
    return md[model == ORC_MODEL_MPI_15 ? 0 : 1];
}
extern "C" int orc_model_num_parts(int model) { return model_desc(model).num_parts; }
extern "C" int orc_model_num_limbs(int model) { return model_desc(model).num_limbs; }
extern "C" int orc_model_num_maps(int model) { return model_desc(model).num_parts + 1 + 2 * model_desc(model).num_limbs; }
extern "C" const int* orc_model_limb_seq(int model) { return model_desc(model).limb_seq.data(); }
extern "C" const int* orc_model_map_idx(int model) { return model_desc(model).map_idx.data(); }
extern "C" const char* orc_model_map_name(int model, int idx) {
    const ModelDesc& d = model_desc(model);
    return (idx >= 0 && idx < (int)d.names.size()) ? d.names[idx].c_str() : "";
}

// ------------------------------------------------------------------------------------------------
// The deploy graph (model/{coco,mpi}/pose_deploy_linevec.prototxt), rebuilt procedurally; the CPU test
// suite compares this table with tests/golden/netspec_*.json (parsed from the prototxts).
// ------------------------------------------------------------------------------------------------
enum LType { L_CONV, L_RELU, L_POOL, L_CONCAT, L_RESIZE, L_NMS };
struct Layer {
    std::string name;
    LType type;
    std::vector<std::string> bottoms;
    std::string top;
    int num_output = 0, kernel = 0, pad = 0, stride = 1;
    int cin = 0;  // conv only, filled at build
    std::vector<float> w, b;
};
struct OrcNet {
    int model;
    std::vector<Layer> layers;
    std::map<std::string, int> channels;  // blob -> channels
};
static const char* ltype_name(LType t) {
    switch (t) {
        case L_CONV: return "Convolution";
        case L_RELU: return "ReLU";
        case L_POOL: return "Pooling";
        case L_CONCAT: return "Concat";
        case L_RESIZE: return "ImResize";
        default: return "Nms";
    }
}
static void add_conv(OrcNet* n, const std::string& name, const std::string& bottom, int cout, int k, const char* relu) {
    Layer l;
    l.name = name; l.type = L_CONV; l.bottoms = {bottom}; l.top = name;
    l.num_output = cout; l.kernel = k; l.pad = k / 2; l.stride = 1;
    l.cin = n->channels[bottom];
    n->channels[name] = cout;
    n->layers.push_back(l);
    if (relu) {
        Layer r;
        r.name = relu; r.type = L_RELU; r.bottoms = {name}; r.top = name;
        n->layers.push_back(r);
    }
}
static void add_pool(OrcNet* n, const std::string& name, const std::string& bottom) {
    Layer l;
    l.name = name; l.type = L_POOL; l.bottoms = {bottom}; l.top = name; l.kernel = 2; l.stride = 2; l.pad = 0;
    n->channels[name] = n->channels[bottom];
    n->layers.push_back(l);
}
static void add_concat(OrcNet* n, const std::string& name, const std::vector<std::string>& bottoms) {
    Layer l;
    l.name = name; l.type = L_CONCAT; l.bottoms = bottoms; l.top = name;
    int c = 0;
    for (auto& b : bottoms) c += n->channels[b];
    n->channels[name] = c;
    n->layers.push_back(l);
}
extern "C" OrcNet* orc_net_create(int model) { return orc_net_create_stages(model, 6); }
// stages = number of CPM stages: 6 = model/{coco,mpi}/pose_deploy_linevec.prototxt, 1 / 2 / 4 = model/mpi/pose_deploy_linevec_{1,2,4}.prototxt
extern "C" OrcNet* orc_net_create_stages(int model, int stages) {
    OrcNet* n = new OrcNet();
    n->model = model;
    const int cL1 = 2 * orc_model_num_limbs(model), cL2 = orc_model_num_parts(model) + 1;
    n->channels["image"] = 3;
    char nm[64], rl[64];
    std::string prev = "image";
    const int vgg[4][2] = {{64, 2}, {128, 2}, {256, 4}, {512, 2}};
    for (int b = 0; b < 4; b++) {
        for (int i = 1; i <= vgg[b][1]; i++) {
            snprintf(nm, 64, "conv%d_%d", b + 1, i);
            snprintf(rl, 64, "relu%d_%d", b + 1, i);
            add_conv(n, nm, prev, vgg[b][0], 3, rl);
            prev = nm;
        }
        if (b < 3) {
            snprintf(nm, 64, "pool%d_stage1", b + 1);
            add_pool(n, nm, prev);
            prev = nm;
        }
    }
    add_conv(n, "conv4_3_CPM", prev, 256, 3, "relu4_3_CPM");
    add_conv(n, "conv4_4_CPM", "conv4_3_CPM", 128, 3, "relu4_4_CPM");
    // stage 1: layers interleaved L1,L2 as in the prototxt
    std::string p1 = "conv4_4_CPM", p2 = "conv4_4_CPM";
    for (int i = 1; i <= 5; i++) {
        for (int br = 1; br <= 2; br++) {
            snprintf(nm, 64, "conv5_%d_CPM_L%d", i, br);
            snprintf(rl, 64, "relu5_%d_CPM_L%d", i, br);
            std::string& p = br == 1 ? p1 : p2;
            if (i <= 3) add_conv(n, nm, p, 128, 3, rl);
            else if (i == 4) add_conv(n, nm, p, 512, 1, rl);
            else add_conv(n, nm, p, br == 1 ? cL1 : cL2, 1, nullptr);
            p = nm;
        }
    }
    for (int s = 2; s <= stages; s++) {
        snprintf(nm, 64, "concat_stage%d", s);
        add_concat(n, nm, {p1, p2, "conv4_4_CPM"});
        p1 = p2 = nm;
        for (int i = 1; i <= 7; i++) {
            for (int br = 1; br <= 2; br++) {
                snprintf(nm, 64, "Mconv%d_stage%d_L%d", i, s, br);
                snprintf(rl, 64, "Mrelu%d_stage%d_L%d", i, s, br);
                std::string& p = br == 1 ? p1 : p2;
                if (i <= 5) add_conv(n, nm, p, 128, 7, rl);
                else if (i == 6) add_conv(n, nm, p, 128, 1, rl);
                else add_conv(n, nm, p, br == 1 ? cL1 : cL2, 1, nullptr);
                p = nm;
            }
        }
    }
    add_concat(n, "concat_stage7", {p2, p1});  // [L2 (parts+bkg), L1 (PAF)]  prototxt :2966-2975
    Layer r; r.name = "resize"; r.type = L_RESIZE; r.bottoms = {"concat_stage7"}; r.top = "resized_map";
    n->layers.push_back(r);
    Layer m; m.name = "nms"; m.type = L_NMS; m.bottoms = {"resized_map"}; m.top = "joints";
    n->layers.push_back(m);
    return n;
}
extern "C" void orc_net_destroy(OrcNet* net) { delete net; }
extern "C" int orc_net_num_layers(const OrcNet* net) { return (int)net->layers.size(); }
extern "C" int orc_net_layer_info(const OrcNet* net, int idx, char* name, char* type, char* bottoms, char* top,
                                  int* num_output, int* kernel, int* pad, int* stride) {
    if (idx < 0 || idx >= (int)net->layers.size()) return -1;
    const Layer& l = net->layers[idx];
    strcpy(name, l.name.c_str());
    strcpy(type, ltype_name(l.type));
    std::string b;
    for (size_t i = 0; i < l.bottoms.size(); i++) b += (i ? "," : "") + l.bottoms[i];
    strcpy(bottoms, b.c_str());
    strcpy(top, l.top.c_str());
    *num_output = l.num_output; *kernel = l.kernel; *pad = l.pad; *stride = l.stride;
    return 0;
}
extern "C" int orc_net_num_convs(const OrcNet* net) {
    int c = 0;
    for (auto& l : net->layers) c += l.type == L_CONV;
    return c;
}
extern "C" int orc_net_conv_info(const OrcNet* net, int conv_idx, char* name, int* cout, int* cin, int* k) {
    int c = 0;
    for (auto& l : net->layers)
        if (l.type == L_CONV && c++ == conv_idx) {
            strcpy(name, l.name.c_str());
            *cout = l.num_output; *cin = l.cin; *k = l.kernel;
            return 0;
        }
    return -1;
}
extern "C" int orc_net_set_weights(OrcNet* net, const char* conv_name, const float* w, const float* b) {
    for (auto& l : net->layers)
        if (l.type == L_CONV && l.name == conv_name) {
            const size_t nw = (size_t)l.num_output * l.cin * l.kernel * l.kernel;
            l.w.assign(w, w + nw);
            l.b.assign(b, b + l.num_output);
            return 0;
        }
    return -1;
}
extern "C" double orc_net_flops(int model, int h, int w) {
    OrcNet* n = orc_net_create(model);
    std::map<std::string, int> hh, ww;
    hh["image"] = h; ww["image"] = w;
    double f = 0;
    for (auto& l : n->layers) {
        if (l.type == L_RESIZE || l.type == L_NMS) break;
        int bh = hh[l.bottoms[0]], bw = ww[l.bottoms[0]];
        if (l.type == L_POOL) { bh = orc_pooled_dim(bh, 2, 2, 0); bw = orc_pooled_dim(bw, 2, 2, 0); }
        if (l.type == L_CONV) f += 2.0 * l.num_output * l.cin * l.kernel * l.kernel * bh * bw;
        hh[l.top] = bh; ww[l.top] = bw;
    }
    orc_net_destroy(n);
    return f;
}

struct Blob { int c = 0, h = 0, w = 0; std::vector<float> d; };

extern "C" int orc_net_forward_blob(OrcNet* net, const float* input, int num, int h, int w, const char* want,
                                    float* blob_out, size_t blob_cap, int* bc, int* bh, int* bw) {
    std::map<std::string, Blob> blobs;
    Blob& in = blobs["image"];
    in.c = 3; in.h = h; in.w = w;
    in.d.assign(input, input + (size_t)num * 3 * h * w);
    for (auto& l : net->layers) {
        if (l.type == L_RESIZE || l.type == L_NMS) break;
        if (l.type == L_CONV) {
            if (l.w.empty()) return -2;
            const Blob& b = blobs[l.bottoms[0]];
            Blob t; t.c = l.num_output; t.h = b.h; t.w = b.w;
            t.d.resize((size_t)num * t.c * t.h * t.w);
            orc_conv2d(b.d.data(), num, b.c, b.h, b.w, l.w.data(), l.b.data(), l.num_output, l.kernel, l.pad, t.d.data());
            blobs[l.top] = std::move(t);
        } else if (l.type == L_RELU) {
            Blob& b = blobs[l.top];
            orc_relu(b.d.data(), b.d.size());
        } else if (l.type == L_POOL) {
            const Blob& b = blobs[l.bottoms[0]];
            Blob t; t.c = b.c; t.h = orc_pooled_dim(b.h, 2, 2, 0); t.w = orc_pooled_dim(b.w, 2, 2, 0);
            t.d.resize((size_t)num * t.c * t.h * t.w);
            orc_maxpool(b.d.data(), num, b.c, b.h, b.w, 2, 2, 0, t.d.data());
            blobs[l.top] = std::move(t);
        } else if (l.type == L_CONCAT) {  // concat_layer.cpp:57-74, axis 1
            Blob t; t.h = blobs[l.bottoms[0]].h; t.w = blobs[l.bottoms[0]].w; t.c = 0;
            for (auto& bn : l.bottoms) t.c += blobs[bn].c;
            const size_t sp = (size_t)t.h * t.w;
            t.d.resize((size_t)num * t.c * sp);
            for (int i = 0; i < num; i++) {
                int off = 0;
                for (auto& bn : l.bottoms) {
                    const Blob& b = blobs[bn];
                    memcpy(&t.d[((size_t)i * t.c + off) * sp], &b.d[(size_t)i * b.c * sp], (size_t)b.c * sp * sizeof(float));
                    off += b.c;
                }
            }
            blobs[l.top] = std::move(t);
        }
        // free blobs no longer needed is skipped: sizes are modest for the oracle
        if (want && blob_out && l.top == want && (l.type != L_CONV || true)) {
            // copy after this layer; a later in-place ReLU on the same top overwrites again below
            const Blob& b = blobs[l.top];
            if (b.d.size() <= blob_cap) memcpy(blob_out, b.d.data(), b.d.size() * sizeof(float));
            if (bc) *bc = b.c;
            if (bh) *bh = b.h;
            if (bw) *bw = b.w;
        }
    }
    return 0;
}
extern "C" int orc_net_forward(OrcNet* net, const float* input, int num, int h, int w, float* out) {
    const int c = orc_model_num_maps(net->model);
    return orc_net_forward_blob(net, input, num, h, w, "concat_stage7", out, (size_t)num * c * (h / 8) * (w / 8), nullptr,
                                nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------------
// ImResize, GPU-kernel arithmetic (imresize_layer.cu:8-18, 98-155).
// FMA placement = what nvcc emits for the reference expressions (SASS of oracle/_ref/libref_cpm.so):
//   a  = fma(v3,.5f, fma(v2,-1.5f, fma(v1,1.5f, -(.5f*v0))))           (float)
//   T1 = ((a*d)*d)*d                                                   (float)
//   b  = fma((double)v3,-.5, fma((double)v2,2., (double)fmaf(v1,-2.5f,v0)))
//   acc= fma((double)d, (double)d*b, (double)T1) + (double)(d*fmaf(v2,.5f,-(.5f*v0))) + (double)v1
// ------------------------------------------------------------------------------------------------
static inline float cubic_ref(float v0, float v1, float v2, float v3, float d) {
    const float h = 0.5f * v0;
    const float a = fmaf(v3, 0.5f, fmaf(v2, -1.5f, fmaf(v1, 1.5f, -h)));
    const float t1 = ((a * d) * d) * d;
    const double b = fma((double)v3, -0.5, fma((double)v2, 2.0, (double)fmaf(v1, -2.5f, v0)));
    const double dd = (double)d;
    double acc = fma(dd, dd * b, (double)t1);
    const float t3 = d * fmaf(v2, 0.5f, -h);
    acc = acc + (double)t3;
    acc = acc + (double)v1;
    return (float)acc;
}

struct ResizeAxis { int i0, i1, i2, i3; float d; };
// one axis of imresize_cubic_kernel for output coordinate `x`, scale index n
static inline ResizeAxis resize_axis(int x, int ori, int t, int n, float start_scale, float scale_gap) {
    const float f = fmaf((float)n, scale_gap, 1 - start_scale);  // (1-start_scale + n*scale_gap), contracted
    const int pad = (int)floorf((float)(ori / 2) * f);
    const int o = ori - 2 * pad;
    const float q = (float)t / (float)o;
    const float offset = fmaf(q, 0.5f, -0.5f);  // tw/float(ow)/2 - 0.5  (double sub narrows to float exactly)
    const float on_ori = ((float)x - offset) * ((float)o / (float)t);
    ResizeAxis r;
    int n1 = (int)((double)on_ori + 1e-5);
    n1 = n1 < 0 ? 0 : n1;
    r.i0 = ((n1 - 1 < 0) ? n1 : (n1 - 1)) + pad;
    int n2 = (n1 + 1 >= o) ? (o - 1) : (n1 + 1);
    r.i3 = ((n2 + 1 >= o) ? (o - 1) : (n2 + 1)) + pad;
    r.d = on_ori - (float)n1;
    r.i1 = n1 + pad;
    r.i2 = n2 + pad;
    return r;
}

static inline float imresize_pixel(const float* src_c, int num, size_t src_offset, int h8, int w8, int th, int tw,
                                   float start_scale, float scale_gap, int y, int x) {
    float sum = 0;
    for (int n = 0; n < num; n++) {
        const ResizeAxis ax = resize_axis(x, w8, tw, n, start_scale, scale_gap);
        const ResizeAxis ay = resize_axis(y, h8, th, n, start_scale, scale_gap);
        const float* s = src_c + (size_t)n * src_offset;
        const int yi[4] = {ay.i0, ay.i1, ay.i2, ay.i3};
        float temp[4];
        for (int i = 0; i < 4; i++) {
            const float* row = s + (size_t)yi[i] * w8;  // row stride (ow + 2*padw) == w8
            temp[i] = cubic_ref(row[ax.i0], row[ax.i1], row[ax.i2], row[ax.i3], ax.d);
        }
        const float d_temp = cubic_ref(temp[0], temp[1], temp[2], temp[3], ay.d);
        sum = sum + d_temp;
    }
    return sum / (float)num;
}

extern "C" void orc_imresize(const float* src, int num, int channels, int h8, int w8, int th, int tw, float start_scale,
                             float scale_gap, float* dst) {
    const size_t src_offset = (size_t)channels * h8 * w8;
#pragma omp parallel for schedule(dynamic, 1)
    for (int cy = 0; cy < channels * th; cy++) {
        const int c = cy / th, y = cy % th;
        const float* src_c = src + (size_t)c * h8 * w8;
        float* d = dst + ((size_t)c * th + y) * tw;
        for (int x = 0; x < tw; x++)
            d[x] = imresize_pixel(src_c, num, src_offset, h8, w8, th, tw, start_scale, scale_gap, y, x);
    }
}
extern "C" float orc_imresize_at(const float* src, int num, int channels, int h8, int w8, int th, int tw,
                                 float start_scale, float scale_gap, int c, int y, int x) {
    return imresize_pixel(src + (size_t)c * h8 * w8, num, (size_t)channels * h8 * w8, h8, w8, th, tw, start_scale,
                          scale_gap, y, x);
}

// ------------------------------------------------------------------------------------------------
// NMS, GPU-kernel arithmetic (nms_layer.cu:14-46 register, :176 exclusive scan, :49-113 write)
// ------------------------------------------------------------------------------------------------
extern "C" void orc_nms(const float* map, int channels, int height, int width, int num_parts, int max_peaks,
                        float threshold, float* peaks) {
    const int offset = height * width, offset_dst = (max_peaks + 1) * 3;
    (void)channels;
    memset(peaks, 0, sizeof(float) * (size_t)num_parts * offset_dst);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < num_parts; c++) {
        const float* src = map + (size_t)c * offset;
        float* out = peaks + (size_t)c * offset_dst;
        int peak_index = 0;  // == exclusive-scan value at the peak
        for (int y = 1; y < height - 1; y++)
            for (int x = 1; x < width - 1; x++) {
                const float v = src[y * width + x];
                if (!(v > threshold)) continue;
                if (!(v > src[(y - 1) * width + x] && v > src[(y + 1) * width + x] && v > src[y * width + x - 1] &&
                      v > src[y * width + x + 1] && v > src[(y - 1) * width + x - 1] && v > src[(y + 1) * width + x - 1] &&
                      v > src[(y + 1) * width + x + 1] && v > src[(y - 1) * width + x + 1]))
                    continue;
                if (peak_index < max_peaks) {
                    float x_acc = 0.f, y_acc = 0.f, score_acc = 0.f;
                    for (int dy = -3; dy < 4; dy++) {
                        if ((y + dy) > 0 && (y + dy) < width) {  // sic: width (nms_layer.cu:79)
                            for (int dx = -3; dx < 4; dx++) {
                                if ((x + dx) > 0 && (x + dx) < width) {
                                    const float score = src[(y + dy) * width + x + dx];  // may alias the next channel
                                    const float fx = (float)(x + dx), fy = (float)(y + dy);
                                    if (score > 0) {
                                        x_acc = fmaf(fx, score, x_acc);  // nvcc contracts x_acc += x*score
                                        y_acc = fmaf(fy, score, y_acc);
                                        score_acc = score_acc + score;
                                    }
                                }
                            }
                        }
                    }
                    const int oi = (peak_index + 1) * 3;
                    out[oi] = x_acc / score_acc;
                    out[oi + 1] = y_acc / score_acc;
                    out[oi + 2] = v;
                }
                peak_index++;
            }
        out[0] = (float)peak_index;  // total, NOT clamped (nms_layer.cu:110)
    }
}

// ------------------------------------------------------------------------------------------------
// connectLimbs (MPI, rtpose.cpp:549-751) / connectLimbsCOCO (rtpose.cpp:808-1076)
// Differences kept: MPI uses pow() (double) for the norm, has no upper clamp of the sample
// coordinates and no duplicate check in the nA==0 / nB==0 branches.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_default_params(int model, float* nms_threshold, OrcConnectParams* p) {
    // rtpose.cpp:212-226
    p->min_subset_cnt = 3;
    p->min_subset_score = 0.4f;
    p->clamp_counts = 1;
    if (model == ORC_MODEL_MPI_15) {
        if (nms_threshold) *nms_threshold = 0.2f;
        p->inter_threshold = 0.01f;
        p->inter_min_above = 8;
    } else {
        if (nms_threshold) *nms_threshold = 0.05f;
        p->inter_threshold = 0.050f;
        p->inter_min_above = 9;
    }
}

struct Cand { double i, j, conn, all; };
struct CandGreater { bool operator()(const Cand& l, const Cand& r) const { return l.conn > r.conn; } };  // ColumnCompare :144-152

extern "C" int orc_connect(int model, const float* heatmap, const float* peaks, int max_peaks, int net_w, int net_h,
                           int disp_w, int disp_h, const OrcConnectParams* prm, float* joints, double* subset_out,
                           int subset_cap, int* subset_rows) {
    const ModelDesc& md = model_desc(model);
    const bool coco = model != ORC_MODEL_MPI_15;
    const int num_parts = md.num_parts, nlimb = md.num_limbs;
    const int S_CNT = num_parts + 2, S_SCORE = num_parts + 1, S_SIZE = num_parts + 3;
    const int peaks_offset = 3 * (max_peaks + 1);
    const size_t plane = (size_t)net_h * net_w;
    std::vector<std::vector<double>> subset;

    for (int k = 0; k < nlimb; k++) {
        const int pa = md.limb_seq[2 * k], pb = md.limb_seq[2 * k + 1];
        const float* map_x = heatmap + (size_t)md.map_idx[2 * k] * plane;
        const float* map_y = heatmap + (size_t)md.map_idx[2 * k + 1] * plane;
        const float* candA = peaks + pa * peaks_offset;
        const float* candB = peaks + pb * peaks_offset;
        int nA = (int)candA[0], nB = (int)candB[0];
        if (prm->clamp_counts) { nA = std::min(nA, max_peaks); nB = std::min(nB, max_peaks); }

        if (nA == 0 && nB == 0) continue;
        if (nA == 0 || nB == 0) {
            const int part = nA == 0 ? pb : pa;
            const float* cand = nA == 0 ? candB : candA;
            const int n = nA == 0 ? nB : nA;
            for (int i = 1; i <= n; i++) {
                const int off = part * peaks_offset + i * 3 + 2;
                int num = 0;
                if (coco)
                    for (size_t j = 0; j < subset.size(); j++)
                        if (subset[j][part] == off) num++;
                if (num == 0) {
                    std::vector<double> row(S_SIZE, 0);
                    row[part] = off;
                    row[S_CNT] = 1;
                    row[S_SCORE] = cand[i * 3 + 2];
                    subset.push_back(row);
                }
            }
            continue;
        }

        std::vector<Cand> temp;
        const int num_inter = 10;
        for (int i = 1; i <= nA; i++)
            for (int j = 1; j <= nB; j++) {
                const float s_x = candA[i * 3], s_y = candA[i * 3 + 1];
                const float d_x = candB[j * 3] - candA[i * 3];
                const float d_y = candB[j * 3 + 1] - candA[i * 3 + 1];
                float norm_vec;
                if (coco) norm_vec = sqrtf(d_x * d_x + d_y * d_y);
                else norm_vec = (float)sqrt(pow((double)d_x, 2) + pow((double)d_y, 2));
                if (norm_vec < 1e-6) continue;
                const float vec_x = d_x / norm_vec, vec_y = d_y / norm_vec;
                float sum = 0;
                int count = 0;
                for (int lm = 0; lm < num_inter; lm++) {
                    int my = (int)roundf(s_y + lm * d_y / num_inter);
                    int mx = (int)roundf(s_x + lm * d_x / num_inter);
                    if (coco) {
                        if (mx >= net_w) mx = net_w - 1;
                        if (my >= net_h) my = net_h - 1;
                    }
                    const int idx = my * net_w + mx;
                    const float score = vec_x * map_x[idx] + vec_y * map_y[idx];
                    if (score > prm->inter_threshold) { sum = sum + score; count++; }
                }
                if (count > prm->inter_min_above) {
                    Cand c;
                    c.all = sum / count + candA[i * 3 + 2] + candB[j * 3 + 2];
                    c.conn = sum / count;
                    c.i = i; c.j = j;
                    temp.push_back(c);
                }
            }
        if (!temp.empty()) std::sort(temp.begin(), temp.end(), CandGreater());

        const int num = std::min(nA, nB);
        int cnt = 0;
        std::vector<int> occurA(nA, 0), occurB(nB, 0);
        struct Conn { double a, b, s; };
        std::vector<Conn> conn_k;
        for (size_t row = 0; row < temp.size(); row++) {
            if (cnt == num) break;
            const int i = (int)temp[row].i, j = (int)temp[row].j;
            const float score = (float)temp[row].conn;
            if (occurA[i - 1] == 0 && occurB[j - 1] == 0) {
                Conn c;
                c.a = pa * peaks_offset + i * 3 + 2;
                c.b = pb * peaks_offset + j * 3 + 2;
                c.s = score;
                conn_k.push_back(c);
                cnt++;
                occurA[i - 1] = 1; occurB[j - 1] = 1;
            }
        }

        if (k == 0) {
            std::vector<double> row(S_SIZE, 0);
            for (size_t i = 0; i < conn_k.size(); i++) {
                row[md.limb_seq[0]] = conn_k[i].a;
                row[md.limb_seq[1]] = conn_k[i].b;
                row[S_CNT] = 2;
                row[S_SCORE] = peaks[(int)conn_k[i].a] + peaks[(int)conn_k[i].b] + conn_k[i].s;
                subset.push_back(row);
            }
        } else {
            if (conn_k.empty()) continue;
            for (size_t i = 0; i < conn_k.size(); i++) {
                int nfound = 0;
                const double indexA = conn_k[i].a, indexB = conn_k[i].b;
                for (size_t j = 0; j < subset.size(); j++) {
                    if (subset[j][pa] == indexA) {
                        subset[j][pb] = indexB;
                        nfound++;
                        subset[j][S_CNT] = subset[j][S_CNT] + 1;
                        subset[j][S_SCORE] = subset[j][S_SCORE] + peaks[(int)indexB] + conn_k[i].s;
                    }
                }
                if (nfound == 0) {
                    std::vector<double> row(S_SIZE, 0);
                    row[pa] = indexA;
                    row[pb] = indexB;
                    row[S_CNT] = 2;
                    row[S_SCORE] = peaks[(int)indexA] + peaks[(int)indexB] + conn_k[i].s;
                    subset.push_back(row);
                }
            }
        }
    }

    int cnt = 0;
    for (size_t i = 0; i < subset.size(); i++) {
        if (subset[i][S_CNT] >= prm->min_subset_cnt && (subset[i][S_SCORE] / subset[i][S_CNT]) > prm->min_subset_score) {
            for (int j = 0; j < num_parts; j++) {
                const int idx = (int)subset[i][j];
                float* o = joints + (size_t)cnt * num_parts * 3 + j * 3;
                if (idx) {
                    o[2] = peaks[idx];
                    o[1] = peaks[idx - 1] * disp_h / (float)net_h;
                    o[0] = peaks[idx - 2] * disp_w / (float)net_w;
                } else {
                    o[0] = o[1] = o[2] = 0;
                }
            }
            cnt++;
            if (cnt == 96) break;  // MAX_PEOPLE
        }
    }
    if (subset_rows) *subset_rows = (int)subset.size();
    if (subset_out)
        for (int i = 0; i < (int)subset.size() && i < subset_cap; i++)
            for (int j = 0; j < S_SIZE; j++) subset_out[i * S_SIZE + j] = subset[i][j];
    return cnt;
}

// ------------------------------------------------------------------------------------------------
// Preprocess: OpenCV INTER_AREA (cv::resize, imgproc/resize.cpp: computeResizeAreaTab + ResizeArea_Invoker,
// WT=float for 8U; integer ratios take resizeAreaFast_).  Third-party arithmetic, pinned by
// tests/golden/area_*.npz generated with cv2 4.13 in the build container.
// ------------------------------------------------------------------------------------------------
struct DecAlpha { int si, di; float alpha; };
static int area_tab(int ssize, int dsize, int cn, double scale, std::vector<DecAlpha>& tab) {
    tab.clear();
    for (int dx = 0; dx < dsize; dx++) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = std::min(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        if (sx1 - fsx1 > 1e-3) tab.push_back({(sx1 - 1) * cn, dx * cn, (float)((sx1 - fsx1) / cell)});
        for (int sx = sx1; sx < sx2; sx++) tab.push_back({sx * cn, dx * cn, (float)(1.0 / cell)});
        if (fsx2 - sx2 > 1e-3) tab.push_back({sx2 * cn, dx * cn, (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)});
    }
    return (int)tab.size();
}
static inline uint8_t sat_u8(float v) {
    long r = lrintf(v);  // cvRound: round half to even
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
extern "C" int orc_resize_area_u8c3(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw) {
    const int cn = 3;
    if (dh == sh && dw == sw) { memcpy(dst, src, (size_t)sh * sw * cn); return 0; }
    const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    const double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
    if (scale_x < 1 || scale_y < 1) {
        // cv::resize leaves the area path as soon as ONE axis enlarges: INTER_AREA then means the fixed-point bilinear
        // resize (INTER_RESIZE_COEF_BITS = 11) with "area mode" sample positions on BOTH axes (imgproc/resize.cpp:
        // sx = floor(dx*scale), fx = (dx+1) - (sx+1)*inv_scale clipped to [0,1); HResizeLinear + VResizeLinear<uchar>).
        struct Lin { int s, a0, a1; };
        auto tab = [](int ssize, int dsize, std::vector<Lin>& t) {
            const double inv = (double)dsize / ssize, scale = 1. / inv;
            t.resize(dsize);
            for (int dx = 0; dx < dsize; dx++) {
                int sx = (int)floor(dx * scale);
                float fx = (float)((dx + 1) - (sx + 1) * inv);
                fx = fx <= 0 ? 0.f : fx - floorf(fx);
                if (sx < 0) { fx = 0; sx = 0; }
                if (sx >= ssize - 1) { fx = 0; sx = ssize - 1; }
                t[dx].s = sx;
                t[dx].a0 = (int)std::min(32767L, std::max(-32768L, lrintf((1.f - fx) * 2048)));
                t[dx].a1 = (int)std::min(32767L, std::max(-32768L, lrintf(fx * 2048)));
            }
        };
        std::vector<Lin> tx, ty;
        tab(sw, dw, tx);
        tab(sh, dh, ty);
        for (int dy = 0; dy < dh; dy++) {
            const int y0 = ty[dy].s, y1 = std::min(y0 + 1, sh - 1);
            for (int dx = 0; dx < dw; dx++) {
                const int x0 = tx[dx].s, x1 = std::min(x0 + 1, sw - 1);
                for (int c = 0; c < cn; c++) {
                    const int h0 = src[((size_t)y0 * sw + x0) * cn + c] * tx[dx].a0 + src[((size_t)y0 * sw + x1) * cn + c] * tx[dx].a1;
                    const int h1 = src[((size_t)y1 * sw + x0) * cn + c] * tx[dx].a0 + src[((size_t)y1 * sw + x1) * cn + c] * tx[dx].a1;
                    const int v = ((ty[dy].a0 * (h0 >> 4)) >> 16) + ((ty[dy].a1 * (h1 >> 4)) >> 16);
                    const int o = (v + 2) >> 2;
                    dst[((size_t)dy * dw + dx) * cn + c] = (uint8_t)(o < 0 ? 0 : (o > 255 ? 255 : o));
                }
            }
        }
        return 0;
    }
    const int iscale_x = (int)lrint(scale_x), iscale_y = (int)lrint(scale_y);
    const bool fast = fabs(scale_x - iscale_x) < 2.220446049250313e-16 && fabs(scale_y - iscale_y) < 2.220446049250313e-16;
    if (fast) {  // resizeAreaFast_<uchar,int>: integer sum, saturate_cast<uchar>(sum*scale) with float scale
        const int area = iscale_x * iscale_y;
        const float scale = 1.f / area;
        for (int dy = 0; dy < dh; dy++)
            for (int dx = 0; dx < dw; dx++)
                for (int c = 0; c < cn; c++) {
                    int sum = 0;
                    for (int yy = 0; yy < iscale_y; yy++)
                        for (int xx = 0; xx < iscale_x; xx++)
                            sum += src[((size_t)(dy * iscale_y + yy) * sw + dx * iscale_x + xx) * cn + c];
                    // 2x2 takes OpenCV's ResizeAreaFastVec (SIMD) specialisation: (sum + 2) >> 2
                    dst[((size_t)dy * dw + dx) * cn + c] =
                        (iscale_x == 2 && iscale_y == 2) ? (uint8_t)((sum + 2) >> 2) : sat_u8((float)sum * scale);
                }
        return 0;
    }
    std::vector<DecAlpha> xtab, ytab;
    area_tab(sw, dw, cn, scale_x, xtab);
    area_tab(sh, dh, 1, scale_y, ytab);
    const int W = dw * cn;
    std::vector<float> buf(W), sum(W, 0.f);
    int prev_dy = ytab[0].di;
    for (size_t j = 0; j < ytab.size(); j++) {
        const float beta = ytab[j].alpha;
        const int dy = ytab[j].di, sy = ytab[j].si;
        const uint8_t* S = src + (size_t)sy * sw * cn;
        std::fill(buf.begin(), buf.end(), 0.f);
        for (size_t k = 0; k < xtab.size(); k++) {
            const int sxn = xtab[k].si, dxn = xtab[k].di;
            const float alpha = xtab[k].alpha;
            buf[dxn] = buf[dxn] + S[sxn] * alpha;
            buf[dxn + 1] = buf[dxn + 1] + S[sxn + 1] * alpha;
            buf[dxn + 2] = buf[dxn + 2] + S[sxn + 2] * alpha;
        }
        if (dy != prev_dy) {
            uint8_t* D = dst + (size_t)prev_dy * W;
            for (int dx = 0; dx < W; dx++) { D[dx] = sat_u8(sum[dx]); sum[dx] = beta * buf[dx]; }
            prev_dy = dy;
        } else {
            for (int dx = 0; dx < W; dx++) sum[dx] = sum[dx] + beta * buf[dx];
        }
    }
    uint8_t* D = dst + (size_t)prev_dy * W;
    for (int dx = 0; dx < W; dx++) D[dx] = sat_u8(sum[dx]);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Display image: cv::warpAffine(frame, diag(s,s), display size, INTER_CUBIC, BORDER_CONSTANT 0)  (rtpose.cpp:474-487).
// Third-party arithmetic (OpenCV imgproc/imgwarp.cpp: fixed-point coordinates AB_BITS=10 / INTER_BITS=5, bicubic
// weights a=-0.75 as 15-bit shorts normalised to sum 32768, rounding (sum + 2^14) >> 15), restated and pinned to
// cv2 4.13 fixtures (tests/golden/warp_cv2.npz).
// ------------------------------------------------------------------------------------------------
static const short* warp_cubic_tab() {
    static short tab[32 * 32 * 16];
    static bool init = false;
    if (!init) {
        float t1[32][4];
        const float A = -0.75f, scale = 1.f / 32;
        for (int i = 0; i < 32; i++) {
            const float x = i * scale;
            t1[i][0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
            t1[i][1] = ((A + 2) * x - (A + 3)) * x * x + 1;
            t1[i][2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
            t1[i][3] = 1.f - t1[i][0] - t1[i][1] - t1[i][2];
        }
        for (int i = 0; i < 32; i++)
            for (int j = 0; j < 32; j++) {
                short* w = tab + (i * 32 + j) * 16;
                int isum = 0;
                for (int k1 = 0; k1 < 4; k1++)
                    for (int k2 = 0; k2 < 4; k2++) {
                        const float v = t1[i][k1] * t1[j][k2];
                        long r = lrintf(v * 32768.f);
                        r = r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
                        w[k1 * 4 + k2] = (short)r;
                        isum += (int)r;
                    }
                if (isum != 32768) {
                    const int diff = isum - 32768;
                    int Mi = 2 * 4 + 2, mi = 2 * 4 + 2;
                    for (int k1 = 2; k1 < 4; k1++)
                        for (int k2 = 2; k2 < 4; k2++) {
                            if (w[k1 * 4 + k2] < w[mi]) mi = k1 * 4 + k2;
                            else if (w[k1 * 4 + k2] > w[Mi]) Mi = k1 * 4 + k2;
                        }
                    if (diff < 0) w[Mi] = (short)(w[Mi] - diff);
                    else w[mi] = (short)(w[mi] - diff);
                }
            }
        init = true;
    }
    return tab;
}

// rtpose.cpp:474-480
extern "C" double orc_display_scale(int cols, int rows, int disp_w, int disp_h) {
    if (cols / (double)rows > disp_w / (double)disp_h) return disp_w / (double)cols;
    return disp_h / (double)rows;
}

extern "C" void orc_warp_affine_cubic_u8c3(const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw, double scale) {
    double M[6] = {scale, 0, 0, 0, scale, 0};
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    const double b1 = -M[0] * M[2] - M[1] * M[5], b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    const short* tab = warp_cubic_tab();
    std::vector<int> adelta(dw), bdelta(dw);
    for (int x = 0; x < dw; x++) { adelta[x] = (int)lrint(M[0] * x * 1024); bdelta[x] = (int)lrint(M[3] * x * 1024); }
    const int round_delta = 1024 / 32 / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; y++) {
        const int X0 = (int)lrint((M[1] * y + M[2]) * 1024) + round_delta, Y0 = (int)lrint((M[4] * y + M[5]) * 1024) + round_delta;
        for (int x = 0; x < dw; x++) {
            const int X = (X0 + adelta[x]) >> 5, Y = (Y0 + bdelta[x]) >> 5;
            const int sx = (X >> 5) - 1, sy = (Y >> 5) - 1;
            const short* w = tab + ((Y & 31) * 32 + (X & 31)) * 16;
            int acc[3] = {0, 0, 0};
            for (int k1 = 0; k1 < 4; k1++) {
                const int yy = sy + k1;
                if (yy < 0 || yy >= sh) continue;
                for (int k2 = 0; k2 < 4; k2++) {
                    const int xx = sx + k2;
                    if (xx < 0 || xx >= sw) continue;
                    const uint8_t* p = src + ((size_t)yy * sw + xx) * 3;
                    const int ww = w[k1 * 4 + k2];
                    acc[0] += p[0] * ww; acc[1] += p[1] * ww; acc[2] += p[2] * ww;
                }
            }
            for (int c = 0; c < 3; c++) {
                const int v = (acc[c] + (1 << 14)) >> 15;
                dst[((size_t)y * dw + x) * 3 + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
    }
}

// rtpose.cpp:508-511: float scale = START_SCALE - i*SCALE_GAP (double math, stored float);
// target = 16*ceil(NET*scale/16)
extern "C" void orc_scale_target(int net_w, int net_h, double start_scale, double scale_gap, int i, int* tw, int* th) {
    const float scale = (float)(start_scale - i * scale_gap);
    *tw = (int)(16 * ceil(net_w * scale / 16));
    *th = (int)(16 * ceil(net_h * scale / 16));
}

extern "C" int orc_preprocess(const uint8_t* disp, int disp_h, int disp_w, int net_h, int net_w, int num_scales,
                              double start_scale, double scale_gap, float* out) {
    const size_t offset = (size_t)3 * net_h * net_w;
    for (int i = 0; i < num_scales; i++) {
        int tw, th;
        orc_scale_target(net_w, net_h, start_scale, scale_gap, i, &tw, &th);
        if (tw > net_w || th > net_h) return -2;  // CHECK_LE rtpose.cpp:513-514
        std::vector<uint8_t> tmp((size_t)tw * th * 3);
        if (orc_resize_area_u8c3(disp, disp_h, disp_w, tmp.data(), th, tw)) return -1;
        // process_and_pad_image(target, image_temp, NET_W, NET_H, normalize=1)  rtpose.cpp:239-269
        const int padw = (net_w - tw) / 2, padh = (net_h - th) / 2;
        float* target = out + i * offset;
        for (int c = 0; c < 3; c++)
            for (int y = 0; y < net_h; y++) {
                const int oy = y - padh;
                for (int x = 0; x < net_w; x++) {
                    const int ox = x - padw;
                    float v = 0;
                    if (ox >= 0 && ox < tw && oy >= 0 && oy < th) v = (float)tmp[((size_t)oy * tw + ox) * 3 + c] / 256.0f - 0.5f;
                    target[(size_t)c * net_h * net_w + (size_t)y * net_w + x] = v;
                }
            }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// JSON (rtpose.cpp:1383-1416); default ostream formatting == "%g" (6 significant digits)
// ------------------------------------------------------------------------------------------------
// =====================================================================================================
// Renderers (TEST ONLY): CPU restatement of render() (examples/rtpose/rtpose.cpp:271-300) and of the six kernels of
// src/rtpose/renderFunctions.cu.  The reference runs these on the GPU only; float vs double promotions follow the
// kernels statement by statement, sums of products are evaluated WITHOUT fusing (the GPU build of the reference fuses
// some of them and its sinf/cosf/atan2f differ from libm in the last place), so GPU parity against this restatement is
// "equal up to 1e-3 on the float canvas except at shape borders"; the bit-level pin is the reference's own kernels
// compiled into oracle/_ref/libref_render.so (tests/test_gpu_render.py).
// =====================================================================================================
namespace {

const int kLimbMpi[] = {0, 1, 2, 3, 3, 4, 5, 6, 6, 7, 8, 9, 9, 10, 11, 12, 12, 13};                       // renderFunctions.cu:7
const int kLimbCoco[] = {1, 2, 1, 5, 2, 3, 3, 4, 5, 6, 6, 7, 1, 8, 8, 9, 9, 10, 1, 11, 11, 12, 12, 13,
                         1, 0, 0, 14, 14, 16, 0, 15, 15, 17};                                               // :9 (NOEAR)
const int kColor9[] = {255, 0, 0, 255, 170, 0, 170, 255, 0, 0, 255, 0, 0, 255, 170, 0, 170, 255, 0, 0, 255, 170, 0, 255, 255, 0, 170};   // :145-153
const int kColor18[] = {255, 0, 0, 255, 85, 0, 255, 170, 0, 255, 255, 0, 170, 255, 0, 85, 255, 0, 0, 255, 0, 0, 255, 85, 0, 255, 170,
                        0, 255, 255, 0, 170, 255, 0, 85, 255, 0, 0, 255, 85, 0, 255, 170, 0, 255, 255, 0, 255, 255, 0, 170, 255, 0, 85};   // :461-479

struct Bgr { float b, g, r; };

inline void mix(Bgr& px, float alpha, float cb, float cg, float cr) {   // b = (1-alpha)*b + alpha*colour
    px.b = (1 - alpha) * px.b + alpha * cb;
    px.g = (1 - alpha) * px.g + alpha * cg;
    px.r = (1 - alpha) * px.r + alpha * cr;
}

// one limb ellipse test (renderFunctions.cu:181-205 / 513-530): returns judge, or -1 when a joint is missing
inline bool limb_judge(const float* pose, int pa, int pb, float thr, int x, int y, float b_sqrt, bool head, float* judge_out) {
    const float x_a = pose[pa * 3], x_b = pose[pb * 3], y_a = pose[pa * 3 + 1], y_b = pose[pb * 3 + 1];
    if (!(pose[pa * 3 + 2] > thr && pose[pb * 3 + 2] > thr)) return false;
    const float x_p = (x_a + x_b) / 2, y_p = (y_a + y_b) / 2;
    const float angle = atan2f(y_b - y_a, x_b - x_a);
    const float sine = sinf(angle), cosine = cosf(angle);
    float a_sqrt = (x_a - x_p) * (x_a - x_p) + (y_a - y_p) * (y_a - y_p);
    if (head) { a_sqrt = (float)(a_sqrt * 1.2); b_sqrt = a_sqrt; }   // MPI limb 0 (:189-193)
    const float A = cosine * (x - x_p) + sine * (y - y_p);
    const float B = sine * (x - x_p) - cosine * (y - y_p);
    *judge_out = A * A / a_sqrt + B * B / b_sqrt;
    return true;
}

void skeleton_mpi(float* canvas, int w, int h, const float* poses, int num_people) {   // render_pose_29parts :124-240
    const int NP = 15;
    const float thr = 0.0f, radius = 3 * h / 200.0f, stick = h / 60.0f;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            Bgr px = {canvas[y * w + x], canvas[w * h + y * w + x], canvas[2 * w * h + y * w + x]};
            for (int p = 0; p < num_people; p++) {
                const float* pose = poses + p * NP * 3;
                for (int l = 0; l < 9; l++) {
                    float judge;
                    if (!limb_judge(pose, kLimbMpi[2 * l], kLimbMpi[2 * l + 1], thr, x, y, stick * stick, l == 0, &judge)) continue;
                    const float minV = l == 0 ? 0.8f : 0.f;
                    if (judge >= minV && judge <= 1) mix(px, 0.6f, (float)kColor9[l * 3 + 2], (float)kColor9[l * 3 + 1], (float)kColor9[l * 3]);
                }
                for (int i = 0; i < NP; i++) {
                    const float jx = pose[i * 3], jy = pose[i * 3 + 1];
                    if (pose[i * 3 + 2] > thr && (x - jx) * (x - jx) + (y - jy) * (y - jy) <= radius * radius) {
                        px.b = (float)(0.6 * px.b + 0.4 * kColor9[(i % 9) * 3 + 2]);   // double arithmetic (:218-220)
                        px.g = (float)(0.6 * px.g + 0.4 * kColor9[(i % 9) * 3 + 1]);
                        px.r = (float)(0.6 * px.r + 0.4 * kColor9[(i % 9) * 3]);
                    }
                }
            }
            canvas[y * w + x] = px.b; canvas[w * h + y * w + x] = px.g; canvas[2 * w * h + y * w + x] = px.r;
        }
}

void skeleton_coco(float* canvas, int w, int h, const float* poses, int num_people, bool googly) {   // :394-636
    const int NP = 18;
    const float thr = 0.01f, radius = 2 * h / 200.0f, stick = h / 120.0f;
    std::vector<float> minx(num_people), miny(num_people), maxx(num_people), maxy(num_people), scale(num_people);
    for (int p = 0; p < num_people; p++) {   // per-person box and size factor (:410-444)
        float mnx = (float)w, mny = (float)h, mxx = 0, mxy = 0;
        for (int part = 0; part < NP; part++) {
            const float jx = poses[p * NP * 3 + part * 3], jy = poses[p * NP * 3 + part * 3 + 1];
            if (poses[p * NP * 3 + part * 3 + 2] > thr) {
                if (jx < mnx) mnx = jx;
                if (jx > mxx) mxx = jx;
                if (jy < mny) mny = jy;
                if (jy > mxy) mxy = jy;
            }
        }
        float s = (float)(((mxx - mnx) + (mxy - mny)) / 2.0);
        if (s < 200) { s = s / 200; if (s < 0.33) s = (float)0.33; } else s = 1.0f;
        maxx[p] = mxx + 50; maxy[p] = mxy + 50; minx[p] = mnx - 50; miny[p] = mny - 50; scale[p] = s;
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            Bgr px = {canvas[y * w + x], canvas[w * h + y * w + x], canvas[2 * w * h + y * w + x]};
            for (int p = 0; p < num_people; p++) {
                if (x > maxx[p] || x < minx[p] || y > maxy[p] || y < miny[p]) continue;
                const float* pose = poses + p * NP * 3;
                const float sc = scale[p];
                for (int l = 0; l < 17; l++) {
                    float judge;
                    if (!limb_judge(pose, kLimbCoco[2 * l], kLimbCoco[2 * l + 1], thr, x, y, sc * sc * stick * stick, false, &judge)) continue;
                    if (judge >= 0 && judge <= 1) mix(px, 0.5f, (float)kColor18[(l % 18) * 3 + 2], (float)kColor18[(l % 18) * 3 + 1], (float)kColor18[(l % 18) * 3]);
                }
                for (int i = 0; i < NP; i++) {
                    const float jx = pose[i * 3], jy = pose[i * 3 + 1];
                    if (!(pose[i * 3 + 2] > thr)) continue;
                    const float dist2 = (x - jx) * (x - jx) + (y - jy) * (y - jy);
                    float cb = (float)kColor18[(i % 18) * 3 + 2], cg = (float)kColor18[(i % 18) * 3 + 1], cr = (float)kColor18[(i % 18) * 3];
                    if (googly && (i == 14 || i == 15)) {   // :589-611
                        const float maxr2 = (float)(sc * sc * 2.5 * 2.5 * radius * radius);
                        const float minr2 = (float)(sc * sc * (2.5 * radius - 2) * (2.5 * radius - 2));
                        cb = cg = cr = 0;
                        if (dist2 <= maxr2) {
                            if (dist2 <= minr2) cb = cg = cr = 255;
                            if (dist2 <= minr2 * 0.6) {
                                const float dist3 = (x - 4 - jx) * (x - 4 - jx) + (y - jy + 4) * (y - jy + 4);
                                if (dist3 > 3.75 * 3.75) cb = cg = cr = 0;
                            }
                            mix(px, 0.9f, cb, cg, cr);
                        }
                    } else {
                        const float maxr2 = sc * sc * radius * radius;
                        if (dist2 >= 0 && dist2 <= maxr2) mix(px, 0.6f, cb, cg, cr);
                    }
                }
            }
            canvas[y * w + x] = px.b; canvas[w * h + y * w + x] = px.g; canvas[2 * w * h + y * w + x] = px.r;
        }
}

void jet(float* c, float v, float vmin, float vmax) {   // getColor :11-44; c = {b, g, r}
    c[0] = c[1] = c[2] = 255;
    if (v < vmin) v = vmin;
    if (v > vmax) v = vmax;
    const float dv = vmax - vmin;
    if (v < (vmin + 0.125 * dv)) { c[0] = (float)(256 * (0.5 + (v * 4))); c[1] = c[2] = 0; }
    else if (v < (vmin + 0.375 * dv)) { c[0] = 255; c[1] = (float)(256 * (v - 0.125) * 4); c[2] = 0; }
    else if (v < (vmin + 0.625 * dv)) { c[0] = (float)(256 * (-4 * v + 2.5)); c[1] = 255; c[2] = (float)(256 * (4 * (v - 0.375))); }
    else if (v < (vmin + 0.875 * dv)) { c[0] = 0; c[1] = (float)(256 * (-4 * v + 3.5)); c[2] = 255; }
    else { c[0] = 0; c[1] = 0; c[2] = (float)(256 * (-4 * v + 4.5)); }
}

void wheel(float* c, float v) {   // getColor2 :46-94 with vmin 0, vmax 1
    if (v < 0) v = 0;
    if (v > 1) v = 1;
    v = 55 * v;
    const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
    if (v < RY) { c[0] = 255; c[1] = 255 * (v / RY); c[2] = 0; }
    else if (v < RY + YG) { c[0] = 255 - 255 * ((v - RY) / YG); c[1] = 255; c[2] = 0; }
    else if (v < RY + YG + GC) { c[0] = 0; c[1] = 255; c[2] = 255 * ((v - RY - YG) / GC); }
    else if (v < RY + YG + GC + CB) { c[0] = 0; c[1] = 255 - 255 * ((v - RY - YG - GC) / CB); c[2] = 255; }
    else if (v < RY + YG + GC + CB + BM) { c[0] = 255 * ((v - RY - YG - GC - CB) / BM); c[1] = 0; c[2] = 255; }
    else if (v < RY + YG + GC + CB + BM + MR) { c[0] = 255; c[1] = 0; c[2] = 255 - 255 * ((v - RY - YG - GC - CB - BM) / MR); }
    else { c[0] = 255; c[1] = 0; c[2] = 0; }
}

void vec_color(float* c, float x, float y) {   // getColorXY :96-112
    float rad = sqrtf(x * x + y * y);
    const float a = (float)(atan2f(-y, -x) / M_PI);
    float fk = (float)((a + 1) / 2.0);
    if (std::isnan(fk)) fk = 0;
    if (rad > 1) rad = 1;
    wheel(c, fk);
    for (int k = 0; k < 3; k++) c[k] = 255 * (rad * (c[k] / 255));
}

float cubic_render(float v0, float v1, float v2, float v3, float dx) {   // :114-122, float/double mix as written there
    return (float)((-0.5f * v0 + 1.5f * v1 - 1.5f * v2 + 0.5f * v3) * dx * dx * dx + (v0 - 2.5f * v1 + 2.0 * v2 - 0.5 * v3) * dx * dx +
                   (-0.5f * v0 + 0.5f * v2) * dx + v1);
}

struct HeatTaps { int xn[4], yn[4]; float dx, dy; bool inside; };
HeatTaps heat_taps(int x, int y, int w_canvas, int h_canvas, int w_net, int h_net) {   // :263-285 (identical in all four views)
    HeatTaps t;
    const float h_inv = (float)h_net / (float)h_canvas, w_inv = (float)w_net / (float)w_canvas;
    const float xb = (float)(w_inv * x + (0.5 * w_inv - 0.5)), yb = (float)(h_inv * y + (0.5 * h_inv - 0.5));
    t.inside = xb >= 0 && xb < w_net && yb >= 0 && yb < h_net;
    auto axis = [](float on, int n, int* nei, float* d) {
        nei[1] = int(on + 1e-5);
        nei[1] = nei[1] < 0 ? 0 : nei[1];
        nei[0] = nei[1] - 1 < 0 ? nei[1] : nei[1] - 1;
        nei[2] = nei[1] + 1 >= n ? n - 1 : nei[1] + 1;
        nei[3] = nei[2] + 1 >= n ? n - 1 : nei[2] + 1;
        *d = on - nei[1];
    };
    axis(xb, w_net, t.xn, &t.dx);
    axis(yb, h_net, t.yn, &t.dy);
    return t;
}

// mode 0: render_pose_29parts_heatmap :242-329; 1: ..._coco_heatmap :638-724; 2: ..._coco_heatmap2 :726-836 (in_part = 0);
// 3: ..._coco_affinity :838-975
void heat_view(float* canvas, int w, int h, int w_net, int h_net, const float* heat, int mode, int part, int accum) {
    const int off2 = w_net * h_net;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            Bgr px = {canvas[y * w + x], canvas[w * h + y * w + x], canvas[2 * w * h + y * w + x]};
            const HeatTaps t = heat_taps(x, y, w, h, w_net, h_net);
            if (mode == 0 || mode == 1) {
                const int NP = mode == 0 ? 15 : 18;
                float value = (part == NP - 1) ? 1.f : 0.f;
                if (t.inside) {
                    const float* src = heat + (size_t)part * off2;
                    float tmp[4];
                    for (int i = 0; i < 4; i++)
                        tmp[i] = cubic_render(src[t.yn[i] * w_net + t.xn[0]], src[t.yn[i] * w_net + t.xn[1]], src[t.yn[i] * w_net + t.xn[2]],
                                              src[t.yn[i] * w_net + t.xn[3]], t.dx);
                    value = cubic_render(tmp[0], tmp[1], tmp[2], tmp[3], t.dy);
                }
                float c[3];
                if (mode == 0) {
                    if (part < 16) jet(c, value, 0, 1); else jet(c, value, -1, 1);
                    px.b = (float)(0.5 * px.b + 0.5 * c[0]); px.g = (float)(0.5 * px.g + 0.5 * c[1]); px.r = (float)(0.5 * px.r + 0.5 * c[2]);
                } else {
                    if (part < NP + 1) jet(c, value, 0, 1); else jet(c, value, -1, 1);
                    mix(px, 0.7f, c[2], c[1], c[0]);
                }
            } else if (mode == 2) {
                float c[3] = {0, 0, 0};
                if (t.inside)
                    for (int p2 = 0; p2 < 18; p2++) {
                        const float value = heat[(size_t)p2 * off2 + t.yn[1] * w_net + t.xn[1]];
                        for (int k = 0; k < 3; k++) c[k] += value * kColor18[(p2 % 18) * 3 + k];
                    }
                mix(px, 0.7f, c[2], c[1], c[0]);
            } else {
                float c[3] = {0, 0, 0};
                if (t.inside)
                    for (int p2 = part; p2 < part + accum * 2; p2 += 2) {
                        const float* h0 = heat + (size_t)p2 * off2;
                        const float* h1 = heat + (size_t)(p2 + 1) * off2;
                        float value, value2;
                        if (accum == 1) {   // bilinear (:912-935)
                            auto bil = [&](const float* m) {
                                const float a = m[t.yn[1] * w_net + t.xn[1]], b = m[t.yn[1] * w_net + t.xn[2]];
                                const float cc = m[t.yn[2] * w_net + t.xn[1]], d = m[t.yn[2] * w_net + t.xn[2]];
                                return (1 - t.dx) * (1 - t.dy) * a + (t.dx) * (1 - t.dy) * b + (1 - t.dx) * (t.dy) * cc + (t.dx) * (t.dy) * d;
                            };
                            value = bil(h0); value2 = bil(h1);
                        } else {
                            value = h0[t.yn[1] * w_net + t.xn[1]];
                            value2 = h1[t.yn[1] * w_net + t.xn[1]];
                        }
                        float c2[3];
                        vec_color(c2, value, value2);
                        for (int k = 0; k < 3; k++) c[k] += c2[k];
                    }
                for (int k = 0; k < 3; k++) if (c[k] > 255) c[k] = 255;
                mix(px, 0.7f, c[2], c[1], c[0]);
            }
            canvas[y * w + x] = px.b; canvas[w * h + y * w + x] = px.g; canvas[2 * w * h + y * w + x] = px.r;
        }
}

}  // namespace

// process_and_pad_image(..., normalize = 0) for an image of the canvas size (rtpose.cpp:239-269, :349)
extern "C" void orc_canvas_from_u8(const uint8_t* bgr, int h, int w, float* canvas) {
    for (int c = 0; c < 3; c++)
        for (int i = 0; i < h * w; i++) canvas[(size_t)c * h * w + i] = float(bgr[(size_t)i * 3 + c]);
}
// postProcessFrame (rtpose.cpp:1286-1296)
extern "C" void orc_canvas_to_u8(const float* canvas, int h, int w, uint8_t* bgr) {
    for (int c = 0; c < 3; c++)
        for (int i = 0; i < h * w; i++) {
            int value = int(canvas[(size_t)c * h * w + i] + 0.5);
            value = value < 0 ? 0 : (value > 255 ? 255 : value);
            bgr[(size_t)i * 3 + c] = (unsigned char)value;
        }
}
// render() (rtpose.cpp:271-300): the launcher a --part_to_show value selects and the two arguments that depend on it.
// out3 = {launcher (0 render_mpi_parts, 1 render_coco_parts, 2 render_coco_aff), its `part` argument, googly_eyes / num_parts_accum}.
// Pinned against the reference's own text compiled with recording launchers (oracle/_ref, tests/test_oracle.py).
extern "C" int orc_render_dispatch(int model, int part_to_show, int googly_eyes, int* out3) {
    const int NP = model_desc(model).num_parts;
    if (NP == 15) { out3[0] = 0; out3[1] = part_to_show; out3[2] = 0; return 0; }
    if (part_to_show - 1 <= NP) { out3[0] = 1; out3[1] = part_to_show; out3[2] = googly_eyes != 0; return 0; }
    int aff_part = ((part_to_show - 1) - NP - 1) * 2, accum = 1;
    if (aff_part == 0) accum = 19; else aff_part -= 2;
    aff_part += 1 + NP;
    out3[0] = 2; out3[1] = aff_part; out3[2] = accum;
    return 0;
}
// render() + the launchers (renderFunctions.cu:331-389, 978-1080).  heatmaps: the full-resolution
// resized_map (num_maps x h_net x w_net), only read when part_to_show > 0.  Returns 0, or -1 for a bad part_to_show.
extern "C" int orc_render(int model, float* canvas, int w_canvas, int h_canvas, int w_net, int h_net, const float* heatmaps,
                          const float* poses, int num_people, int part_to_show, int googly_eyes) {
    const ModelDesc& md = model_desc(model);
    const int NP = md.num_parts, num_maps = NP + 1 + 2 * md.num_limbs;
    if (part_to_show < 0) return -1;
    int d[3];
    orc_render_dispatch(model, part_to_show, googly_eyes, d);
    const int part = d[1];
    if (d[0] == 0) {          // render_mpi_parts (renderFunctions.cu:331-389)
        if (part == 0) { if (num_people != 0) skeleton_mpi(canvas, w_canvas, h_canvas, poses, num_people); }
        else { if (part - 1 >= num_maps) return -1; heat_view(canvas, w_canvas, h_canvas, w_net, h_net, heatmaps, 0, part - 1, 0); }
    } else if (d[0] == 1) {   // render_coco_parts (:978-1036)
        if (part == 0) { if (num_people != 0) skeleton_coco(canvas, w_canvas, h_canvas, poses, num_people, d[2] != 0); }
        else if (part - 1 == NP) heat_view(canvas, w_canvas, h_canvas, w_net, h_net, heatmaps, 2, 0, 0);
        else heat_view(canvas, w_canvas, h_canvas, w_net, h_net, heatmaps, 1, part - 1, 0);
    } else {                  // render_coco_aff (:1038-1080)
        if (part + 2 * d[2] > num_maps) return -1;
        heat_view(canvas, w_canvas, h_canvas, w_net, h_net, heatmaps, 3, part, d[2]);
    }
    return 0;
}

extern "C" int orc_json(const float* joints, int num_people, int num_parts, double frame_scale, char* buf, int cap) {
    std::string s;
    char t[64];
    const double scale = 1.0 / frame_scale;
    s += "{\n\"version\":0.1,\n\"bodies\":[\n";
    for (int ip = 0; ip < num_people; ip++) {
        s += "{\n\"joints\":[";
        for (int ij = 0; ij < num_parts; ij++) {
            const float* j = joints + (size_t)ip * num_parts * 3 + ij * 3;
            snprintf(t, sizeof t, "%g,", scale * j[0]); s += t;
            snprintf(t, sizeof t, "%g,", scale * j[1]); s += t;
            snprintf(t, sizeof t, "%g", (double)j[2]); s += t;
            if (ij < num_parts - 1) s += ",";
        }
        s += "]\n}";
        if (ip < num_people - 1) s += ",\n";
    }
    s += "]\n}\n";
    if ((int)s.size() < cap) memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// ------------------------------------------------------------------------------------------------
// Whole frame (processFrame rtpose.cpp:1099-1166, minus rendering)
// ------------------------------------------------------------------------------------------------
extern "C" int orc_process_frame(OrcNet* net, int model, const uint8_t* disp, int disp_h, int disp_w, int net_h,
                                 int net_w, int num_scales, double start_scale, double scale_gap, float nms_threshold,
                                 const OrcConnectParams* p, float* joints, float* peaks_out, float* maps8_out) {
    const int C = orc_model_num_maps(model), P = orc_model_num_parts(model);
    const int max_peaks = model == ORC_MODEL_MPI_15 ? 20 : 64;
    const int h8 = net_h / 8, w8 = net_w / 8;
    std::vector<float> input((size_t)num_scales * 3 * net_h * net_w);
    if (orc_preprocess(disp, disp_h, disp_w, net_h, net_w, num_scales, start_scale, scale_gap, input.data())) return -1;
    std::vector<float> maps8((size_t)num_scales * C * h8 * w8);
    if (orc_net_forward(net, input.data(), num_scales, net_h, net_w, maps8.data())) return -2;
    if (maps8_out) memcpy(maps8_out, maps8.data(), maps8.size() * sizeof(float));
    std::vector<float> full((size_t)C * net_h * net_w);
    orc_imresize(maps8.data(), num_scales, C, h8, w8, net_h, net_w, (float)start_scale, (float)scale_gap, full.data());
    std::vector<float> peaks((size_t)P * (max_peaks + 1) * 3);
    orc_nms(full.data(), C, net_h, net_w, P, max_peaks, nms_threshold, peaks.data());
    if (peaks_out) memcpy(peaks_out, peaks.data(), peaks.size() * sizeof(float));
    return orc_connect(model, full.data(), peaks.data(), max_peaks, net_w, net_h, disp_w, disp_h, p, joints, nullptr, 0,
                       nullptr);
}
