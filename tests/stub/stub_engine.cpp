// Test double for the GPU-touching entry points of include/poseengine.h, linked IN FRONT of libposeengine.so into a
// ThreadSanitizer build of host/rtpose.cpp (tests/test_host_pipeline.py): the host pipeline of rtpose.bin - producers, worker
// threads, re-orderer / writer, frame-drop policy, run-time keys, error exits (examples/rtpose/rtpose.cpp:1459-1549 topology) - runs
// on a machine without a GPU, under a race detector.  Everything that is host code in the library (codecs, AVI reader, JSON writer,
// prototxt reader, model tables) still comes from the real libposeengine.so.  TEST INFRASTRUCTURE: nothing here is shipped.
//
// The "forward" encodes each frame's first pixel and size into the joints, so a result that reaches the wrong JSON file, a frame
// read after its buffer went back to the pool, or two frames swapped by the re-orderer shows up in the files.
//   STUB_FORWARD_MS  sleep per pe_forward_* call (the "GPU time")
//   STUB_FAIL_AT     the k-th forward of the process (1-based) fails with PE_ERR_CUDA
//   STUB_LOG         file that receives one line per create / forward / parameter change
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "poseengine.h"

namespace {
std::mutex g_log_mutex;
std::atomic<int> g_forwards{0}, g_engines{0};
std::string g_create_error;

void logf(const char* fmt, ...) {
    const char* path = getenv("STUB_LOG");
    if (!path) return;
    std::lock_guard<std::mutex> l(g_log_mutex);
    FILE* f = fopen(path, "a");
    if (!f) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(f, fmt, ap);
    va_end(ap);
    fputc('\n', f);
    fclose(f);
}
int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
}  // namespace

struct pe_engine {
    pe_config cfg;
    int id = 0, parts = 18;
    float nms_threshold = 0.05f, start_scale = 1.f, scale_gap = 0.3f;
    int min_subset_cnt = 3, inter_min_above = 9;
    float min_subset_score = 0.4f, inter_threshold = 0.05f;
    bool weights = false, committed = false;
    std::string err;
    std::atomic<int> busy{0};   // "calls on one handle must not be concurrent" (poseengine.h): checked, not assumed
    struct Result { int people; std::vector<float> joints; std::vector<uint8_t> frame; int w, h; };
    std::vector<Result> results;
};

namespace {
struct Exclusive {
    pe_engine* e;
    explicit Exclusive(pe_engine* e_) : e(e_) {
        if (e->busy.fetch_add(1) != 0) { fprintf(stderr, "STUB: concurrent calls on engine %d\n", e->id); abort(); }
    }
    ~Exclusive() { e->busy.fetch_sub(1); }
};
int fail(pe_engine* e, int code, const char* msg) { (e ? e->err : g_create_error) = msg; return code; }

int forward(pe_engine* e, const uint8_t* const* frames, int n, int w, int h, const char* how) {
    Exclusive x(e);
    if (!e->committed) return fail(e, PE_ERR_STATE, "pe_commit_weights has not been called");
    if (n < 1 || n > e->cfg.max_batch) return fail(e, PE_ERR_INVALID, "n outside 1..max_batch");
    const int k = ++g_forwards;
    if (k == env_int("STUB_FAIL_AT", -1)) return fail(e, PE_ERR_CUDA, "stub: injected device failure");
    const int ms = env_int("STUB_FORWARD_MS", 0);
    if (ms > 0) std::this_thread::sleep_for(std::chrono::milliseconds(ms));
    e->results.clear();
    for (int i = 0; i < n; i++) {
        const uint8_t* p = frames[i];
        unsigned sum = 0;
        for (size_t j = 0; j < (size_t)w * h * 3; j++) sum += p[j];   // reads the whole buffer: a recycled / freed frame is a sanitizer finding
        pe_engine::Result r;
        r.people = 1 + p[0] % 3;
        r.w = w; r.h = h;
        r.joints.assign((size_t)r.people * e->parts * 3, 0.f);
        for (int q = 0; q < r.people; q++)
            for (int j = 0; j < e->parts; j++) {
                float* d = &r.joints[((size_t)q * e->parts + j) * 3];
                d[0] = (float)(p[0] + 256 * p[1]);   // x: the frame's identity (first two bytes)
                d[1] = (float)(q * 100 + j);         // y: person and part
                d[2] = (float)((sum % 1000) / 1000.0 + 0.0005);
            }
        if (getenv("STUB_KEEP_FRAMES")) r.frame.assign(p, p + (size_t)w * h * 3);
        e->results.push_back(std::move(r));
    }
    logf("forward engine=%d call=%d n=%d %s size=%dx%d nms=%.4f connect=%d,%.4f,%.4f,%d exact=%.9g,%.9g,%.9g", e->id, k, n, how, w, h, e->nms_threshold,
         e->min_subset_cnt, e->min_subset_score, e->inter_threshold, e->inter_min_above, e->nms_threshold, e->min_subset_score, e->inter_threshold);
    return PE_OK;
}
}  // namespace

extern "C" {
int pe_create(const pe_config* cfg, pe_engine** out) {
    if (!cfg || !out) return fail(nullptr, PE_ERR_INVALID, "null argument");
    if (cfg->device >= env_int("STUB_NUM_DEVICES", 8)) return fail(nullptr, PE_ERR_CUDA, "stub: invalid device ordinal");
    pe_engine* e = new pe_engine;
    e->cfg = *cfg;
    e->id = g_engines++;
    e->parts = cfg->model == PE_MODEL_MPI_15 ? 15 : 18;
    if (cfg->model == PE_MODEL_MPI_15) { e->nms_threshold = 0.2f; e->inter_threshold = 0.01f; e->inter_min_above = 8; }
    logf("create engine=%d device=%d model=%d net=%dx%d disp=%dx%d scales=%d batch=%d precision=%d", e->id, cfg->device, cfg->model, cfg->net_w,
         cfg->net_h, cfg->disp_w, cfg->disp_h, cfg->num_scales, cfg->max_batch, cfg->precision);
    *out = e;
    return PE_OK;
}
int pe_create_from_prototxt(const pe_config* cfg, const char* path, pe_engine** out) {
    logf("prototxt %s", path);
    return pe_create(cfg, out);
}
void pe_destroy(pe_engine* e) { if (e) { logf("destroy engine=%d", e->id); delete e; } }
const char* pe_last_error(const pe_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int pe_num_conv_layers(const pe_engine*) { return 2; }
int pe_conv_layer_info(const pe_engine*, int idx, char* name64, int* cout, int* cin, int* ksize) {
    if (idx < 0 || idx > 1) return PE_ERR_INVALID;
    snprintf(name64, 64, "conv%d", idx);
    *cout = 4; *cin = 3; *ksize = 3;
    return PE_OK;
}
int pe_set_conv_weights(pe_engine* e, const char*, const float*, size_t nw, const float*, size_t nb) {
    if (nw != 4 * 3 * 3 * 3 || nb != 4) return fail(e, PE_ERR_INVALID, "stub: weight count");
    e->weights = true;
    return PE_OK;
}
int pe_commit_weights(pe_engine* e) {
    if (!e->weights) return fail(e, PE_ERR_STATE, "weights were never set");
    e->committed = true;
    return PE_OK;
}
int pe_calibrate(pe_engine* e, const uint8_t* const* frames, int n) {
    logf("calibrate engine=%d", e->id);
    return forward(e, frames, n, e->cfg.disp_w, e->cfg.disp_h, "calibrate");
}
int pe_share_weights(pe_engine* from, pe_engine* to) {
    if (!from->committed) return fail(from, PE_ERR_STATE, "the source handle has no committed weights");
    if (from->cfg.device != to->cfg.device) return fail(from, PE_ERR_INVALID, "handles are on different devices");
    to->weights = to->committed = true;
    logf("share %d -> %d", from->id, to->id);
    return PE_OK;
}
int pe_broadcast_weights(pe_engine* const* engines, int n) {
    if (getenv("STUB_NO_NCCL")) return fail(engines[0], PE_ERR_STATE, "NCCL is not available (stub)");
    if (!engines[0]->committed) return fail(engines[0], PE_ERR_STATE, "engines[0] has no committed weights to broadcast");
    for (int i = 1; i < n; i++) engines[i]->weights = engines[i]->committed = true;
    logf("broadcast n=%d", n);
    return PE_OK;
}
size_t pe_packed_weights_bytes(const pe_engine*) { return 1000000; }

int pe_nms_get_max_peaks(const pe_engine*) { return 64; }
int pe_nms_get_num_parts(const pe_engine* e) { return e->parts; }
float pe_nms_get_threshold(const pe_engine* e) { return e->nms_threshold; }
int pe_nms_set_threshold(pe_engine* e, float t) { Exclusive x(e); e->nms_threshold = t; return PE_OK; }
int pe_resize_set_start_scale(pe_engine* e, float s) { e->start_scale = s; return PE_OK; }
int pe_resize_set_scale_gap(pe_engine* e, float s) { e->scale_gap = s; return PE_OK; }
float pe_resize_get_start_scale(const pe_engine* e) { return e->start_scale; }
float pe_resize_get_scale_gap(const pe_engine* e) { return e->scale_gap; }
int pe_set_connect_params(pe_engine* e, int a, float b, float c, int d) {
    Exclusive x(e);
    e->min_subset_cnt = a; e->min_subset_score = b; e->inter_threshold = c; e->inter_min_above = d;
    return PE_OK;
}

int pe_forward_frames(pe_engine* e, const uint8_t* const* frames, int n) { return forward(e, frames, n, e->cfg.disp_w, e->cfg.disp_h, "display"); }
int pe_forward_camera_frames(pe_engine* e, const uint8_t* const* frames, int n, int orig_w, int orig_h, double* scale) {
    if (scale) *scale = std::min((double)e->cfg.disp_w / orig_w, (double)e->cfg.disp_h / orig_h);   // rtpose.cpp:474-480
    return forward(e, frames, n, orig_w, orig_h, "camera");
}
int pe_fetch(pe_engine* e, int idx, float* joints, int* num_people, float*) {
    Exclusive x(e);
    if (idx < 0 || idx >= (int)e->results.size()) return fail(e, PE_ERR_INVALID, "idx outside the last forward");
    const auto& r = e->results[idx];
    *num_people = r.people;
    memcpy(joints, r.joints.data(), r.joints.size() * sizeof(float));
    return PE_OK;
}
int pe_render(pe_engine* e, int idx, int part_to_show, int googly_eyes, const uint8_t*, float*, uint8_t* bgr) {
    Exclusive x(e);
    if (idx < 0 || idx >= (int)e->results.size()) return fail(e, PE_ERR_INVALID, "idx outside the last forward");
    const auto& r = e->results[idx];
    const size_t bytes = (size_t)e->cfg.disp_w * e->cfg.disp_h * 3;
    if (r.frame.size() == bytes) memcpy(bgr, r.frame.data(), bytes); else memset(bgr, 40, bytes);
    bgr[0] = (uint8_t)part_to_show; bgr[1] = (uint8_t)googly_eyes;
    logf("render engine=%d idx=%d part=%d googly=%d", e->id, idx, part_to_show, googly_eyes);
    return PE_OK;
}
void* pe_host_alloc(size_t bytes) { return getenv("STUB_NO_PINNED") ? nullptr : malloc(bytes); }
void pe_host_free(void* p) { free(p); }
}
