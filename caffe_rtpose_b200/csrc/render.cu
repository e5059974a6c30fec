// Renderers: skeleton / heat-map / PAF overlays on the display frame.
//
// What it replaces:
//   render()                      examples/rtpose/rtpose.cpp:271-300   (dispatch on model and --part_to_show)
//   render_mpi_parts / render_coco_parts / render_coco_aff and their six kernels
//                                 src/rtpose/renderFunctions.cu:124-975
//   canvas upload (float planar BGR copy of the display frame)   rtpose.cpp:239-269 (normalize = 0), :1127
//   float canvas -> uint8 BGR frame (postProcessFrame)           rtpose.cpp:1286-1296
//
// B200-first differences.  The reference uploads a 11 MB float canvas per frame, renders, downloads 11 MB and converts
// to uint8 on one host thread; its heat-map views read the 55 MB full-resolution map.  Here the display frame is
// already on the device (pe_forward_frames), the joints never left it, the canvas lives in HBM, only the uint8 image
// (2.8 MB) returns, and the heat-map views evaluate the few channels they show from the stride-8 maps (fullres.cuh)
// into a scratch - the full-resolution map still does not exist.  The reference launches its kernels with grid and
// block swapped (`<<<threadsPerBlock, numBlocks>>>`, renderFunctions.cu:362,1009): block = (w/32, h/32) threads, so
// displays above ~1024*32*32/... pixels (e.g. 1920x1080 -> 2040 threads) fail to launch and nothing is drawn, and on
// tiny canvases people beyond the block size are read from uninitialised shared memory.  This implementation uses a
// regular launch; results are identical wherever the reference's launch is valid.
//
// Arithmetic: per-pixel expressions follow the reference statement by statement (float vs double promotions
// included) so that nvcc contracts them the same way; tests/test_gpu_render.py compares against the reference's own
// kernels compiled for sm_100a by the test infrastructure.
#include "common.h"
#include "kernels.h"
#include "fullres.cuh"

namespace pe {

// ------------------------------------------------------------------------------------------------
// canvas <-> uint8
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) canvas_fill_kernel(const uint8_t* __restrict__ bgr, float* __restrict__ canvas, int w, int h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const uint8_t* p = bgr + (size_t)i * 3;
    canvas[i] = (float)p[0];
    canvas[(size_t)w * h + i] = (float)p[1];
    canvas[(size_t)2 * w * h + i] = (float)p[2];
}

__global__ void __launch_bounds__(256) canvas_to_u8_kernel(const float* __restrict__ canvas, uint8_t* __restrict__ bgr, int w, int h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int v = __double2int_rz(__dadd_rn((double)canvas[(size_t)c * w * h + i], 0.5));   // int(value + 0.5), rtpose.cpp:1291
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        bgr[(size_t)i * 3 + c] = (uint8_t)v;
    }
}

// ------------------------------------------------------------------------------------------------
// materialise `nch` channels (ch0, ch0+step, ...) of the reference's resized_map for one frame
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fullres_fill_kernel(PostDev pd, int frame, int ch0, int nch, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int c = blockIdx.z;
    if (x >= pd.p.net_w || c >= nch) return;
    const FullRes fr = make_fullres(pd, frame);
    out[((size_t)c * pd.p.net_h + y) * pd.p.net_w + x] = fullres_at(fr, ch0 + c, y, x);
}

// ------------------------------------------------------------------------------------------------
// skeleton overlays
// ------------------------------------------------------------------------------------------------
__constant__ int c_limb_coco[34] = {1, 2, 1, 5, 2, 3, 3, 4, 5, 6, 6, 7, 1, 8, 8, 9, 9, 10, 1, 11, 11, 12, 12, 13, 1, 0, 0, 14, 14, 16, 0, 15, 15, 17};
__constant__ int c_color18[54] = {255, 0, 0, 255, 85, 0, 255, 170, 0, 255, 255, 0, 170, 255, 0, 85, 255, 0, 0, 255, 0, 0, 255, 85, 0, 255, 170,
                                  0, 255, 255, 0, 170, 255, 0, 85, 255, 0, 0, 255, 85, 0, 255, 170, 0, 255, 255, 0, 255, 255, 0, 170, 255, 0, 85};

struct SkelArgs {
    float* canvas; int w, h;
    const float* poses;        // [PE_MAX_PEOPLE][parts][3] of this frame (device)
    const int* num_people;     // device
    int googly;
};

// render_pose_29parts (renderFunctions.cu:124-240)
__global__ void __launch_bounds__(256) skeleton_mpi_kernel(SkelArgs a) {
    constexpr int NP = 15;
    __shared__ float sp[NP * 3 * PE_MAX_PEOPLE];
    const int np = min(*a.num_people, PE_MAX_PEOPLE);
    if (np <= 0) return;   // the reference does not launch (renderFunctions.cu:358)
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < np * NP * 3; i += blockDim.x * blockDim.y) sp[i] = a.poses[i];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int w_canvas = a.w, h_canvas = a.h;
    if (x >= w_canvas || y >= h_canvas) return;
    const float threshold = 0.0f;
    const int limb[] = {0, 1, 2, 3, 3, 4, 5, 6, 6, 7, 8, 9, 9, 10, 11, 12, 12, 13};   // head-neck, arms, legs (LIMB_MPI)
    const int nlimb = sizeof(limb) / (2 * sizeof(int));
    int color[27] = {255, 0, 0, 255, 170, 0, 170, 255, 0, 0, 255, 0, 0, 255, 170, 0, 170, 255, 0, 0, 255, 170, 0, 255, 255, 0, 170};
    float radius = 3 * h_canvas / 200.0f;
    float stickwidth = h_canvas / 60.0f;
    float b = a.canvas[y * w_canvas + x];
    float g = a.canvas[w_canvas * h_canvas + y * w_canvas + x];
    float r = a.canvas[2 * w_canvas * h_canvas + y * w_canvas + x];
    for (int p = 0; p < np; p++) {
        const float* pose = sp + p * NP * 3;
        for (int l = 0; l < nlimb; l++) {
            float b_sqrt = stickwidth * stickwidth;
            float alpha = 0.6;
            const int pa = limb[2 * l], pb = limb[2 * l + 1];
            const float x_a = pose[pa * 3], x_b = pose[pb * 3], y_a = pose[pa * 3 + 1], y_b = pose[pb * 3 + 1];
            if (pose[pa * 3 + 2] > threshold && pose[pb * 3 + 2] > threshold) {
                float x_p = (x_a + x_b) / 2;
                float y_p = (y_a + y_b) / 2;
                float angle = atan2f(y_b - y_a, x_b - x_a);
                float sine = sinf(angle);
                float cosine = cosf(angle);
                float a_sqrt = (x_a - x_p) * (x_a - x_p) + (y_a - y_p) * (y_a - y_p);
                if (l == 0) {
                    a_sqrt *= 1.2;
                    b_sqrt = a_sqrt;
                }
                float A = cosine * (x - x_p) + sine * (y - y_p);
                float B = sine * (x - x_p) - cosine * (y - y_p);
                float judge = A * A / a_sqrt + B * B / b_sqrt;
                float minV = 0;
                if (l == 0) minV = 0.8;
                if (judge >= minV && judge <= 1) {
                    b = (1 - alpha) * b + alpha * color[l * 3 + 2];
                    g = (1 - alpha) * g + alpha * color[l * 3 + 1];
                    r = (1 - alpha) * r + alpha * color[l * 3];
                }
            }
        }
        for (int i = 0; i < NP; i++) {
            const float px = pose[i * 3], py = pose[i * 3 + 1];
            if (pose[i * 3 + 2] > threshold) {
                if ((x - px) * (x - px) + (y - py) * (y - py) <= radius * radius) {
                    b = 0.6 * b + 0.4 * color[(i % 9) * 3 + 2];
                    g = 0.6 * g + 0.4 * color[(i % 9) * 3 + 1];
                    r = 0.6 * r + 0.4 * color[(i % 9) * 3];
                }
            }
        }
    }
    a.canvas[y * w_canvas + x] = b;
    a.canvas[w_canvas * h_canvas + y * w_canvas + x] = g;
    a.canvas[2 * w_canvas * h_canvas + y * w_canvas + x] = r;
}

// render_pose_coco_parts (renderFunctions.cu:394-636); the `if (0 && ...)` branches of the reference are dead code
__global__ void __launch_bounds__(256) skeleton_coco_kernel(SkelArgs a) {
    constexpr int NP = 18;
    __shared__ float sp[NP * 3 * PE_MAX_PEOPLE];
    __shared__ float2 s_min[PE_MAX_PEOPLE], s_max[PE_MAX_PEOPLE];
    __shared__ float s_scale[PE_MAX_PEOPLE];
    const int np = min(*a.num_people, PE_MAX_PEOPLE);
    if (np <= 0) return;   // the reference does not launch (renderFunctions.cu:1006)
    const int w_canvas = a.w, h_canvas = a.h;
    const float threshold = 0.01f;
    for (int p = threadIdx.y * blockDim.x + threadIdx.x; p < np; p += blockDim.x * blockDim.y) {   // per-person box and size
        float2 mn = make_float2((float)w_canvas, (float)h_canvas), mx = make_float2(0.f, 0.f);
        for (int part = 0; part < NP; part++) {
            const float px = a.poses[p * NP * 3 + part * 3], py = a.poses[p * NP * 3 + part * 3 + 1], pz = a.poses[p * NP * 3 + part * 3 + 2];
            sp[p * NP * 3 + part * 3] = px; sp[p * NP * 3 + part * 3 + 1] = py; sp[p * NP * 3 + part * 3 + 2] = pz;
            if (pz > threshold) {
                if (px < mn.x) mn.x = px;
                if (px > mx.x) mx.x = px;
                if (py < mn.y) mn.y = py;
                if (py > mx.y) mx.y = py;
            }
        }
        float sx = mx.x - mn.x, sy = mx.y - mn.y;
        sx = (sx + sy) / 2.0;
        if (sx < 200) {
            sx = sx / 200;
            if (sx < 0.33) sx = 0.33;
        } else {
            sx = 1.0;
        }
        mx.x += 50; mx.y += 50; mn.x -= 50; mn.y -= 50;
        s_min[p] = mn; s_max[p] = mx; s_scale[p] = sx;
    }
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w_canvas || y >= h_canvas) return;
    float radius = 2 * h_canvas / 200.0f;
    float stickwidth = h_canvas / 120.0f;
    float b = a.canvas[y * w_canvas + x];
    float g = a.canvas[w_canvas * h_canvas + y * w_canvas + x];
    float r = a.canvas[2 * w_canvas * h_canvas + y * w_canvas + x];
    const bool googly_eyes = a.googly != 0;
    for (int p = 0; p < np; p++) {
        if (x > s_max[p].x || x < s_min[p].x || y > s_max[p].y || y < s_min[p].y) continue;
        const float* pose = sp + p * NP * 3;
        const float sc = s_scale[p];
        for (int l = 0; l < 17; l++) {
            float b_sqrt = sc * sc * stickwidth * stickwidth;
            float alpha = 0.5;
            const int pa = c_limb_coco[2 * l], pb = c_limb_coco[2 * l + 1];
            const float x_a = pose[pa * 3], x_b = pose[pb * 3], y_a = pose[pa * 3 + 1], y_b = pose[pb * 3 + 1];
            if (pose[pa * 3 + 2] > threshold && pose[pb * 3 + 2] > threshold) {
                float x_p = (x_a + x_b) / 2;
                float y_p = (y_a + y_b) / 2;
                float angle = atan2f(y_b - y_a, x_b - x_a);
                float sine = sinf(angle);
                float cosine = cosf(angle);
                float a_sqrt = (x_a - x_p) * (x_a - x_p) + (y_a - y_p) * (y_a - y_p);
                float A = cosine * (x - x_p) + sine * (y - y_p);
                float B = sine * (x - x_p) - cosine * (y - y_p);
                float judge = A * A / a_sqrt + B * B / b_sqrt;
                float minV = 0;
                float maxV = 1;
                float3 co;
                co.x = c_color18[(l % 18) * 3 + 0];
                co.y = c_color18[(l % 18) * 3 + 1];
                co.z = c_color18[(l % 18) * 3 + 2];
                if (judge >= minV && judge <= maxV) {
                    b = (1 - alpha) * b + alpha * co.z;
                    g = (1 - alpha) * g + alpha * co.y;
                    r = (1 - alpha) * r + alpha * co.x;
                }
            }
        }
        for (int i = 0; i < NP; i++) {
            const float local_x = pose[i * 3], local_y = pose[i * 3 + 1];
            if (pose[i * 3 + 2] > threshold) {
                float dist2 = (x - local_x) * (x - local_x) + (y - local_y) * (y - local_y);
                float minr2 = 0;
                float maxr2 = sc * sc * radius * radius;
                float alpha = 0.6;
                float3 co;
                co.x = c_color18[(i % 18) * 3 + 0];
                co.y = c_color18[(i % 18) * 3 + 1];
                co.z = c_color18[(i % 18) * 3 + 2];
                if (googly_eyes && (i == 14 || i == 15)) {
                    maxr2 = sc * sc * 2.5 * 2.5 * radius * radius;
                    minr2 = sc * sc * (2.5 * radius - 2) * (2.5 * radius - 2);
                    alpha = 0.9;
                    co.x = 0; co.y = 0; co.z = 0;
                    if (dist2 <= maxr2) {
                        if (dist2 <= minr2) { co.x = 255; co.y = 255; co.z = 255; }
                        if (dist2 <= minr2 * 0.6) {
                            float dist3 = (x - 4 - local_x) * (x - 4 - local_x) + (y - local_y + 4) * (y - local_y + 4);
                            if (dist3 > 3.75 * 3.75) { co.x = 0; co.y = 0; co.z = 0; }
                        }
                        b = (1 - alpha) * b + alpha * co.z;
                        g = (1 - alpha) * g + alpha * co.y;
                        r = (1 - alpha) * r + alpha * co.x;
                    }
                } else if (dist2 >= minr2 && dist2 <= maxr2) {
                    b = (1 - alpha) * b + alpha * co.z;
                    g = (1 - alpha) * g + alpha * co.y;
                    r = (1 - alpha) * r + alpha * co.x;
                }
            }
        }
    }
    a.canvas[y * w_canvas + x] = b;
    a.canvas[w_canvas * h_canvas + y * w_canvas + x] = g;
    a.canvas[2 * w_canvas * h_canvas + y * w_canvas + x] = r;
}

// ------------------------------------------------------------------------------------------------
// heat-map / PAF views
// ------------------------------------------------------------------------------------------------
// getColor (renderFunctions.cu:11-44): jet-like map of v in [vmin, vmax]; c = {b, g, r}
__device__ __forceinline__ void jet_color(float* c, float v, float vmin, float vmax) {
    c[0] = c[1] = c[2] = 255;
    float dv;
    if (v < vmin) v = vmin;
    if (v > vmax) v = vmax;
    dv = vmax - vmin;
    if (v < (vmin + 0.125 * dv)) {
        c[0] = 256 * (0.5 + (v * 4));
        c[1] = c[2] = 0;
    } else if (v < (vmin + 0.375 * dv)) {
        c[0] = 255;
        c[1] = 256 * (v - 0.125) * 4;
        c[2] = 0;
    } else if (v < (vmin + 0.625 * dv)) {
        c[0] = 256 * (-4 * v + 2.5);
        c[1] = 255;
        c[2] = 256 * (4 * (v - 0.375));
    } else if (v < (vmin + 0.875 * dv)) {
        c[0] = 0;
        c[1] = 256 * (-4 * v + 3.5);
        c[2] = 255;
    } else {
        c[0] = 0;
        c[1] = 0;
        c[2] = 256 * (-4 * v + 4.5);
    }
}

// getColor2 (renderFunctions.cu:46-94): 55-step optical-flow colour wheel
__device__ __forceinline__ void wheel_color(float* c, float v, float vmin, float vmax) {
    c[0] = c[1] = c[2] = 255;
    if (v < vmin) v = vmin;
    if (v > vmax) v = vmax;
    v = 55 * v;
    const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
    if (v < RY) {
        c[0] = 255;
        c[1] = 255 * (v / (RY));
        c[2] = 0;
    } else if (v < RY + YG) {
        c[0] = 255 - 255 * ((v - RY) / (YG));
        c[1] = 255;
        c[2] = 0;
    } else if (v < RY + YG + GC) {
        c[0] = 0;
        c[1] = 255;
        c[2] = 255 * ((v - RY - YG) / (GC));
    } else if (v < RY + YG + GC + CB) {
        c[0] = 0;
        c[1] = 255 - 255 * ((v - RY - YG - GC) / (CB));
        c[2] = 255;
    } else if (v < RY + YG + GC + CB + BM) {
        c[0] = 255 * ((v - RY - YG - GC - CB) / (BM));
        c[1] = 0;
        c[2] = 255;
    } else if (v < RY + YG + GC + CB + BM + MR) {
        c[0] = 255;
        c[1] = 0;
        c[2] = 255 - 255 * ((v - RY - YG - GC - CB - BM) / (MR));
    } else {
        c[0] = 255;
        c[1] = 0;
        c[2] = 0;
    }
}

// getColorXY (renderFunctions.cu:96-112): direction -> hue, magnitude -> brightness
__device__ __forceinline__ void vector_color(float* c, float x, float y) {
    float rad = sqrt(x * x + y * y);
    float a = atan2(-y, -x) / M_PI;
    float fk = (a + 1) / 2.0;
    if (::isnan(fk)) fk = 0;
    if (rad > 1) rad = 1;
    wheel_color(c, fk, 0, 1);
    c[0] = 255 * (rad * (c[0] / 255));
    c[1] = 255 * (rad * (c[1] / 255));
    c[2] = 255 * (rad * (c[2] / 255));
}

// cubic_interpolation of renderFunctions.cu:114-122 (same expression as the ImResize one)
__device__ __forceinline__ void render_cubic(float& out, float& v0, float& v1, float& v2, float& v3, float dx) {
    out = (-0.5f * v0 + 1.5f * v1 - 1.5f * v2 + 0.5f * v3) * dx * dx * dx
        + (v0 - 2.5f * v1 + 2.0 * v2 - 0.5 * v3) * dx * dx
        + (-0.5f * v0 + 0.5f * v2) * dx
        + v1;
}

struct HeatArgs {
    float* canvas; int w, h;
    float* heat;           // [nch][h_net][w_net]: the channels this view shows, in view order (read only)
    int w_net, h_net;
    int mode;              // 0: MPI part map, 1: COCO part map, 2: COCO all parts (nearest), 3: COCO PAF
    int part;              // channel number in the reference's numbering (colour range / colour index)
    int nch;               // mode 2: 18 parts; mode 3: 2 * num_parts_accum
};

struct Taps { int xn[4], yn[4]; float dx, dy; bool inside; };
__device__ __forceinline__ Taps heat_taps(int x, int y, int w_canvas, int h_canvas, int w_net, int h_net) {
    Taps t;
    float h_inv = (float)h_net / (float)h_canvas;
    float w_inv = (float)w_net / (float)w_canvas;
    float x_on_box = w_inv * x + (0.5 * w_inv - 0.5);
    float y_on_box = h_inv * y + (0.5 * h_inv - 0.5);
    t.inside = x_on_box >= 0 && x_on_box < w_net && y_on_box >= 0 && y_on_box < h_net;
    t.xn[1] = int(x_on_box + 1e-5);
    t.xn[1] = (t.xn[1] < 0) ? 0 : t.xn[1];
    t.xn[0] = (t.xn[1] - 1 < 0) ? t.xn[1] : (t.xn[1] - 1);
    t.xn[2] = (t.xn[1] + 1 >= w_net) ? (w_net - 1) : (t.xn[1] + 1);
    t.xn[3] = (t.xn[2] + 1 >= w_net) ? (w_net - 1) : (t.xn[2] + 1);
    t.dx = x_on_box - t.xn[1];
    t.yn[1] = int(y_on_box + 1e-5);
    t.yn[1] = (t.yn[1] < 0) ? 0 : t.yn[1];
    t.yn[0] = (t.yn[1] - 1 < 0) ? t.yn[1] : (t.yn[1] - 1);
    t.yn[2] = (t.yn[1] + 1 >= h_net) ? (h_net - 1) : (t.yn[1] + 1);
    t.yn[3] = (t.yn[2] + 1 >= h_net) ? (h_net - 1) : (t.yn[2] + 1);
    t.dy = y_on_box - t.yn[1];
    return t;
}

// render_pose_29parts_heatmap :242-329, render_pose_coco_heatmap :638-724, ..._heatmap2 :726-836, ..._affinity :838-975
template <int MODE>
__global__ void __launch_bounds__(256) heat_view_kernel(HeatArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int w_canvas = a.w, h_canvas = a.h, w_net = a.w_net, h_net = a.h_net;
    if (x >= w_canvas || y >= h_canvas) return;
    const int offset2 = w_net * h_net;
    float b = a.canvas[y * w_canvas + x];
    float g = a.canvas[w_canvas * h_canvas + y * w_canvas + x];
    float r = a.canvas[2 * w_canvas * h_canvas + y * w_canvas + x];
    const Taps t = heat_taps(x, y, w_canvas, h_canvas, w_net, h_net);
    if (MODE == 0 || MODE == 1) {
        const int num_parts = MODE == 0 ? 15 : 18;
        float value = (a.part == num_parts - 1) ? 1 : 0;
        if (t.inside) {
            float temp[4];
            for (int i = 0; i < 4; i++)
                render_cubic(temp[i], a.heat[t.yn[i] * w_net + t.xn[0]], a.heat[t.yn[i] * w_net + t.xn[1]],
                             a.heat[t.yn[i] * w_net + t.xn[2]], a.heat[t.yn[i] * w_net + t.xn[3]], t.dx);
            render_cubic(value, temp[0], temp[1], temp[2], temp[3], t.dy);
        }
        float c[3];
        if (MODE == 0) {
            if (a.part < 16) jet_color(c, value, 0, 1); else jet_color(c, value, -1, 1);
            b = 0.5 * b + 0.5 * c[0];
            g = 0.5 * g + 0.5 * c[1];
            r = 0.5 * r + 0.5 * c[2];
        } else {
            if (a.part < num_parts + 1) jet_color(c, value, 0, 1); else jet_color(c, value, -1, 1);
            float alpha = 0.7;
            b = (1 - alpha) * b + alpha * c[2];
            g = (1 - alpha) * g + alpha * c[1];
            r = (1 - alpha) * r + alpha * c[0];
        }
    } else if (MODE == 2) {
        float c[3] = {0, 0, 0};
        for (int part = 0; part < a.nch; part++) {
            if (t.inside) {
                const float value = a.heat[part * offset2 + t.yn[1] * w_net + t.xn[1]];
                c[0] += value * c_color18[(part % 18) * 3 + 0];
                c[1] += value * c_color18[(part % 18) * 3 + 1];
                c[2] += value * c_color18[(part % 18) * 3 + 2];
            }
        }
        float alpha = 0.7;
        b = (1 - alpha) * b + alpha * c[2];
        g = (1 - alpha) * g + alpha * c[1];
        r = (1 - alpha) * r + alpha * c[0];
    } else {
        float c[3] = {0, 0, 0};
        const int num_parts_accum = a.nch / 2;
        for (int k = 0; k < num_parts_accum; k++) {
            if (t.inside) {
                const float* h0 = a.heat + (2 * k) * offset2;
                const float* h1 = a.heat + (2 * k + 1) * offset2;
                float value, value2;
                const float dx = t.dx, dy = t.dy;
                if (num_parts_accum == 1) {
                    {
                        float a_ = h0[t.yn[1] * w_net + t.xn[1]], b_ = h0[t.yn[1] * w_net + t.xn[2]];
                        float c_ = h0[t.yn[2] * w_net + t.xn[1]], d_ = h0[t.yn[2] * w_net + t.xn[2]];
                        value = (1 - dx) * (1 - dy) * a_ + (dx) * (1 - dy) * b_ + (1 - dx) * (dy) * c_ + (dx) * (dy) * d_;
                    }
                    {
                        float a_ = h1[t.yn[1] * w_net + t.xn[1]], b_ = h1[t.yn[1] * w_net + t.xn[2]];
                        float c_ = h1[t.yn[2] * w_net + t.xn[1]], d_ = h1[t.yn[2] * w_net + t.xn[2]];
                        value2 = (1 - dx) * (1 - dy) * a_ + (dx) * (1 - dy) * b_ + (1 - dx) * (dy) * c_ + (dx) * (dy) * d_;
                    }
                } else {
                    value = h0[t.yn[1] * w_net + t.xn[1]];
                    value2 = h1[t.yn[1] * w_net + t.xn[1]];
                }
                float c2[3];
                vector_color(c2, value, value2);
                c[0] += c2[0];
                c[1] += c2[1];
                c[2] += c2[2];
            }
        }
        if (c[0] > 255) c[0] = 255;
        if (c[1] > 255) c[1] = 255;
        if (c[2] > 255) c[2] = 255;
        float alpha = 0.7;
        b = (1 - alpha) * b + alpha * c[2];
        g = (1 - alpha) * g + alpha * c[1];
        r = (1 - alpha) * r + alpha * c[0];
    }
    a.canvas[y * w_canvas + x] = b;
    a.canvas[w_canvas * h_canvas + y * w_canvas + x] = g;
    a.canvas[2 * w_canvas * h_canvas + y * w_canvas + x] = r;
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int launch_canvas_fill(const uint8_t* bgr, float* canvas, int w, int h, cudaStream_t st) {
    canvas_fill_kernel<<<(w * h + 255) / 256, 256, 0, st>>>(bgr, canvas, w, h);
    return 1;
}
int launch_canvas_to_u8(const float* canvas, uint8_t* bgr, int w, int h, cudaStream_t st) {
    canvas_to_u8_kernel<<<(w * h + 255) / 256, 256, 0, st>>>(canvas, bgr, w, h);
    return 1;
}
int launch_fullres_fill(const PostDev& pd, int frame, int ch0, int nch, float* out, cudaStream_t st) {
    fullres_fill_kernel<<<dim3((pd.p.net_w + 255) / 256, pd.p.net_h, nch), 256, 0, st>>>(pd, frame, ch0, nch, out);
    return 1;
}
int launch_skeleton(int model, float* canvas, int w, int h, const float* poses, const int* num_people, int googly, cudaStream_t st) {
    SkelArgs a;
    a.canvas = canvas; a.w = w; a.h = h; a.poses = poses; a.num_people = num_people; a.googly = googly;
    const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    if (model == PE_MODEL_MPI_15) skeleton_mpi_kernel<<<grid, block, 0, st>>>(a);
    else skeleton_coco_kernel<<<grid, block, 0, st>>>(a);
    return 1;
}
int launch_heat_view(float* canvas, int w, int h, float* heat, int w_net, int h_net, int mode, int part, int nch, cudaStream_t st) {
    HeatArgs a;
    a.canvas = canvas; a.w = w; a.h = h; a.heat = heat; a.w_net = w_net; a.h_net = h_net; a.mode = mode; a.part = part; a.nch = nch;
    const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    switch (mode) {
        case 0: heat_view_kernel<0><<<grid, block, 0, st>>>(a); break;
        case 1: heat_view_kernel<1><<<grid, block, 0, st>>>(a); break;
        case 2: heat_view_kernel<2><<<grid, block, 0, st>>>(a); break;
        default: heat_view_kernel<3><<<grid, block, 0, st>>>(a); break;
    }
    return 1;
}

}  // namespace pe
