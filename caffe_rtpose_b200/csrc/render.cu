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
// Arithmetic: the per-pixel tests and blends use the reference's operand types and operation order (float vs double
// promotions included) so that the canvas is bit-identical; tests/test_gpu_render.py compares against the reference's own
// kernels compiled for sm_100a by the test infrastructure.  The kernels themselves are organised differently (see the
// skeleton section: per-CTA limb records, person culling).
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
// skeleton overlays (what render_pose_29parts / render_pose_coco_parts draw, renderFunctions.cu:124-240, 394-636)
//
// Own design: the reference evaluates atan2f / sinf / cosf, the limb midpoint and the ellipse axes of EVERY (person, limb)
// for EVERY pixel.  None of that depends on the pixel, so here each CTA first builds, for a chunk of persons, a table of
// limb records in shared memory (one thread per (person, limb): 256x fewer transcendentals for a 32x8 tile), culls persons
// whose extent cannot reach the tile - for both models; the reference has a per-pixel box for COCO only - and then every
// pixel walks the surviving records in person / limb / joint order (alpha blending is order dependent).  The per-pixel
// ellipse test and the blends use the same float / double operations on the same operands as the reference kernels, so the
// canvas is bit-identical to theirs (tests/test_gpu_render.py runs the reference's kernels next to this one).
// ------------------------------------------------------------------------------------------------
__constant__ int c_limb_coco[34] = {1, 2, 1, 5, 2, 3, 3, 4, 5, 6, 6, 7, 1, 8, 8, 9, 9, 10, 1, 11, 11, 12, 12, 13, 1, 0, 0, 14, 14, 16, 0, 15, 15, 17};
__constant__ int c_limb_mpi[18] = {0, 1, 2, 3, 3, 4, 5, 6, 6, 7, 8, 9, 9, 10, 11, 12, 12, 13};
__constant__ int c_color18[54] = {255, 0, 0, 255, 85, 0, 255, 170, 0, 255, 255, 0, 170, 255, 0, 85, 255, 0, 0, 255, 0, 0, 255, 85, 0, 255, 170,
                                  0, 255, 255, 0, 170, 255, 0, 85, 255, 0, 0, 255, 85, 0, 255, 170, 0, 255, 255, 0, 255, 255, 0, 170, 255, 0, 85};
__constant__ int c_color9[27] = {255, 0, 0, 255, 170, 0, 170, 255, 0, 0, 255, 0, 0, 255, 170, 0, 170, 255, 0, 0, 255, 170, 0, 255, 255, 0, 170};

struct SkelArgs {
    float* canvas; int w, h;
    const float* poses;        // [PE_MAX_PEOPLE][parts][3] of this frame (device)
    const int* num_people;     // device (engine path) or null: num_people_host then holds the count (device-pointer API)
    int num_people_host;
    int googly;
};

struct LimbRec {   // pixel-independent part of one limb ellipse
    float mx, my;              // midpoint of the two joints
    float sn, cs;              // sine / cosine of the limb direction
    float major2, minor2;      // squared semi-axes
    float lo;                  // the ellipse is drawn where lo <= judge <= 1 (the MPI head is a ring)
    int draw;                  // both joints above the confidence threshold
};

template <int MODEL> struct SkelTraits;
template <> struct SkelTraits<PE_MODEL_MPI_15> { static constexpr int NP = 15, NL = 9; };
template <> struct SkelTraits<PE_MODEL_COCO_18> { static constexpr int NP = 18, NL = 17; };

constexpr int SK_CHUNK = 16;   // persons per shared-memory pass

template <int MODEL>
__global__ void __launch_bounds__(256) skeleton_kernel(SkelArgs a) {
    constexpr int NP = SkelTraits<MODEL>::NP, NL = SkelTraits<MODEL>::NL;
    constexpr bool COCO = MODEL == PE_MODEL_COCO_18;
    __shared__ float s_pose[SK_CHUNK][NP * 3];
    __shared__ LimbRec s_limb[SK_CHUNK][NL];
    __shared__ float4 s_box[SK_CHUNK];          // xmin, ymin, xmax, ymax of the person (COCO: the reference's +-50 px box)
    __shared__ float s_size[SK_CHUNK];          // COCO: per-person drawing scale
    __shared__ int s_hit[SK_CHUNK];             // the person can touch this tile
    const int np = min(a.num_people ? *a.num_people : a.num_people_host, PE_MAX_PEOPLE);
    if (np <= 0) return;                        // the reference does not launch (renderFunctions.cu:358, 1006)
    const int W = a.w, H = a.h;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const bool inside = x < W && y < H;
    const float conf_min = COCO ? 0.01f : 0.0f;
    const float radius = (COCO ? 2 : 3) * H / 200.0f;
    const float stickwidth = COCO ? H / 120.0f : H / 60.0f;
    // tile rectangle, for culling
    const float tx0 = (float)(blockIdx.x * blockDim.x), ty0 = (float)(blockIdx.y * blockDim.y);
    const float tx1 = tx0 + (float)(blockDim.x - 1), ty1 = ty0 + (float)(blockDim.y - 1);
    float b = 0.f, g = 0.f, r = 0.f;
    if (inside) {
        b = a.canvas[y * W + x];
        g = a.canvas[W * H + y * W + x];
        r = a.canvas[2 * W * H + y * W + x];
    }
    for (int p0 = 0; p0 < np; p0 += SK_CHUNK) {
        const int nq = min(SK_CHUNK, np - p0);
        __syncthreads();
        // ---- per person: joints, extent, scale, tile test
        for (int q = tid; q < nq; q += nthr) {
            const float* src = a.poses + (size_t)(p0 + q) * NP * 3;
            float xmin = (float)W, ymin = (float)H, xmax = 0.f, ymax = 0.f;
            for (int j = 0; j < NP; j++) {
                const float px = src[j * 3], py = src[j * 3 + 1], pc = src[j * 3 + 2];
                s_pose[q][j * 3] = px; s_pose[q][j * 3 + 1] = py; s_pose[q][j * 3 + 2] = pc;
                if (pc > conf_min) {
                    if (px < xmin) xmin = px;
                    if (px > xmax) xmax = px;
                    if (py < ymin) ymin = py;
                    if (py > ymax) ymax = py;
                }
            }
            float size = 1.f;
            float margin;
            if (COCO) {   // person size -> drawing scale, and the +-50 px box outside of which the person is not drawn (:424-452)
                float sx = xmax - xmin, sy = ymax - ymin;
                sx = (sx + sy) / 2.0;
                if (sx < 200) {
                    sx = sx / 200;
                    if (sx < 0.33) sx = 0.33;
                } else {
                    sx = 1.0;
                }
                size = sx;
                xmax += 50; ymax += 50; xmin -= 50; ymin -= 50;
                margin = 0.f;
            } else {      // no box in the reference: a conservative reach of any ellipse / circle of this person
                const float dx = xmax - xmin, dy = ymax - ymin;
                margin = fmaxf(stickwidth, radius) + 0.1f * sqrtf(dx * dx + dy * dy) + 2.f;
            }
            s_box[q] = make_float4(xmin, ymin, xmax, ymax);
            s_size[q] = size;
            s_hit[q] = !(tx0 > xmax + margin || tx1 < xmin - margin || ty0 > ymax + margin || ty1 < ymin - margin);
        }
        __syncthreads();
        // ---- per (person, limb): the pixel-independent half of the ellipse test
        for (int i = tid; i < nq * NL; i += nthr) {
            const int q = i / NL, l = i % NL;
            if (!s_hit[q]) continue;
            const int ja = COCO ? c_limb_coco[2 * l] : c_limb_mpi[2 * l], jb = COCO ? c_limb_coco[2 * l + 1] : c_limb_mpi[2 * l + 1];
            const float x_a = s_pose[q][ja * 3], y_a = s_pose[q][ja * 3 + 1], x_b = s_pose[q][jb * 3], y_b = s_pose[q][jb * 3 + 1];
            LimbRec rec;
            rec.draw = s_pose[q][ja * 3 + 2] > conf_min && s_pose[q][jb * 3 + 2] > conf_min;
            float x_p = (x_a + x_b) / 2;
            float y_p = (y_a + y_b) / 2;
            float angle = atan2f(y_b - y_a, x_b - x_a);
            rec.sn = sinf(angle);
            rec.cs = cosf(angle);
            float a_sqrt = (x_a - x_p) * (x_a - x_p) + (y_a - y_p) * (y_a - y_p);
            float b_sqrt = COCO ? s_size[q] * s_size[q] * stickwidth * stickwidth : stickwidth * stickwidth;
            float lo = 0;
            if (!COCO && l == 0) {   // MPI head: a ring, 1.2x the neck-head distance
                a_sqrt *= 1.2;
                b_sqrt = a_sqrt;
                lo = 0.8;
            }
            rec.mx = x_p; rec.my = y_p; rec.major2 = a_sqrt; rec.minor2 = b_sqrt; rec.lo = lo;
            s_limb[q][l] = rec;
        }
        __syncthreads();
        if (!inside) continue;
        // ---- per pixel, in person / limb / joint order
        for (int q = 0; q < nq; q++) {
            if (!s_hit[q]) continue;
            if (COCO) {
                const float4 bx = s_box[q];
                if (x > bx.z || x < bx.x || y > bx.w || y < bx.y) continue;
            }
            const float sc = s_size[q];
            for (int l = 0; l < NL; l++) {
                const LimbRec& rec = s_limb[q][l];
                if (!rec.draw) continue;
                const float x_p = rec.mx, y_p = rec.my, sine = rec.sn, cosine = rec.cs, a_sqrt = rec.major2, b_sqrt = rec.minor2;
                float A = cosine * (x - x_p) + sine * (y - y_p);
                float B = sine * (x - x_p) - cosine * (y - y_p);
                float judge = A * A / a_sqrt + B * B / b_sqrt;
                if (judge >= rec.lo && judge <= 1) {
                    if (COCO) {
                        float alpha = 0.5;
                        b = (1 - alpha) * b + alpha * (float)c_color18[(l % 18) * 3 + 2];
                        g = (1 - alpha) * g + alpha * (float)c_color18[(l % 18) * 3 + 1];
                        r = (1 - alpha) * r + alpha * (float)c_color18[(l % 18) * 3 + 0];
                    } else {
                        float alpha = 0.6;
                        b = (1 - alpha) * b + alpha * c_color9[l * 3 + 2];
                        g = (1 - alpha) * g + alpha * c_color9[l * 3 + 1];
                        r = (1 - alpha) * r + alpha * c_color9[l * 3];
                    }
                }
            }
            for (int j = 0; j < NP; j++) {
                if (!(s_pose[q][j * 3 + 2] > conf_min)) continue;
                const float jx = s_pose[q][j * 3], jy = s_pose[q][j * 3 + 1];
                if (!COCO) {
                    if ((x - jx) * (x - jx) + (y - jy) * (y - jy) <= radius * radius) {
                        b = 0.6 * b + 0.4 * c_color9[(j % 9) * 3 + 2];
                        g = 0.6 * g + 0.4 * c_color9[(j % 9) * 3 + 1];
                        r = 0.6 * r + 0.4 * c_color9[(j % 9) * 3];
                    }
                    continue;
                }
                float dist2 = (x - jx) * (x - jx) + (y - jy) * (y - jy);
                float minr2 = 0;
                float maxr2 = sc * sc * radius * radius;
                float alpha = 0.6;
                float cr = c_color18[(j % 18) * 3 + 0], cg = c_color18[(j % 18) * 3 + 1], cb = c_color18[(j % 18) * 3 + 2];
                if (a.googly && (j == 14 || j == 15)) {   // eyes: white disc, black rim, pupil offset by (4, -4)
                    maxr2 = sc * sc * 2.5 * 2.5 * radius * radius;
                    minr2 = sc * sc * (2.5 * radius - 2) * (2.5 * radius - 2);
                    alpha = 0.9;
                    cr = 0; cg = 0; cb = 0;
                    if (dist2 <= maxr2) {
                        if (dist2 <= minr2) { cr = 255; cg = 255; cb = 255; }
                        if (dist2 <= minr2 * 0.6) {
                            float dist3 = (x - 4 - jx) * (x - 4 - jx) + (y - jy + 4) * (y - jy + 4);
                            if (dist3 > 3.75 * 3.75) { cr = 0; cg = 0; cb = 0; }
                        }
                        b = (1 - alpha) * b + alpha * cb;
                        g = (1 - alpha) * g + alpha * cg;
                        r = (1 - alpha) * r + alpha * cr;
                    }
                } else if (dist2 >= minr2 && dist2 <= maxr2) {
                    b = (1 - alpha) * b + alpha * cb;
                    g = (1 - alpha) * g + alpha * cg;
                    r = (1 - alpha) * r + alpha * cr;
                }
            }
        }
    }
    if (inside) {
        a.canvas[y * W + x] = b;
        a.canvas[W * H + y * W + x] = g;
        a.canvas[2 * W * H + y * W + x] = r;
    }
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
int launch_skeleton(int model, float* canvas, int w, int h, const float* poses, const int* num_people, int googly, cudaStream_t st,
                    int num_people_host) {
    SkelArgs a;
    a.canvas = canvas; a.w = w; a.h = h; a.poses = poses; a.num_people = num_people; a.num_people_host = num_people_host; a.googly = googly;
    const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    if (model == PE_MODEL_MPI_15) skeleton_kernel<PE_MODEL_MPI_15><<<grid, block, 0, st>>>(a);
    else skeleton_kernel<PE_MODEL_COCO_18><<<grid, block, 0, st>>>(a);
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
