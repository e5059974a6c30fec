// fp32 implicit-GEMM convolution on CUDA cores (PE_PREC_FP32_SIMT).
//
// Role: the exact-fp32 arithmetic reference on the device (same rounding class as Caffe's
// im2col + cblas_sgemm, base_conv_layer.cpp:257-279) and the bring-up baseline that the tcgen05 kernel
// (conv_tc.cu) is checked against.  Same data layout as the tensor-core path: flat padded NHWC, a filter
// tap is a constant row shift (common.h), bias + ReLU fused in the epilogue, concat by channel slice,
// final stage written straight as the planar concat_stage7 blob.
//
//   out[m][co] = bias[co] + sum_{tap} sum_{c} in[m + shift(tap)][c] * w[(tap*cin_pad + c)][co]
//
// Tile: 128 rows x BN channels per CTA (256 threads, 8 x BN/16 outputs per thread), BK = 16, register
// prefetch of the next K-slab while the current one is consumed from shared memory.
#include "common.h"
#include "kernels.h"

namespace pe {

#define SIMT_BM 128
#define SIMT_BK 16

template <int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
    constexpr int TN = BN / 16;
    __shared__ __align__(16) float As[2][SIMT_BK][SIMT_BM + 4];
    __shared__ __align__(16) float Bs[2][SIMT_BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const long long m0 = (long long)blockIdx.x * SIMT_BM;
    const int n0 = blockIdx.y * BN;
    const float* in = (const float*)a.in;
    const float* w = (const float*)a.w;
    const int taps = a.ksize * a.ksize;
    const int kblocks_per_tap = a.cin_pad / SIMT_BK;
    const int nk = taps * kblocks_per_tap;

    // loader mapping: A: row = tid/2, 8 consecutive k at (tid%2)*8 ; B: BN/4 float4 per k-row
    const int a_row = tid >> 1, a_k = (tid & 1) * 8;
    float4 ra0, ra1;
    float4 rb[BN == 128 ? 2 : 1];

    auto load_global = [&](int kb) {
        const int tap = kb / kblocks_per_tap, c0 = (kb % kblocks_per_tap) * SIMT_BK;
        const int r = tap / a.ksize, s = tap % a.ksize;
        const long long row = m0 + a_row + (long long)(r - a.pad) * a.Wp + (s - a.pad);
        if (row >= 0 && row < a.M) {
            const float4* p = (const float4*)(in + row * a.in_pitch + c0 + a_k);
            ra0 = __ldg(p); ra1 = __ldg(p + 1);
        } else {
            ra0 = make_float4(0.f, 0.f, 0.f, 0.f); ra1 = ra0;
        }
        const long long krow = (long long)tap * a.cin_pad + c0;
#pragma unroll
        for (int i = 0; i < (BN == 128 ? 2 : 1); i++) {
            const int f = tid + i * 256;           // float4 index within the 16 x BN slab
            const int kk = f / (BN / 4), nn = (f % (BN / 4)) * 4;
            rb[i] = __ldg((const float4*)(w + (krow + kk) * a.cout_pad + n0 + nn));
        }
    };
    auto store_smem = [&](int buf) {
        As[buf][a_k + 0][a_row] = ra0.x; As[buf][a_k + 1][a_row] = ra0.y; As[buf][a_k + 2][a_row] = ra0.z; As[buf][a_k + 3][a_row] = ra0.w;
        As[buf][a_k + 4][a_row] = ra1.x; As[buf][a_k + 5][a_row] = ra1.y; As[buf][a_k + 6][a_row] = ra1.z; As[buf][a_k + 7][a_row] = ra1.w;
#pragma unroll
        for (int i = 0; i < (BN == 128 ? 2 : 1); i++) {
            const int f = tid + i * 256;
            const int kk = f / (BN / 4), nn = (f % (BN / 4)) * 4;
            *(float4*)&Bs[buf][kk][nn] = rb[i];
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int kb = 0; kb < nk; kb++) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_global(kb + 1);
#pragma unroll
        for (int k = 0; k < SIMT_BK; k++) {
            float av[8], bv[TN];
            const float4 a0 = *(const float4*)&As[buf][k][ty * 8];
            const float4 a1 = *(const float4*)&As[buf][k][ty * 8 + 4];
            av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 b = *(const float4*)&Bs[buf][k][tx * TN + j];
                bv[j] = b.x; bv[j + 1] = b.y; bv[j + 2] = b.z; bv[j + 3] = b.w;
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) store_smem(buf ^ 1);
        __syncthreads();
    }

    // epilogue: bias + ReLU, store valid rows only (gap rows must stay zero)
    const int per_img = a.Hs * a.Wp;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long long m = m0 + ty * 8 + i;
        if (m >= a.M) continue;
        const int n = (int)(m / per_img);
        const int rem = (int)(m % per_img);
        const int y = rem / a.Wp, x = rem % a.Wp;
        if (x >= a.W || y >= a.H) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int co = n0 + tx * TN + j;
            if (co >= a.cout) continue;
            float v = acc[i][j] + a.bias[co];
            if (a.relu) v = fmaxf(v, 0.f);
            if (a.planar) a.planar[(((size_t)n * a.planar_C + a.planar_coff + co) * a.H + y) * a.W + x] = v;
            else ((float*)a.out)[m * a.out_pitch + a.out_coff + co] = v;
        }
    }
}

int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
    const unsigned gm = (unsigned)((a.M + SIMT_BM - 1) / SIMT_BM);
    if (a.cout_pad % 128 == 0) {
        conv_simt_kernel<128><<<dim3(gm, a.cout_pad / 128), 256, 0, st>>>(a);
    } else {
        conv_simt_kernel<64><<<dim3(gm, a.cout_pad / 64), 256, 0, st>>>(a);
    }
    return 1;
}

}  // namespace pe
