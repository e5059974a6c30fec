// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (B200), hand-written PTX.
//
// Replaces cudnnConvolutionForward + cudnnAddTensor + ReLU (src/caffe/layers/cudnn_conv_layer.cu:11-46,
// cudnn_relu_layer.cu:19) and the Concat copies (concat_layer.cu:40-43) for every 3x3 / 7x7 / 1x1
// convolution of the deploy graph.  Arithmetic contract = Caffe's conv (base_conv_layer.cpp:257-279):
// out = W * im2col(in) + bias, zero padding, stride 1.
//
// Mapping to the hardware
//   * GEMM view: D[m][co] = sum_{tap} sum_{c} A[m + shift(tap)][c] * Wt[co][tap*C + c], M = flat padded
//     pixels (common.h), so the A tile of one (tap, 64-channel block) is ONE 2-D TMA box {64 ch x 128 rows}
//     at a shifted row coordinate; zero padding = TMA out-of-bounds fill + never-written gap rows.
//   * TMA (cp.async.bulk.tensor.3d, SWIZZLE_128B) stages A and B tiles into shared memory, completion on
//     mbarriers; rings of slots overlap loads with MMAs.  The k taps of one filter row share one 136-row A window.
//   * One elected lane issues tcgen05.mma.kind::f16 (K=16 per instruction), accumulators in TMEM, fp32.
//   * Split precision: activations and weights are stored as P 16-bit "planes" whose sum is the fp32 value
//     (hi / lo); the kernel issues the cross products hi*hi, hi*lo, lo*hi on the tensor core.  P=1 plain bf16,
//     P=2 = parity mode on fp16 planes (kernels.h), ~1e-5 over the whole net.
//   * Epilogue warps read TMEM with tcgen05.ld, sum the accumulation chunks in registers (round to nearest), add bias,
//     ReLU, re-split into planes and store NHWC planes (channel-slice stores implement Concat; TMA stores from a
//     swizzled staging tile in the pair kernel), or - for the last stage - the planar fp32 concat_stage7 blob.
// Three kernels, newest first:
//   conv_tcp_kernel  CTA pairs (cta_group::2, M = 256), the product kernel of the parity mode          (round 2)
//   conv_tcw_kernel  one CTA per 128-row tile, persistent, window trick, N-concatenated split precision  (round 1)
//   conv_tc_kernel   one TMA tile per tap, non-persistent; baseline kept for A/B and the 3-plane mode
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include <string>

#include "common.h"
#include "conv_tc.h"
#include "kernels.h"

namespace pe {

struct TcArgs {
    const float* bias;
    __nv_bfloat16* out; int out_pitch, out_coff; long long out_plane;
    float* planar; int planar_C, planar_coff;
    int cout, relu;
    const float* out_scale;                   // -> 2^-k in the packed buffer: the weights of this layer carry a factor 2^k (fp16 planes)
    int ksize, pad, kblocks_per_tap, cin_k;   // cin_k = channels per tap in the weight K ordering
    int W, H, Wp, Hs;
    long long M;
    int n_tiles_n; long long total_tiles;   // persistent window kernel: tile = (m_tile, n_tile), n fastest
    int chunk_steps;                        // (channel block, filter row) steps per TMEM accumulation chunk (window kernel)
    unsigned* range;                        // [0]: running max |stored value| of this layer as float bits (atomicMax; values are >= 0), or null
    int a_hi_only;                          // pair kernel: the lo plane of the input is identically zero (conv1_1 on uint8 frames: k/256 - 0.5 is exact in
                                            // fp16): only the hi window is loaded and the A_lo x B_hi product is not issued - same result bit for bit
    int tma_store;                          // pair kernel: the epilogue stages 32-row x HALF-channel boxes in shared memory and stores them with TMA
    int dbg;                                // PE_TC_DBG bit mask, TIMING EXPERIMENTS ONLY (results are wrong): 1 no TMEM loads in the chunk drains,
                                            // 2 no epilogue math / stores, 4 weight tiles loaded once per slot only, 8 A windows loaded once per slot only, 16 one-lane issue loop,
                                            // 32 epilogue complete except the TMA store instructions
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4
//   (8 rows x 128 B = 1024 B -> 64) | [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// A row-shifted view of a larger tile (start address not on a 1024-byte swizzle-atom boundary) uses the same
// descriptor with the shifted start address: on B200 the swizzle phase follows the absolute shared-memory address
// bits, and the base_offset field [49,52) must stay 0 (measured: setting it to (addr>>7)&7 gives wrong results).
// instruction descriptor kind::f16 (InstrDescriptor): c_format F32 (bit 4), a/b format BF16 (bits 7,10),
// K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool f16 = false) {   // a/b format: 0 = F16, 1 = BF16
    return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// One lane of a CONVERGED warp.  The MMA-issuing warp runs its loop with all 32 lanes (barrier waits, index math) and
// only the issue itself under elect_one(): the compiler then keeps descriptors in uniform registers and emits plain
// UTCHMMA.  With the role guarded by `lane == 0` it wrapped every UTCHMMA in an ELECT / BRA.U.ANY uniformisation loop
// plus R2UR moves, ~115 SASS instructions per filter tap, and the single issuing thread - not the tensor pipe - set the
// pace (ncu r2a: issuer never blocked on a barrier, tensor pipe 74 % on the 7x7 layers, 65 % on 3x3).
__device__ __forceinline__ bool elect_one(uint32_t mask = 0xffffffffu) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, %1;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred) : "r"(mask));
    return pred != 0;
}

// Programmatic dependent launch (PDL): a conv kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream drains.  Everything before pdl_wait() - barrier init, TMEM allocation,
// tensor-map prefetch, bias staging, L2 prefetch of the first weight tiles - overlaps the predecessor's tail;
// pdl_wait() returns once the predecessor grid has completed and its writes are visible.  pdl_launch_dependents()
// lets the NEXT kernel's CTAs be scheduled as soon as SMs free up.  No-ops without the launch attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// Range tracking of the 16-bit planes: every epilogue warp folds the largest |value| it stores into one word per layer.  The
// host reads it to calibrate per-layer power-of-two activation scales (engine.cu, pe_calibrate) and to report values that
// left the fp16 range instead of letting inf / flushed zeros poison the following layers silently.
__device__ __forceinline__ void range_publish(unsigned* range, float m) {
    if (!range) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(range, __float_as_uint(m));
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;          // bf16 elements per K block = one 128-byte swizzle row
constexpr int TC_THREADS = 192;
constexpr int TCW_THREADS = 320;      // window kernel: TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quadrant)

template <int BN, int PLANES, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    constexpr int A_BYTES = TC_BM * 128;
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
    constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : 128);
    constexpr uint32_t IDESC = umma_idesc(TC_BM, BN, planes_are_fp16(PLANES));

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024 B
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * TC_BM;
    const int n0 = blockIdx.y * BN;
    const int taps = a.ksize * a.ksize;
    const int num_k = taps * a.kblocks_per_tap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int it = 0; it < num_k; it++) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            const int tap = it / a.kblocks_per_tap, kb = it % a.kblocks_per_tap;
            const int r = tap / a.ksize, q = tap % a.ksize;
            const int row0 = (int)(m0 + (long long)(r - a.pad) * a.Wp + (q - a.pad));
            uint8_t* st = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
            for (int p = 0; p < PLANES; p++) tma_load_3d(st + p * A_BYTES, &tmA, &full_bar[s], kb * TC_BK, row0, p);
#pragma unroll
            for (int p = 0; p < PLANES; p++)
                tma_load_3d(st + PLANES * A_BYTES + p * B_BYTES, &tmB, &full_bar[s], tap * a.cin_k + kb * TC_BK, n0, p);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        for (int it = 0; it < num_k; it++) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
            const uint32_t sb = sa + PLANES * A_BYTES;
            bool first = (it == 0);
#pragma unroll
            for (int pa = 0; pa < PLANES; pa++)
#pragma unroll
                for (int pb = 0; pb < PLANES - pa; pb++) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++) {
                        const uint64_t da = umma_desc(sa + pa * A_BYTES + k * 32);
                        const uint64_t db = umma_desc(sb + pb * B_BYTES + k * 32);
                        umma_bf16(tmem_base, da, db, IDESC, first ? 0u : 1u);
                        first = false;
                    }
                }
            umma_commit(&empty_bar[s]);   // frees the smem slot when these MMAs have read it
        }
        umma_commit(&tmem_full_bar);      // accumulator complete
    } else if (warp >= 2) {
        // ===== epilogue =====
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int quad = warp & 3;
        const long long m = m0 + quad * 32 + lane;
        const int per_img = a.Hs * a.Wp;
        bool valid = m < a.M;
        int n = 0, y = 0, x = 0;
        if (valid) {
            n = (int)(m / per_img);
            const int rem = (int)(m % per_img);
            y = rem / a.Wp; x = rem % a.Wp;
            valid = (x < a.W) && (y < a.H);
        }
        const int cout8 = (a.cout + 7) & ~7;
        const float out_scale = __ldg(a.out_scale);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t r[16];
            __syncwarp();
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);   // warp-collective
            if (valid) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int co = n0 + c0 + j;
                    float t = __fmaf_rn(__uint_as_float(r[j]), out_scale, co < a.cout ? __ldg(a.bias + co) : 0.f);
                    if (a.relu) t = fmaxf(t, 0.f);
                    v[j] = t;
                }
                if (a.planar) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int co = n0 + c0 + j;
                        if (co < a.cout) a.planar[(((size_t)n * a.planar_C + a.planar_coff + co) * a.H + y) * a.W + x] = v[j];
                    }
                } else {
                    uint32_t pk[PLANES][8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        float r0 = v[j], r1 = v[j + 1];
#pragma unroll
                        for (int p = 0; p < PLANES; p++) pk[p][j / 2] = split_pair<planes_are_fp16(PLANES)>(r0, r1);
                    }
                    __nv_bfloat16* orow = a.out + (size_t)m * a.out_pitch + a.out_coff + n0 + c0;
#pragma unroll
                    for (int p = 0; p < PLANES; p++) {
                        uint4* dst = (uint4*)(orow + (size_t)p * a.out_plane);
                        if (n0 + c0 < cout8) dst[0] = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
                        if (n0 + c0 + 8 < cout8) dst[1] = make_uint4(pk[p][4], pk[p][5], pk[p][6], pk[p][7]);
                    }
                }
            }
        }
        __syncwarp();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// Window variant (default): the A tiles of the k taps of one filter row are 1-row shifts of each other in the
// flat padded layout, so ONE TMA box of 128+k-1 (rounded to 136) rows per (filter row, channel block) feeds
// all k taps through row-shifted UMMA descriptors.  A traffic from L2 drops k-fold (7x for the 7x7 layers that
// are 68 % of the FLOPs); A windows and per-tap B tiles run in two independent mbarrier rings.
// Split precision P=2 uses N-concatenation: MMA1 = A_hi x [B_hi;B_lo] (N = 2*BN, TMEM columns [0,BN) = hi*hi,
// [BN,2BN) = hi*lo) and MMA2 = A_lo x B_hi accumulated into [BN,2BN): 2 instructions instead of 3 per K=16
// step, A_hi read once, and the large hi*hi chain sits alone in its accumulator (the tensor core's fp32
// accumulation truncates, so a shorter chain per accumulator is also more accurate).  The epilogue adds the
// two accumulators in fp32.
// ------------------------------------------------------------------------------------------------
constexpr int TCW_ROWS = 136;                 // 128 + 7 - 1 rounded up to a multiple of 8
constexpr int TCW_A_BYTES = TCW_ROWS * 128;   // 17408 = 17 * 1024

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// PERSISTENT: grid = min(#tiles, #SMs); every role loops over tiles t = blockIdx.x, +gridDim.x, ...  The TMA and
// MMA rings run straight through tile boundaries, and the TMEM accumulator is double-buffered (2 x P*BN
// columns) so the epilogue of tile i overlaps the main loop of tile i+1 - this removes the per-tile
// prologue/epilogue latency that dominated the short-K layers (conv1_x, conv2_x).
template <int BN, int PLANES, int NA, int NB, int ROWB>
__global__ void __launch_bounds__(TCW_THREADS, 1)
conv_tcw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    static_assert(PLANES == 1 || PLANES == 2, "window kernel supports 1 or 2 planes");
    constexpr int B_BYTES = BN * 128;
    constexpr int A_SLOT = PLANES * TCW_A_BYTES;
    constexpr int B_TAP = PLANES * B_BYTES;                 // one tap: [B_hi ; B_lo]
    constexpr int B_SLOT = (ROWB ? 3 : 1) * B_TAP;          // ROWB: all taps of a filter row (ksize <= 3) share a slot
    constexpr int ACC_COLS = PLANES * BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
    constexpr bool F16 = planes_are_fp16(PLANES);
    constexpr uint32_t IDESC1 = umma_idesc(TC_BM, PLANES * BN, F16);   // A_hi x [B_hi;B_lo]
    constexpr uint32_t IDESC2 = umma_idesc(TC_BM, BN, F16);            // A_lo x B_hi

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + NA * A_SLOT;
    __shared__ __align__(8) uint64_t a_full[NA], a_empty[NA], b_full[NB], b_empty[NB];
    __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_bias[512];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ks = a.ksize;
    for (int i = threadIdx.x; i < 512; i += TCW_THREADS) s_bias[i] = (i < a.cout) ? a.bias[i] : 0.f;
    const int n_tiles_n = a.n_tiles_n;
    const long long total_tiles = a.total_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NB; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
    if (warp == 0 && lane == 0 && blockIdx.x < total_tiles) {   // weights do not depend on the previous kernel: warm L2 with the first taps
        const int n0 = (int)(blockIdx.x % n_tiles_n) * BN;
        for (int q = 0; q < (ks < 4 ? ks : 4); q++) tma_prefetch_3d(&tmB, q * a.cin_k, n0, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: windows and per-tap weight tiles in consumption order =====
        int aw = 0, bt = 0;
        for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const long long m0 = (t / n_tiles_n) * TC_BM;
            const int n0 = (int)(t % n_tiles_n) * BN;
            for (int kb = 0; kb < a.kblocks_per_tap; kb++)
                for (int r = 0; r < ks; r++) {
                    {
                        const int s = aw % NA;
                        mbar_wait(&a_empty[s], ((uint32_t)(aw / NA) & 1u) ^ 1u);
                        const int row0 = (int)(m0 + (long long)(r - a.pad) * a.Wp - a.pad);
                        if ((a.dbg & 8) && aw >= NA) mbar_arrive(&a_full[s]);
                        else {
                            mbar_expect_tx(&a_full[s], A_SLOT);
                            tma_load_3d(smem_a + s * A_SLOT, &tmA, &a_full[s], kb * TC_BK, row0, 0);   // box {64, 136, PLANES}
                        }
                        aw++;
                    }
                    if (ROWB) {   // short-K / narrow-N layers: one barrier round trip per filter row instead of per tap
                        const int s = bt % NB;
                        mbar_wait(&b_empty[s], ((uint32_t)(bt / NB) & 1u) ^ 1u);
                        mbar_expect_tx(&b_full[s], ks * B_TAP);
                        for (int q = 0; q < ks; q++)
                            tma_load_3d(smem_b + s * B_SLOT + q * B_TAP, &tmB, &b_full[s], (r * ks + q) * a.cin_k + kb * TC_BK, n0, 0);
                        bt++;
                    } else {
                        for (int q = 0; q < ks; q++) {
                            const int s = bt % NB;
                            mbar_wait(&b_empty[s], ((uint32_t)(bt / NB) & 1u) ^ 1u);
                            const int tap = r * ks + q;
                            if ((a.dbg & 4) && bt >= NB) mbar_arrive(&b_full[s]);
                            else {
                                mbar_expect_tx(&b_full[s], B_TAP);
                                tma_load_3d(smem_b + s * B_SLOT, &tmB, &b_full[s], tap * a.cin_k + kb * TC_BK, n0, 0);   // box {64, BN, PLANES}
                            }
                            bt++;
                        }
                    }
                }
        }
    } else if (warp == 1 && (((a.dbg & 16) ? 1u : 0xffffffffu) >> lane & 1u)) {
        const uint32_t wmask = (a.dbg & 16) ? 1u : 0xffffffffu;   // PE_TC_DBG bit 4: only lane 0 runs the issue loop (experiment)
        // ===== MMA issuer: the whole warp runs the loop converged, one elected lane issues (see elect_one) =====
        // The tensor core adds every K=16 step into the fp32 TMEM accumulator with truncation, so a long chain drifts
        // (measured: 1.9e-4 relative over the net with one accumulator, 5.7e-5 with hi*hi alone in its accumulator).
        // The chain is therefore cut into CHUNKS of a.chunk_steps steps: every chunk starts from zero in one of the two
        // TMEM buffers and the epilogue warps add the chunks in registers with round-to-nearest while the next chunk runs.
        int aw = 0, bt = 0;
        uint32_t ci = 0;
        const int nsteps = a.kblocks_per_tap * ks, cs = a.chunk_steps;
        const uint32_t sa_base = smem_u32(smem_a), sb_base = smem_u32(smem_b);
        for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            int step = 0, as = 0;
            uint32_t acc = 0;
            uint32_t first = 0;   // accumulate flag of the next hi*hi MMA (0 at the start of a chunk)
            for (int kb = 0; kb < a.kblocks_per_tap; kb++)
                for (int r = 0; r < ks; r++) {
                    if (step % cs == 0) {
                        as = (int)(ci & 1u);
                        mbar_wait(&tmem_empty[as], ((ci >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
                        acc = tmem_base + (uint32_t)(as * ACC_COLS);
                        first = 0;
                    }
                    const int sa_slot = aw % NA;
                    mbar_wait(&a_full[sa_slot], (uint32_t)(aw / NA) & 1u);
                    // descriptors are built once per operand; K steps and row shifts only move the 14-bit start address
                    const uint64_t dA0 = umma_desc(sa_base + sa_slot * A_SLOT), dA1 = dA0 + (uint64_t)(TCW_A_BYTES >> 4);
                    if (ROWB) {
                        const int sb_slot = bt % NB;
                        mbar_wait(&b_full[sb_slot], (uint32_t)(bt / NB) & 1u);
                        tc_fence_after();
                        const uint64_t dB = umma_desc(sb_base + sb_slot * B_SLOT);
                        if (elect_one(wmask)) {
                            for (int q = 0; q < ks; q++) {
#pragma unroll
                                for (int k = 0; k < TC_BK / 16; k++) {
                                    const uint64_t db = dB + (uint64_t)(q * (B_TAP >> 4) + k * 2);
                                    umma_bf16(acc, dA0 + (uint64_t)(q * 8 + k * 2), db, IDESC1, (q | k) ? 1u : first);
                                    if (PLANES == 2) umma_bf16(acc + BN, dA1 + (uint64_t)(q * 8 + k * 2), db, IDESC2, 1u);
                                }
                            }
                            umma_commit(&b_empty[sb_slot]);
                        }
                        __syncwarp(wmask);
                        first = 1;
                        bt++;
                    } else {
                        for (int q = 0; q < ks; q++) {
                            const int sb_slot = bt % NB;
                            mbar_wait(&b_full[sb_slot], (uint32_t)(bt / NB) & 1u);
                            tc_fence_after();
                            const uint64_t dB = umma_desc(sb_base + sb_slot * B_SLOT);
                            if (elect_one(wmask)) {
#pragma unroll
                                for (int k = 0; k < TC_BK / 16; k++) {
                                    // row-shifted view of the window: start address + q rows (q*128 B = q*8 units); the swizzle
                                    // phase follows the absolute smem address (verified on B200: base_offset must stay 0)
                                    umma_bf16(acc, dA0 + (uint64_t)(q * 8 + k * 2), dB + (uint64_t)(k * 2), IDESC1, k ? 1u : first);
                                    if (PLANES == 2) umma_bf16(acc + BN, dA1 + (uint64_t)(q * 8 + k * 2), dB + (uint64_t)(k * 2), IDESC2, 1u);
                                }
                                umma_commit(&b_empty[sb_slot]);
                            }
                            __syncwarp(wmask);
                            first = 1;
                            bt++;
                        }
                    }
                    step++;
                    if (elect_one(wmask)) {
                        umma_commit(&a_empty[sa_slot]);
                        if (step % cs == 0 || step == nsteps) umma_commit(&tmem_full[as]);
                    }
                    __syncwarp(wmask);
                    if (step % cs == 0 || step == nsteps) ci++;
                    aw++;
                }
        }
    } else if (warp >= 2) {
        // ===== epilogue: 8 warps; warp w owns TMEM lanes 32*(w%4).. and the column half (w-2)/4 =====
        constexpr bool SPLIT = (BN / 2) % 16 == 0;            // BN = 128, 64, 32: two warps share a lane quadrant
        constexpr int HALF = SPLIT ? BN / 2 : BN;             // columns per warp
        constexpr int NCHUNK = HALF / 16;
        const int quad = warp & 3;
        const int half = SPLIT ? (warp - 2) / 4 : 0;
        const bool active_half = SPLIT || (warp - 2) / 4 == 0;
        const int per_img = a.Hs * a.Wp;
        const int cout8 = (a.cout + 7) & ~7;
        const float out_scale = __ldg(a.out_scale);
        uint32_t ci = 0;
        float range_max = 0.f;   // largest |value| this warp stores (range tracking of the fp16 planes, see TcArgs::range)
        const int nsteps = a.kblocks_per_tap * ks;
        const int nchunks = (nsteps + a.chunk_steps - 1) / a.chunk_steps;
        for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const long long m0 = (t / n_tiles_n) * TC_BM;
            const int n0 = (int)(t % n_tiles_n) * BN;
            const long long m = m0 + quad * 32 + lane;
            bool valid = m < a.M;
            int n = 0, y = 0, x = 0;
            if (valid) {
                n = (int)(m / per_img);
                const int rem = (int)(m % per_img);
                y = rem / a.Wp; x = rem % a.Wp;
                valid = (x < a.W) && (y < a.H);
            }
            // ---- sum the chunks of this tile in registers (fp32 round-to-nearest) ----
            float accv[HALF];
            const int ndrain = nchunks;
            for (int c = 0; c < ndrain; c++, ci++) {
                const int as = (int)(ci & 1u);
                mbar_wait(&tmem_full[as], (ci >> 1) & 1u);
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * ACC_COLS + half * HALF);
                if (active_half && !((a.dbg & 1) && c > 0)) {
                    uint32_t r[2][16], r2[2][16];
                    __syncwarp();
                    tmem_ld16_nowait(trow, r[0]);
                    if (PLANES == 2) tmem_ld16_nowait(trow + BN, r2[0]);
                    tmem_ld_wait();
#pragma unroll
                    for (int pc = 0; pc < NCHUNK; pc++) {
                        const int cur = pc & 1;
                        if (pc + 1 < NCHUNK) {   // next 16 columns in flight while these are added
                            tmem_ld16_nowait(trow + (uint32_t)((pc + 1) * 16), r[cur ^ 1]);
                            if (PLANES == 2) tmem_ld16_nowait(trow + (uint32_t)(BN + (pc + 1) * 16), r2[cur ^ 1]);
                        }
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float tv = __uint_as_float(r[cur][j]);
                            if (PLANES == 2) tv = __fadd_rn(tv, __uint_as_float(r2[cur][j]));
                            accv[pc * 16 + j] = (c == 0) ? tv : __fadd_rn(accv[pc * 16 + j], tv);
                        }
                        if (pc + 1 < NCHUNK) tmem_ld_wait();
                    }
                }
                __syncwarp();
                tc_fence_before();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);   // 8 epilogue warps -> accumulator free for chunk ci+2
            }
            // ---- bias, ReLU, re-split, store ----
            {
            if (active_half && valid && !(a.dbg & 2)) {
#pragma unroll
                for (int pc = 0; pc < NCHUNK; pc++) {
                    const int cb = n0 + half * HALF + pc * 16;   // first output channel of this piece
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float tv = __fmaf_rn(accv[pc * 16 + j], out_scale, s_bias[cb + j]);   // out_scale is a power of two: exact
                        if (a.relu) tv = fmaxf(tv, 0.f);
                        v[j] = tv;
                        if (cb + j < a.cout) range_max = fmaxf(range_max, fabsf(tv));
                    }
                    if (a.planar) {
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            if (cb + j < a.cout) a.planar[(((size_t)n * a.planar_C + a.planar_coff + cb + j) * a.H + y) * a.W + x] = v[j];
                    } else {
                        uint32_t pk[PLANES][8];
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            float r0 = v[j], r1 = v[j + 1];
#pragma unroll
                            for (int p = 0; p < PLANES; p++) pk[p][j / 2] = split_pair<F16>(r0, r1);   // one packed conversion
                        }
                        __nv_bfloat16* orow = a.out + (size_t)m * a.out_pitch + a.out_coff + cb;
#pragma unroll
                        for (int p = 0; p < PLANES; p++) {
                            uint4* dst = (uint4*)(orow + (size_t)p * a.out_plane);
                            if (cb < cout8) dst[0] = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
                            if (cb + 8 < cout8) dst[1] = make_uint4(pk[p][4], pk[p][5], pk[p][6], pk[p][7]);
                        }
                    }
                }
            }
            }
        }
        range_publish(a.range, range_max);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster (one TPC) compute ONE 256-row tile.  Each CTA stages its own
// 128(+k-1)-row A window and only HALF of every weight tile (rows [rank*BN/2, rank*BN/2 + BN/2) of B_hi and of B_lo):
// the tensor core of each SM reads the other half from its partner.  Per SM this halves the weight bytes pulled from
// L2 and held in shared memory (16 KB per tap instead of 32 KB at BN = 128 -> 7 B slots + 3 A windows) and it halves
// the B bytes each MMA reads from shared memory, which is what bounded the single-CTA kernel (M=128 x N=128 reads
// 8 KB per 64 cycles).  With that bottleneck gone the three partial products are three plain M=256 x N=BN MMAs:
//   H += A_hi x B_hi   (large terms; chunked, see below)      C += A_hi x B_lo      C += A_lo x B_hi
// TMEM (per CTA: its 128 rows): H0 | H1 | C0 | C1, BN columns each.  The hi*hi chain is cut into chunks that alternate
// between H0 and H1 and are summed in registers with round-to-nearest by the epilogue warps (tensor-core accumulation
// truncates); the cross terms are 2^-11 smaller, so their chain runs over the whole tile in C[tile & 1] and is drained
// once.  Only the leader CTA (cluster rank 0) issues MMAs; both CTAs run a TMA producer and 8 epilogue warps.
// Barriers: a_full/b_full live in the leader (the peer's TMA completes its bytes there, cp.async.bulk.tensor
// .cta_group::2), a_empty/b_empty/h_full/c_full exist in both CTAs and are signalled by multicast tcgen05.commit,
// h_empty/c_empty live in the leader and collect the 16 epilogue warps of the pair.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {   // warps arrive reconverged; the non-.aligned forms tolerate stragglers anyway
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrive on a barrier of the pair's leader.  RELAXED: what the arrival publishes is "my tcgen05.ld of this TMEM buffer has completed"
// (tcgen05.wait::ld + tcgen05.fence::before_thread_sync order that), no global or shared data.  With .release.cluster the compiler
// emitted MEMBAR.ALL.GPU + ERRBAR in front of every arrive - 17 % of the epilogue warps' stall samples (ncu r2k) on the path that
// frees the accumulator for the MMA warp.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar) {   // arrives on this barrier offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 16-byte store to a shared-memory address (explicit state space: through a pointer rebuilt from an integer the compiler emits generic ST.E)
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int TCP_BM = 256;   // rows per pair tile

// ST2: both planes of a tile are staged at once (64 KB at BN = 128, fewer weight slots).
// ALT (BN = 64, short tiles: conv1_1 / conv1_2): the two epilogue warp groups take alternate tiles, each with its own pair of
// hi*hi accumulators (4 + 2 accumulators = 384 TMEM columns), instead of half the columns of every tile.
template <int BN, int NA, int NB, int ST2, int ALT_>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TCW_THREADS, 1)
conv_tcp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                const TcArgs a) {
    constexpr int A_SLOT = 2 * TCW_A_BYTES;                 // hi + lo window of this CTA's 128 rows
    constexpr int B_HALF = (BN / 2) * 128;                  // this CTA's rows of one plane of one tap
    constexpr int B_TAP = 2 * B_HALF;                       // [B_hi half ; B_lo half]
    constexpr bool ALT = ALT_ != 0;
    static_assert(!ALT || BN == 64, "alternating epilogue groups: 64-wide tiles only");
    constexpr int NH = ALT ? 4 : 2;                         // hi*hi accumulators
    constexpr int TCOLS = (NH + 2) * BN;
    constexpr int TMEM_COLS = TCOLS <= 32 ? 32 : (TCOLS <= 64 ? 64 : (TCOLS <= 128 ? 128 : (TCOLS <= 256 ? 256 : 512)));
    constexpr uint32_t IDESC = umma_idesc(TCP_BM, BN, true);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + NA * A_SLOT;
    uint8_t* smem_o = smem_b + NB * B_TAP;                  // epilogue staging: 8 warps x (32 rows x HALF channels x 2 B), see tma_store
    __shared__ __align__(8) uint64_t a_full[NA], a_empty[NA], b_full[NB], b_empty[NB];
    __shared__ __align__(8) uint64_t h_full[NH], h_empty[NH], c_full[2], c_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_bias[512];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int ks = a.ksize;
    for (int i = threadIdx.x; i < 512; i += TCW_THREADS) s_bias[i] = (i < a.cout) ? a.bias[i] : 0.f;
    const int n_tiles_n = a.n_tiles_n;
    const long long total_tiles = a.total_tiles;
    const long long tile0 = blockIdx.x >> 1, tile_stride = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NB; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        // h_empty / c_empty collect the epilogue warps of BOTH CTAs that work on a tile: 16, or 8 when the warp groups alternate tiles (ALT)
        for (int s = 0; s < NH; s++) { mbar_init(&h_full[s], 1); mbar_init(&h_empty[s], ALT ? 8 : 16); }
        for (int s = 0; s < 2; s++) { mbar_init(&c_full[s], 1); mbar_init(&c_empty[s], ALT ? 8 : 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {   // the same warp of both CTAs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0 && tile0 < total_tiles) {   // weights do not depend on the previous kernel: warm L2 with the first taps
        const int n0 = (int)(tile0 % n_tiles_n) * BN + (int)rank * (BN / 2);
        for (int q = 0; q < (ks < 4 ? ks : 4); q++) tma_prefetch_3d(&tmB, q * a.cin_k, n0, 0);
    }
    tc_fence_before();
    cluster_sync_all();      // barrier inits and the TMEM allocation of both CTAs are visible to the pair
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tm_h = tmem_base, tm_c = tmem_base + NH * BN;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer (both CTAs): own A window, own half of every weight tile; bytes land on the leader's barriers =====
        int aw = 0, bt = 0;
        for (long long t = tile0; t < total_tiles; t += tile_stride) {
            const long long m0 = (t / n_tiles_n) * TCP_BM + (long long)rank * TC_BM;
            const int n0 = (int)(t % n_tiles_n) * BN + (int)rank * (BN / 2);
            for (int kb = 0; kb < a.kblocks_per_tap; kb++)
                for (int r = 0; r < ks; r++) {
                    {
                        const int s = aw % NA;
                        mbar_wait(&a_empty[s], ((uint32_t)(aw / NA) & 1u) ^ 1u);
                        const int row0 = (int)(m0 + (long long)(r - a.pad) * a.Wp - a.pad);
                        if ((a.dbg & 8) && aw >= NA) { if (rank == 0) mbar_arrive(&a_full[s]); }
                        else {
                            if (rank == 0) mbar_expect_tx(&a_full[s], a.a_hi_only ? 2 * TCW_A_BYTES : 2 * A_SLOT);   // tmA's box holds one plane then
                            tma_load_3d_2sm(smem_a + s * A_SLOT, &tmA, mapa_u32(smem_u32(&a_full[s]), 0), kb * TC_BK, row0, 0);
                        }
                        aw++;
                    }
                    for (int q = 0; q < ks; q++) {
                        const int s = bt % NB;
                        mbar_wait(&b_empty[s], ((uint32_t)(bt / NB) & 1u) ^ 1u);
                        const int tap = r * ks + q;
                        if ((a.dbg & 4) && bt >= NB) { if (rank == 0) mbar_arrive(&b_full[s]); }
                        else {
                            if (rank == 0) mbar_expect_tx(&b_full[s], 2 * B_TAP);
                            tma_load_3d_2sm(smem_b + s * B_TAP, &tmB, mapa_u32(smem_u32(&b_full[s]), 0), tap * a.cin_k + kb * TC_BK, n0, 0);   // box {64, BN/2, 2}
                        }
                        bt++;
                    }
                }
        }
    } else if (warp == 1 && rank == 0 && (((a.dbg & 16) ? 1u : 0xffffffffu) >> lane & 1u)) {
        const uint32_t wmask = (a.dbg & 16) ? 1u : 0xffffffffu;   // PE_TC_DBG bit 4: only lane 0 runs the issue loop (experiment)
        // ===== MMA issuer (leader CTA only): converged warp, one elected lane issues =====
        int aw = 0, bt = 0;
        uint32_t ci = 0, ti = 0;
        const int nsteps = a.kblocks_per_tap * ks, cs = a.chunk_steps;
        const uint32_t nchunks = (uint32_t)((nsteps + cs - 1) / cs);
        const uint32_t sa_base = smem_u32(smem_a), sb_base = smem_u32(smem_b);
        for (long long t = tile0; t < total_tiles; t += tile_stride, ti++) {
            const uint32_t cbuf = ti & 1u;
            mbar_wait(&c_empty[cbuf], ((ti >> 1) & 1u) ^ 1u);       // both CTAs' epilogues have drained this cross accumulator
            const uint32_t accC = tm_c + cbuf * BN;
            uint32_t accH = tm_h;
            int step = 0, hs = 0;
            uint32_t firstH = 0, firstC = 0;   // accumulate flags of the next H / C MMA
            for (int kb = 0; kb < a.kblocks_per_tap; kb++)
                for (int r = 0; r < ks; r++) {
                    if (step % cs == 0) {
                        // accumulator and its use count: two that all tiles share, or (ALT) two per warp group (= tile parity)
                        const uint32_t c = (uint32_t)(step / cs);
                        const uint32_t use = ALT ? (ti >> 1) * ((nchunks + 1u - (c & 1u)) >> 1) + (c >> 1) : (ci >> 1);
                        hs = ALT ? (int)((ti & 1u) * 2u + (c & 1u)) : (int)(ci & 1u);
                        mbar_wait(&h_empty[hs], (use & 1u) ^ 1u);
                        accH = tm_h + (uint32_t)(hs * BN);
                        firstH = 0;
                    }
                    const int sa_slot = aw % NA;
                    mbar_wait(&a_full[sa_slot], (uint32_t)(aw / NA) & 1u);
                    const uint64_t dA0 = umma_desc(sa_base + sa_slot * A_SLOT), dA1 = dA0 + (uint64_t)(TCW_A_BYTES >> 4);
                    for (int q = 0; q < ks; q++) {
                        const int sb_slot = bt % NB;
                        mbar_wait(&b_full[sb_slot], (uint32_t)(bt / NB) & 1u);
                        tc_fence_after();
                        const uint64_t dBh = umma_desc(sb_base + sb_slot * B_TAP), dBl = dBh + (uint64_t)(B_HALF >> 4);
                        if (elect_one(wmask)) {
#pragma unroll
                            for (int k = 0; k < TC_BK / 16; k++) {
                                const uint64_t ao = (uint64_t)(q * 8 + k * 2), bo = (uint64_t)(k * 2);
                                umma_f16_2sm(accH, dA0 + ao, dBh + bo, IDESC, k ? 1u : firstH);
                                umma_f16_2sm(accC, dA0 + ao, dBl + bo, IDESC, k ? 1u : firstC);
                                if (!a.a_hi_only) umma_f16_2sm(accC, dA1 + ao, dBh + bo, IDESC, 1u);
                            }
                            umma_commit_mc(&b_empty[sb_slot]);
                        }
                        __syncwarp(wmask);
                        firstH = 1; firstC = 1;
                        bt++;
                    }
                    step++;
                    const bool chunk_end = step % cs == 0 || step == nsteps;
                    if (elect_one(wmask)) {
                        umma_commit_mc(&a_empty[sa_slot]);
                        if (chunk_end) umma_commit_mc(&h_full[hs]);
                        if (step == nsteps) umma_commit_mc(&c_full[cbuf]);
                    }
                    __syncwarp(wmask);
                    if (chunk_end) ci++;
                    aw++;
                }
        }
    } else if (warp >= 2) {
        // ===== epilogue (both CTAs): 8 warps; warp w owns TMEM lanes 32*(w%4).. (this CTA's rows) and column half (w-2)/4 =====
        // ALT (BN = 64: conv1_1 / conv1_2, 8500 short tiles each): the two warp groups take ALTERNATE tiles instead of half the columns
        // of every tile.  A 64-wide tile's MMAs last 0.2 - 2 us but its epilogue (drain, convert, stage, store) ~3 us of latency, so
        // with all 8 warps on one tile the kernel ran at the epilogue's latency; two tiles in flight double the rate.
        constexpr bool SPLIT = !ALT && (BN / 2) % 16 == 0;
        constexpr int HALF = SPLIT ? BN / 2 : BN;
        constexpr int NCHUNK = HALF / 16;
        const int quad = warp & 3;
        const int group = (warp - 2) / 4;
        const int half = SPLIT ? group : 0;
        const bool active_half = SPLIT || ALT || group == 0;
        const int per_img = a.Hs * a.Wp;
        const int cout8 = (a.cout + 7) & ~7;
        const float out_scale = __ldg(a.out_scale);
        const uint32_t h_empty_leader = mapa_u32(smem_u32(&h_empty[0]), 0), c_empty_leader = mapa_u32(smem_u32(&c_empty[0]), 0);
        uint32_t ci = 0, ti = 0;
        float range_max = 0.f;
        const int nsteps = a.kblocks_per_tap * ks;
        const int nchunks = (nsteps + a.chunk_steps - 1) / a.chunk_steps;
        for (long long t = tile0; t < total_tiles; t += tile_stride, ti++) {
            if (ALT && (int)(ti & 1u) != group) continue;     // the other group's tile
            const long long m0 = (t / n_tiles_n) * TCP_BM + (long long)rank * TC_BM;
            const int n0 = (int)(t % n_tiles_n) * BN;
            const long long m = m0 + quad * 32 + lane;
            bool valid = m < a.M;
            int n = 0, y = 0, x = 0;
            if (valid) {
                n = (int)(m / per_img);
                const int rem = (int)(m % per_img);
                y = rem / a.Wp; x = rem % a.Wp;
                valid = (x < a.W) && (y < a.H);
            }
            float accv[HALF];
            // ---- hi*hi chunks, summed in registers with round-to-nearest ----
            for (int c = 0; c < nchunks; c++, ci++) {
                const uint32_t use = ALT ? (ti >> 1) * (uint32_t)((nchunks + 1 - (c & 1)) >> 1) + (uint32_t)(c >> 1) : (ci >> 1);
                const int hs = ALT ? group * 2 + (c & 1) : (int)(ci & 1u);
                mbar_wait(&h_full[hs], use & 1u);
                tc_fence_after();
                const uint32_t trow = tm_h + ((uint32_t)(quad * 32) << 16) + (uint32_t)(hs * BN + half * HALF);
                if (active_half && !((a.dbg & 1) && c > 0)) {
                    uint32_t r[2][16];
                    __syncwarp();
                    tmem_ld16_nowait(trow, r[0]);
                    tmem_ld_wait();
#pragma unroll
                    for (int pc = 0; pc < NCHUNK; pc++) {
                        const int cur = pc & 1;
                        if (pc + 1 < NCHUNK) tmem_ld16_nowait(trow + (uint32_t)((pc + 1) * 16), r[cur ^ 1]);
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const float tv = __uint_as_float(r[cur][j]);
                            accv[pc * 16 + j] = (c == 0) ? tv : __fadd_rn(accv[pc * 16 + j], tv);
                        }
                        if (pc + 1 < NCHUNK) tmem_ld_wait();
                    }
                }
                __syncwarp();
                tc_fence_before();
                if (lane == 0) mbar_arrive_cluster(h_empty_leader + (uint32_t)(hs * 8));
            }
            // ---- cross terms of the whole tile, then bias / ReLU / re-split / store piece by piece ----
            // Stores: a lane owns one output row, i.e. 2*HALF contiguous bytes per plane; written directly, every 16-byte
            // store instruction of a warp touches 32 different 128-byte lines and the LSU serialises them (measured with
            // PE_TC_DBG=2: the store phase cost 16 % of the conv time and sat on the critical path of the next tile's
            // chunk drains).  With a.tma_store the warp instead stages its 32 x HALF box of one plane in shared memory
            // (SWIZZLE_128B chunk order when a row is 128 bytes: conflict-free 16-byte STS) and one lane hands it to the
            // TMA store engine; the lo plane waits in registers until the engine has read the hi plane.
            {
                constexpr int RB = HALF * 2;                        // bytes of one staged row
                constexpr bool TS_OK = (SPLIT || ALT) && (HALF == 64 || HALF == 32);
                const bool ts = TS_OK && a.tma_store;
                uint8_t* stage = smem_o + (warp - 2) * (32 * RB) * (ST2 ? 2 : 1);
                const uint32_t stage_s = smem_u32(stage);
                const uint32_t cbuf = ti & 1u;
                mbar_wait(&c_full[cbuf], (ti >> 1) & 1u);
                tc_fence_after();
                const uint32_t trow = tm_c + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cbuf * BN + half * HALF);
                if (active_half) {
                    uint32_t r[2][16];
                    uint32_t plo[(TS_OK && !ST2) ? NCHUNK : 1][8];
                    __syncwarp();
                    tmem_ld16_nowait(trow, r[0]);
                    if (ts) { if (lane == 0) bulk_wait_read(); }      // the engine has read what the previous tile staged
                    tmem_ld_wait();
#pragma unroll
                    for (int pc = 0; pc < NCHUNK; pc++) {
                        const int cur = pc & 1;
                        if (pc + 1 < NCHUNK) { __syncwarp(); tmem_ld16_nowait(trow + (uint32_t)((pc + 1) * 16), r[cur ^ 1]); }
                        const int cb = n0 + half * HALF + pc * 16;   // first output channel of this piece
                        if ((valid || ts) && !(a.dbg & 2)) {
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                float tv = __fadd_rn(accv[pc * 16 + j], __uint_as_float(r[cur][j]));
                                tv = __fmaf_rn(tv, out_scale, s_bias[cb + j]);   // out_scale is a power of two: exact
                                if (a.relu) tv = fmaxf(tv, 0.f);
                                v[j] = valid ? tv : 0.f;                        // gap rows are written as the zeros they hold
                                if (cb + j < a.cout) range_max = fmaxf(range_max, fabsf(v[j]));
                            }
                            if (a.planar) {
#pragma unroll
                                for (int j = 0; j < 16; j++)
                                    if (cb + j < a.cout) a.planar[(((size_t)n * a.planar_C + a.planar_coff + cb + j) * a.H + y) * a.W + x] = v[j];
                            } else {
                                uint32_t pk[2][8];
#pragma unroll
                                for (int j = 0; j < 16; j += 2) {
                                    float r0 = v[j], r1 = v[j + 1];
#pragma unroll
                                    for (int p = 0; p < 2; p++) pk[p][j / 2] = split_pair<true>(r0, r1);
                                }
                                if (TS_OK && ts) {
#pragma unroll
                                    for (int h2 = 0; h2 < 2; h2++) {
                                        const int chunk = pc * 2 + h2;
                                        const int phys = RB == 128 ? (chunk ^ (lane & 7)) : chunk;
                                        sts128(stage_s + lane * RB + phys * 16, pk[0][h2 * 4], pk[0][h2 * 4 + 1], pk[0][h2 * 4 + 2], pk[0][h2 * 4 + 3]);
                                    }
                                    if (ST2) {
#pragma unroll
                                        for (int h2 = 0; h2 < 2; h2++) {
                                            const int chunk = pc * 2 + h2;
                                            const int phys = RB == 128 ? (chunk ^ (lane & 7)) : chunk;
                                            sts128(stage_s + 32 * RB + lane * RB + phys * 16, pk[1][h2 * 4], pk[1][h2 * 4 + 1], pk[1][h2 * 4 + 2], pk[1][h2 * 4 + 3]);
                                        }
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 8; j++) plo[(TS_OK && !ST2) ? pc : 0][j] = pk[1][j];
                                    }
                                } else {
                                    __nv_bfloat16* orow = a.out + (size_t)m * a.out_pitch + a.out_coff + cb;
#pragma unroll
                                    for (int p = 0; p < 2; p++) {
                                        uint4* dst = (uint4*)(orow + (size_t)p * a.out_plane);
                                        if (cb < cout8) dst[0] = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
                                        if (cb + 8 < cout8) dst[1] = make_uint4(pk[p][4], pk[p][5], pk[p][6], pk[p][7]);
                                    }
                                }
                            }
                        }
                        if (pc + 1 < NCHUNK) { __syncwarp(); tmem_ld_wait(); }
                    }
                    // the cross accumulator is in registers: release it before the stores
                    __syncwarp();
                    tc_fence_before();
                    if (lane == 0) mbar_arrive_cluster(c_empty_leader + (uint32_t)(cbuf * 8));
                    if (TS_OK && ts && !(a.dbg & 2)) {
                        const int oc0 = a.out_coff + n0 + half * HALF, orow0 = (int)(m0 + quad * 32);
                        fence_async_smem();
                        __syncwarp();
                        if (ST2) {
                            if (lane == 0 && !(a.dbg & 32)) { tma_store_3d(&tmO, stage, oc0, orow0, 0); tma_store_3d(&tmO, stage + 32 * RB, oc0, orow0, 1); }
                        } else {
                            if (lane == 0 && !(a.dbg & 32)) { tma_store_3d(&tmO, stage, oc0, orow0, 0); bulk_wait_read(); }
                            __syncwarp();
#pragma unroll
                            for (int pc = 0; pc < NCHUNK; pc++)
#pragma unroll
                                for (int h2 = 0; h2 < 2; h2++) {
                                    const int chunk = pc * 2 + h2;
                                    const int phys = RB == 128 ? (chunk ^ (lane & 7)) : chunk;
                                    constexpr bool PL = TS_OK && !ST2;
                                    sts128(stage_s + lane * RB + phys * 16, plo[PL ? pc : 0][h2 * 4], plo[PL ? pc : 0][h2 * 4 + 1], plo[PL ? pc : 0][h2 * 4 + 2],
                                           plo[PL ? pc : 0][h2 * 4 + 3]);
                                }
                            fence_async_smem();
                            __syncwarp();
                            if (lane == 0 && !(a.dbg & 32)) tma_store_3d(&tmO, stage, oc0, orow0, 1);
                        }
                    }
                } else {
                    __syncwarp();
                    tc_fence_before();
                    if (lane == 0) mbar_arrive_cluster(c_empty_leader + (uint32_t)(cbuf * 8));
                }
            }
        }
        if (lane == 0) bulk_wait_read();   // staged boxes have been read out of shared memory before the CTA may exit
        range_publish(a.range, range_max);
    }
    tc_fence_before();
    cluster_sync_all();      // no CTA of the pair exits (or frees TMEM) while its partner can still signal it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int tc_cout_pad(int cout) {
    if (cout >= 128) return (cout + 127) / 128 * 128;
    if (cout > 48) return 64;
    if (cout > 32) return 48;
    if (cout > 16) return 32;
    return 16;
}
static int tc_bn(int cout_pad) { return cout_pad >= 128 ? 128 : cout_pad; }

static int stage_bytes(int bn, int planes) { return planes * (TC_BM * 128 + bn * 128); }
static int pick_stages(int bn, int planes) {
    const int s = (200 * 1024) / stage_bytes(bn, planes);
    return s >= 6 ? 6 : (s >= 4 ? 4 : (s >= 3 ? 3 : 2));
}

template <int BN, int PLANES, int STAGES>
static int launch_inst(const TcLayer& l, const TcArgs& a, dim3 grid, cudaStream_t st) {
    auto kern = conv_tc_kernel<BN, PLANES, STAGES>;
    const int smem = STAGES * stage_bytes(BN, PLANES) + 1024;
    static std::atomic<unsigned long long> attr_done{0};   // bit d: attribute set on device d (it is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_done.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    const CUtensorMap* maps = (const CUtensorMap*)l.maps;
    kern<<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], a);
    return 1;
}

template <int BN>
static int launch_bn(const TcLayer& l, const TcArgs& a, dim3 grid, cudaStream_t st) {
    switch (l.d.planes) {
        case 1: return launch_inst<BN, 1, (200 * 1024) / (1 * (TC_BM * 128 + BN * 128)) >= 6 ? 6 : 4>(l, a, grid, st);
        case 2: return launch_inst<BN, 2, (200 * 1024) / (2 * (TC_BM * 128 + BN * 128)) >= 4 ? 4 : 3>(l, a, grid, st);
        default: return launch_inst<BN, 3, (200 * 1024) / (3 * (TC_BM * 128 + BN * 128)) >= 3 ? 3 : 2>(l, a, grid, st);
    }
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Launch with the programmatic-stream-serialization attribute (PDL, see pdl_wait): back-to-back conv kernels overlap the
// next one's prologue with the previous one's tail.  PE_TC_PDL=0 launches plainly.
template <typename Kern, typename... Args>
static int launch_pdl(Kern kern, dim3 grid, int smem, cudaStream_t st, const Args&... args) {
    static const int pdl = env_int("PE_TC_PDL", 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(TCW_THREADS, 1, 1); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, args...);
    return 1;
}

template <int BN, int PLANES, int NA, int NB, int ROWB>
static int launch_win_inst(const TcLayer& l, const TcArgs& a, dim3 grid, cudaStream_t st, int bmap) {
    auto kern = conv_tcw_kernel<BN, PLANES, NA, NB, ROWB>;
    const int smem = NA * PLANES * TCW_A_BYTES + NB * (ROWB ? 3 : 1) * PLANES * BN * 128 + 1024;
    static std::atomic<unsigned long long> attr_done{0};   // bit d: attribute set on device d (it is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_done.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    const CUtensorMap* maps = (const CUtensorMap*)l.maps;
    return launch_pdl(kern, grid, smem, st, maps[2], maps[bmap], a);
}
template <int BN, int ST2, int ALT>
static int launch_pair_st(const TcLayer& l, const TcArgs& a, dim3 grid, cudaStream_t st, int bmap) {
    constexpr int NA = 2;
    constexpr int B_TAP = BN * 128;
    // epilogue staging for TMA stores: 8 warps x 32 rows x (BN/2 channels, or all 64 when the warp groups alternate tiles) x 2 B (x 2 planes)
    constexpr int STAGE = (BN / 2) % 16 == 0 ? 8 * 32 * BN * (ALT ? 2 : 1) * (ST2 ? 2 : 1) : 0;
    constexpr int BUDGET = 227 * 1024 - 1024 - 3072 - NA * 2 * TCW_A_BYTES - STAGE;
    constexpr int NB = BUDGET / B_TAP >= 12 ? 12 : BUDGET / B_TAP;   // BN = 128: 7 slots (5 with both planes staged)
    auto kern = conv_tcp_kernel<BN, NA, NB, ST2, ALT>;
    const int smem = NA * 2 * TCW_A_BYTES + NB * B_TAP + STAGE + 1024;
    static std::atomic<unsigned long long> attr_done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_done.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    const CUtensorMap* maps = (const CUtensorMap*)l.maps;
    return launch_pdl(kern, grid, smem, st, maps[a.a_hi_only ? 11 : 2], maps[bmap], maps[(BN == 128 || ALT) ? 9 : 10], a);   // 64-channel boxes (ALT: one warp stores all 64 columns of its rows)
}
template <int BN>
static int launch_pair(const TcLayer& l, const TcArgs& a, dim3 grid, cudaStream_t st, int bmap) {
    static const int st2 = env_int("PE_TC_ST2", 1);   // both planes staged at once (r2f: 814 vs 794 frames/s with 7 -> 5 weight slots)
    // alternating epilogue groups for 64-wide tiles with few accumulation chunks (conv1_1, conv1_2): their epilogue latency, not the MMAs, set the pace
    static const int alt = env_int("PE_TC_ALT", 1);
    if (BN == 64 && alt) {
        const int nsteps = a.kblocks_per_tap * a.ksize;
        if ((nsteps + a.chunk_steps - 1) / a.chunk_steps <= 4) return launch_pair_st<BN, 1, BN == 64 ? 1 : 0>(l, a, grid, st, bmap);
    }
    if (st2 && (BN == 128 || BN == 64)) return launch_pair_st<BN, (BN == 128 || BN == 64) ? 1 : 0, 0>(l, a, grid, st, bmap);
    return launch_pair_st<BN, 0, 0>(l, a, grid, st, bmap);
}

template <int BN>
static int launch_win_bn(const TcLayer& l, const TcArgs& a, dim3 grid, cudaStream_t st, int bmap) {
    if (BN <= 64 && l.d.ksize <= 3) {   // narrow-N layers (conv1_x, the 1x1 heads): filter-row B slots
        if (l.d.planes == 1) return launch_win_inst<(BN <= 64 ? BN : 64), 1, 3, 4, 1>(l, a, grid, st, bmap);
        return launch_win_inst<(BN <= 64 ? BN : 64), 2, 2, 3, 1>(l, a, grid, st, bmap);
    }
    if (l.d.planes == 1) return launch_win_inst<BN, 1, 3, (BN >= 128 ? 8 : 10), 0>(l, a, grid, st, bmap);
    return launch_win_inst<BN, 2, 2, (BN >= 128 ? 4 : 6), 0>(l, a, grid, st, bmap);
}


int tc_layer_create(const TcLayerDesc& d, TcLayer& out, std::string& err) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { err = "cuTensorMapEncodeTiled unavailable"; return -1; }
    if (d.in_cused % TC_BK) { err = "input channels not a multiple of 64"; return -1; }
    out.d = d;
    out.bn = tc_bn(d.cout_pad);
    CUtensorMap* maps = nullptr;
    if (posix_memalign((void**)&maps, 64, 13 * sizeof(CUtensorMap))) { err = "alloc"; return -1; }
    const int taps = d.ksize * d.ksize;
    const cuuint64_t K = (cuuint64_t)taps * d.in_cused;
    {   // A: [planes][M][pitch] bf16, box {64, 128, 1}
        cuuint64_t dims[3] = {(cuuint64_t)d.in_cused, (cuuint64_t)d.geo.M, (cuuint64_t)d.planes};
        cuuint64_t strides[2] = {(cuuint64_t)d.in_pitch * 2, (cuuint64_t)d.in_plane * 2};
        cuuint32_t box[3] = {TC_BK, TC_BM, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d.in, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r); free(maps); return -1; }
    }
    {   // A window: same tensor, box {64, 136, 1} (128 + k - 1 rows serve the k taps of one filter row)
        cuuint64_t dims[3] = {(cuuint64_t)d.in_cused, (cuuint64_t)d.geo.M, (cuuint64_t)d.planes};
        cuuint64_t strides[2] = {(cuuint64_t)d.in_pitch * 2, (cuuint64_t)d.in_plane * 2};
        cuuint32_t box[3] = {TC_BK, TCW_ROWS, (cuuint32_t)(d.planes <= 2 ? d.planes : 1)};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d.in, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(A window) failed: " + std::to_string((int)r); free(maps); return -1; }
    }
    {   // A window, hi plane only (TcArgs::a_hi_only): box {64, 136, 1}
        cuuint64_t dims[3] = {(cuuint64_t)d.in_cused, (cuuint64_t)d.geo.M, (cuuint64_t)d.planes};
        cuuint64_t strides[2] = {(cuuint64_t)d.in_pitch * 2, (cuuint64_t)d.in_plane * 2};
        cuuint32_t box[3] = {TC_BK, TCW_ROWS, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps[11], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d.in, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(A window, one plane) failed: " + std::to_string((int)r); free(maps); return -1; }
    }
    {   // B: [planes][cout_pad][K] bf16, box {64, BN, 1}
        cuuint64_t dims[3] = {K, (cuuint64_t)d.cout_pad, (cuuint64_t)d.planes};
        cuuint64_t strides[2] = {K * 2, K * 2 * (cuuint64_t)d.cout_pad};
        cuuint32_t box[3] = {TC_BK, (cuuint32_t)out.bn, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d.w, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r); free(maps); return -1; }
    }
    {   // B for the window kernel: box {64, BN, planes} -> one TMA op brings [B_hi ; B_lo]
        cuuint64_t dims[3] = {K, (cuuint64_t)d.cout_pad, (cuuint64_t)d.planes};
        cuuint64_t strides[2] = {K * 2, K * 2 * (cuuint64_t)d.cout_pad};
        cuuint32_t box[3] = {TC_BK, (cuuint32_t)out.bn, (cuuint32_t)(d.planes <= 2 ? d.planes : 1)};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps[3], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d.w, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(B window) failed: " + std::to_string((int)r); free(maps); return -1; }
    }
    // narrower N tiles (64, 32) of the same weight matrix for small problems (one frame per forward): more CTAs
    for (int v = 0; v < 2 && out.bn == 128; v++) {
        cuuint64_t dims[3] = {K, (cuuint64_t)d.cout_pad, (cuuint64_t)d.planes};
        cuuint64_t strides[2] = {K * 2, K * 2 * (cuuint64_t)d.cout_pad};
        cuuint32_t box[3] = {TC_BK, (cuuint32_t)(v == 0 ? 64 : 32), (cuuint32_t)(d.planes <= 2 ? d.planes : 1)};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps[4 + v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d.w, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(B narrow) failed: " + std::to_string((int)r); free(maps); return -1; }
    }
    // pair kernel (cta_group::2): each CTA of the pair loads HALF of the N rows of both planes, box {64, bn/2, 2}
    for (int v = 0; v < 3 && d.planes == 2; v++) {
        const int bn = v == 0 ? out.bn : (v == 1 ? 64 : 32);
        if (v > 0 && out.bn != 128) break;
        cuuint64_t dims[3] = {K, (cuuint64_t)d.cout_pad, (cuuint64_t)d.planes};
        cuuint64_t strides[2] = {K * 2, K * 2 * (cuuint64_t)d.cout_pad};
        cuuint32_t box[3] = {TC_BK, (cuuint32_t)(bn / 2), 2};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps[6 + v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d.w, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(B pair) failed: " + std::to_string((int)r); free(maps); return -1; }
    }
    // pair kernel epilogue: TMA stores of 32-row boxes, 64 channels wide (128-byte rows, SWIZZLE_128B) or 32 wide (64-byte rows, linear)
    memset(&maps[9], 0, 2 * sizeof(CUtensorMap));
    for (int v = 0; v < 2 && d.planes == 2 && d.out; v++) {
        const int wch = v == 0 ? 64 : 32;
        if (d.out_pitch < wch) continue;
        cuuint64_t dims[3] = {(cuuint64_t)d.out_pitch, (cuuint64_t)d.geo.M, (cuuint64_t)d.planes};
        cuuint64_t strides[2] = {(cuuint64_t)d.out_pitch * 2, (cuuint64_t)d.out_plane * 2};
        cuuint32_t box[3] = {(cuuint32_t)wch, 32, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&maps[9 + v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d.out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         v == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled(out) failed: " + std::to_string((int)r); free(maps); return -1; }
    }
    out.maps = maps;
    out.stages = pick_stages(out.bn, d.planes);
    out.smem_bytes = out.stages * stage_bytes(out.bn, d.planes) + 1024;
    return 0;
}

void tc_layer_destroy(TcLayer& l) {
    if (l.maps) free(l.maps);
    l.maps = nullptr;
}

int tc_layer_launch(const TcLayer& l, int nimg, cudaStream_t st, int share, int a_hi_only) {
    const TcLayerDesc& d = l.d;
    TcArgs a;
    a.bias = d.bias;
    a.out = (__nv_bfloat16*)d.out; a.out_pitch = d.out_pitch; a.out_coff = d.out_coff; a.out_plane = d.out_plane;
    a.planar = d.planar; a.planar_C = d.planar_C; a.planar_coff = d.planar_coff;
    a.cout = d.cout; a.relu = d.relu; a.out_scale = d.out_scale; a.range = d.range;
    a.ksize = d.ksize; a.pad = d.pad; a.kblocks_per_tap = d.in_cused / TC_BK; a.cin_k = d.in_cused;
    a.W = d.geo.W; a.H = d.geo.H; a.Wp = d.geo.Wp; a.Hs = d.geo.Hs;
    a.M = (long long)nimg * d.geo.Hs * d.geo.Wp;
    a.chunk_steps = 1 << 30;
    static const int dbg = env_int("PE_TC_DBG", 0);
    a.dbg = dbg;
    a.tma_store = 0;
    static const int hi_only_env = env_int("PE_TC_HIONLY", 1);   // 0: always load both input planes (A/B measurement)
    a.a_hi_only = 0;
    dim3 grid((unsigned)((a.M + TC_BM - 1) / TC_BM), (unsigned)(d.cout_pad / l.bn));
    static const int variant = env_int("PE_TC_VARIANT", 1);     // 1: window kernel, 0: one TMA tile per tap
    if (variant == 1 && d.planes <= 2) {
        static int nsm = 0;
        if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
        // Tile width: 128 output channels per CTA is the efficient shape, but a single frame at the 46x82 level has
        // only 33 row tiles for 148 SMs.  Pick the width that maximises (CTAs that can run at once) x (relative
        // efficiency of that MMA shape: the A operand is re-read per N tile, so narrow tiles are smem-bound).
        static const int pair = env_int("PE_TC_PAIR", 1);   // CTA-pair kernel (cta_group::2) for the fp16-plane parity mode; 0: single-CTA window kernel
        if (pair && d.planes == 2 && planes_are_fp16(2) && nsm >= 2) {
            const long long mt = (a.M + TCP_BM - 1) / TCP_BM;
            const int npairs = nsm / 2;
            const int npairs_eff = std::max(1, npairs / (share < 1 ? 1 : share));   // pairs this layer can count on next to its sibling branch
            int bn = l.bn, bmap = 6;
            if (l.bn == 128) {   // same trade-off as below, in units of CTA pairs
                const double s128 = (double)std::min<long long>(mt * (d.cout_pad / 128), npairs_eff) * 1.00;
                const double s64 = (double)std::min<long long>(mt * (d.cout_pad / 64), npairs_eff) * 0.80;
                const double s32 = (double)std::min<long long>(mt * (d.cout_pad / 32), npairs_eff) * 0.55;
                if (s64 > s128 && s64 >= s32) { bn = 64; bmap = 7; }
                else if (s32 > s128 && s32 > s64) { bn = 32; bmap = 8; }
            }
            a.n_tiles_n = d.cout_pad / bn;
            a.total_tiles = mt * a.n_tiles_n;
            a.a_hi_only = a_hi_only && hi_only_env;
            const int nsteps = a.kblocks_per_tap * a.ksize;
            static const int chunk_env = env_int("PE_TC_CHUNK", -1);
            static const int chunk_mul = env_int("PE_TC_CHUNK_MUL", 1);
            int cs = (a.ksize >= 7 ? 1 : (a.ksize >= 3 ? 2 : 4)) * (chunk_mul < 1 ? 1 : chunk_mul);
            if (chunk_env == 0) cs = nsteps;
            else if (chunk_env > 0) cs = chunk_env;
            a.chunk_steps = cs < 1 ? 1 : (cs > nsteps ? nsteps : cs);
            static const int tstore = env_int("PE_TC_TMASTORE", 1);
            a.tma_store = tstore && d.out && (bn == 128 || bn == 64) && d.cout % bn == 0 && d.out_coff % 8 == 0;
            grid = dim3((unsigned)(2 * std::min<long long>(a.total_tiles, npairs)), 1, 1);
            switch (bn) {
                case 128: return launch_pair<128>(l, a, grid, st, bmap);
                case 64: return launch_pair<64>(l, a, grid, st, bmap);
                case 48: return launch_pair<48>(l, a, grid, st, bmap);
                case 32: return launch_pair<32>(l, a, grid, st, bmap);
                default: return launch_pair<16>(l, a, grid, st, bmap);
            }
        }
        int bn = l.bn, bmap = 3;
        if (l.bn == 128) {
            static const int narrow = env_int("PE_TC_NARROW", 1);
            const long long mt = grid.x;
            const double s128 = (double)std::min<long long>(mt * (d.cout_pad / 128), nsm) * 1.00;
            const double s64 = (double)std::min<long long>(mt * (d.cout_pad / 64), nsm) * 0.80;
            const double s32 = (double)std::min<long long>(mt * (d.cout_pad / 32), nsm) * 0.55;
            if (narrow && s64 > s128 && s64 >= s32) { bn = 64; bmap = 4; }
            else if (narrow && s32 > s128 && s32 > s64) { bn = 32; bmap = 5; }
        }
        a.n_tiles_n = d.cout_pad / bn;
        a.total_tiles = (long long)grid.x * a.n_tiles_n;
        {   // accumulation chunk (see the MMA issuer): ~24-28 K=16 steps per chunk in the parity mode, whole tile otherwise
            static const int chunk_env = env_int("PE_TC_CHUNK", -1);   // -1: default, 0: one chunk per tile, n: n steps
            const int nsteps = a.kblocks_per_tap * a.ksize;
            int cs = nsteps;
            static const int chunk_mul = env_int("PE_TC_CHUNK_MUL", 1);   // experiments: longer chunks
            if (d.planes == 2) cs = (a.ksize >= 7 ? 1 : (a.ksize >= 3 ? 2 : 4)) * (chunk_mul < 1 ? 1 : chunk_mul);
            if (chunk_env == 0) cs = nsteps;
            else if (chunk_env > 0) cs = chunk_env;
            a.chunk_steps = cs < 1 ? 1 : (cs > nsteps ? nsteps : cs);
        }
        grid = dim3((unsigned)(a.total_tiles < nsm ? a.total_tiles : nsm), 1, 1);
        switch (bn) {
            case 128: return launch_win_bn<128>(l, a, grid, st, bmap);
            case 64: return launch_win_bn<64>(l, a, grid, st, bmap);
            case 48: return launch_win_bn<48>(l, a, grid, st, bmap);
            case 32: return launch_win_bn<32>(l, a, grid, st, bmap);
            default: return launch_win_bn<16>(l, a, grid, st, bmap);
        }
    }
    switch (l.bn) {
        case 128: return launch_bn<128>(l, a, grid, st);
        case 64: return launch_bn<64>(l, a, grid, st);
        case 48: return launch_bn<48>(l, a, grid, st);
        case 32: return launch_bn<32>(l, a, grid, st);
        default: return launch_bn<16>(l, a, grid, st);
    }
}

}  // namespace pe
