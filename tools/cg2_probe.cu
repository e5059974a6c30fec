// cg2_probe - smallest possible check of the cta_group::2 ("CTA pair") tcgen05 forms used by the pair conv kernel
// (csrc/conv_tc.cu, conv_tcp_kernel), run on the GPU box BEFORE the real kernel so that a wrong PTX form shows up
// as a readable FAIL line instead of a hung forward:
//   * tcgen05.alloc / dealloc .cta_group::2 issued by the same warp of both CTAs
//   * cp.async.bulk.tensor.3d.cta_group::2 from each CTA into its OWN shared memory, completing on the LEADER's mbarrier
//   * one tcgen05.mma.cta_group::2.kind::f16 chain (M=256, N in {256,128,64}): which CTA supplies which rows of B, and
//     which CTA's TMEM receives which rows of D
//   * tcgen05.commit ... multicast::cluster to the same barrier offset in both CTAs
//   * mbarrier.arrive.shared::cluster on the leader's barrier from the peer
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o cg2_probe tools/cg2_probe.cu     Run: timeout 60 ./cg2_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ int mbar_wait_bounded(uint64_t* bar, uint32_t parity, long long max_cycles) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > max_cycles) return 0;
    }
    return 1;
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int THREADS = 192;
struct Out { float d[256 * 256]; unsigned info[16]; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int N, Out* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;                 // 128 rows x 128 B
    uint8_t* smem_b = smem + 128 * 128;     // N/2 rows x 128 B
    __shared__ __align__(8) uint64_t full_bar, done_bar, peer_bar;
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const long long LIMIT = 200000000LL;    // ~0.1 s: a wrong form reports a timeout instead of hanging the box

    if (threadIdx.x == 0) {
        mbar_init(&full_bar, 1);
        mbar_init(&done_bar, 1);
        mbar_init(&peer_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // both CTAs, same warp id
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t full_leader = mapa(smem_u32(&full_bar), 0);
    const uint32_t peer_leader = mapa(smem_u32(&peer_bar), 0);
    if (threadIdx.x == 0) {
        out->info[rank * 8 + 0] = smem_u32(&full_bar);
        out->info[rank * 8 + 1] = full_leader;
        out->info[rank * 8 + 2] = tmem_base;
        out->info[rank * 8 + 3] = smem_u32(smem_a);
    }
    unsigned status = 0;
    if (warp == 0 && lane == 0) {
        // producer of BOTH CTAs: own halves into own smem, bytes counted on the leader's barrier
        const uint32_t bytes = 128 * 128 + (N / 2) * 128;
        if (rank == 0) mbar_expect_tx(&full_bar, 2 * bytes);
        tma_load_3d_2sm(smem_a, &tmA, full_leader, 0, (int)rank * 128, 0);
        tma_load_3d_2sm(smem_b, &tmB, full_leader, 0, (int)rank * (N / 2), 0);
    } else if (warp == 2 && lane == 0 && rank == 1) {
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(peer_leader) : "memory");   // peer -> leader
    } else if (warp == 1 && lane == 0 && rank == 0) {
        if (!mbar_wait_bounded(&peer_bar, 0, LIMIT)) status |= 1;
        if (!mbar_wait_bounded(&full_bar, 0, LIMIT)) status |= 2;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!status) {
            const uint32_t idesc = umma_idesc_f16(256, N);
            const uint64_t da = umma_desc(smem_u32(smem_a)), db = umma_desc(smem_u32(smem_b));
            for (int k = 0; k < 4; k++) {
                const uint32_t accum = k > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem_base), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"(accum) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&done_bar)), "h"((uint16_t)3) : "memory");
        out->info[4] = status;
    }
    if (warp >= 2) {
        const int ok = mbar_wait_bounded(&done_bar, 0, 2 * LIMIT);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok && lane == 0) atomicOr(&out->info[rank * 8 + 5], 1u << warp);
        const int quad = warp & 3;
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = (int)rank * 128 + quad * 32 + lane;
            for (int j = 0; j < 16; j++) out->d[row * 256 + c0 + j] = ok ? __uint_as_float(r[j]) : -12345.f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    int fails = 0;
    for (int N : {256, 128, 64}) {
        std::vector<__half> A(256 * 64), B(256 * 64);
        for (int r = 0; r < 256; r++)
            for (int k = 0; k < 64; k++) {
                A[r * 64 + k] = __float2half((float)(((r * 7 + k * 3) % 5) - 2));
                B[r * 64 + k] = __float2half((float)(((r * 5 + k) % 7) - 3));
            }
        __half *dA, *dB; Out* dOut;
        CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dOut, sizeof(Out)));
        CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemset(dOut, 0xff, sizeof(Out)));
        CUtensorMap mA, mB;
        cuuint64_t dims[3] = {64, 256, 1}; cuuint64_t strides[2] = {128, 128 * 256}; cuuint32_t es[3] = {1, 1, 1};
        cuuint32_t boxA[3] = {64, 128, 1}, boxB[3] = {64, (cuuint32_t)(N / 2), 1};
        if (enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
            enc(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, dims, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return 2; }
        const int smem = 128 * 128 + 128 * 128 + 1024;
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe_kernel<<<2, THREADS, smem>>>(mA, mB, N, dOut);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<Out> h(1);
        CK(cudaMemcpy(h.data(), dOut, sizeof(Out), cudaMemcpyDeviceToHost));
        const Out& o = h[0];
        printf("N=%d  rank0: full_bar 0x%x leader-addr 0x%x tmem 0x%x smem_a 0x%x | rank1: full_bar 0x%x leader-addr 0x%x tmem 0x%x smem_a 0x%x | "
               "issuer status %u, epilogue timeouts %x %x\n", N, o.info[0], o.info[1], o.info[2], o.info[3], o.info[8], o.info[9], o.info[10], o.info[11],
               o.info[4], o.info[5], o.info[13]);
        // expectation: D[r][n] = sum_k A[r][k] * B[n][k]; B row n of the instruction = global row n (CTA0: rows [0,N/2), CTA1: [N/2,N))
        long bad = 0, bad_swapped = 0;
        for (int r = 0; r < 256; r++)
            for (int n = 0; n < N; n++) {
                float s = 0, s2 = 0;
                const int n2 = (n + N / 2) % N;   // hypothesis "halves swapped"
                for (int k = 0; k < 64; k++) {
                    s += __half2float(A[r * 64 + k]) * __half2float(B[n * 64 + k]);
                    s2 += __half2float(A[r * 64 + k]) * __half2float(B[n2 * 64 + k]);
                }
                const float got = o.d[r * 256 + n];
                if (got != s) { if (bad < 5) printf("  mismatch r=%d n=%d got %g want %g\n", r, n, got, s); bad++; }
                if (got != s2) bad_swapped++;
            }
        printf("N=%d: %s (%ld mismatches; 'B halves swapped' hypothesis: %ld mismatches)\n", N, bad ? "FAIL" : "PASS", bad, bad_swapped);
        fails += bad != 0;
        cudaFree(dA); cudaFree(dB); cudaFree(dOut);
    }
    printf(fails ? "cg2_probe: FAIL\n" : "cg2_probe: PASS\n");
    return fails ? 1 : 0;
}
