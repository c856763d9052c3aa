// Microbenchmark (experiments only, not part of the library): cycles per tcgen05.mma kind::f16 instruction as a function
// of N, for cta_group::1 (M = 128) and cta_group::2 (M = 256), with the operand strides of the pair kernels
// (no-swizzle K-major core matrices; A: 8-pixel rows of a 9-pixel-wide halo, i.e. 144 B between row groups).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_floor tools/mma_floor.cu && tools/mma_floor
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// One case = up to four "units" (N, TMEM column) issued round-robin, four K = 16 steps each, all compile-time constants so
// that the issue loop is a handful of uniform-datapath instructions per MMA and the tensor pipe, not the issuer, is timed.
template <int PAIR>
__device__ __forceinline__ void mma(uint32_t lead, uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    if (PAIR)
        asm volatile("{\n\t.reg .pred q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 q, %6, 0;\n\t"
                     "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n\t}"
                     ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(a_hi), "r"(b_hi), "r"(lead) : "memory");
    else
        asm volatile("{\n\t.reg .pred q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 q, %6, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t}"
                     ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(a_hi), "r"(b_hi), "r"(lead) : "memory");
}

template <int PAIR, int N0, int N1, int N2, int N3, int SBO, int LBO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_floor(int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (uint32_t)i % 7u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 0 && (!PAIR || rank == 0)) {
        const uint32_t lead = elect_one() ? 1u : 0u;
        const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 96 * 1024) >> 4;
        constexpr uint32_t a_hi = ((uint32_t)SBO >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
        constexpr uint32_t mbits = PAIR ? (256u >> 4) : (128u >> 4);
        constexpr int NS[4] = {N0, N1, N2, N3};
        constexpr int COL[4] = {0, N0 >= 256 ? 0 : 128, 256, 384};
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (NS[u] == 0) continue;
                constexpr uint32_t dummy = 0; (void)dummy;
                const uint32_t n = (uint32_t)NS[u];
                const uint32_t nh = PAIR ? n / 2 : n;
                const uint32_t bk = nh * 16u;
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | (mbits << 24);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t ka = (uint32_t)(2 * k) * (uint32_t)LBO + (uint32_t)u * 16u;
                    const uint32_t kb = (uint32_t)(2 * k) * bk + (uint32_t)u * 4096u;
                    mma<PAIR>(lead, tmem + (uint32_t)COL[u], a16 + (ka >> 4) + (((uint32_t)LBO >> 4) << 16), a_hi,
                              b16 + (kb >> 4) + ((bk >> 4) << 16), b_hi, idesc);
                }
            }
        }
        const long long t1 = clock64();
        if (elect_one()) {
            if (PAIR)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        while (!mbar_try(&bar, 0)) {}
        const long long t2 = clock64();
        constexpr int per_rep = 4 * ((N0 > 0) + (N1 > 0) + (N2 > 0) + (N3 > 0));
        if (threadIdx.x == 0) { out[blockIdx.x * 4] = t2 - t0; out[blockIdx.x * 4 + 1] = t1 - t0; out[blockIdx.x * 4 + 2] = (long long)reps * per_rep; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int PAIR, int N0, int N1, int N2, int N3, int SBO, int LBO>
static void run(const char* name, int grid) {
    static long long* d = nullptr;
    if (!d) cudaMalloc(&d, 4 * 8 * 1024);
    cudaMemset(d, 0, 4 * 8 * 1024);
    const int smem = 160 * 1024;
    auto kern = k_floor<PAIR, N0, N1, N2, N3, SBO, LBO>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<grid, 128, smem>>>(64, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    static long long h[4 * 1024];
    cudaMemcpy(h, d, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
    double tot = 0, iss = 0, cnt = 0; int nb = 0;
    for (int b = 0; b < grid; ++b) if (h[b * 4 + 2] > 0) { tot += (double)h[b * 4]; iss += (double)h[b * 4 + 1]; cnt = (double)h[b * 4 + 2]; ++nb; }
    printf("%s %-44s grid %3d: %6.1f cycles per MMA (issue loop alone %6.1f), %d instructions\n", PAIR ? "pair M=256  " : "single M=128", name, grid,
           tot / nb / cnt, iss / nb / cnt, (int)cnt);
}

template <int PAIR>
static void suite(int grid) {
    constexpr int SH = 144, LH = 17 * 9 * 16, SA = 128, LA = 128 * 16;
    run<PAIR, 256, 0, 0, 0, SH, LH>("N=256, halo strides", grid);
    run<PAIR, 128, 0, 0, 0, SH, LH>("N=128, halo strides", grid);
    run<PAIR, 64, 0, 0, 0, SH, LH>("N=64, halo strides", grid);
    run<PAIR, 32, 0, 0, 0, SH, LH>("N=32, halo strides", grid);
    run<PAIR, 128, 0, 0, 0, SA, LA>("N=128, 128-B aligned row groups", grid);
    run<PAIR, 64, 0, 0, 0, SA, LA>("N=64, 128-B aligned row groups", grid);
    run<PAIR, 32, 0, 0, 0, SA, LA>("N=32, 128-B aligned row groups", grid);
    run<PAIR, 128, 64, 64, 32, SH, LH>("ct3 mix N=128,64,64,32", grid);
    run<PAIR, 128, 64, 128, 64, SH, LH>("ct2 mix N=128,64,128,64", grid);
    run<PAIR, 128, 128, 128, 128, SH, LH>("N=128 x 4 accumulators", grid);
    run<PAIR, 64, 64, 64, 64, SH, LH>("N=64 x 4 accumulators", grid);
}

int main() {
    for (int grid : {2, 148}) {
        suite<1>(grid);
        suite<0>(grid);
    }
    return 0;
}
