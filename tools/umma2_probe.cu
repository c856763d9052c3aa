// Probe of the 2-CTA MMA (tcgen05.mma.cta_group::2, M = 256 over a CTA pair): checks the operand split (each CTA
// supplies its 128 rows of A and HALF of B's N rows) on known values and measures cycles per MMA as a function of N,
// next to the single-CTA cost max(N/2, (4096 + 32 N)/128).  Round-2 groundwork for pair versions of ct1 / ct3 / FC4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma2_probe tools/umma2_probe.cu && tools/umma2_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_m(int m, int n) {      // D f32, A = B = bf16, K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 2000000000ll) return false;      // never hang the GPU
    }
    return true;
}

// out[0] = cycles (leader), out[1] = status; chk[rank*2 + {0,1}] = D[row 0][col 0], D[row 0][col n/2] of that CTA
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(int n, int iters, long long* out, float* chk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t rank = cta_rank();
    const int warp = threadIdx.x >> 5;
    uint16_t* A = reinterpret_cast<uint16_t*>(smem);                 // [2 k-chunks][128 rows][8]
    uint16_t* B = reinterpret_cast<uint16_t*>(smem + 8192);          // [2 k-chunks][n/2 rows][8]
    const uint16_t av = rank == 0 ? 0x3F80 : 0x4040;                 // bf16 1.0 | 3.0
    const uint16_t bv = rank == 0 ? 0x3F80 : 0x4000;                 // bf16 1.0 | 2.0
    for (int i = threadIdx.x; i < 2 * 128 * 8; i += blockDim.x) A[i] = av;
    for (int i = threadIdx.x; i < 2 * (n / 2) * 8; i += blockDim.x) B[i] = bv;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    long long cycles = 0;
    if (rank == 0 && threadIdx.x == 0) {
        const uint64_t a = umma_desc(smem_u32(A), 2048, 128);
        const uint64_t b = umma_desc(smem_u32(B), (uint32_t)(n / 2) * 16u, 128);
        const uint32_t idesc = idesc_m(256, n);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(i > 0 ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        cycles = clock64() - t0;
    }
    __shared__ int okflag;
    if (threadIdx.x == 0) {
        okflag = wait_bar(&bar, 0) ? 1 : 0;
        if (rank == 0) out[0] = cycles;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (okflag && warp == 0) {
        uint32_t r0, r1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(tmem) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r1) : "r"(tmem + (uint32_t)(n / 2)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (threadIdx.x == 0) { chk[rank * 2] = __uint_as_float(r0); chk[rank * 2 + 1] = __uint_as_float(r1); }
    }
    if (threadIdx.x == 0 && rank == 0) out[1] = okflag;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// timing variant: leader measures issue of `iters` MMAs until all complete
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_time(int n, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t rank = cta_rank();
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (rank == 0 && threadIdx.x == 0) {
        const uint64_t a = umma_desc(smem_u32(smem), 2048, 128);
        const uint64_t b = umma_desc(smem_u32(smem + 8192), (uint32_t)(n / 2) * 16u, 128);
        const uint32_t idesc = idesc_m(256, n);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(i > 0 ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        const bool ok = wait_bar(&bar, 0);
        out[blockIdx.x / 2 * 2] = clock64() - t0;
        out[blockIdx.x / 2 * 2 + 1] = ok ? 1 : 0;
    } else if (threadIdx.x == 0) {
        wait_bar(&bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* d; float* c;
    cudaMalloc(&d, 8 * 512); cudaMalloc(&c, 64);
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(probe2_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    printf("operand split check (A rows: CTA0 = 1, CTA1 = 3; B rows: CTA0 half = 1, CTA1 half = 2; K = 16, 1 MMA)\n");
    for (int n : {32, 64, 128, 256}) {
        cudaMemset(c, 0, 64); cudaMemset(d, 0, 64);
        probe2<<<2, 128, 64 * 1024>>>(n, 1, d, c);
        cudaError_t e = cudaDeviceSynchronize();
        float h[4]; long long hs[2];
        cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost); cudaMemcpy(hs, d, 16, cudaMemcpyDeviceToHost);
        printf("N %3d: %s status %lld | CTA0 D[0][0] %g D[0][N/2] %g | CTA1 D[0][0] %g D[0][N/2] %g   (expect 16 32 | 48 96)\n",
               n, cudaGetErrorString(e), hs[1], h[0], h[1], h[2], h[3]);
        if (e != cudaSuccess) return 1;
    }
    printf("cycles per MMA, cta_group::2 M=256 (per pair) vs the single-CTA model\n%5s %6s %12s %14s\n", "N", "pairs", "cyc/MMA", "1-CTA model");
    const int iters = 4096;
    for (int n : {32, 64, 96, 128, 192, 256}) {
        for (int pairs : {1, 74}) {
            long long h[512];
            for (int rep = 0; rep < 2; ++rep) {
                probe2_time<<<2 * pairs, 128, 64 * 1024>>>(n, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 8 * 2 * pairs, cudaMemcpyDeviceToHost);
            }
            long long mx = 0; int ok = 1;
            for (int i = 0; i < pairs; ++i) { mx = h[2 * i] > mx ? h[2 * i] : mx; ok &= (int)h[2 * i + 1]; }
            const double m1 = n / 2.0 > (4096 + 32.0 * n) / 128 ? n / 2.0 : (4096 + 32.0 * n) / 128;
            printf("%5d %6d %12.1f %14.1f %s\n", n, pairs, (double)mx / iters, m1, ok ? "" : "(TIMEOUT)");
        }
    }
    return 0;
}
