// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, no-swizzle K-major operands) as a function of
// N, of the accumulator dependency pattern and of operand reuse.  Guides the MMA ordering in dai_tc.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode bits: 1 = alternate accumulator (independent D), 2 = alternate B operand, 4 = alternate A operand
__global__ void probe(int n, int mode, int iters, long long* out, int a_shift, int a_sbo, int a_lbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), a1 = a0 + 32768, b0 = a0 + 65536, b1 = b0 + 32768;
        const uint32_t idesc = umma_idesc(n);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + ((mode & 1) ? (uint32_t)((i & 1) * 256) : 0u);
            const uint64_t a = umma_desc((((mode & 4) && (i & 1)) ? a1 : a0) + (uint32_t)a_shift, (uint32_t)a_lbo, (uint32_t)a_sbo);
            const uint64_t b = umma_desc(((mode & 2) && (i & 1)) ? b1 : b0, (uint32_t)n * 16u, 128);
            umma(d, a, b, idesc, i > 1 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

__global__ void probe_commit(int n, int per, int ncommit, int tiles, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 65536;
        const uint32_t idesc = umma_idesc(n);
        const uint64_t a = umma_desc(a0, 2048, 128), b = umma_desc(b0, (uint32_t)n * 16u, 128);
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            for (int i = 0; i < per; ++i) umma(tmem + (uint32_t)((t & 1) * 256), a, b, idesc, i > 0 ? 1u : 0u);
            for (int c = 0; c < ncommit; ++c)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[c])) : "memory");
            // barriers complete a phase per commit (count 1): nobody waits on them here
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[7])) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bars[7])) : "memory");
        }
        out[0] = t1 - t0;
        out[1] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int iters = 4096;
    cudaFuncSetAttribute(probe_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    printf("commit cost: tiles of `per` MMAs (N=64) followed by k commits; cycles per tile (issue loop | until all done)\n");
    for (int per : {4, 16, 48}) {
        for (int nc : {0, 1, 3}) {
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) {
                probe_commit<<<1, 128, 160 * 1024>>>(64, per, nc, 256, d);
                cudaDeviceSynchronize();
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            }
            printf("per %3d commits %d: issue %8.1f  total %8.1f   (MMA-only model %d)\n", per, nc, h[0] / 256.0, h[1] / 256.0, per * 48);
        }
    }
    printf("all-SM contention: cycles per MMA (max over CTAs) vs grid size\n%5s %6s %12s\n", "N", "grid", "cyc/MMA");
    long long* dd; cudaMalloc(&dd, 8 * 512);
    for (int n : {32, 64, 128, 256}) {
        for (int grid : {1, 2, 4, 74, 148, 296}) {
            long long h[512];
            for (int rep = 0; rep < 2; ++rep) {
                probe<<<grid, 128, 160 * 1024>>>(n, 0, iters, dd, 0, 128, 2048);
                cudaDeviceSynchronize();
                cudaMemcpy(h, dd, 8 * grid, cudaMemcpyDeviceToHost);
            }
            long long mx = 0, mn = 1ll << 60;
            for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%5d %6d %12.1f (min %.1f)\n", n, grid, (double)mx / iters, (double)mn / iters);
        }
    }
    printf("A-operand geometry (halo windows): cycles per MMA\n%5s %8s %6s %6s %12s\n", "N", "shift", "SBO", "LBO", "cyc/MMA");
    const int geo[][3] = {{0, 128, 2048}, {16, 128, 2048}, {0, 144, 2448}, {16, 144, 2448}, {0, 160, 2880}, {32, 160, 2880},
                          {0, 256, 4096}, {16, 256, 4096}, {0, 1024, 128}};
    for (int n : {32, 64, 128, 256}) {
        for (auto& g : geo) {
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) {
                probe<<<1, 128, 160 * 1024>>>(n, 0, iters, d, g[0], g[1], g[2]);
                cudaDeviceSynchronize();
                cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%5d %8d %6d %6d %12.1f\n", n, g[0], g[1], g[2], (double)c / iters);
        }
    }
    return 0;
}
