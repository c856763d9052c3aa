// Probe of the A-from-TMEM MMA (tcgen05.mma with [a_tmem]) fed by tcgen05.cp (shared memory -> TMEM): validates the
// result on known values, then times (1) the copy alone, (2) the TS-mode MMA alone as a function of N, (3) copy and
// MMAs interleaved the way a conv tile would issue them (1 copy per 1..3 MMAs).  Round-2 groundwork: in SS mode an
// M=128 MMA costs max(N/2, (4096+32N)/128) cycles because of its 4 KB A fetch; with A in TMEM the floor should be N/2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_ts_probe tools/umma_ts_probe.cu && tools/umma_ts_probe
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 2000000000ll) return false;
    }
    return true;
}

// mode 0: validation (1 copy + 1 TS MMA; chk[row] = D[row][0]); 1: copies only; 2: TS MMAs only; 3: SS MMAs only;
// 4: `per` TS MMAs per copy; 5: `per` SS MMAs (reference for mode 4)
__global__ void __launch_bounds__(128, 1) probe(int mode, int n, int per, int iters, long long* out, float* chk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    uint16_t* A = reinterpret_cast<uint16_t*>(smem);                 // [2 k-chunks][128 rows][8]: K-major no-swizzle
    uint16_t* B = reinterpret_cast<uint16_t*>(smem + 8192);          // [2 k-chunks][256 rows][8]
    // A[row][k] = 1 + (row % 4) (exact in bf16), B = 1  ->  D[row][*] = 16 * (1 + row % 4)
    const uint16_t vals[4] = {0x3F80, 0x4000, 0x4040, 0x4080};
    for (int i = threadIdx.x; i < 2 * 128 * 8; i += blockDim.x) A[i] = vals[((i >> 3) & 127) & 3];
    for (int i = threadIdx.x; i < 2 * 256 * 8; i += blockDim.x) B[i] = 0x3F80;
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
    const uint32_t a_tm = tmem + 256;                                // A staging: 8 columns per K=16 slice, 2 buffers
    __shared__ int okflag;
    if (threadIdx.x == 0) {
        const uint64_t a = umma_desc(smem_u32(A), 2048, 128);
        const uint64_t b = umma_desc(smem_u32(B), (uint32_t)n * 16u, 128);
        const uint32_t idesc = umma_idesc(n);
        const long long t0 = clock64();
        if (mode == 0) {
            cp_128x256b(a_tm, a);
            mma_ts(tmem, a_tm, b, idesc, 0u);
        } else if (mode == 1) {
            for (int i = 0; i < iters; ++i) cp_128x256b(a_tm + (uint32_t)((i & 1) * 8), a);
        } else if (mode == 2) {
            cp_128x256b(a_tm, a);
            for (int i = 0; i < iters; ++i) mma_ts(tmem, a_tm, b, idesc, i > 0 ? 1u : 0u);
        } else if (mode == 3) {
            for (int i = 0; i < iters; ++i) mma_ss(tmem, a, b, idesc, i > 0 ? 1u : 0u);
        } else if (mode == 4) {
            for (int i = 0; i < iters; ++i) {
                const uint32_t at = a_tm + (uint32_t)((i & 1) * 8);
                cp_128x256b(at, a);
                for (int j = 0; j < per; ++j) mma_ts(tmem, at, b, idesc, (i | j) ? 1u : 0u);
            }
        } else {
            for (int i = 0; i < iters; ++i)
                for (int j = 0; j < per; ++j) mma_ss(tmem, a, b, idesc, (i | j) ? 1u : 0u);
        }
        commit(&bar);
        okflag = wait_bar(&bar, 0) ? 1 : 0;
        out[0] = clock64() - t0;
        out[1] = okflag;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 0 && okflag) {
        uint32_t r0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        chk[threadIdx.x] = __uint_as_float(r0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static double run(int mode, int n, int per, int iters, long long* d, float* c) {
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 2; ++rep) {
        probe<<<1, 128, 32 * 1024>>>(mode, n, per, iters, d, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    if (!h[1]) { printf("(timeout) "); }
    return (double)h[0];
}

int main() {
    long long* d; float* c;
    cudaMalloc(&d, 64); cudaMalloc(&c, 512);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    cudaMemset(c, 0, 512);
    run(0, 64, 1, 1, d, c);
    float h[128];
    cudaMemcpy(h, c, 512, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r) bad += h[r] != 16.0f * (1 + (r & 3));
    printf("TS-mode validation (A via tcgen05.cp.128x256b, D[row][0] = 16 * (1 + row %% 4)): rows 0..7 = %g %g %g %g %g %g %g %g, %d of 128 rows wrong\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], bad);
    const int iters = 4096;
    printf("tcgen05.cp.128x256b (4 KB): %.1f cycles per copy\n", run(1, 64, 1, iters, d, c) / iters);
    printf("%5s %12s %12s\n", "N", "TS cyc/MMA", "SS cyc/MMA");
    for (int n : {32, 64, 128, 256}) printf("%5d %12.1f %12.1f\n", n, run(2, n, 1, iters, d, c) / iters, run(3, n, 1, iters, d, c) / iters);
    printf("one copy per `per` MMAs (cycles per group): TS (copy + MMAs) vs SS (MMAs only)\n%5s %5s %12s %12s\n", "N", "per", "TS", "SS");
    for (int n : {32, 64, 128})
        for (int per : {1, 2, 3}) printf("%5d %5d %12.1f %12.1f\n", n, per, run(4, n, per, iters, d, c) / iters, run(5, n, per, iters, d, c) / iters);
    return 0;
}
