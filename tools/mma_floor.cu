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

struct Unit { uint32_t n, col; };
struct Case {
    int pair;            // 1: cta_group::2 (M = 256), 0: cta_group::1 (M = 128)
    int nunits;
    Unit u[4];
    uint32_t a_sbo, a_lbo;   // bytes
    int reps;
    int distinct;        // 1: A/B addresses advance per instruction as in the kernels, 0: the same operands every time
};

template <int PAIR>
__device__ __forceinline__ void mma(uint32_t lead, uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    if (PAIR)
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %7, 0;\n\t"
                     "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
                     ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(a_hi), "r"(b_hi), "r"(lead) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %7, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                     ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(a_hi), "r"(b_hi), "r"(lead) : "memory");
}

template <int PAIR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_floor(Case c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
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
        const uint32_t a_hi = (c.a_sbo >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
        const uint32_t mbits = PAIR ? (256u >> 4) : (128u >> 4);
        const long long t0 = clock64();
        int count = 0;
        for (int r = 0; r < c.reps; ++r) {
#pragma unroll 1
            for (int u = 0; u < c.nunits; ++u) {
                const uint32_t n = c.u[u].n;
                const uint32_t nh = PAIR ? n / 2 : n;
                const uint32_t bk = nh * 16u;
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | (mbits << 24);
                const uint32_t d = tmem + c.u[u].col;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t ka = c.distinct ? (uint32_t)(2 * k) * c.a_lbo + (uint32_t)u * 16u : 0u;
                    const uint32_t kb = c.distinct ? (uint32_t)(2 * k) * bk + (uint32_t)u * 4096u : 0u;
                    mma<PAIR>(lead, d, a16 + (ka >> 4) + ((c.a_lbo >> 4) << 16), a_hi, b16 + (kb >> 4) + ((bk >> 4) << 16), b_hi, idesc, 1u);
                    ++count;
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
        if (threadIdx.x == 0) { out[blockIdx.x * 4] = t2 - t0; out[blockIdx.x * 4 + 1] = t1 - t0; out[blockIdx.x * 4 + 2] = count; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

static void run(const char* name, Case c, int grid) {
    static long long* d = nullptr;
    if (!d) cudaMalloc(&d, 4 * 8 * 1024);
    cudaMemset(d, 0, 4 * 8 * 1024);
    const int smem = 160 * 1024;
    if (c.pair) {
        cudaFuncSetAttribute(k_floor<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_floor<1><<<grid, 128, smem>>>(c, d);
    } else {
        cudaFuncSetAttribute(k_floor<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_floor<0><<<grid, 128, smem>>>(c, d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    static long long h[4 * 1024];
    cudaMemcpy(h, d, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
    double tot = 0, iss = 0, cnt = 0; int nb = 0;
    for (int b = 0; b < grid; ++b) if (h[b * 4 + 2] > 0) { tot += (double)h[b * 4]; iss += (double)h[b * 4 + 1]; cnt = (double)h[b * 4 + 2]; ++nb; }
    printf("%-58s grid %3d: %6.1f cycles per MMA (issue loop alone %6.1f), %d instructions\n", name, grid, tot / nb / cnt, iss / nb / cnt, (int)cnt);
}

int main() {
    const uint32_t sbo_halo = 144, lbo_halo = 17 * 9 * 16, sbo_al = 128, lbo_al = 128 * 16;
    for (int grid : {2, 148}) {
        for (int pair : {1, 0}) {
            for (uint32_t n : {256u, 128u, 64u, 32u, 16u}) {
                if (pair && n < 32) continue;
                char nm[128];
                snprintf(nm, sizeof nm, "%s N=%u, halo strides, distinct operands", pair ? "pair M=256" : "single M=128", n);
                run(nm, Case{pair, 1, {{n, 0}, {0, 0}, {0, 0}, {0, 0}}, sbo_halo, lbo_halo, 64, 1}, grid);
            }
            char nm[128];
            snprintf(nm, sizeof nm, "%s N=64, halo strides, SAME operands", pair ? "pair M=256" : "single M=128");
            run(nm, Case{pair, 1, {{64, 0}, {0, 0}, {0, 0}, {0, 0}}, sbo_halo, lbo_halo, 64, 0}, grid);
            snprintf(nm, sizeof nm, "%s N=64, 128-B aligned row groups", pair ? "pair M=256" : "single M=128");
            run(nm, Case{pair, 1, {{64, 0}, {0, 0}, {0, 0}, {0, 0}}, sbo_al, lbo_al, 64, 1}, grid);
            snprintf(nm, sizeof nm, "%s N=128, 128-B aligned row groups", pair ? "pair M=256" : "single M=128");
            run(nm, Case{pair, 1, {{128, 0}, {0, 0}, {0, 0}, {0, 0}}, sbo_al, lbo_al, 64, 1}, grid);
            snprintf(nm, sizeof nm, "%s ct3 mix N=128,64,64,32 (model 202 per 4)", pair ? "pair M=256" : "single M=128");
            run(nm, Case{pair, 4, {{128, 0}, {64, 32}, {64, 64}, {32, 64}}, sbo_halo, lbo_halo, 32, 1}, grid);
            snprintf(nm, sizeof nm, "%s ct2 mix N=128,64,128,64", pair ? "pair M=256" : "single M=128");
            run(nm, Case{pair, 4, {{128, 0}, {64, 0}, {128, 0}, {64, 0}}, sbo_halo, lbo_halo, 32, 1}, grid);
        }
    }
    return 0;
}
