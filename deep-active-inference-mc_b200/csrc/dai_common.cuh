// Shared device helpers: keyed Philox noise, row maps, reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dai {

constexpr int S_DIM = 10;
constexpr int PI_DIM = 4;
constexpr int IMG = 4096;            // 64*64*1

// noise sites, one per reference RNG draw inside one MC sample (SURVEY.md §8 a5 order);
// must match oracle/philox.py SITES.
constexpr int SITE_PS_A = 0, SITE_PO_A = 4, SITE_QS_A = 8, SITE_PS_B = 12, SITE_PO_B1 = 16,
              SITE_RP_B = 20, SITE_PO_B2 = 21, SITE_QS_ROOT = 32, SITE_CAT = 40;

struct NoiseKey {
    uint32_t k0, k1;     // seed + call_index
    uint32_t step;
    int32_t training;    // 0: dropout is identity
    const uint32_t* dyn; // non-null: {k0, k1, base step} are read from device memory instead and `step` is an offset from that
                         // base (replayed CUDA graphs: the captured kernels keep their parameters, the key of the call / step is
                         // written before each launch)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += W0; k1 += W1;
    }
    return c;
}

// 128 random bits for (site, block) of (row, sample) at this key/step.  `sample` carries the MC sample in its low 24 bits
// and, for launches that batch several horizon steps (RowMap::sps), the row's step offset in the high 8: the counter the
// reference stream is keyed on is (.., sample, nk.step + offset), the same as a launch of that step alone would use.
constexpr uint32_t SAMPLE_BITS = 24, SAMPLE_MASK = (1u << SAMPLE_BITS) - 1u;
__device__ __forceinline__ uint4 noise_block(const NoiseKey& nk, uint32_t site, uint32_t blk,
                                             uint32_t row, uint32_t sample) {
    uint32_t k0 = nk.k0, k1 = nk.k1, step = nk.step;
    if (nk.dyn) { k0 = __ldg(nk.dyn); k1 = __ldg(nk.dyn + 1); step += __ldg(nk.dyn + 2); }   // nk.step: offset from the replay's base step
    return philox4x32_10(make_uint4(blk | (site << 16), row, sample & SAMPLE_MASK, step + (sample >> SAMPLE_BITS)), k0, k1);
}

// standard normal for element e: Box-Muller in fp64 on words 0,1 of block e, rounded once
__device__ __forceinline__ float noise_normal(const NoiseKey& nk, uint32_t site, uint32_t e,
                                              uint32_t row, uint32_t sample) {
    const uint4 w = noise_block(nk, site, e, row, sample);
    const double u1 = ((double)w.x + 0.5) * (1.0 / 4294967296.0);
    const double u2 = ((double)w.y + 0.5) * (1.0 / 4294967296.0);
    return (float)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
}

__device__ __forceinline__ float noise_uniform24(const NoiseKey& nk, uint32_t site, uint32_t row,
                                                 uint32_t sample) {
    const uint4 w = noise_block(nk, site, 0, row, sample);
    return (float)(w.x >> 8) * (1.0f / 16777216.0f);
}

// reparameterize (src/torchmodel.py:54-56): eps * exp(0.5*logvar) + mean, mul then add like torch
__device__ __forceinline__ float reparam(float eps, float mean, float logvar) {
    return __fadd_rn(__fmul_rn(eps, expf(__fmul_rn(logvar, 0.5f))), mean);
}

// Rows of a batched net launch are ordered (set, local sample, b); set selects the noise
// site base (which of the reference's decoder/transition calls this row stands for).
struct RowMap {
    int32_t B;          // (state, action) rows of the call
    int32_t Sl;         // sample slots per set held by this rank
    int32_t sample0;    // global index of local sample slot 0
    int32_t nsets;
    int32_t site[3];    // noise site base per set
    int32_t b0;         // row offset of this launch inside the call's batch (encoder row chunks: noise row = b0 + local row)
    int32_t slot0;      // slot offset of this launch inside the call's slots (encoder slot chunks)
    int32_t sps;        // > 0: the slots are (horizon step, sample) pairs, `sps` samples per step (time-batched rollout:
                        // slot g = t * sps + j stands for local sample j of step nk.step + t); 0: one step
    __device__ __forceinline__ void decode(int r, int& set, int& slot, int& b) const {
        b = r % B + b0;
        const int q = r / B;
        slot = q % Sl;
        set = q / Sl;
    }
    // step offset of a (launch-local) slot and its sample slot within that step
    __device__ __forceinline__ void split(int slot, int& t, int& j) const {
        const int g = slot + slot0;
        t = sps > 0 ? g / sps : 0;
        j = g - t * sps;
    }
    // global MC sample of a slot, with the step offset in the high bits (see noise_block)
    __device__ __forceinline__ uint32_t sample_of(int slot) const {
        int t, j;
        split(slot, t, j);
        return (uint32_t)(sample0 + j) | ((uint32_t)t << SAMPLE_BITS);
    }
    __host__ __device__ int rows() const { return nsets * Sl * B; }
};

// Noise identity of the rows of a batched MLP launch: up to three consecutive row sets, each (slot, b)-ordered,
// each standing for one of the reference's net calls (site base); slot -> global MC sample.
struct NoiseRows {
    int32_t B;
    int32_t set_end[3];     // cumulative row count at the end of each set
    int32_t site[3];
    int32_t sample0;        // global sample of slot 0
    int32_t extra_slot;     // set 0 only: slot that stands for global sample `extra_sample` (or -1)
    int32_t extra_sample;
    int32_t b0;             // see RowMap::b0
    int32_t slot0, sps;     // see RowMap::slot0, RowMap::sps
    __device__ __forceinline__ void decode(int r, int& site_out, int& b, uint32_t& sample) const {
        const int set = r < set_end[0] ? 0 : (r < set_end[1] ? 1 : 2);
        const int q = r - (set == 0 ? 0 : set_end[set - 1]);
        const int slot = q / B;
        b = q - slot * B + b0;
        site_out = site[set];
        const int g = slot + slot0;
        const int t = sps > 0 ? g / sps : 0;
        const int j = g - t * sps;
        sample = ((set == 0 && j == extra_slot) ? (uint32_t)extra_sample : (uint32_t)(sample0 + j)) | ((uint32_t)t << SAMPLE_BITS);
    }
};

// Programmatic dependent launch.  A kernel launched with the programmatic-serialization attribute (launch_dep in
// dai_kernels.h) may start while the kernel before it in the stream is still running: whatever it does before
// pdl_wait() — barrier and tensor-memory set-up, weight loads — overlaps that kernel's tail; pdl_wait() returns once
// the preceding grid has completed and its memory is visible.  Every kernel that can be launched that way calls it
// before its first access to anything another kernel of the step writes or reads (so completion is transitive along
// the chain); without the attribute both instructions are no-ops.  pdl_trigger(): this CTA no longer minds the
// next kernel's CTAs being scheduled next to it (they only become resident where resources are free).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dai
