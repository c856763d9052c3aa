// Tensor-core path of the decoder's contraction layers: tcgen05.mma (kind::f16, bf16 operands,
// fp32 accumulators in TMEM) fed by TMA, warp-specialised, persistent.
//
// Transposed convolutions are implicit GEMMs over a halo tile that is loaded ONCE per
// 128-pixel output tile and re-read by every tap through shifted shared-memory descriptors:
//   activations in HBM are channel-blocked   [plane hi|lo][row][kc = C/8][H][W][8] bf16
//   a TMA box (W' x H' x 8 kc x 2 planes) lands as [plane][kc][y][x][8] = the UMMA
//   no-swizzle K-major canonical layout (8 pixels x 16 B core matrices, LBO = kc plane,
//   SBO = halo row pitch), so a tap shift is just a start-address offset of 16 B per pixel.
// Precision: x = hi + lo (bf16 each); D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (DAI_PREC_BF16X3)
// or the first product only (DAI_PREC_BF16X1).
#include "dai_tc.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/dai_b200.h"

namespace dai {

namespace {

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive without memory-ordering semantics: used where the barrier hands over a TMEM accumulator that has already been
// read into registers (tcgen05.wait::ld + tcgen05.fence::before_thread_sync order the tensor-memory side).  The default
// .release arrive makes the warp wait for all of its outstanding global stores first (MEMBAR + ERRBAR in SASS; 11 % of
// the pair kernel's samples before this change).
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must fail the launch (trap -> CUDA error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// one lane of a CONVERGED warp (the issue pattern the compiler turns into a single predicated UTCHMMA; issuing from
// inside `if (lane == 0)` instead costs ~10 extra predicate / branch instructions per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, M = 128, K = 16, bf16 x bf16 -> f32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (elect_one())
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptors given as (low word in a register, compile-time high word): the high words become uniform-register
// immediates instead of two more register -> uniform-register moves on the issuing thread
template <uint32_t A_TOP, uint32_t B_TOP>
__device__ __forceinline__ void umma_bf16_split(uint32_t lead, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
    // `lead` = 1 on the one lane elected for the whole tile (the warp stays converged; the MMA is predicated)
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(A_TOP), "n"(B_TOP), "r"(lead)
        : "memory");
}
// arrive on an mbarrier once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if (elect_one())
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// this warp's 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle: 8-row x 16-byte core matrices;
// LBO = bytes between the two 16-byte K chunks of one MMA, SBO = bytes between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---------------------------------------------------------------------------------------
// layer geometry
// ---------------------------------------------------------------------------------------
struct Unit {            // one MMA group of a tile: a tap shift of the halo x a block of weights
    int16_t oy, ox;      // halo-local pixel offset of the A window
    int16_t col;         // TMEM column offset inside the tile's accumulator block
    int16_t n;           // N of the MMA
    int16_t init;        // 1: its first MMA overwrites the accumulator
    int16_t kc0;         // first 8-channel plane of the halo this unit reads (parity plane of a strided conv)
    int32_t woff;        // byte offset of its weight block in the resident weight image
};

constexpr int MAX_UNITS = 9;

struct ConvParams {
    Unit units[MAX_UNITS];
    int32_t nunits;
    int32_t nrows;           // images in this launch
    int32_t nprod;           // 3: bf16x3, 1: bf16x1
    const uint8_t* wpack;    // packed weights (see pack_layer)
    const float* bias;       // [Cout]
    void* out;               // blocked bf16 planes (ct1, ct2) or the last deconv's row planes + border terms [row][PROJ_ROW_FLOATS] (ct3)
    float2 w4[288];          // ct3: last deconv's weights [c 32][tap 9], each duplicated (w, w) for packed fp32x2 FMAs over the
                             // (left, right) output pixel pair (kernel params = constant bank).  GEN layers: conv1's weights as
                             // channel pairs [tap 9][16] and its bias [16] (entries 0..159)
    int32_t two_pass;        // 1: per tile all hi-plane MMAs, then all lo-plane MMAs (each ring slot is released as soon
                             // as its plane is done); 0: both planes waited for, products interleaved per k step
    const float* gen_img;    // GEN layers: fp32 images [nrows][64][64] (conv1's weights and bias travel in w4)
    int32_t dbg;             // experiments only (env DAI_TC_DBG): 1 = epilogue does no work, 2 = no MMAs issued
    long long* counters;     // experiments only: per CTA {mma_total, mma_wait_acc, mma_wait_a, epi_total, epi_wait, tiles, 0, 0}
};

enum { OUT_BLOCKED = 0, OUT_PROJ = 1, OUT_PARITY = 2, OUT_NHWC_F32 = 3 };

// Layer traits:
//   MODE     0: convT k3 s1 p1; 1: convT k3 s2 p1 op1 (4 sub-pixel phases); 2: conv k3 s2 valid (parity-split input)
//   NPH      Cout;  KCIN  Cin/8
//   GH,GW    m-grid the 16x8 tiles cover (output grid for MODE 0/2, input grid for MODE 1); VH,VW its valid part
//   PH,PW    height/width of one input plane in HBM (MODE 2: of one parity plane)
//   NA       halo ring slots (one slot = one bf16 plane, hi or lo, of one tile)
//   TWO_PASS (default of ConvParams::two_pass; env DAI_TC_TWO_PASS = bit mask over layer IDs overrides it) per tile all
//            MMAs on the hi plane first, then all on the lo plane, each ring slot released when its plane is done;
//            otherwise both planes are waited for and the products of one (tap, k) step are issued back to back
//   CONCAT   (bf16x3) a weight block stores hi rows then lo rows: A_hi*[B_hi;B_lo] is ONE MMA of doubled N — the fixed
//            ~46-cycle cost of a small-N MMA is paid twice per tap instead of three times; the hi*lo part lands in a
//            second column block that the epilogue adds
struct TrCt1 { static constexpr int ID = 0, MODE = 0, NPH = 64, KCIN = 8, GH = 16, GW = 16, VH = 16, VW = 16, PH = 16, PW = 16, NA = 3,
               EPI_WARPS = 8, OUT = OUT_BLOCKED; static constexpr bool TWO_PASS = true, CONCAT = true, GEN = false; };
struct TrCt2 { static constexpr int ID = 1, MODE = 1, NPH = 64, KCIN = 8, GH = 16, GW = 16, VH = 16, VW = 16, PH = 16, PW = 16, NA = 4,
               EPI_WARPS = 8, OUT = OUT_BLOCKED; static constexpr bool TWO_PASS = false, CONCAT = false, GEN = false; };
struct TrCt3 { static constexpr int ID = 2, MODE = 1, NPH = 32, KCIN = 8, GH = 32, GW = 32, VH = 32, VW = 32, PH = 32, PW = 32, NA = 7,
               EPI_WARPS = 8, OUT = OUT_PROJ; static constexpr bool TWO_PASS = false, CONCAT = false, GEN = false; };
// encoder: Conv2d 32->32 (31x31 -> 15x15) and 32->64 (15x15 -> 7x7), k3 s2 valid
struct TrQc2 { static constexpr int ID = 3, MODE = 2, NPH = 32, KCIN = 4, GH = 16, GW = 16, VH = 15, VW = 15, PH = 16, PW = 16, NA = 4,
               EPI_WARPS = 8, OUT = OUT_PARITY; static constexpr bool TWO_PASS = false, CONCAT = false, GEN = false; };
struct TrQc3 { static constexpr int ID = 4, MODE = 2, NPH = 64, KCIN = 4, GH = 16, GW = 8, VH = 7, VW = 7, PH = 8, PW = 8, NA = 4,
               EPI_WARPS = 8, OUT = OUT_NHWC_F32; static constexpr bool TWO_PASS = false, CONCAT = false, GEN = false; };
// conv2 with conv1 (Conv2d 1->32, k3 s2 valid, 64x64 -> 31x31, + ReLU) computed IN the kernel: GEN_WARPS generator warps
// write each tile's halo planes (conv1's output around the tile, bf16 hi/lo, parity-split) straight into the ring slots the
// MMAs read, from the fp32 image — the 128 KB/row activation between the two layers (conv1 was bound by writing it to HBM,
// conv2 by reading it back) never exists in global memory.
struct TrQc2G : TrQc2 { static constexpr int ID = 5; static constexpr bool GEN = true; };

template <class T>
struct Cfg : T {
    static constexpr int GEN_WARPS = T::GEN ? 12 : 0;                    // generator warps (after the four single-role warps)
    static constexpr int THREADS = 128 + 32 * T::EPI_WARPS + 32 * GEN_WARPS;
    static constexpr int TH = 16, TW = 8;    // tile of the m-grid: 128 pixels
    static constexpr int HY = T::MODE == 0 ? TH + 2 : TH + 1;
    static constexpr int HX = T::MODE == 0 ? TW + 2 : TW + 1;
    static constexpr int KPLANES = T::MODE == 2 ? 4 * T::KCIN : T::KCIN;   // 8-channel planes in the halo (x4 parities)
    static constexpr int KSTEPS = T::KCIN / 2;                            // k16 steps per tap
    static constexpr int NUNITS = T::MODE == 1 ? 4 : 9;                   // MMA groups per tile (unit_at)
    static constexpr int KC_STRIDE = HY * HX * 16;             // bytes of one 8-channel plane of the halo
    static constexpr int PLANE_A = KPLANES * KC_STRIDE;        // one bf16 plane (hi or lo) = one ring slot
    static constexpr int TILES_X = (T::GW + TW - 1) / TW, TILES_Y = (T::GH + TH - 1) / TH;
    static constexpr int TILES = TILES_X * TILES_Y;
    static constexpr int ACC_COLS = (T::MODE == 1 ? 4 * T::NPH : T::NPH) * (T::CONCAT ? 2 : 1);
    // accumulator buffers in TMEM: as many as fit (<= 4).  The MMA-complete -> epilogue -> buffer-free round trip is
    // long compared with a tile's MMA time, so two buffers leave the tensor pipe idle between tiles.
    static constexpr int NACC = ACC_COLS * 4 <= 512 ? 4 : 2;
    static constexpr int TMEM_COLS = ACC_COLS * NACC < 32 ? 32 : ACC_COLS * NACC;
    static constexpr int W_BYTES = 9 * T::NPH * T::KCIN * 32;  // all 9 taps, hi + lo, resident
    static constexpr int SMEM_A = T::NA * PLANE_A;
    static constexpr int SMEM_TAIL = 1024;                     // barriers, tmem slot, bias
    static constexpr int SMEM_GEN = T::GEN ? 16384 : 0;        // the current tile's fp32 image (generator warps)
    static constexpr int SMEM_BYTES = W_BYTES + SMEM_A + SMEM_TAIL + SMEM_GEN;
};

// The unit table of a layer as a function of compile-time traits (build_layer packs the weights in this order and
// checks itself against it).  The MMA issuer unrolls over it, so every per-unit quantity below is an immediate.
template <class C>
__host__ __device__ constexpr Unit unit_at(int u) {
    constexpr int CW = C::NPH * C::KCIN * 32;          // bytes of one tap's weights, hi + lo
    Unit r{};
    if (C::MODE == 1) {
        // grouped sub-pixel phases, TMEM column slots [00, 01, 11, 10]
        r.oy = (int16_t)(u >> 1); r.ox = (int16_t)(u & 1);
        r.col = (int16_t)(u == 0 ? 0 : (u == 1 ? C::NPH : 2 * C::NPH));
        r.n = (int16_t)(u == 0 ? 4 * C::NPH : (u == 3 ? C::NPH : 2 * C::NPH));
        r.init = (int16_t)(u == 0);
        r.kc0 = 0;
        r.woff = (u == 0 ? 0 : (u == 1 ? 4 : (u == 2 ? 6 : 8))) * CW;
    } else {
        const int kh = u / 3, kw = u % 3;
        r.oy = (int16_t)(C::MODE == 0 ? 2 - kh : kh >> 1); r.ox = (int16_t)(C::MODE == 0 ? 2 - kw : kw >> 1);
        r.col = 0; r.n = (int16_t)C::NPH; r.init = (int16_t)(u == 0);
        r.kc0 = (int16_t)(C::MODE == 2 ? ((kh & 1) * 2 + (kw & 1)) * C::KCIN : 0);
        r.woff = u * CW;
    }
    return r;
}

// All MMAs of one tile.  a_hi16 / a_lo16 / w16: shared-memory addresses >> 4 of the two halo planes and of the resident
// weights; d0: TMEM address of the tile's accumulator block.  SEL 0: the products that read the hi plane, 1: the
// product that reads the lo plane, 2: all of them, interleaved per k step.  A descriptor is (constant high word,
// base + immediate low word): one integer add per operand per MMA on the issuing thread, which matters — at N <= 128
// an MMA takes 46..64 cycles and the issue stream of ONE thread has to stay ahead of that.
template <class C, int SEL, bool X3>
__device__ __forceinline__ void issue_tile(uint32_t a_hi16, uint32_t a_lo16, uint32_t w16, uint32_t d0) {
    constexpr uint32_t A_TOP = (uint32_t)((C::HX * 16) >> 4) | (1u << 14);   // SBO = halo row pitch; descriptor bit 46
    constexpr uint32_t B_TOP = (uint32_t)(128 >> 4) | (1u << 14);            // SBO = 8 rows x 16 B
    constexpr uint32_t A_LBO = (uint32_t)(C::KC_STRIDE >> 4) << 16;
    const uint32_t lead = elect_one() ? 1u : 0u;
#pragma unroll
    for (int u = 0; u < C::NUNITS; ++u) {
        const Unit un = unit_at<C>(u);
        const uint32_t n = (uint32_t)un.n;
        const uint32_t bk = C::CONCAT ? 2u * n * 16u : n * 16u;       // bytes between kc planes of the weight block
        const uint32_t b_plane = (uint32_t)C::KCIN * n * 16u;         // hi -> lo plane of a non-concatenated block
        const uint32_t a_off = (uint32_t)un.kc0 * C::KC_STRIDE + (uint32_t)(un.oy * C::HX + un.ox) * 16u;
        const uint32_t d = d0 + (uint32_t)un.col;
#pragma unroll
        for (int k = 0; k < C::KSTEPS; ++k) {
            const uint32_t acc0 = (un.init && k == 0) ? 0u : 1u;
            const uint32_t a_imm = ((a_off + (uint32_t)(2 * k) * C::KC_STRIDE) >> 4) + A_LBO;
            const uint32_t b_imm = (((uint32_t)un.woff + (uint32_t)(2 * k) * bk) >> 4) + ((bk >> 4) << 16);
            const uint32_t a_hi = a_hi16 + a_imm, a_lo = a_lo16 + a_imm;
            const uint32_t b_hi = w16 + b_imm, b_lo = w16 + b_imm + (b_plane >> 4);
            if (!X3) {
                if (SEL != 1) umma_bf16_split<A_TOP, B_TOP>(lead, d, a_hi, b_hi, umma_idesc(un.n), acc0);
            } else if (C::CONCAT) {
                if (SEL != 1) umma_bf16_split<A_TOP, B_TOP>(lead, d, a_hi, b_hi, umma_idesc(2 * un.n), acc0);   // [A_hi*B_hi | A_hi*B_lo]
                if (SEL != 0) umma_bf16_split<A_TOP, B_TOP>(lead, d, a_lo, b_hi, umma_idesc(un.n), 1u);         // A_lo*B_hi (first n rows)
            } else {
                if (SEL != 1) umma_bf16_split<A_TOP, B_TOP>(lead, d, a_hi, b_hi, umma_idesc(un.n), acc0);
                if (SEL != 0) umma_bf16_split<A_TOP, B_TOP>(lead, d, a_lo, b_hi, umma_idesc(un.n), 1u);
                if (SEL != 1) umma_bf16_split<A_TOP, B_TOP>(lead, d, a_hi, b_lo, umma_idesc(un.n), 1u);
            }
        }
    }
}

// relu(acc + bias) for two neighbouring channels -> packed bf16 hi pair and lo pair (x = hi + lo)
__device__ __forceinline__ void split2(uint32_t r0, uint32_t r1, float b0, float b1, float scale0, float scale1,
                                       uint32_t& hi, uint32_t& lo) {
    const float v0 = fmaxf(__uint_as_float(r0) + b0, 0.0f) * scale0;
    const float v1 = fmaxf(__uint_as_float(r1) + b1, 0.0f) * scale1;
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float f0 = __uint_as_float(hi << 16), f1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - f0, v1 - f1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// packed fp32x2 FMA (FFMA2): d = a * b + c on both halves
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}

__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&a)[4], const uint32_t (&b)[4]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]),
                 "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3])
                 : "memory");
}
// L2 eviction priorities: the pair kernel's scratch images should stay in L2 between the ct2 epilogue's stores, ct3's loads
// and the next overwrite two stages later (evict_last); its streaming traffic (act1 in, projection rows out) should not
// push them out (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_global_256_hint(void* p, const uint32_t (&a)[4], const uint32_t (&b)[4], uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]),
                 "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "l"(pol)
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// the kernel: warps 0..7 = epilogue (TMEM -> registers -> HBM; TMEM lane quarter = warp % 4), then
// warp 8 = halo TMA producer, warp 9 = MMA issuer, warp 10 = TMEM allocator, warp 11 = weight loader (once).
// The single-thread roles sit in the HIGHEST warp ids on purpose: the warp scheduler favours higher warp ids among
// eligible warps, and an MMA issuer that loses arbitration to the FFMA-heavy epilogue warps starves the tensor pipe.
// Per tile the issuer makes two passes over the tap units: pass 1 on the hi plane of the halo
// (A_hi*B_hi, A_hi*B_lo), pass 2 on the lo plane (A_lo*B_hi); the planes travel through the ring
// separately so three 20 KB slots are enough to keep the next tile's data in flight.
// ---------------------------------------------------------------------------------------
// DBG = true is the experiments build of the same kernel (in-kernel cycle counters, `p.dbg` switches; env DAI_TC_DBG /
// DAI_TC_COUNTERS select it at launch); the production instantiation carries none of it.
template <class C, bool DBG>
__global__ void __launch_bounds__(C::THREADS, 1) k_tc_conv(const __grid_constant__ CUtensorMap tmapA, const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smW = smem;
    uint8_t* smA = smem + C::W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::W_BYTES + C::SMEM_A);
    uint64_t* a_full = bars;                  // [NA]
    uint64_t* a_empty = a_full + C::NA;       // [NA]
    uint64_t* acc_full = a_empty + C::NA;     // [NACC]
    uint64_t* acc_empty = acc_full + C::NACC; // [NACC]
    uint64_t* w_full = acc_empty + C::NACC;   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 2);   // [NPH]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int W_PROD = C::EPI_WARPS, W_MMA = C::EPI_WARPS + 1, W_ALLOC = C::EPI_WARPS + 2, W_WGT = C::EPI_WARPS + 3;
    const int nplanes = p.nprod == 3 ? 2 : 1;
    float* simg = reinterpret_cast<float*>(smem + C::W_BYTES + C::SMEM_A + C::SMEM_TAIL);   // GEN: the current tile's image
    if (threadIdx.x == 0) {
        // a_full: one arrival with the TMA's transaction bytes, or one per generator warp
        for (int i = 0; i < C::NA; ++i) { mbar_init(&a_full[i], C::GEN ? C::GEN_WARPS : 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < C::NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], C::EPI_WARPS); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == W_ALLOC) tmem_alloc(tmem_slot, C::TMEM_COLS);
    if (threadIdx.x < C::NPH) sbias[threadIdx.x] = p.bias[threadIdx.x];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int ntiles = p.nrows * C::TILES;
    // programmatic dependent launch: everything above (and the weight loads below) may overlap the tail of the kernel
    // before; the halo planes it wrote and the buffers this kernel writes may be touched only after it has completed
    if (warp != W_WGT) pdl_wait();
    pdl_trigger();

    if (C::GEN && warp > W_WGT) {
        // ===== halo generators (GEN layers): conv1 + ReLU of the tile's surroundings -> the ring slots, hi and lo =====
        // One thread per halo position (parity plane, y, x): 9 image pixels, 32 channels x 9 FMAs in the order of
        // k_qs_conv1 (the same bits), eight 16-byte shared-memory stores (4 channel groups x hi/lo; consecutive positions
        // are consecutive 16-byte units: conflict-free).  Positions outside the parity plane or outside conv1's 31 x 31
        // output are zero, like the TMA's out-of-bounds fill of the two-kernel path.
        // The tile's image is staged in shared memory by one coalesced sweep (the next tile's is in flight, in registers,
        // while this tile is computed): read per position straight from global memory, the nine pixel loads of each of a
        // thread's two or three positions were serialised L2 round trips and the generators, not the MMAs, paced the kernel.
        const int gtid = (warp - W_WGT - 1) * 32 + lane;
        constexpr int NPOS = 4 * C::HY * C::HX, GT = 32 * C::GEN_WARPS;
        constexpr int NPRE = C::GEN ? (1024 + GT - 1) / (GT > 0 ? GT : 1) : 1;     // float4 of the 16 KB image per generator thread
        float4 pre[NPRE];
        auto fetch = [&](int tile) {
            const float4* src = reinterpret_cast<const float4*>(p.gen_img + (size_t)(tile / C::TILES) * 4096);
#pragma unroll
            for (int i = 0; i < NPRE; ++i)
                if (gtid + i * GT < 1024) pre[i] = __ldg(src + gtid + i * GT);
        };
        if ((int)blockIdx.x < ntiles) fetch(blockIdx.x);
        int cnt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            asm volatile("bar.sync 1, %0;" ::"n"(GT > 0 ? GT : 32) : "memory");      // every generator is done with the previous image
#pragma unroll
            for (int i = 0; i < NPRE; ++i)
                if (gtid + i * GT < 1024) reinterpret_cast<float4*>(simg)[gtid + i * GT] = pre[i];
            asm volatile("bar.sync 1, %0;" ::"n"(GT > 0 ? GT : 32) : "memory");
            if (tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x);
            const int t = tile % C::TILES;
            const int y0 = (t / C::TILES_X) * C::TH, x0 = (t % C::TILES_X) * C::TW;
            const int s_hi = cnt % C::NA;
            const uint32_t ph_hi = (uint32_t)(cnt / C::NA) & 1u;
            ++cnt;
            int s_lo = s_hi;
            uint32_t ph_lo = ph_hi;
            if (nplanes == 2) { s_lo = cnt % C::NA; ph_lo = (uint32_t)(cnt / C::NA) & 1u; ++cnt; }
            if (lane == 0) {
                mbar_wait(&a_empty[s_hi], ph_hi ^ 1u);
                if (nplanes == 2) mbar_wait(&a_empty[s_lo], ph_lo ^ 1u);
            }
            __syncwarp();
            uint8_t* dhi = smA + (size_t)s_hi * C::PLANE_A;
            uint8_t* dlo = smA + (size_t)s_lo * C::PLANE_A;
            // Positions are enumerated (row parity, y, c = 2x + column parity): consecutive lanes read consecutive pixel pairs
            // of the image (8-byte loads, conflict-free) and write 16-byte units that alternate between two parity planes 64
            // bytes apart modulo 128 — conflict-free as well.  A thread's (up to) two positions are computed together with
            // packed fp32x2 FMAs (two channels per instruction, the same IEEE fma per lane as k_qs_conv1's) whose weight
            // operands come from the constant bank: fetched from shared memory (warp-uniform 16-byte loads, four passes
            // each) the weights alone kept the shared-memory pipe busy for 3.8k cycles per tile, longer than the tile's MMAs.
            constexpr int NP = C::GEN ? (NPOS + GT - 1) / (GT > 0 ? GT : 1) : 1;
            const unsigned long long* gw = reinterpret_cast<const unsigned long long*>(p.w4);
            float v[NP][9];
            uint32_t off[NP];
            bool valid[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int pos = gtid + q * GT;
                const int py = pos / (C::HY * 2 * C::HX), rem = pos - py * (C::HY * 2 * C::HX);
                const int hy = rem / (2 * C::HX), c = rem - hy * (2 * C::HX);
                const int hx = c >> 1, px = c & 1;
                const int Y = y0 + hy, X = x0 + hx;                       // position in the parity plane
                const int oy = 2 * Y + py, ox = 2 * X + px;               // conv1 output pixel
                valid[q] = pos < NPOS && Y < C::PH && X < C::PW && oy < 31 && ox < 31;
                off[q] = (uint32_t)((py * 2 + px) * 4) * C::KC_STRIDE + (uint32_t)(hy * C::HX + hx) * 16u;
                const float* in = simg + (valid[q] ? (2 * oy) * 64 + 2 * ox : 0);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const float2 a = *reinterpret_cast<const float2*>(in + kh * 64);
                    v[q][kh * 3 + 0] = a.x; v[q][kh * 3 + 1] = a.y; v[q][kh * 3 + 2] = in[kh * 64 + 2];
                }
            }
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                unsigned long long acc[NP][4];
#pragma unroll
                for (int q = 0; q < NP; ++q)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[q][e] = gw[144 + kc * 4 + e];
#pragma unroll
                for (int i = 0; i < 9; ++i)
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const unsigned long long vv = pack_f32x2(v[q][i], v[q][i]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[q][e] = ffma2(vv, gw[i * 16 + kc * 4 + e], acc[q][e]);
                    }
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (gtid + q * GT >= NPOS) continue;
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float v0 = fmaxf(__uint_as_float((uint32_t)acc[q][e]), 0.0f), v1 = fmaxf(__uint_as_float((uint32_t)(acc[q][e] >> 32)), 0.0f);
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
                        hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - __uint_as_float(hi[e] << 16), v1 - __uint_as_float(hi[e] & 0xffff0000u));
                        lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    if (!valid[q]) { hi[0] = hi[1] = hi[2] = hi[3] = 0u; lo[0] = lo[1] = lo[2] = lo[3] = 0u; }
                    const uint32_t o = off[q] + (uint32_t)kc * C::KC_STRIDE;
                    *reinterpret_cast<uint4*>(dhi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    if (nplanes == 2) *reinterpret_cast<uint4*>(dlo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads, then one arrival per warp
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_full[s_hi]);
                if (nplanes == 2) mbar_arrive(&a_full[s_lo]);
            }
        }
    } else if (warp == W_PROD && !C::GEN) {
        // ===== halo producer: one TMA box per (tile, plane) =====
        if (lane == 0) {
            int cnt = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int row = tile / C::TILES, t = tile % C::TILES;
                const int y0 = (t / C::TILES_X) * C::TH, x0 = (t % C::TILES_X) * C::TW;
                const int hy0 = C::MODE == 0 ? y0 - 1 : y0, hx0 = C::MODE == 0 ? x0 - 1 : x0;   // MODE 2: parity-plane coords
                for (int pl = 0; pl < nplanes; ++pl, ++cnt) {
                    const int s = cnt % C::NA;
                    const uint32_t ph = (uint32_t)(cnt / C::NA) & 1u;
                    mbar_wait(&a_empty[s], ph ^ 1u);
                    mbar_expect_tx(&a_full[s], C::PLANE_A);
                    tma_load_5d(smA + (size_t)s * C::PLANE_A, &tmapA, &a_full[s], hx0 * 8, hy0, 0, row, pl);
                }
            }
        }
    } else if (warp == W_WGT) {
        // ===== weights: all taps, once =====
        if (lane == 0) {
            mbar_expect_tx(w_full, C::W_BYTES);
            for (int off = 0; off < C::W_BYTES; off += 4096) bulk_load(smW + off, p.wpack + off, 4096, w_full);
        }
    } else if (warp == W_MMA) {
        // ===== MMA issuer: the whole warp runs the loop converged, one elected lane issues =====
        {
            mbar_wait(w_full, 0);
            tc_fence_after();
            const uint32_t w16 = smem_u32(smW) >> 4, a16 = smem_u32(smA) >> 4;
            int it = 0, cnt = 0;
            long long t_begin = 0, w_acc = 0, w_a = 0, tw = 0;
            unsigned long long ns_begin = 0;
            if constexpr (DBG) {
                t_begin = clock64();
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
            }
            const bool issue = !DBG || !(p.dbg & 2);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int buf = it % C::NACC;
                const uint32_t aph = (uint32_t)(it / C::NACC) & 1u;
                if constexpr (DBG) tw = clock64();
                mbar_wait(&acc_empty[buf], aph ^ 1u);
                if constexpr (DBG) w_acc += clock64() - tw;
                const uint32_t d0 = tmem_base + (uint32_t)(buf * C::ACC_COLS);
                const int s0 = cnt % C::NA;
                if constexpr (DBG) tw = clock64();
                mbar_wait(&a_full[s0], (uint32_t)(cnt / C::NA) & 1u);
                ++cnt;
                if (nplanes == 1) {
                    if constexpr (DBG) w_a += clock64() - tw;
                    tc_fence_after();
                    if (issue) issue_tile<C, 2, false>(a16 + (uint32_t)s0 * (C::PLANE_A >> 4), 0u, w16, d0);
                    umma_commit(&a_empty[s0]);
                } else if (p.two_pass) {
                    if constexpr (DBG) w_a += clock64() - tw;
                    tc_fence_after();
                    if (issue) issue_tile<C, 0, true>(a16 + (uint32_t)s0 * (C::PLANE_A >> 4), 0u, w16, d0);
                    umma_commit(&a_empty[s0]);
                    const int s1 = cnt % C::NA;
                    if constexpr (DBG) tw = clock64();
                    mbar_wait(&a_full[s1], (uint32_t)(cnt / C::NA) & 1u);
                    ++cnt;
                    if constexpr (DBG) w_a += clock64() - tw;
                    tc_fence_after();
                    if (issue) issue_tile<C, 1, true>(0u, a16 + (uint32_t)s1 * (C::PLANE_A >> 4), w16, d0);
                    umma_commit(&a_empty[s1]);
                } else {
                    const int s1 = cnt % C::NA;
                    mbar_wait(&a_full[s1], (uint32_t)(cnt / C::NA) & 1u);
                    ++cnt;
                    if constexpr (DBG) w_a += clock64() - tw;
                    tc_fence_after();
                    if (issue)
                        issue_tile<C, 2, true>(a16 + (uint32_t)s0 * (C::PLANE_A >> 4), a16 + (uint32_t)s1 * (C::PLANE_A >> 4), w16, d0);
                    umma_commit(&a_empty[s0]);
                    umma_commit(&a_empty[s1]);
                }
                umma_commit(&acc_full[buf]);
            }
            if constexpr (DBG) {
                if (p.counters && lane == 0) {
                    long long* c = p.counters + (size_t)blockIdx.x * 8;
                    c[0] = clock64() - t_begin; c[1] = w_acc; c[2] = w_a; c[5] = it;
                    unsigned long long ns_end;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
                    c[6] = (long long)(ns_end - ns_begin);
                }
            }
        }
    } else if (warp < C::EPI_WARPS) {
        // ===== epilogue: lane = pixel of the tile; `half` = which half of the tile's outputs this warp owns =====
        const int ew = warp & 3;                  // the TMEM lane quarter this warp may read
        const int grp = warp >> 2;          // 0..EPI_WARPS/4-1
        const int half = grp & 1;
        const int m = ew * 32 + lane;
        const int ty = m >> 3, tx = m & 7;
        int it = 0;
        long long e_begin = 0, e_wait = 0, ew0 = 0;
        if constexpr (DBG) e_begin = clock64();
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int buf = it % C::NACC;
            const uint32_t aph = (uint32_t)(it / C::NACC) & 1u;
            const int row = tile / C::TILES, t = tile % C::TILES;
            const int y = (t / C::TILES_X) * C::TH + ty, x = (t % C::TILES_X) * C::TW + tx;
            if constexpr (DBG) ew0 = clock64();
            mbar_wait(&acc_full[buf], aph);
            if constexpr (DBG) e_wait += clock64() - ew0;
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * C::ACC_COLS);
            // the accumulator buffer goes back to the MMA issuer as soon as this warp's part of it sits in registers — before
            // the arithmetic and the stores, and without waiting for them (relaxed arrive)
            auto release_acc = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(&acc_empty[buf]);
            };
            if (DBG && (p.dbg & 1)) {
                release_acc();
            } else if (C::MODE != 1) {
                // 32 of the NPH output channels of this pixel (half = which 32, when NPH = 64)
                const int c0 = (C::NPH == 64) ? half * 32 : 0;
                const bool active = (C::NPH == 64 || half == 0) && y < C::VH && x < C::VW;
                uint32_t r[32];
                tmem_ld32(tbase + c0, r);
                if (C::CONCAT && nplanes == 2) {
                    uint32_t r2[32];
                    tmem_ld32(tbase + C::NPH + c0, r2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                }
                release_acc();
                if (active && C::OUT == OUT_NHWC_F32) {
                    float* out = reinterpret_cast<float*>(p.out) + (((size_t)row * C::VH + y) * C::VW + x) * C::NPH + c0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 v;
                        v.x = fmaxf(__uint_as_float(r[q * 4 + 0]) + sbias[c0 + q * 4 + 0], 0.0f);
                        v.y = fmaxf(__uint_as_float(r[q * 4 + 1]) + sbias[c0 + q * 4 + 1], 0.0f);
                        v.z = fmaxf(__uint_as_float(r[q * 4 + 2]) + sbias[c0 + q * 4 + 2], 0.0f);
                        v.w = fmaxf(__uint_as_float(r[q * 4 + 3]) + sbias[c0 + q * 4 + 3], 0.0f);
                        *reinterpret_cast<float4*>(out + q * 4) = v;
                    }
                } else if (active) {
                    // blocked bf16 hi/lo [plane][row][kc][H][W][8], or its parity-split form
                    // [plane][row][parity][kc][H/2][W/2][8] for a following strided conv
                    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
                    constexpr int KCO = C::NPH / 8;
                    constexpr int OH = C::OUT == OUT_PARITY ? (C::VH + 1) / 2 : C::GH, OW = C::OUT == OUT_PARITY ? (C::VW + 1) / 2 : C::GW;
                    constexpr int NPAR = C::OUT == OUT_PARITY ? 4 : 1;
                    const size_t plane = (size_t)p.nrows * NPAR * KCO * OH * OW * 8;
                    const int par = C::OUT == OUT_PARITY ? ((y & 1) * 2 + (x & 1)) : 0;
                    const int py = C::OUT == OUT_PARITY ? (y >> 1) : y, px = C::OUT == OUT_PARITY ? (x >> 1) : x;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            split2(r[q * 8 + 2 * e], r[q * 8 + 2 * e + 1], sbias[c0 + q * 8 + 2 * e], sbias[c0 + q * 8 + 2 * e + 1],
                                   1.0f, 1.0f, hi[e], lo[e]);
                        const int kc = (c0 >> 3) + q;
                        const size_t o = (((((size_t)row * NPAR + par) * KCO + kc) * OH + py) * OW + px) * 8;
                        *reinterpret_cast<uint4*>(out + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(out + plane + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            } else {
                // TMEM column slots are [00, 01, 11, 10] (host unit table): this warp takes output row parity
                // py = half and both column parities, so every store covers two neighbouring output pixels.
                constexpr int HO = 2 * C::GH, WO = 2 * C::GW;
                const int py = half;
                const int slot_l = py == 0 ? 0 : 3, slot_r = py == 0 ? 1 : 2;   // px = 0, px = 1
                const int oy = 2 * y + py, ox = 2 * x;
                if (C::NPH == 32) {
                    // Last tensor-core layer.  The next layer (ConvT 32->1, k3 s1 p1) is linear in this output,
                    // so its channel contraction is done here, in registers, per output pixel:
                    //   d[t] = sum_c relu(acc[c] + b[c]) * w4[c][t],  t = kh*3+kw
                    // and its kw sum too, across the lanes of a tile row (lane = tx + 8*ty_local):
                    //   e[kh][oy][ox] = d[kh,0][oy][ox+1] + d[kh,1][oy][ox] + d[kh,2][oy][ox-1]
                    // Only the 3 row planes leave the SM ([row][3][HO][WO] fp32, 12 B/pixel instead of 128), plus,
                    // per tile row, the two terms that cross the tile's left / right border (PROJ_EDGE floats per
                    // image after the planes: [kh][oy][tile x][0: d[kh,0] of the tile's first column, wanted by the
                    // tile to the left; 1: d[kh,2] of its last column, wanted by the tile to the right]).  The pixel
                    // kernel (k_ct4_rows) adds them and finishes sum over kh, sigmoid, entropy, reward.
                    float* out = reinterpret_cast<float*>(p.out) + (size_t)row * PROJ_ROW_FLOATS;
                    const size_t o = (size_t)oy * WO + ox;
                    uint32_t rl[32], rr[32];
                    tmem_ld32(tbase + slot_l * 32, rl);
                    tmem_ld32(tbase + slot_r * 32, rr);
                    release_acc();
                    if (DBG && (p.dbg & 8)) {       // experiment: TMEM loads only
                        if (rl[0] == 0x12345678u && rr[5] == 0x9abcdef0u) out[o] = 1.0f;
                    } else
                    {
                        // (left, right) pixel pair per packed fp32x2 FMA: 9 FFMA2 per channel instead of 18 FFMA
                        unsigned long long acc[9];
#pragma unroll
                        for (int t9 = 0; t9 < 9; ++t9) acc[t9] = 0ull;
                        const unsigned long long* w4p = reinterpret_cast<const unsigned long long*>(p.w4);
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const float vl = fmaxf(__uint_as_float(rl[c]) + sbias[c], 0.0f);
                            const float vr = fmaxf(__uint_as_float(rr[c]) + sbias[c], 0.0f);
                            const unsigned long long v = pack_f32x2(vl, vr);
#pragma unroll
                            for (int t9 = 0; t9 < 9; ++t9) acc[t9] = ffma2(v, w4p[c * 9 + t9], acc[t9]);
                        }
                        if (DBG && (p.dbg & 4) && acc[0] == 123ull) {
                        } else {
                            const int txl = lane & 7;
                            float* edge = out + 3 * HO * WO + ((size_t)oy * C::TILES_X + (t % C::TILES_X)) * 2;
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
                                const float dl0 = __uint_as_float((uint32_t)acc[kh * 3 + 0]), dr0 = __uint_as_float((uint32_t)(acc[kh * 3 + 0] >> 32));
                                const float dl1 = __uint_as_float((uint32_t)acc[kh * 3 + 1]), dr1 = __uint_as_float((uint32_t)(acc[kh * 3 + 1] >> 32));
                                const float dl2 = __uint_as_float((uint32_t)acc[kh * 3 + 2]), dr2 = __uint_as_float((uint32_t)(acc[kh * 3 + 2] >> 32));
                                const float from_left = __shfl_up_sync(0xffffffffu, dr2, 1, 8);     // d[kh,2] of pixel ox-1
                                const float from_right = __shfl_down_sync(0xffffffffu, dl0, 1, 8);  // d[kh,0] of pixel ox+2
                                float2 e;
                                e.x = (dr0 + dl1) + (txl > 0 ? from_left : 0.0f);
                                e.y = (dr1 + dl2) + (txl < 7 ? from_right : 0.0f);
                                *reinterpret_cast<float2*>(out + (size_t)kh * HO * WO + o) = e;
                                if (txl == 0) edge[(size_t)kh * HO * C::TILES_X * 2] = dl0;
                                if (txl == 7) edge[(size_t)kh * HO * C::TILES_X * 2 + 1] = dr2;
                            }
                        }
                    }
                } else {
                    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
                    const size_t plane = (size_t)p.nrows * C::NPH * HO * WO;
#pragma unroll 1
                    for (int c0 = 0; c0 < C::NPH; c0 += 32) {
                        uint32_t rl[32], rr[32];
                        tmem_ld32(tbase + slot_l * C::NPH + c0, rl);
                        tmem_ld32(tbase + slot_r * C::NPH + c0, rr);
                        if (c0 + 32 >= C::NPH) release_acc();
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t hl[4], ll[4], hr[4], lr[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float b0 = sbias[c0 + q * 8 + 2 * e], b1 = sbias[c0 + q * 8 + 2 * e + 1];
                                split2(rl[q * 8 + 2 * e], rl[q * 8 + 2 * e + 1], b0, b1, 1.0f, 1.0f, hl[e], ll[e]);
                                split2(rr[q * 8 + 2 * e], rr[q * 8 + 2 * e + 1], b0, b1, 1.0f, 1.0f, hr[e], lr[e]);
                            }
                            const int kc = (c0 >> 3) + q;
                            const size_t o = ((((size_t)row * (C::NPH / 8) + kc) * HO + oy) * WO + ox) * 8;
                            st_global_256(out + o, hl, hr);               // pixels (oy, 2x) and (oy, 2x+1): 32 B
                            st_global_256(out + plane + o, ll, lr);
                        }
                    }
                }
            }
        }
        if constexpr (DBG) {
            if (p.counters && warp == 0 && lane == 0) {
                long long* c = p.counters + (size_t)blockIdx.x * 8;
                c[3] = clock64() - e_begin; c[4] = e_wait;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------
// Dense layers as tcgen05 GEMMs: tile 128 rows x NT columns, K in chunks of 64 through a 2-3 stage
// ring.  A = activations in K-blocked bf16 hi/lo planes [plane][K/8][rows_pad][8], B = weights
// pre-packed per (n-tile, k-chunk) in the UMMA no-swizzle layout; both arrive by plain bulk copies
// (UBLKCP).  Two epilogues:
//   FC4    (256 -> 16384, NT = 256): bias + ReLU + MC-dropout bit (from the permuted bit planes) +
//          hi/lo split -> channel-blocked activation planes for ct1;
//   HIDDEN (512->512, 256->256, 576->256; NT = 128): bias + ReLU + MC-dropout bit generated in-register
//          from the keyed Philox stream + hi/lo split -> K-blocked planes for the next dense layer.
// ---------------------------------------------------------------------------------------
struct DenseParams {
    const __nv_bfloat16* a;      // K-blocked input, already offset to the launch's first row
    size_t a_kc_stride;          // elements between kc planes  (rows_pad * 8)
    size_t a_plane;              // elements between hi and lo  (K/8 * rows_pad * 8)
    const uint8_t* wpack;        // [n_tile][k_chunk] blocks of NT*256 bytes
    const float* bias;
    int32_t nrows, nprod, kchunks, ntn;
    // FC4 epilogue
    const uint32_t* mask;        // [nrows][512] dropout bits in the GEMM's column order, or null (eval mode)
    __nv_bfloat16* out;          // FC4: blocked [plane][nrows][kc 8][256 px][8]; HIDDEN: K-blocked [plane][N/8][rows_pad][8]
    // HIDDEN epilogue
    size_t out_kc_stride, out_plane;
    NoiseKey nk;
    NoiseRows nr;
    int32_t layer;               // dropout site = nr.site[set] + layer
    long long* counters;         // experiments only (DAI_TC_COUNTERS, pair FC4): per CTA {mma loop, wait acc_empty, wait own B, wait peer B, wait A, tiles}
};

// FC4 epilogue for 32 accumulator columns of one row: bias + ReLU + MC-dropout bit + hi/lo split + stores.
// FC4 column order (chosen on the host): n = ((pg*8 + kc)*4 + pl)*8 + e for pixel 4*pg + pl, channel 8*kc + e.  A
// 256-column tile is one group of 4 pixels x 64 channels and 32 consecutive columns are 4 pixels x 8 channels of one
// kc: 64 contiguous bytes of the blocked plane -> two 256-bit stores per plane.
// The inputs of that epilogue which do not depend on the accumulator, requested BEFORE the warp waits for the tile's MMAs:
// the row's dropout words of the warp's 128 columns (one 16-byte load) and, per 32-column chunk, this lane's one bias value
// (lane j holds bias[n + j]; the chunk code broadcasts it with shuffles).  Loading them inside the chunk loop left their
// global-memory latency — a scattered 2 KB-stride mask word and 32 broadcast bias loads per chunk — exposed four times per tile,
// and with two accumulator buffers that epilogue, not the operand stream, paced the kernel (measured: a version with A resident
// and half the B traffic ran no faster).
struct Fc4Pre {
    uint4 mw;
    float b[4];
};
__device__ __forceinline__ Fc4Pre fc4_prefetch(const DenseParams& p, int row, int n_base, int lane) {
    Fc4Pre f;
    f.mw = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (p.mask && row < p.nrows) f.mw = __ldg(reinterpret_cast<const uint4*>(p.mask + (size_t)row * 512 + (n_base >> 5)));
#pragma unroll
    for (int c = 0; c < 4; ++c) f.b[c] = __ldg(p.bias + n_base + c * 32 + lane);
    return f;
}
__device__ __forceinline__ void fc4_store32(const DenseParams& p, const uint32_t (&r)[32], int row, int nt, int n0, uint32_t mw, float bias_lane,
                                            uint64_t pol = 0) {
    const size_t plane = (size_t)p.nrows * 16384;
    const float s2 = p.mask ? 2.0f : 1.0f;
    const int kc = (n0 >> 5) & 7;
    uint32_t hi[4][4], lo[4][4];
#pragma unroll
    for (int pl = 0; pl < 4; ++pl)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j0 = pl * 8 + 2 * e, j1 = j0 + 1;
            split2(r[j0], r[j1], __shfl_sync(0xffffffffu, bias_lane, j0), __shfl_sync(0xffffffffu, bias_lane, j1),
                   ((mw >> j0) & 1u) ? s2 : 0.0f, ((mw >> j1) & 1u) ? s2 : 0.0f, hi[pl][e], lo[pl][e]);
        }
    if (row >= p.nrows) return;
    const size_t o = (((size_t)row * 8 + kc) * 256 + nt * 4) * 8;
    if (pol) {      // streaming output: L2 evict_first for the 64 KB/row act0 planes
        st_global_256_hint(p.out + o, hi[0], hi[1], pol);
        st_global_256_hint(p.out + o + 16, hi[2], hi[3], pol);
        st_global_256_hint(p.out + plane + o, lo[0], lo[1], pol);
        st_global_256_hint(p.out + plane + o + 16, lo[2], lo[3], pol);
        return;
    }
    st_global_256(p.out + o, hi[0], hi[1]);
    st_global_256(p.out + o + 16, hi[2], hi[3]);
    st_global_256(p.out + plane + o, lo[0], lo[1]);
    st_global_256(p.out + plane + o + 16, lo[2], lo[3]);
}

enum { EPI_FC4 = 0, EPI_HIDDEN = 1, EPI_CONV4 = 2 };

template <int NT, int EPI>
struct DenseCfg {
    static constexpr int NS = NT == 256 ? 2 : (NT == 128 ? 3 : 4);   // pipeline stages
    static constexpr int A_BYTES = 2 * 8 * 128 * 16;        // 32 KB: hi+lo, 8 kc, 128 rows
    static constexpr int B_BYTES = 2 * 8 * NT * 16;         // hi+lo, 8 kc, NT columns
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int SMEM = NS * STAGE + 1024;
};

template <int NT, int EPI>
__global__ void __launch_bounds__(384, 1) k_tc_dense(const DenseParams p) {
    using D = DenseCfg<NT, EPI>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + D::NS * D::STAGE);
    uint64_t* full = bars;                 // [NS]
    uint64_t* empty = full + D::NS;        // [NS]
    uint64_t* acc_full = empty + D::NS;    // [2]
    uint64_t* acc_empty = acc_full + 2;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < D::NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        fence_barrier_init();
    }
    if (warp == 10) tmem_alloc(tmem_slot, 2 * NT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int mtiles = (p.nrows + 127) / 128;
    const int ntiles = mtiles * p.ntn;
    const int nplanes = p.nprod == 3 ? 2 : 1;
    pdl_wait();               // programmatic dependent launch (see k_tc_conv): set-up above overlaps the previous kernel
    pdl_trigger();

    if (warp == 8) {          // producer (single-thread roles use the highest warp ids: see k_tc_conv)
        if (lane == 0) {
            int cnt = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int nt = tile / mtiles, mt = tile % mtiles;
                for (int kch = 0; kch < p.kchunks; ++kch, ++cnt) {
                    const int s = cnt % D::NS;
                    const uint32_t ph = (uint32_t)(cnt / D::NS) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    uint8_t* sa = smem + (size_t)s * D::STAGE;
                    uint8_t* sb = sa + D::A_BYTES;
                    mbar_expect_tx(&full[s], nplanes == 2 ? D::STAGE : D::STAGE / 2);
                    for (int pl = 0; pl < nplanes; ++pl)
                        for (int kc = 0; kc < 8; ++kc)
                            bulk_load(sa + (pl * 8 + kc) * 2048,
                                      p.a + pl * p.a_plane + (size_t)(kch * 8 + kc) * p.a_kc_stride + (size_t)mt * 128 * 8, 2048, &full[s]);
                    const uint8_t* wsrc = p.wpack + ((size_t)nt * p.kchunks + kch) * D::B_BYTES;
                    for (int off = 0; off < (nplanes == 2 ? D::B_BYTES : D::B_BYTES / 2); off += 4096) bulk_load(sb + off, wsrc + off, 4096, &full[s]);
                }
            }
        }
    } else if (warp == 9) {   // MMA issuer: converged warp, one elected lane issues
        {
            const uint32_t idesc = umma_idesc(NT);
            int it = 0, cnt = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t aph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(&acc_empty[buf], aph ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * NT);
                for (int kch = 0; kch < p.kchunks; ++kch, ++cnt) {
                    const int s = cnt % D::NS;
                    const uint32_t ph = (uint32_t)(cnt / D::NS) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(smem + (size_t)s * D::STAGE);
                    const uint32_t b_base = a_base + D::A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t a_hi = umma_desc(a_base + (uint32_t)(2 * k) * 2048u, 2048, 128);
                        const uint64_t b_hi = umma_desc(b_base + (uint32_t)(2 * k) * (NT * 16u), NT * 16u, 128);
                        umma_bf16(d, a_hi, b_hi, idesc, (kch == 0 && k == 0) ? 0u : 1u);
                        if (nplanes == 2) {
                            const uint64_t a_lo = umma_desc(a_base + 16384u + (uint32_t)(2 * k) * 2048u, 2048, 128);
                            const uint64_t b_lo = umma_desc(b_base + 8u * NT * 16u + (uint32_t)(2 * k) * (NT * 16u), NT * 16u, 128);
                            umma_bf16(d, a_lo, b_hi, idesc, 1u);
                            umma_bf16(d, a_hi, b_lo, idesc, 1u);
                        }
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp < 8) {    // epilogue
        const int ew = warp & 3, half = warp >> 2;       // lane quarter; column half of the tile
        const int m = ew * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            const int nt = tile / mtiles, mt = tile % mtiles;
            const int row = mt * 128 + m;
            Fc4Pre pre{};
            if (EPI == EPI_FC4) pre = fc4_prefetch(p, row, nt * NT + half * (NT / 2), lane);
            mbar_wait(&acc_full[buf], aph);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * NT + half * (NT / 2));
            uint4 drop = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            float sc = 1.0f;
            if (EPI == EPI_HIDDEN && p.nk.training && row < p.nrows) {
                int site, b;
                uint32_t sample;
                p.nr.decode(row, site, b, sample);
                // NT = 128: the tile is exactly one 128-bit Philox block of this row's mask at this layer
                drop = noise_block(p.nk, (uint32_t)(site + p.layer), (uint32_t)nt, (uint32_t)b, sample);
                sc = 2.0f;
            }
#pragma unroll 1
            for (int c0 = 0; c0 < NT / 2; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tbase + c0, r);
                if (c0 + 32 >= NT / 2) {          // last columns of this warp's share are in registers: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_relaxed(&acc_empty[buf]);
                }
                const int n0 = nt * NT + half * (NT / 2) + c0;
                if (EPI == EPI_FC4) {
                    const int c = c0 >> 5;
                    fc4_store32(p, r, row, nt, n0, c == 0 ? pre.mw.x : c == 1 ? pre.mw.y : c == 2 ? pre.mw.z : pre.mw.w,
                                c == 0 ? pre.b[0] : c == 1 ? pre.b[1] : c == 2 ? pre.b[2] : pre.b[3]);
                    continue;
                }
                if (row >= p.nrows) continue;
                {
                    const int wsel = (half * (NT / 2) + c0) >> 5;       // which 32-bit word of the Philox block
                    const uint32_t mw = wsel == 0 ? drop.x : wsel == 1 ? drop.y : wsel == 2 ? drop.z : drop.w;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j0 = q * 8 + 2 * e, j1 = j0 + 1;
                            split2(r[j0], r[j1], __ldg(p.bias + n0 + j0), __ldg(p.bias + n0 + j1),
                                   ((mw >> j0) & 1u) ? sc : 0.0f, ((mw >> j1) & 1u) ? sc : 0.0f, hi[e], lo[e]);
                        }
                        // HIDDEN: output k = column, row = row.  CONV4: GEMM row = image * 9 + output pixel, column = channel
                        // -> the encoder FC1's operand, k = pixel * 64 + channel (NHWC flatten), row = image
                        const int img = EPI == EPI_CONV4 ? row / 9 : row;
                        const int kc_out = EPI == EPI_CONV4 ? (row - img * 9) * 8 + (n0 >> 3) + q : (n0 >> 3) + q;
                        const size_t o = (size_t)kc_out * p.out_kc_stride + (size_t)img * 8;
                        *reinterpret_cast<uint4*>(p.out + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(p.out + p.out_plane + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 10) tmem_dealloc(tmem_base, 2 * NT);
}

// =======================================================================================
// ct2 -> ct3 FUSED on CTA pairs (tcgen05.mma.cta_group::2): the 256 KB/row activation between the two stride-2
// transposed convolutions never travels to HBM (SURVEY.md §8 d "fusing").
//
// Why pairs: both layers' weights must be resident (every tile re-reads all of them) — 144 KB + 72 KB does not fit one
// CTA next to the halo ring.  In a cta_group::2 MMA each CTA supplies its own 128 A rows (its own image's tile) and only
// HALF of the B rows, so a CTA holds 72 + 36 = 108 KB of weights and a 6-slot halo ring (115 KB).
// Why not DSMEM hand-off: ct3 needs (with halo) all of an image's ct2 output before most of its tiles can start, and a
// 32x32x64 hi/lo activation (256 KB) fits no shared memory.  Instead every CTA owns two 256 KB scratch images in global
// memory that it rewrites for every image it processes: 148 x 512 KB = 76 MB of lines that are overwritten while still
// resident in the 126 MB L2, read back by the same SM's TMA a few microseconds later, and never needed in HBM.
//
// Schedule (both CTAs of a pair run the same item sequence, each on its own image; rank 0 issues every MMA):
//     stage j:  ct2 tiles 0,1 of image j   then   ct3 tiles 0..7 of image j-1
// so ct3's inputs were written one stage earlier and no role ever waits for the ct2 epilogue -> store -> TMA round trip.
// TMEM: four 128-column slots used round-robin; a ct2 tile (256 columns) takes an aligned slot pair, a ct3 tile one
// slot; one stage is exactly 3 turns of the ring.
// Barriers: in the leader a_full[s] (BOTH CTAs' halo planes of slot s landed: the peer's TMA loads complete on the leader's
// barrier, cp.async.bulk.tensor .cta_group::2) and acc_empty[q] (8 epilogue warps of each CTA); per CTA a_empty[s] /
// acc_full[group][q] (multicast commits from the leader) and act2_ready[parity] (the ct2-epilogue warps: this CTA's scratch
// image is complete and visible to TMA).  (The first version announced the peer's planes through a relay thread in
// rank 1 and a second barrier per slot: 527 against 532 rollouts/s.)
// =======================================================================================
struct FusedParams {
    const uint8_t* wpack2;   // ct2 weights, [rank][half image]: per unit the rank's half of the B rows
    const uint8_t* wpack3;   // ct3 weights, same form
    const float* bias2;      // [64]
    const float* bias3;      // [32]
    void* scratch;           // act2 scratch, blocked bf16 planes [plane 2][srows][kc 8][32][32][8]
    float* out;              // ct3 projection rows [nrows][PROJ_ROW_FLOATS]
    int32_t nrows, nprod, srows;
    long long* counters;     // experiments only (DAI_TC_COUNTERS): per CTA {mma loop cycles, wait acc_empty, wait a_full, items}
    float2 w4[288];          // last deconv's weights [c 32][tap 9] duplicated (constant bank operands of the FFMA2s)
};

struct F23 {
    using C2 = Cfg<TrCt2>;
    using C3 = Cfg<TrCt3>;
    static constexpr int NA = 6;                                   // halo ring slots (one bf16 plane of one tile each)
    static constexpr int PLANE = C2::PLANE_A;                      // 19,584 B; ct3's halo plane has the same shape
    static constexpr int W2_HALF = C2::W_BYTES / 2, W3_HALF = C3::W_BYTES / 2;
    static constexpr int W_BYTES = W2_HALF + W3_HALF;              // 110,592 B per CTA
    static constexpr int SMEM_A = NA * PLANE;
    static constexpr int SMEM_BYTES = W_BYTES + SMEM_A + 1024;
    static constexpr int THREADS = 512;
    // scratch image = ct2's output as blocked bf16 planes [kc 8][32][32][8], hi and lo: what ct3's tensor-map boxes read.
    // (Storing it as ct3's eight halo tiles instead — one contiguous 19.6 KB block per (tile, plane), loaded with a single bulk
    // copy — was measured: the scattered 16-byte stores with their duplicates for the halo rows/columns made the ct2 epilogue
    // 66 % slower, and the issuer's waits for halos did not shrink: those were the relay's fences, not the TMA boxes.)
    static constexpr int SCR_IMAGE = 2 * 64 * 32 * 32 * 2;         // 262,144 B
    static_assert(C2::PLANE_A == C3::PLANE_A, "the two layers share the halo ring");
    static_assert(PLANE % 128 == 0 && W_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Wait on a local mbarrier whose arrivals come from the peer CTA, WITHOUT acquire semantics.  Everything the MMA issuer
// waits for this way is either in tensor memory (ordered by tcgen05.fence::after_thread_sync) or was written by the async
// proxy and is read by it; an .acquire.cluster wait instead compiles to a CCTL.IVALL (invalidate all of L1) on every
// success — three per tile in the issuer's loop, ~150 cycles each.  Bounded like mbar_wait.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (t0 == 0) t0 = clock64();                       // the clock is read only once the first probe has failed
        else if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// The same wait for a CONVERGED warp (the MMA issuer), leaving the loop on a warp-uniform vote: with a per-thread exit
// condition the compiler treats everything the issuer's loop carries (slot and buffer counters, operand addresses) as
// thread-varying and pays a VIADD + R2UR.BROADCAST for each descriptor of each MMA — 12 instructions per MMA on a warp that
// shares its scheduler with three epilogue warps; with uniform control flow they stay in uniform registers.
__device__ __forceinline__ void mbar_wait_cluster_warp(uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (__all_sync(0xffffffffu, ok != 0)) return;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// one non-blocking probe of the same kind, by a converged warp
__device__ __forceinline__ bool mbar_test_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return __all_sync(0xffffffffu, ok != 0) != 0;          // whole warp, uniform result (see mbar_wait_cluster_warp)
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster, without memory-ordering semantics.
// Everything handed over this way was written by the async proxy (TMA / bulk copies, whose completion the sender observed on
// its own mbarrier) or sits in tensor memory and is read by the tensor core; a .release.cluster arrive instead costs a
// MEMBAR.ALL.GPU + ERRBAR per arrival — in the relay thread that was most of the MMA issuer's wait for halos.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// the barrier at this offset in BOTH CTAs gets one arrival once every MMA issued so far has completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    if (elect_one())
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <uint32_t A_TOP, uint32_t B_TOP>
__device__ __forceinline__ void umma2_bf16_split(uint32_t lead, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(A_TOP), "n"(B_TOP), "r"(lead)
        : "memory");
}
// instruction descriptor of the pair MMA: M = 256 (128 rows per CTA), N = n (n/2 B rows from each CTA)
__host__ __device__ constexpr uint32_t umma2_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// A 5-D box load with an L2 eviction hint, issued by either CTA of a pair, its completion signalled on the LEADER's barrier (cta_group::2 lets the
// mbarrier live in the peer CTA): the leader's a_full[s] then counts both CTAs' planes and no relay thread is needed.
__device__ __forceinline__ void tma_load_5d_hint_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                                      int c4, uint64_t pol) {
    uint32_t lead_bar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lead_bar) : "r"(smem_u32(bar)), "r"(0u));
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(lead_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void st_global_f2_hint(float* p, float2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

// ---- unit tables of the pair kernel ---------------------------------------------------------------------------
// Every item of the pair kernel accumulates into ONE 128-column TMEM buffer, so all items share a 4-buffer ring:
//   L = 0: ct2, output-row parity 0 (phases 00 | 01, 64 channels each)          2 units
//   L = 1: ct2, output-row parity 1 (phases 11 | 10)                             4 units
//   L = 2: ct3, all four phases (00 | 01 | 11 | 10, 32 channels each)            4 units
// A unit = one tap shift (oy, ox) of the halo x the weight rows of the phases that read it (sub s = tap (kh[s], kw[s]),
// rows s*Cout .. s*Cout+Cout-1), accumulated at column `col`.  Splitting ct2 by row parity costs 10 % more ct2 MMA time
// (its 256-row unit becomes two 128-row MMAs at the same rate, its 128-row unit two 64-row MMAs at the per-instruction
// floor) and buys buffers of one size: ct2's store-bound epilogue then holds 128 columns, not 256 of the 512.
struct PU { int oy, ox, col, n, init, nsub; int kh[4], kw[4]; };
template <int L>
__host__ __device__ constexpr int pu_count() { return L == 0 ? 2 : 4; }
template <int L>
__host__ __device__ constexpr int pu_cout() { return L == 2 ? 32 : 64; }
template <int L>
__host__ __device__ constexpr PU pu_at(int u) {
    constexpr int C = pu_cout<L>();
    if (L == 0) {
        if (u == 0) return PU{0, 0, 0, 2 * C, 1, 2, {1, 1, 0, 0}, {1, 2, 0, 0}};
        return PU{0, 1, C, C, 0, 1, {1, 0, 0, 0}, {0, 0, 0, 0}};
    }
    if (L == 1) {
        if (u == 0) return PU{0, 0, 0, 2 * C, 1, 2, {2, 2, 0, 0}, {2, 1, 0, 0}};
        if (u == 1) return PU{0, 1, 0, C, 0, 1, {2, 0, 0, 0}, {0, 0, 0, 0}};
        if (u == 2) return PU{1, 0, 0, 2 * C, 0, 2, {0, 0, 0, 0}, {2, 1, 0, 0}};
        return PU{1, 1, 0, C, 0, 1, {0, 0, 0, 0}, {0, 0, 0, 0}};
    }
    if (u == 0) return PU{0, 0, 0, 4 * C, 1, 4, {1, 1, 2, 2}, {1, 2, 2, 1}};
    if (u == 1) return PU{0, 1, C, 2 * C, 0, 2, {1, 2, 0, 0}, {0, 0, 0, 0}};
    if (u == 2) return PU{1, 0, 2 * C, 2 * C, 0, 2, {0, 0, 0, 0}, {2, 1, 0, 0}};
    return PU{1, 1, 2 * C, C, 0, 1, {0, 0, 0, 0}, {0, 0, 0, 0}};
}
// byte offset of unit u's weight block in ONE rank's image of layer L: blocks are [plane hi|lo][kc 8][n/2][8] bf16
template <int L>
__host__ __device__ constexpr int pu_woff(int u) {
    int off = 0;
    for (int i = 0; i < u; ++i) off += pu_at<L>(i).n / 2 * 64 * 2 * 2;
    return off;
}
template <int L>
__host__ __device__ constexpr int pu_bytes() { return pu_woff<L>(pu_count<L>()); }

// The pair MMAs of one item of layer L that read ONE halo plane.  a16 / w16: shared-memory addresses >> 4 of that plane
// and of this CTA's weight image of layer L; d0: TMEM address of the item's buffer.
//   PASS 0: the products on the hi plane (A_hi*B_hi, which initialises the accumulator, and — bf16x3 — A_hi*B_lo)
//   PASS 1: the product on the lo plane (A_lo*B_hi)
// Plane-major order lets the hi plane's ring slot go back to the producer after two thirds of a tile's MMAs and gives the
// lo plane's load that much more time: the ring holds only three tiles and the issuer runs about one tile ahead of the
// tensor pipe, so with both planes waited for together the issuer spent 18 % of its time waiting for halos.
template <int L, int PASS, bool X3>
__device__ __forceinline__ void issue_item_pair(uint32_t a16, uint32_t w16, uint32_t d0) {
    constexpr int HX = 9, KC_STRIDE = 17 * 9 * 16, KCIN = 8, KSTEPS = 4;
    constexpr uint32_t A_TOP = (uint32_t)((HX * 16) >> 4) | (1u << 14);
    constexpr uint32_t B_TOP = (uint32_t)(128 >> 4) | (1u << 14);
    constexpr uint32_t A_LBO = (uint32_t)(KC_STRIDE >> 4) << 16;
    const uint32_t lead = elect_one() ? 1u : 0u;
#pragma unroll
    for (int u = 0; u < pu_count<L>(); ++u) {
        const PU un = pu_at<L>(u);
        const uint32_t nh = (uint32_t)un.n / 2u;
        const uint32_t bk = nh * 16u;                                  // bytes between kc planes of the half block
        const uint32_t b_plane = (uint32_t)KCIN * nh * 16u;            // hi -> lo plane
        const uint32_t a_off = (uint32_t)(un.oy * HX + un.ox) * 16u;
        const uint32_t woff = (uint32_t)pu_woff<L>(u);
        const uint32_t d = d0 + (uint32_t)un.col;
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
            const uint32_t a_imm = ((a_off + (uint32_t)(2 * k) * KC_STRIDE) >> 4) + A_LBO;
            const uint32_t b_imm = ((woff + (uint32_t)(2 * k) * bk) >> 4) + ((bk >> 4) << 16);
            const uint32_t a = a16 + a_imm;
            const uint32_t b_hi = w16 + b_imm, b_lo = w16 + b_imm + (b_plane >> 4);
            if (PASS == 0) {
                umma2_bf16_split<A_TOP, B_TOP>(lead, d, a, b_hi, umma2_idesc(un.n), (un.init && k == 0) ? 0u : 1u);
                if (X3) umma2_bf16_split<A_TOP, B_TOP>(lead, d, a, b_lo, umma2_idesc(un.n), 1u);
            } else {
                umma2_bf16_split<A_TOP, B_TOP>(lead, d, a, b_hi, umma2_idesc(un.n), 1u);
            }
        }
    }
}

// Item k of a stage (12 items, all roles walk the same list):
//   0: ct2 tile 0 parity 0   1: ct2 tile 0 parity 1   2: ct3 tile 0
//   3: ct2 tile 1 parity 0   4: ct2 tile 1 parity 1   5..11: ct3 tiles 1..7
// The two parities of a ct2 tile are adjacent (they share the tile's halo planes).  The ct2 items come EARLY in the stage:
// the four ct2-epilogue warps work through them back to back (store-bound, ~4.5k cycles each) and the scratch image must
// be complete well before the producer wants to prefetch the next stage's first ct3 halo (measured: with the ct2 tiles
// spread over the stage the MMA issuer waited 23 % of its time for halos).
constexpr int F23_ITEMS = 12;
constexpr int F23_LAST_CT2 = 4;
__device__ __forceinline__ int item_layer(int k) { return (k == 0 || k == 3) ? 0 : ((k == 1 || k == 4) ? 1 : 2); }
__device__ __forceinline__ int item_tile(int k) { return k < 3 ? 0 : (k < 5 ? 1 : k - 4); }

// DBG = true is the experiments build (DAI_TC_COUNTERS): in the production instantiation no role carries a counter — a
// single `if (timing && lane == 0)` inside the MMA issuer's loop makes the compiler give up on the warp being converged and
// turns every descriptor of every MMA into VIADD + predicated R2UR.BROADCAST (10.8 instead of 5.3 instructions per MMA).
template <bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F23::THREADS, 1)
k_tc_ct23(const __grid_constant__ CUtensorMap tmapIn, const __grid_constant__ CUtensorMap tmapScr, const FusedParams p) {
    using C2 = F23::C2;
    using C3 = F23::C3;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smW = smem;
    uint8_t* smA = smem + F23::W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F23::W_BYTES + F23::SMEM_A);
    uint64_t* a_full = bars;                      // [NA]
    uint64_t* a_empty = a_full + F23::NA;         // [NA]
    uint64_t* acc_full = a_empty + F23::NA;       // [2][4]: per epilogue group (0: ct2 items, 1: ct3 items) and TMEM buffer.  Each
                                                  // group waits only for its own items, and a parity wait must never skip a
                                                  // phase — so the two groups cannot share a barrier.
    uint64_t* acc_empty = acc_full + 8;           // [4]    (leader; only the MMA issuer waits, item by item)
    uint64_t* act2_ready = acc_empty + 4;         // [2]
    uint64_t* w_full = act2_ready + 2;            // [1]
    uint64_t* peer_w = w_full + 1;                // [1]    (leader)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_w + 1);
    float* sbias2 = reinterpret_cast<float*>(tmem_slot + 2);   // [64]
    float* sbias3 = sbias2 + 64;                                // [32]
    long long* whist = reinterpret_cast<long long*>(sbias3 + 32);   // [36] experiments only: the issuer's waits per item of the stage

    // warps 0-7: ct3 epilogue; 8-11: ct2 epilogue; 12: halo producer; 13: MMA issuer (leader);
    // 14: TMEM allocator; 15: weight loader.  (TMEM lane quarter of an epilogue warp = warp % 4; the single-thread roles
    // sit in the highest warp ids, which the scheduler favours.)
    constexpr int W_PROD = 12, W_MMA = 13, W_ALLOC = 14, W_WGT = 15;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
    const uint32_t rank = cluster_ctarank();
    const int nplanes = p.nprod == 3 ? 2 : 1;
    if (threadIdx.x == 0) {
        for (int i = 0; i < F23::NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&acc_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&acc_empty[i], 16);     // 8 arrivals per CTA, whichever group owns the item
        for (int i = 0; i < 2; ++i) mbar_init(&act2_ready[i], 4);     // the 4 ct2-epilogue warps, once per image
        mbar_init(w_full, 1); mbar_init(peer_w, 1);
        fence_barrier_init();
    }
    if (warp == W_ALLOC) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x < 64) sbias2[threadIdx.x] = p.bias2[threadIdx.x];
    if (threadIdx.x >= 64 && threadIdx.x < 96) sbias3[threadIdx.x - 64] = p.bias3[threadIdx.x - 64];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // images of this CTA: row(j) = 2 * (pair + j * npairs) + rank, j = 0 .. J-1 (J from the leader's rows; the last
    // image of an odd batch has no partner: rank 1 then recomputes the last row and stores nothing).
    // Stage j = ct2 items of image j (j < J) interleaved with the ct3 items of image j-1 (j >= 1).
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int pairs_total = (p.nrows + 1) / 2;
    const int J = pair < pairs_total ? (pairs_total - pair + npairs - 1) / npairs : 0;
    const int srow0 = (int)blockIdx.x * 2;           // this CTA's two scratch images
    const bool timing = DBG && p.counters != nullptr;
    if (warp != W_WGT) pdl_wait();     // programmatic dependent launch (see k_tc_conv); no early trigger: the kernel after this
                                       // one has no set-up worth hiding, its CTAs would only sit next to these for 1.4 ms
    // Every role numbers the items it walks with `seq` (items that exist in this stage, in list order): item seq uses
    // TMEM buffer seq & 3, in its (seq >> 2)-th use.
#define F23_FOR_ITEMS(j, k, L, t)                                                   \
    for (int k = 0; k < F23_ITEMS; ++k)                                             \
        if (const int L = item_layer(k); (L == 2) ? ((j) >= 1) : ((j) < J))         \
            if (const int t = item_tile(k); true)

    if (warp == W_PROD) {
        // ===== halo producer: one TMA box per plane per TILE (the two parities of a ct2 tile share it) =====
        if (lane == 0) {
            int cnt = 0;
            const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
            for (int j = 0; j <= J; ++j) {
                const int row = min(2 * (pair + j * npairs) + (int)rank, p.nrows - 1);
                const int par = (j - 1) & 1;
                bool scratch_ok = false;
                F23_FOR_ITEMS(j, k, L, t) {
                    if (L == 1) continue;                               // same halo as the preceding parity-0 item
                    if (L == 2 && !scratch_ok) {
                        mbar_wait(&act2_ready[par], (uint32_t)((j - 1) >> 1) & 1u);   // this CTA's ct2 output of image j-1 is in the scratch
                        scratch_ok = true;
                    }
                    const int y0 = L == 2 ? (t / C3::TILES_X) * C3::TH : (t / C2::TILES_X) * C2::TH;
                    const int x0 = L == 2 ? (t % C3::TILES_X) * C3::TW : (t % C2::TILES_X) * C2::TW;
                    for (int pl = 0; pl < nplanes; ++pl, ++cnt) {
                        const int s = cnt % F23::NA;
                        mbar_wait(&a_empty[s], ((uint32_t)(cnt / F23::NA) & 1u) ^ 1u);
                        uint8_t* dst = smA + (size_t)s * F23::PLANE;
                        // both CTAs' planes complete on the leader's a_full[s] (the peer's bytes may land before the leader
                        // has posted the expectation of this phase: the transaction count is signed)
                        if (rank == 0) mbar_expect_tx(&a_full[s], 2 * F23::PLANE);
                        if (L == 2) tma_load_5d_hint_pair(dst, &tmapScr, &a_full[s], x0 * 8, y0, 0, srow0 + par, pl, pol_keep);
                        else tma_load_5d_hint_pair(dst, &tmapIn, &a_full[s], x0 * 8, y0, 0, row, pl, pol_stream);
                    }
                }
            }
        }
    } else if (warp == W_WGT) {
        // ===== weights: this CTA's halves of both layers, once =====
        if (lane == 0) {
            mbar_expect_tx(w_full, F23::W_BYTES);
            const uint8_t* s2 = p.wpack2 + (size_t)rank * F23::W2_HALF;
            const uint8_t* s3 = p.wpack3 + (size_t)rank * F23::W3_HALF;
            for (int off = 0; off < F23::W2_HALF; off += 4096) bulk_load(smW + off, s2 + off, 4096, w_full);
            for (int off = 0; off < F23::W3_HALF; off += 4096) bulk_load(smW + F23::W2_HALF + off, s3 + off, 4096, w_full);
            if (rank == 1) {
                mbar_wait(w_full, 0);
                mbar_arrive_remote_relaxed(peer_w, 0);
            }
        }
    } else if (warp == W_MMA && rank == 0) {
        // ===== MMA issuer (leader): the whole warp runs the loop converged, one elected lane issues =====
        mbar_wait_cluster_warp(w_full, 0);
        mbar_wait_cluster_warp(peer_w, 0);
        tc_fence_after();
        const uint32_t w2a_16 = smem_u32(smW) >> 4, w2b_16 = smem_u32(smW + pu_bytes<0>()) >> 4,
                       w3_16 = smem_u32(smW + F23::W2_HALF) >> 4, a16 = smem_u32(smA) >> 4;
        // The loop is software-pipelined over the passes.  Before a pass is issued the barriers of the NEXT pass are probed
        // (item i's lo plane before its pass 0; item i+1's accumulator and hi plane before item i's pass 1) with a
        // non-blocking test: the ring runs two tiles ahead, so the probe normally succeeds, and the next pass then follows the
        // commits directly — the ~200 cycles of wait + commit + loop code per pass (measured: uniform over all items, i.e.
        // latency of the code path, not starvation) overlap the tail of the queued MMAs instead of draining the tensor pipe.
        // A probe that fails is waited for only AFTER everything already possible has been issued and committed: blocking
        // earlier would delay this item's commits (and, on a one-image schedule, deadlock: the first ct3 halo needs the
        // ct2 epilogue of the item being issued).
        int cnt = 0, seq = 0;
        long long t_begin = 0, w_acc = 0, w_a = 0, w_a2 = 0, tw = 0;
        if (timing) {
            if (lane == 0) for (int i = 0; i < 36; ++i) whist[i] = 0;
            __syncwarp();
            t_begin = clock64();
        }
        auto exists = [&](int j, int k) { return item_layer(k) == 2 ? j >= 1 : j < J; };
        auto advance = [&](int& j, int& k) {               // next item of the schedule; j > J when there is none
            do {
                if (++k == F23_ITEMS) { k = 0; ++j; }
            } while (j <= J && !exists(j, k));
        };
        struct Prep { int buf, s_hi, k; uint32_t d0, ah, ph_acc, ph_hi; bool acc_ok, hi_ok; };
        Prep cur{}, nxt{};
        // accumulator buffer and hi-plane slot of the next item in the schedule (no waiting); an item of ct2's parity 1 reads
        // the planes of the item before it
        auto begin = [&](int k, const Prep& prev, Prep& out) {
            out.k = k;
            out.buf = seq & 3;
            out.ph_acc = ((uint32_t)(seq >> 2) & 1u) ^ 1u;
            ++seq;
            out.d0 = tmem_base + (uint32_t)(out.buf * 128);
            out.acc_ok = false;
            if (item_layer(k) != 1) {
                out.s_hi = cnt % F23::NA;
                out.ph_hi = (uint32_t)(cnt / F23::NA) & 1u;
                ++cnt;
                out.ah = a16 + (uint32_t)out.s_hi * (F23::PLANE >> 4);
                out.hi_ok = false;
            } else {
                out.s_hi = prev.s_hi; out.ah = prev.ah; out.ph_hi = prev.ph_hi;
                out.hi_ok = true;
            }
        };
        auto probe = [&](Prep& q) {
            if (!q.acc_ok) q.acc_ok = mbar_test_cluster(&acc_empty[q.buf], q.ph_acc);
            if (!q.hi_ok) q.hi_ok = mbar_test_cluster(&a_full[q.s_hi], q.ph_hi);
        };
        auto finish = [&](Prep& q) {
            if (!q.acc_ok) {
                if (timing) tw = clock64();
                mbar_wait_cluster_warp(&acc_empty[q.buf], q.ph_acc);
                if (timing) { const long long dt = clock64() - tw; w_acc += dt; if (lane == 0) whist[24 + q.k] += dt; }
            }
            if (!q.hi_ok) {
                if (timing) tw = clock64();
                mbar_wait_cluster_warp(&a_full[q.s_hi], q.ph_hi);
                if (timing) { const long long dt = clock64() - tw; w_a += dt; if (item_layer(q.k) != 2) w_a2 += dt; if (lane == 0) whist[q.k] += dt; }
            }
            q.acc_ok = q.hi_ok = true;
        };
        int j = 0, k = -1;
        advance(j, k);
        if (j <= J) { begin(k, cur, cur); finish(cur); }
        int s_lo = 0;
        uint32_t al = 0;
        while (j <= J) {
            const int L = item_layer(k);
            bool lo_ok = true;
            uint32_t ph_lo = 0;
            if (nplanes == 2 && L != 1) {                              // this item's lo plane: probed before pass 0 goes out
                s_lo = cnt % F23::NA;
                ph_lo = (uint32_t)(cnt / F23::NA) & 1u;
                ++cnt;
                al = a16 + (uint32_t)s_lo * (F23::PLANE >> 4);
                lo_ok = mbar_test_cluster(&a_full[s_lo], ph_lo);
            }
            tc_fence_after();
            if (nplanes == 2) {
                if (L == 0) issue_item_pair<0, 0, true>(cur.ah, w2a_16, cur.d0);
                else if (L == 1) issue_item_pair<1, 0, true>(cur.ah, w2b_16, cur.d0);
                else issue_item_pair<2, 0, true>(cur.ah, w3_16, cur.d0);
            } else {
                if (L == 0) issue_item_pair<0, 0, false>(cur.ah, w2a_16, cur.d0);
                else if (L == 1) issue_item_pair<1, 0, false>(cur.ah, w2b_16, cur.d0);
                else issue_item_pair<2, 0, false>(cur.ah, w3_16, cur.d0);
            }
            if (L != 0) umma2_commit(&a_empty[cur.s_hi]);              // the hi plane is done with (after parity 1 for ct2)
            int jn = j, kn = k;
            advance(jn, kn);
            if (jn <= J) begin(kn, cur, nxt);
            if (nplanes == 2) {
                if (!lo_ok) {
                    if (timing) tw = clock64();
                    mbar_wait_cluster_warp(&a_full[s_lo], ph_lo);
                    if (timing) { const long long dt = clock64() - tw; w_a += dt; if (L != 2) w_a2 += dt; if (lane == 0) whist[12 + k] += dt; }
                }
                if (jn <= J) probe(nxt);
                tc_fence_after();
                if (L == 0) issue_item_pair<0, 1, true>(al, w2a_16, cur.d0);
                else if (L == 1) issue_item_pair<1, 1, true>(al, w2b_16, cur.d0);
                else issue_item_pair<2, 1, true>(al, w3_16, cur.d0);
                if (L != 0) umma2_commit(&a_empty[s_lo]);
            }
            umma2_commit(&acc_full[(L == 2 ? 4 : 0) + cur.buf]);
            if (jn <= J) finish(nxt);
            j = jn; k = kn; cur = nxt;
        }
        if (timing && lane == 0) {
            long long* c = p.counters + (size_t)blockIdx.x * 8;
            c[0] = clock64() - t_begin; c[1] = w_acc; c[2] = w_a; c[5] = J; c[7] = w_a2;
            long long* hst = p.counters + 8 * 512 + (size_t)(blockIdx.x >> 1) * 36;
            for (int i = 0; i < 36; ++i) hst[i] = whist[i];
        }
    } else if (warp >= 8 && warp < 12) {
        // ===== ct2 epilogue (both CTAs, 4 warps): lane = pixel of the tile, one output-row parity per item =====
        // -> this CTA's scratch image (j & 1) as blocked bf16 hi/lo planes [plane][srow][kc 8][32][32][8]
        const int ew = warp & 3;
        const int m = ew * 32 + lane;
        const int ty = m >> 3, tx = m & 7;
        __nv_bfloat16* scr = reinterpret_cast<__nv_bfloat16*>(p.scratch);
        const size_t scr_plane = (size_t)p.srows * 64 * 32 * 32;
        int seq = 0;
        const uint64_t pol_keep = l2_policy_evict_last();
        uint32_t own[4] = {0, 0, 0, 0};            // this group's uses of each TMEM buffer so far (phase of its barrier)
        long long e_begin = 0, e_wait = 0, e_fence = 0, tw = 0;
        if (timing) e_begin = clock64();
        for (int j = 0; j <= J; ++j) {
            const int srow = srow0 + (j & 1);
            F23_FOR_ITEMS(j, k, L, t) {
                const int buf = seq & 3;
                ++seq;
                if (L == 2) continue;
                if (timing) tw = clock64();
                mbar_wait(&acc_full[buf], own[buf] & 1u);
                ++own[buf];
                if (timing) e_wait += clock64() - tw;
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 128);
                const int py = L;                                        // parity 0: columns [00 | 01]; parity 1: [11 | 10]
                const int col_l = py == 0 ? 0 : 64, col_r = py == 0 ? 64 : 0;      // left = even output column (px = 0)
                const int y = (t / C2::TILES_X) * C2::TH + ty, x = (t % C2::TILES_X) * C2::TW + tx;
                const int oy = 2 * y + py, ox = 2 * x;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t rl[32], rr[32];
                    tmem_ld32(tbase + col_l + c0, rl);
                    tmem_ld32(tbase + col_r + c0, rr);
                    if (c0 == 32) {
                        // the accumulator is free as soon as its last columns sit in registers (two arrivals per warp: the
                        // barrier counts 8 per CTA for either epilogue group)
                        tc_fence_before();
                        __syncwarp();
                        if (lane < 2) {
                            if (rank == 0) mbar_arrive_relaxed(&acc_empty[buf]);
                            else mbar_arrive_remote_relaxed(&acc_empty[buf], 0);
                        }
                    }
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        uint32_t hl[4], ll[4], hr[4], lr[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float b0 = sbias2[c0 + qq * 8 + 2 * e], b1 = sbias2[c0 + qq * 8 + 2 * e + 1];
                            split2(rl[qq * 8 + 2 * e], rl[qq * 8 + 2 * e + 1], b0, b1, 1.0f, 1.0f, hl[e], ll[e]);
                            split2(rr[qq * 8 + 2 * e], rr[qq * 8 + 2 * e + 1], b0, b1, 1.0f, 1.0f, hr[e], lr[e]);
                        }
                        const int kc = (c0 >> 3) + qq;
                        const size_t o = ((((size_t)srow * 8 + kc) * 32 + oy) * 32 + ox) * 8;
                        st_global_256_hint(scr + o, hl, hr, pol_keep);               // pixels (oy, 2x) and (oy, 2x+1): 32 B
                        st_global_256_hint(scr + scr_plane + o, ll, lr, pol_keep);
                    }
                }
                if (k == F23_LAST_CT2) {
                    // last ct2 item of the image: the scratch image is ready once this warp's stores (of all four items)
                    // are visible to the async proxy — TMA reads them back
                    if (timing) tw = clock64();
                    fence_proxy_async();
                    if (timing) e_fence += clock64() - tw;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&act2_ready[j & 1]);
                }
            }
        }
        if (timing && warp == 8 && lane == 0) {
            long long* c = p.counters + (size_t)blockIdx.x * 8;
            c[6] = clock64() - e_begin - e_wait; (void)e_fence;
        }
    } else if (warp < 8) {
        // ===== ct3 epilogue (both CTAs, 8 warps): lane = pixel; warps 0-3 output-row parity 0, warps 4-7 parity 1 =====
        // -> last deconv's row planes + border terms (see k_tc_conv, OUT_PROJ)
        const int ew = warp & 3, py = (warp >> 2) & 1;
        const int m = ew * 32 + lane;
        const int ty = m >> 3, tx = m & 7;
        const int slot_l = py == 0 ? 0 : 3, slot_r = py == 0 ? 1 : 2;      // TMEM column slots [00, 01, 11, 10]
        int seq = 0;
        const uint64_t pol_stream = l2_policy_evict_first();
        uint32_t own[4] = {0, 0, 0, 0};
        long long e_begin = 0, e_wait = 0, tw = 0;
        if (timing) e_begin = clock64();
        constexpr int HO = 64, WO = 64;
        for (int j = 0; j <= J; ++j) {
            const int row = 2 * (pair + (j - 1) * npairs) + (int)rank;
            const bool live = j >= 1 && row < p.nrows;
            float* out = p.out + (size_t)(live ? row : 0) * PROJ_ROW_FLOATS;
            F23_FOR_ITEMS(j, k, L, t) {
                const int buf = seq & 3;
                ++seq;
                if (L != 2) continue;
                if (timing) tw = clock64();
                mbar_wait(&acc_full[4 + buf], own[buf] & 1u);
                ++own[buf];
                if (timing) e_wait += clock64() - tw;
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 128);
                const int y = (t / C3::TILES_X) * C3::TH + ty, x = (t % C3::TILES_X) * C3::TW + tx;
                const int oy = 2 * y + py, ox = 2 * x;
                const size_t o = (size_t)oy * WO + ox;
                uint32_t rl[32], rr[32];
                tmem_ld32(tbase + slot_l * 32, rl);
                tmem_ld32(tbase + slot_r * 32, rr);
                // the accumulator buffer is free once it sits in registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) mbar_arrive_relaxed(&acc_empty[buf]);
                    else mbar_arrive_remote_relaxed(&acc_empty[buf], 0);
                }
                unsigned long long acc[9];
#pragma unroll
                for (int t9 = 0; t9 < 9; ++t9) acc[t9] = 0ull;
                const unsigned long long* w4p = reinterpret_cast<const unsigned long long*>(p.w4);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float vl = fmaxf(__uint_as_float(rl[c]) + sbias3[c], 0.0f);
                    const float vr = fmaxf(__uint_as_float(rr[c]) + sbias3[c], 0.0f);
                    const unsigned long long v = pack_f32x2(vl, vr);
#pragma unroll
                    for (int t9 = 0; t9 < 9; ++t9) acc[t9] = ffma2(v, w4p[c * 9 + t9], acc[t9]);
                }
                const int txl = lane & 7;
                float* edge = out + 3 * HO * WO + ((size_t)oy * C3::TILES_X + (t % C3::TILES_X)) * 2;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const float dl0 = __uint_as_float((uint32_t)acc[kh * 3 + 0]), dr0 = __uint_as_float((uint32_t)(acc[kh * 3 + 0] >> 32));
                    const float dl1 = __uint_as_float((uint32_t)acc[kh * 3 + 1]), dr1 = __uint_as_float((uint32_t)(acc[kh * 3 + 1] >> 32));
                    const float dl2 = __uint_as_float((uint32_t)acc[kh * 3 + 2]), dr2 = __uint_as_float((uint32_t)(acc[kh * 3 + 2] >> 32));
                    const float from_left = __shfl_up_sync(0xffffffffu, dr2, 1, 8);     // d[kh,2] of pixel ox-1
                    const float from_right = __shfl_down_sync(0xffffffffu, dl0, 1, 8);  // d[kh,0] of pixel ox+2
                    float2 e;
                    e.x = (dr0 + dl1) + (txl > 0 ? from_left : 0.0f);
                    e.y = (dr1 + dl2) + (txl < 7 ? from_right : 0.0f);
                    if (live) {
                        st_global_f2_hint(out + (size_t)kh * HO * WO + o, e, pol_stream);
                        if (txl == 0) edge[(size_t)kh * HO * C3::TILES_X * 2] = dl0;
                        if (txl == 7) edge[(size_t)kh * HO * C3::TILES_X * 2 + 1] = dr2;
                    }
                }
            }
        }
        if (timing && warp == 0 && lane == 0) {
            long long* c = p.counters + (size_t)blockIdx.x * 8;
            c[3] = clock64() - e_begin; c[4] = e_wait;
        }
    }
#undef F23_FOR_ITEMS
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == W_ALLOC) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// =======================================================================================
// FC4 (256 -> 16384) on CTA pairs with the A operand RESIDENT: k_tc_fc4_pair.
//
// The generic kernel (k_tc_dense<256, FC4>) moves 384 KB of operands per 128 x 256 tile through L2 for 6.1k cycles of MMA
// work — 63 B/clk per SM where the L2 path delivers ~32 — and ran at 44 % tensor-pipe activity.  Here a pair of CTAs computes
// 256 rows x 256 columns per tile with ONE cta_group::2 MMA stream: each CTA keeps ITS 128 rows of A (all of K = 256, hi and
// lo planes: 128 KB) resident in shared memory for as long as the pair stays on the same row block, and streams only its HALF
// of the tile's B columns (32 KB per K = 64 stage, 3 stages): 128 KB per tile per CTA, 21 B/clk.  Tiles are dealt out as one
// contiguous run per pair in (row-block, column-tile) order, so A is reloaded about once per 16 tiles.
//   barriers per CTA: full[s] (own B stage landed), empty[s] / acc_full[b] / a_free (multicast commits of the leader),
//   a_full (own A landed); in the leader also peer_full[s], peer_a (relay from rank 1) and acc_empty[b] (8 epilogue warps of
//   each CTA).
// =======================================================================================
struct Fc4Pair {
    static constexpr int NT = 256, NS = 3;
    static constexpr int A_CHUNK = 2 * 8 * 128 * 16;        // 32 KB: one K = 64 chunk of this CTA's 128 rows, [hi|lo][kc 8][128][8]
    static constexpr int A_BYTES = 4 * A_CHUNK;             // 128 KB
    static constexpr int B_STAGE = 2 * 8 * 128 * 16;        // 32 KB: [hi|lo][kc 8][this CTA's 128 of the 256 columns][8]
    static constexpr int SMEM = A_BYTES + NS * B_STAGE + 1024;
    static_assert(SMEM <= 232448, "shared memory budget");
};

__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (elect_one())
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <bool DBG>      // true: experiments build with the issuer's cycle counters (see k_tc_ct23)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) k_tc_fc4_pair(const DenseParams p) {
    using P = Fc4Pair;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smAres = smem;
    uint8_t* smB = smem + P::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::A_BYTES + P::NS * P::B_STAGE);
    uint64_t* full = bars;                    // [NS]
    uint64_t* peer_full = full + P::NS;       // [NS]  (leader)
    uint64_t* empty = peer_full + P::NS;      // [NS]
    uint64_t* acc_full = empty + P::NS;       // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]   (leader)
    uint64_t* a_full = acc_empty + 2;         // [1]
    uint64_t* peer_a = a_full + 1;            // [1]   (leader)
    uint64_t* a_free = peer_a + 1;            // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        for (int i = 0; i < P::NS; ++i) { mbar_init(&full[i], 1); mbar_init(&peer_full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
        mbar_init(a_full, 1); mbar_init(peer_a, 1); mbar_init(a_free, 1);
        fence_barrier_init();
    }
    if (warp == 10) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * P::NT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tiles in (row block of 256, column tile) order; this pair's contiguous run [t0, t1)
    const int mtiles = (p.nrows + 127) / 128, mpairs = (mtiles + 1) / 2;
    const long long total = (long long)mpairs * p.ntn;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int t0 = (int)(total * pair / npairs), t1 = (int)(total * (pair + 1) / npairs);
    const int nplanes = p.nprod == 3 ? 2 : 1;
    pdl_wait();               // programmatic dependent launch (see k_tc_conv)
    pdl_trigger();

    if (warp == 8) {          // producer (both CTAs): own A rows when the row block changes, own half of the B columns per stage
        if (lane == 0) {
            int cnt = 0, loads = 0, cur_mp = -1;
            for (int tile = t0; tile < t1; ++tile) {
                const int mp = tile / p.ntn, nt = tile % p.ntn;
                if (mp != cur_mp) {
                    cur_mp = mp;
                    if (loads > 0) mbar_wait(a_free, (uint32_t)(loads - 1) & 1u);      // every MMA on the previous rows has completed
                    ++loads;
                    const int mt = 2 * mp + (int)rank;
                    mbar_expect_tx(a_full, nplanes == 2 ? P::A_BYTES : P::A_BYTES / 2);
                    for (int kch = 0; kch < 4; ++kch)
                        for (int pl = 0; pl < nplanes; ++pl)
                            for (int kc = 0; kc < 8; ++kc)
                                bulk_load(smAres + kch * P::A_CHUNK + (pl * 8 + kc) * 2048,
                                          p.a + pl * p.a_plane + (size_t)(kch * 8 + kc) * p.a_kc_stride + (size_t)mt * 128 * 8, 2048, a_full);
                }
                for (int kch = 0; kch < 4; ++kch, ++cnt) {
                    const int s = cnt % P::NS;
                    mbar_wait(&empty[s], (((uint32_t)(cnt / P::NS)) & 1u) ^ 1u);
                    uint8_t* sb = smB + (size_t)s * P::B_STAGE;
                    mbar_expect_tx(&full[s], nplanes == 2 ? P::B_STAGE : P::B_STAGE / 2);
                    const uint8_t* blk = p.wpack + ((size_t)nt * 4 + kch) * 65536;      // [hi 32 KB | lo 32 KB], each [kc 8][256][8]
                    for (int pl = 0; pl < nplanes; ++pl)
                        for (int kc = 0; kc < 8; ++kc)
                            bulk_load(sb + (pl * 8 + kc) * 2048, blk + (size_t)pl * 32768 + ((size_t)kc * 256 + 128 * rank) * 16, 2048, &full[s]);
                }
            }
        }
    } else if (warp == 9 && rank == 1) {   // relay (peer): tell the leader what has landed here
        if (lane == 0) {
            int cnt = 0, loads = 0, cur_mp = -1;
            for (int tile = t0; tile < t1; ++tile) {
                const int mp = tile / p.ntn;
                if (mp != cur_mp) {
                    cur_mp = mp;
                    mbar_wait(a_full, (uint32_t)loads & 1u);
                    ++loads;
                    mbar_arrive_remote_relaxed(peer_a, 0);
                }
                for (int kch = 0; kch < 4; ++kch, ++cnt) {
                    const int s = cnt % P::NS;
                    mbar_wait(&full[s], ((uint32_t)(cnt / P::NS)) & 1u);
                    mbar_arrive_remote_relaxed(&peer_full[s], 0);
                }
            }
        }
    } else if (warp == 9) {   // MMA issuer (leader): converged warp, one elected lane issues
        const uint32_t idesc = umma2_idesc(P::NT);
        const uint32_t a_res = smem_u32(smAres);
        int it = 0, cnt = 0, loads = 0, cur_mp = -1;
        const bool timing = DBG && p.counters != nullptr;
        long long t_begin = 0, w_acc = 0, w_b = 0, w_pb = 0, w_a = 0, tw = 0;
        if (timing) t_begin = clock64();
        for (int tile = t0; tile < t1; ++tile, ++it) {
            const int mp = tile / p.ntn;
            const int buf = it & 1;
            if (timing) tw = clock64();
            mbar_wait_cluster(&acc_empty[buf], (((uint32_t)(it >> 1)) & 1u) ^ 1u);
            if (timing) { w_acc += clock64() - tw; tw = clock64(); }
            if (mp != cur_mp) {
                cur_mp = mp;
                mbar_wait(a_full, (uint32_t)loads & 1u);
                mbar_wait_cluster(peer_a, (uint32_t)loads & 1u);
                ++loads;
            }
            if (timing) w_a += clock64() - tw;
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)(buf * P::NT);
            for (int kch = 0; kch < 4; ++kch, ++cnt) {
                const int s = cnt % P::NS;
                const uint32_t ph = ((uint32_t)(cnt / P::NS)) & 1u;
                if (timing) tw = clock64();
                mbar_wait(&full[s], ph);
                if (timing) { w_b += clock64() - tw; tw = clock64(); }
                mbar_wait_cluster(&peer_full[s], ph);
                if (timing) w_pb += clock64() - tw;
                tc_fence_after();
                const uint32_t a_base = a_res + (uint32_t)(kch * P::A_CHUNK);
                const uint32_t b_base = smem_u32(smB + (size_t)s * P::B_STAGE);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t a_hi = umma_desc(a_base + (uint32_t)(2 * k) * 2048u, 2048, 128);
                    const uint64_t b_hi = umma_desc(b_base + (uint32_t)(2 * k) * 2048u, 2048, 128);
                    umma2_bf16(d, a_hi, b_hi, idesc, (kch == 0 && k == 0) ? 0u : 1u);
                    if (nplanes == 2) {
                        const uint64_t a_lo = umma_desc(a_base + 16384u + (uint32_t)(2 * k) * 2048u, 2048, 128);
                        const uint64_t b_lo = umma_desc(b_base + 16384u + (uint32_t)(2 * k) * 2048u, 2048, 128);
                        umma2_bf16(d, a_lo, b_hi, idesc, 1u);
                        umma2_bf16(d, a_hi, b_lo, idesc, 1u);
                    }
                }
                umma2_commit(&empty[s]);
            }
            umma2_commit(&acc_full[buf]);
            // last tile on these rows: once its MMAs are through, the resident A may be replaced
            if (tile + 1 < t1 && (tile + 1) / p.ntn != mp) umma2_commit(a_free);
        }
        if (timing && lane == 0) {
            long long* c = p.counters + (size_t)blockIdx.x * 8;
            c[0] = clock64() - t_begin; c[1] = w_acc; c[2] = w_b; c[3] = w_pb; c[4] = w_a; c[5] = t1 - t0;
        }
    } else if (warp < 8) {    // epilogue (both CTAs): this CTA's 128 rows x 256 columns
        const int ew = warp & 3, half = warp >> 2;
        const int m = ew * 32 + lane;
        const uint64_t pol_out = p.layer == 1 ? l2_policy_evict_first() : 0;     // DenseParams::layer (unused by FC4) carries the switch
        int it = 0;
        for (int tile = t0; tile < t1; ++tile, ++it) {
            const int buf = it & 1;
            const int mp = tile / p.ntn, nt = tile % p.ntn;
            const int row = (2 * mp + (int)rank) * 128 + m;
            const Fc4Pre pre = fc4_prefetch(p, row, nt * P::NT + half * (P::NT / 2), lane);
            mbar_wait(&acc_full[buf], ((uint32_t)(it >> 1)) & 1u);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * P::NT + half * (P::NT / 2));
#pragma unroll 1
            for (int c0 = 0; c0 < P::NT / 2; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tbase + c0, r);
                if (c0 + 32 >= P::NT / 2) {       // this warp's share of the accumulator sits in registers: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (rank == 0) mbar_arrive_relaxed(&acc_empty[buf]);
                        else mbar_arrive_remote_relaxed(&acc_empty[buf], 0);
                    }
                }
                const int c = c0 >> 5;
                fc4_store32(p, r, row, nt, nt * P::NT + half * (P::NT / 2) + c0,
                            c == 0 ? pre.mw.x : c == 1 ? pre.mw.y : c == 2 ? pre.mw.z : pre.mw.w,
                            c == 0 ? pre.b[0] : c == 1 ? pre.b[1] : c == 2 ? pre.b[2] : pre.b[3], pol_out);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 10) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * P::NT) : "memory");
}

using CfgCt1 = Cfg<TrCt1>;   // 144 KB of weights + 3 x 22.5 KB halo planes
using CfgCt2 = Cfg<TrCt2>;   // 144 KB of weights + 4 x 19.1 KB halo planes (2 tiles in flight)
using CfgCt3 = Cfg<TrCt3>;   //  72 KB of weights + 7 x 19.1 KB halo planes (3.5 tiles in flight)
using CfgQc2 = Cfg<TrQc2>;   //  36 KB of weights + 4 x 38.3 KB halo planes (4 parities x 4 kc)
using CfgQc3 = Cfg<TrQc3>;   //  72 KB of weights + 4 x 38.3 KB halo planes
using CfgQc2G = Cfg<TrQc2G>; //  conv1 -> conv2 in one kernel: as CfgQc2, the halo planes generated in shared memory

// ---------------------------------------------------------------------------------------
// host: weight packing, tensor maps, launches
// ---------------------------------------------------------------------------------------
struct LayerPack {
    uint8_t* wpack = nullptr;    // device
    uint8_t* wpair = nullptr;    // device, fused pair kernel: [rank 2][half image] (each CTA's half of every unit's B rows)
    Unit units[MAX_UNITS];
    int nunits = 0;
};

struct TcImpl {
    LayerPack ct1, ct2, ct3, qc2, qc3;
    float w4[288];               // po_net.19.weight as [c][tap]
    float genw[320];             // qs_net.0.weight as [tap][c] (288), then qs_net.0.bias (32): the fused conv1 of k_tc_conv<CfgQc2G>
    uint8_t* dense_w[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // TC_PS1 .. TC_QS2, TC_QC4
    float* dense_b[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int dense_k[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dense_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t* fc4_wpack = nullptr;
    float* fc4_bias = nullptr;     // po_net.9.bias in the tensor-core FC4's column order
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    bool attrs_set = false;
};

struct Sub { int kh, kw; };

// One weight block = [plane hi|lo][kc Cin/8][n][8] bf16 with n = sub * Cout + co: the UMMA no-swizzle K-major layout
// of an (n x Cin) operand (concat: [kc][2n rows: hi rows then lo rows][8]).  ConvTranspose2d weights are
// (Cin,Cout,3,3), Conv2d weights (Cout,Cin,3,3).
// Packing is described, not performed, on the host: dst[i] holds the index of the fp32 source element that bf16
// element i of the packed image is made from (REPACK_LO set: its lo part), and a gather kernel (k_repack, dai_simt.cu)
// builds the image on the device whenever that weight tensor changes.
void map_block(int Cin, int Cout, const Sub* subs, int nsub, uint32_t* dst, bool concat, bool conv_layout) {
    const int n = nsub * Cout, KC = Cin / 8;
    for (int s = 0; s < nsub; ++s)
        for (int co = 0; co < Cout; ++co)
            for (int ci = 0; ci < Cin; ++ci) {
                const uint32_t wi = (uint32_t)(conv_layout ? (((size_t)co * Cin + ci) * 3 + subs[s].kh) * 3 + subs[s].kw
                                                           : (((size_t)ci * Cout + co) * 3 + subs[s].kh) * 3 + subs[s].kw);
                const int nn = s * Cout + co, kc = ci >> 3, e = ci & 7;
                if (concat) {
                    dst[((size_t)kc * 2 * n + nn) * 8 + e] = wi;
                    dst[((size_t)kc * 2 * n + n + nn) * 8 + e] = wi | REPACK_LO;
                } else {
                    const size_t o = ((size_t)kc * n + nn) * 8 + e;
                    dst[o] = wi;
                    dst[(size_t)KC * n * 8 + o] = wi | REPACK_LO;
                }
            }
}

// sub-pixel decomposition of ConvTranspose2d(k3, s2, p1, op1): output (2y+py, 2x+px) reads input
// (y+dy, x+dx) with kh = 1 (py = 0); kh = 0 for dy = 1 and kh = 2 for dy = 0 (py = 1); same in x.
inline int k_of(int parity, int d) { return parity == 0 ? 1 : (d == 1 ? 0 : 2); }

int build_layer(const char* key, int mode, int Cin, int Cout, bool grouped, bool concat, LayerPack* lp,
                std::vector<RepackJob>* jobs) {
    std::vector<uint32_t> host((size_t)9 * Cout * Cin * 2, REPACK_NONE);   // 9 taps * Cout * Cin * 2 planes
    const int KC = Cin / 8;
    int nu = 0;
    size_t off = 0;                                            // in uint16 elements
    auto add = [&](int oy, int ox, int col, const Sub* subs, int nsub, int init, int kc0) {
        Unit& u = lp->units[nu++];
        u.oy = (int16_t)oy; u.ox = (int16_t)ox; u.col = (int16_t)col; u.n = (int16_t)(nsub * Cout); u.init = (int16_t)init;
        u.kc0 = (int16_t)kc0;
        u.woff = (int32_t)(off * 2);
        map_block(Cin, Cout, subs, nsub, host.data() + off, concat, mode == 2);
        off += (size_t)nsub * Cout * Cin * 2;
    };
    if (mode == 0) {
        // convT s1 p1: oy = iy - 1 + kh; halo origin (y0-1, x0-1) => local row = ty + 2 - kh
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                Sub s{kh, kw};
                add(2 - kh, 2 - kw, 0, &s, 1, nu == 0, 0);
            }
    } else if (mode == 2) {
        // conv s2 valid: input (2*oy+kh, 2*ox+kw) = parity plane (kh&1, kw&1) at (oy + (kh>>1), ox + (kw>>1))
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                Sub s{kh, kw};
                add(kh >> 1, kw >> 1, 0, &s, 1, nu == 0, ((kh & 1) * 2 + (kw & 1)) * KC);
            }
    } else if (!grouped) {
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                bool first = true;
                for (int dy = 0; dy <= py; ++dy)
                    for (int dx = 0; dx <= px; ++dx) {
                        Sub s{k_of(py, dy), k_of(px, dx)};
                        add(dy, dx, (py * 2 + px) * Cout, &s, 1, first, 0);
                        first = false;
                    }
            }
    } else {
        // column slots [00, 01, 11, 10]; every shift of the halo feeds all phases that read it in one MMA
        const Sub g00[4] = {{k_of(0, 0), k_of(0, 0)}, {k_of(0, 0), k_of(1, 0)}, {k_of(1, 0), k_of(1, 0)}, {k_of(1, 0), k_of(0, 0)}};
        add(0, 0, 0, g00, 4, 1, 0);
        const Sub g01[2] = {{k_of(0, 0), k_of(1, 1)}, {k_of(1, 0), k_of(1, 1)}};        // phases 01, 11
        add(0, 1, 1 * Cout, g01, 2, 0, 0);
        const Sub g10[2] = {{k_of(1, 1), k_of(1, 0)}, {k_of(1, 1), k_of(0, 0)}};        // phases 11, 10
        add(1, 0, 2 * Cout, g10, 2, 0, 0);
        const Sub g11[1] = {{k_of(1, 1), k_of(1, 1)}};                                  // phase 11
        add(1, 1, 2 * Cout, g11, 1, 0, 0);
    }
    lp->nunits = nu;
    jobs->push_back(RepackJob{key, std::move(host), 1, false, reinterpret_cast<void**>(&lp->wpack)});
    return 0;
}

// blocked bf16 activation tensor [plane 2][rows][kc 8][H][W][8] -> 5-D map (W*8, H, kc, rows, plane), box = halo
int make_map(TcImpl* im, const void* base, int rows, int H, int W, int kplanes, int HX, int HY, CUtensorMap* map,
             std::string* err) {
    cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)kplanes, (cuuint64_t)rows, 2};
    cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)kplanes * H * W * 16,
                             (cuuint64_t)rows * kplanes * H * W * 16};
    cuuint32_t box[5] = {(cuuint32_t)HX * 8, (cuuint32_t)HY, (cuuint32_t)kplanes, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = im->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        *err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
        return -1;
    }
    return 0;
}

template <class C>
int launch_conv(TcImpl* im, const LayerPack& lp, const float* bias, int precision, const void* in, void* out, int nrows,
                cudaStream_t st, std::string* err, const float* w4 = nullptr, const float* gen_img = nullptr) {
    CUtensorMap map;
    if (make_map(im, in, nrows, C::PH, C::PW, C::KPLANES, C::HX, C::HY, &map, err) != 0) return -1;
    ConvParams p{};
    p.gen_img = gen_img;
    for (int i = 0; i < lp.nunits; ++i) p.units[i] = lp.units[i];
    p.nunits = lp.nunits; p.nrows = nrows; p.nprod = precision == DAI_PREC_BF16X1 ? 1 : 3;
    p.wpack = lp.wpack; p.bias = bias; p.out = out;
    if (w4) for (int i = 0; i < 288; ++i) p.w4[i] = make_float2(w4[i], w4[i]);
    if (C::GEN)      // conv1's weights / bias as channel pairs [tap 9][16], [16]: constant-bank operands of the generators' packed FMAs
        for (int i = 0; i < 160; ++i) p.w4[i] = make_float2(im->genw[2 * i], im->genw[2 * i + 1]);
    for (int i = 0; i < C::NUNITS; ++i) {
        const Unit a = lp.units[i], b = unit_at<C>(i);
        if (lp.nunits != C::NUNITS || a.oy != b.oy || a.ox != b.ox || a.col != b.col || a.n != b.n || a.init != b.init ||
            a.kc0 != b.kc0 || a.woff != b.woff) {
            *err = "tensor-core layer: packed unit table differs from the compile-time table";
            return -1;
        }
    }
    static const int env_two_pass = getenv("DAI_TC_TWO_PASS") ? atoi(getenv("DAI_TC_TWO_PASS")) : -1;
    p.two_pass = env_two_pass >= 0 ? ((env_two_pass >> C::ID) & 1) : (C::TWO_PASS ? 1 : 0);
    static const int env_dbg = getenv("DAI_TC_DBG") ? atoi(getenv("DAI_TC_DBG")) : 0;      // experiments only
    p.dbg = env_dbg;
    static long long* dbg_counters = nullptr;
    static const bool want_counters = getenv("DAI_TC_COUNTERS") != nullptr;                // experiments only
    if (want_counters) {
        if (!dbg_counters) cudaMalloc(&dbg_counters, 8 * 8 * 512);
        cudaMemsetAsync(dbg_counters, 0, 8 * 8 * 512, st);
        p.counters = dbg_counters;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = nrows * C::TILES;
    const int grid = ntiles < sms ? ntiles : sms;
    if (env_dbg || want_counters) {
        static bool attr_dbg = false;
        if (!attr_dbg) { cudaFuncSetAttribute(k_tc_conv<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES); attr_dbg = true; }
        k_tc_conv<C, true><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(map, p);
    } else {
        launch_dep(k_tc_conv<C, false>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, true, map, p);
    }
    if (want_counters) {
        static int printed = 0;
        cudaStreamSynchronize(st);
        long long h[8 * 512];
        cudaMemcpy(h, dbg_counters, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost);
        double a[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < grid; ++i) for (int j = 0; j < 7; ++j) a[j] += (double)h[i * 8 + j] / grid;
        if (printed++ % 23 == 3)
            fprintf(stderr, "[tc counters] mode %d nph %d rows %d: per CTA cycles: mma loop %.0f (wait acc_empty %.0f, wait a_full %.0f) | "
                    "epilogue loop %.0f (wait acc_full %.0f) | tiles %.1f | mma loop %.1f us => SM clock %.0f MHz\n", C::MODE, C::NPH, nrows, a[0], a[1], a[2], a[3], a[4], a[5],
                    a[6] * 1e-3, a[6] > 0 ? a[0] / a[6] * 1e3 : 0.0);
    }
    return 1;
}

__global__ void k_to_blocked(const float* __restrict__ in, int rows, int hw, int C, __nv_bfloat16* __restrict__ out) {
    const size_t n = (size_t)rows * hw * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t px = (i / C) % hw, row = i / ((size_t)C * hw);
        __nv_bfloat16 hi, lo;
        split_bf16(in[i], hi, lo);
        const size_t o = ((row * (C / 8) + (c >> 3)) * hw + px) * 8 + (c & 7);
        out[o] = hi;
        out[n + o] = lo;
    }
}

__global__ void k_from_blocked(const __nv_bfloat16* __restrict__ in, int rows, int hw, int C, float* __restrict__ out) {
    const size_t n = (size_t)rows * hw * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t px = (i / C) % hw, row = i / ((size_t)C * hw);
        const size_t o = ((row * (C / 8) + (c >> 3)) * hw + px) * 8 + (c & 7);
        out[i] = __bfloat162float(in[o]) + __bfloat162float(in[n + o]);
    }
}

}  // namespace

int tc_to_blocked(const float* nhwc, int rows, int hw, int C, void* blocked, cudaStream_t st) {
    k_to_blocked<<<1024, 256, 0, st>>>(nhwc, rows, hw, C, static_cast<__nv_bfloat16*>(blocked));
    return 1;
}

int tc_from_blocked(const void* blocked, int rows, int hw, int C, float* nhwc, cudaStream_t st) {
    k_from_blocked<<<1024, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(blocked), rows, hw, C, nhwc);
    return 1;
}

namespace {
// Weight image of layer L of the pair kernel for both ranks: [rank][unit blocks], rank r holding rows
// [r*n/2, (r+1)*n/2) of every unit as [plane hi|lo][kc 8][n/2][8] (pu_at / pu_woff, device side).  ConvTranspose2d
// weights are (Cin = 64, Cout, 3, 3).
template <int L>
void map_pair_layer(uint32_t* dst, size_t rank_stride_elems) {
    constexpr int Cout = pu_cout<L>(), Cin = 64, KC = 8;
    for (int u = 0; u < pu_count<L>(); ++u) {
        const PU un = pu_at<L>(u);
        const int nh = un.n / 2;
        const size_t base = (size_t)pu_woff<L>(u) / 2;                 // elements
        for (int nn = 0; nn < un.n; ++nn) {
            const int sub = nn / Cout, co = nn % Cout, r = nn / nh, loc = nn % nh;
            for (int ci = 0; ci < Cin; ++ci) {
                const uint32_t wi = (uint32_t)((((size_t)ci * Cout + co) * 3 + un.kh[sub]) * 3 + un.kw[sub]);
                const size_t o = (size_t)r * rank_stride_elems + base + ((size_t)(ci >> 3) * nh + loc) * 8 + (ci & 7);
                dst[o] = wi;
                dst[o + (size_t)KC * nh * 8] = wi | REPACK_LO;
            }
        }
    }
}
}  // namespace

namespace {
// torch Linear weight (N,K) -> [n_tile = N/NT][k_chunk = K/64] blocks of [plane hi|lo][kc 8][NT][8] bf16.
// src(n, k) = index of the source element for GEMM column n, contraction index k.
template <class F>
void plan_dense(TcImpl* im, int which, const char* wkey, const char* bkey, int N, int K, F src, std::vector<RepackJob>* jobs, int NT = 128) {
    const int ntn = N / NT, kch = K / 64;
    const size_t blk = (size_t)2 * 8 * NT * 8;               // elements per block
    std::vector<uint32_t> host((size_t)ntn * kch * blk, REPACK_NONE);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const uint32_t wi = (uint32_t)src(n, k);
            const size_t o = ((size_t)(n / NT) * kch + k / 64) * blk + ((size_t)((k >> 3) & 7) * NT + (n % NT)) * 8 + (k & 7);
            host[o] = wi;
            host[o + (size_t)8 * NT * 8] = wi | REPACK_LO;
        }
    jobs->push_back(RepackJob{wkey, std::move(host), 1, false, reinterpret_cast<void**>(&im->dense_w[which])});
    jobs->push_back(RepackJob{bkey, {}, 0, true, reinterpret_cast<void**>(&im->dense_b[which])});     // the bias is used as stored
    im->dense_k[which] = K; im->dense_n[which] = N;
}
}  // namespace

int tc_plan_weights(TcWeights* out, std::vector<RepackJob>* jobs, std::string* err) {
    TcImpl* im = static_cast<TcImpl*>(out->impl);
    if (!im) { im = new TcImpl(); out->impl = im; }
    if (!im->encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            *err = "cuTensorMapEncodeTiled not available from the driver";
            return -1;
        }
        im->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    if (!im->attrs_set) {
        if (cudaFuncSetAttribute(k_tc_conv<CfgCt1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgCt1::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_conv<CfgCt2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgCt2::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_conv<CfgCt3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgCt3::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_conv<CfgQc2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgQc2::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_conv<CfgQc3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgQc3::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_conv<CfgQc2G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgQc2G::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_ct23<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F23::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_ct23<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F23::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_fc4_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fc4Pair::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_fc4_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fc4Pair::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_dense<256, EPI_FC4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseCfg<256, EPI_FC4>::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_dense<128, EPI_HIDDEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseCfg<128, EPI_HIDDEN>::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_tc_dense<64, EPI_CONV4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DenseCfg<64, EPI_CONV4>::SMEM) != cudaSuccess) {
            *err = std::string("cudaFuncSetAttribute(max dynamic smem): ") + cudaGetErrorString(cudaGetLastError());
            return -1;
        }
        im->attrs_set = true;
    }
    build_layer("po_net.13.weight", 0, 64, 64, false, CfgCt1::CONCAT, &im->ct1, jobs);
    build_layer("po_net.15.weight", 1, 64, 64, true, CfgCt2::CONCAT, &im->ct2, jobs);
    build_layer("po_net.17.weight", 1, 64, 32, true, CfgCt3::CONCAT, &im->ct3, jobs);
    {   // the fused ct2 -> ct3 pair kernel's images
        std::vector<uint32_t> m2((size_t)F23::W2_HALF, REPACK_NONE), m3((size_t)F23::W3_HALF, REPACK_NONE);   // 2 ranks x half bytes / 2 B
        map_pair_layer<0>(m2.data(), F23::W2_HALF / 2);
        map_pair_layer<1>(m2.data() + pu_bytes<0>() / 2, F23::W2_HALF / 2);
        map_pair_layer<2>(m3.data(), F23::W3_HALF / 2);
        jobs->push_back(RepackJob{"po_net.15.weight", std::move(m2), 1, false, reinterpret_cast<void**>(&im->ct2.wpair)});
        jobs->push_back(RepackJob{"po_net.17.weight", std::move(m3), 1, false, reinterpret_cast<void**>(&im->ct3.wpair)});
    }
    {
        static const std::string wk[7] = {"ps_net.3.weight", "ps_net.6.weight", "po_net.3.weight", "po_net.6.weight", "qs_net.9.weight",
                                          "qs_net.12.weight", "qs_net.15.weight"};
        static const std::string bk[7] = {"ps_net.3.bias", "ps_net.6.bias", "po_net.3.bias", "po_net.6.bias", "qs_net.9.bias",
                                          "qs_net.12.bias", "qs_net.15.bias"};
        const int N[7] = {512, 512, 256, 256, 256, 256, 256}, K[7] = {512, 512, 256, 256, 576, 256, 256};
        for (int i = 0; i < 7; ++i) {
            const int Ki = K[i];
            if (i == TC_QS0)      // encoder FC1: GEMM k = pixel*64 + c (NHWC flatten) <- reference c*9 + pixel
                plan_dense(im, i, wk[i].c_str(), bk[i].c_str(), N[i], Ki, [Ki](int n, int k) { return (size_t)n * Ki + (size_t)(k & 63) * 9 + (k >> 6); }, jobs);
            else
                plan_dense(im, i, wk[i].c_str(), bk[i].c_str(), N[i], Ki, [Ki](int n, int k) { return (size_t)n * Ki + k; }, jobs);
        }
    }
    // encoder conv4 (Conv2d 64->64, k3 s2, 7x7 -> 3x3) as a GEMM over im2col rows: B[n = co][k = (kh*3+kw)*64 + ci] = W[co][ci][tap]
    plan_dense(im, TC_QC4, "qs_net.6.weight", "qs_net.6.bias", 64, 576,
               [](int co, int k) { return ((size_t)co * 64 + (k & 63)) * 9 + (k >> 6); }, jobs, 64);
    build_layer("qs_net.2.weight", 2, 32, 32, false, false, &im->qc2, jobs);
    build_layer("qs_net.4.weight", 2, 32, 64, false, false, &im->qc3, jobs);
    {   // FC4 (16384, 256): reference row e = c*256 + p -> column n = ((pg*8 + kc)*4 + pl)*8 + ce for pixel 4*pg + pl,
        // channel 8*kc + ce; blocks [n_tile][k_chunk] of [plane][kc 8][256 n][8]
        std::vector<uint32_t> host((size_t)16384 * 256 * 2, REPACK_NONE), bias((size_t)16384);
        for (int n = 0; n < 16384; ++n) {
            const int ce = n & 7, pl = (n >> 3) & 3, kc8 = (n >> 5) & 7, pg = n >> 8;
            const int px = pg * 4 + pl, c = kc8 * 8 + ce, e = c * 256 + px;
            bias[n] = (uint32_t)e;
            const int nt = n >> 8, nl = n & 255;
            for (int k = 0; k < 256; ++k) {
                const uint32_t wi = (uint32_t)((size_t)e * 256 + k);
                const int kch = k >> 6, kc = (k >> 3) & 7, ke = k & 7;
                const size_t blk = ((size_t)nt * 4 + kch) * ((DenseCfg<256, EPI_FC4>::B_BYTES / 2));
                const size_t o = blk + ((size_t)kc * 256 + nl) * 8 + ke;
                host[o] = wi;
                host[o + (size_t)8 * 256 * 8] = wi | REPACK_LO;
            }
        }
        jobs->push_back(RepackJob{"po_net.9.weight", std::move(host), 1, false, reinterpret_cast<void**>(&im->fc4_wpack)});
        jobs->push_back(RepackJob{"po_net.9.bias", std::move(bias), 0, false, reinterpret_cast<void**>(&im->fc4_bias)});
    }
    return 0;
}

// po_net.19.weight (Cin 32, Cout 1, 3, 3) = [c][tap]: travels as a kernel parameter of ct3 (constant bank)
void tc_set_conv1(TcWeights* w, const float* w_raw /*[32][9]*/, const float* bias /*[32]*/) {
    TcImpl* im = static_cast<TcImpl*>(w->impl);
    if (!im) return;
    for (int c = 0; c < 32; ++c)
        for (int t = 0; t < 9; ++t) im->genw[t * 32 + c] = w_raw[c * 9 + t];
    memcpy(im->genw + 288, bias, 32 * sizeof(float));
}

void tc_set_w4(TcWeights* w, const float* w19) {
    TcImpl* im = static_cast<TcImpl*>(w->impl);
    if (im) memcpy(im->w4, w19, sizeof(im->w4));
}

void tc_release(TcWeights* w) {
    delete static_cast<TcImpl*>(w->impl);
    w->impl = nullptr;
}

int tc_layer(const TcWeights& tw, const DevWeights& w, int precision, int layer, const void* in, void* out, int nrows,
             cudaStream_t st, std::string* err) {
    TcImpl* im = static_cast<TcImpl*>(tw.impl);
    if (!im) { *err = "tensor-core weights not packed"; return -1; }
    switch (layer) {
        case 1: return launch_conv<CfgCt1>(im, im->ct1, w.ct1_b, precision, in, out, nrows, st, err);
        case 2: return launch_conv<CfgCt2>(im, im->ct2, w.ct2_b, precision, in, out, nrows, st, err);
        case 3: return launch_conv<CfgCt3>(im, im->ct3, w.ct3_b, precision, in, out, nrows, st, err, im->w4);
    }
    *err = "unknown tensor-core layer";
    return -1;
}

// One hidden dense layer (bias + ReLU + keyed dropout) on tensor cores: K-blocked in -> K-blocked out.
int tc_dense_hidden(const TcWeights& tw, int which, int precision, const void* in, void* out, int rows, size_t rows_pad,
                    const NoiseKey& nk, const NoiseRows& nr, int layer, cudaStream_t st, std::string* err) {
    TcImpl* im = static_cast<TcImpl*>(tw.impl);
    if (!im || !im->dense_w[which]) { *err = "dense tensor-core weights not packed"; return -1; }
    const int K = im->dense_k[which], N = im->dense_n[which];
    DenseParams p{};
    p.a = static_cast<const __nv_bfloat16*>(in);
    p.a_kc_stride = rows_pad * 8; p.a_plane = (size_t)(K / 8) * rows_pad * 8;
    p.wpack = im->dense_w[which]; p.bias = im->dense_b[which];
    p.nrows = rows; p.nprod = precision == DAI_PREC_BF16X1 ? 1 : 3; p.kchunks = K / 64; p.ntn = N / 128;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.out_kc_stride = rows_pad * 8; p.out_plane = (size_t)(N / 8) * rows_pad * 8;
    p.nk = nk; p.nr = nr; p.layer = layer;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = ((rows + 127) / 128) * p.ntn;
    launch_dep(k_tc_dense<128, EPI_HIDDEN>, dim3(ntiles < sms ? ntiles : sms), dim3(384), DenseCfg<128, EPI_HIDDEN>::SMEM, st, true, p);
    return 1;
}

// im2col of the encoder's conv4 (k3 s2 valid, 7x7x64 -> 3x3): c3 fp32 NHWC (rows,7,7,64) -> GEMM operand, K-blocked bf16
// hi/lo [plane][kc 72][m_pad][8] with row m = image*9 + oy*3+ox and k = (kh*3+kw)*64 + ci.  One thread per (kc, m).
__global__ void __launch_bounds__(256) k_conv4_im2col(const float* __restrict__ c3, int M, size_t m_pad, __nv_bfloat16* __restrict__ out) {
    const size_t n = (size_t)M * 72;
    const size_t plane = 72 * m_pad * 8;
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M), kc = (int)(i / M);
        const int img = m / 9, px = m - img * 9, oy = px / 3, ox = px - oy * 3;
        const int tap = kc >> 3, kh = tap / 3, kw = tap - kh * 3, c0 = (kc & 7) * 8;
        const float* src = c3 + (((size_t)img * 7 + 2 * oy + kh) * 7 + 2 * ox + kw) * 64 + c0;
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(v[2 * e], h0, l0);
            split_bf16(v[2 * e + 1], h1, l1);
            hi[e] = pack_bf16(h0, h1);
            lo[e] = pack_bf16(l0, l1);
        }
        const size_t o = ((size_t)kc * m_pad + m) * 8;
        *reinterpret_cast<uint4*>(out + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(out + plane + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// encoder conv2 + conv3 on tensor cores: c1 = conv1 output in parity-split blocked planes
// [plane][row][parity 4][kc 4][16][16][8]; c2 = same form of the 15x15x32 map ([..][8][8][8]); c3 = fp32 NHWC (7,7,64)
int tc_qs_convs(const TcWeights& tw, const DevWeights& w, int precision, const void* c1, void* c2, float* c3, int rows,
                cudaStream_t st, std::string* err, const float* img) {
    TcImpl* im = static_cast<TcImpl*>(tw.impl);
    if (!im) { *err = "tensor-core weights not packed"; return -1; }
    if (img) {
        // conv1 computed inside conv2's kernel (c1 only backs the unused tensor map)
        if (launch_conv<CfgQc2G>(im, im->qc2, w.qc2_b, precision, c1, c2, rows, st, err, nullptr, img) < 0) return -1;
    } else if (launch_conv<CfgQc2>(im, im->qc2, w.qc2_b, precision, c1, c2, rows, st, err) < 0) return -1;
    if (launch_conv<CfgQc3>(im, im->qc3, w.qc3_b, precision, c2, c3, rows, st, err) < 0) return -1;
    return 2;
}

// encoder conv4 on tensor cores: c3 fp32 NHWC (rows,7,7,64) -> im2col (scratch: (rows*9 padded) x 576 x 2 planes bf16) ->
// GEMM 576 -> 64 with bias + ReLU -> the K-blocked operand of the encoder's FC1 (k = pixel*64 + c), rows_pad rows
size_t tc_qs_conv4_scratch_bytes(int rows) {
    const size_t m_pad = ((size_t)rows * 9 + 127) / 128 * 128 + 128;
    return m_pad * 576 * 2 * sizeof(__nv_bfloat16);
}

int tc_qs_conv4(const TcWeights& tw, int precision, const float* c3, int rows, void* scratch, size_t rows_pad, void* out,
                cudaStream_t st, std::string* err) {
    TcImpl* im = static_cast<TcImpl*>(tw.impl);
    if (!im || !im->dense_w[TC_QC4]) { *err = "conv4 tensor-core weights not packed"; return -1; }
    const int M = rows * 9;
    const size_t m_pad = ((size_t)M + 127) / 128 * 128 + 128;
    launch_dep(k_conv4_im2col, dim3(1184), dim3(256), 0, st, true, c3, M, m_pad, static_cast<__nv_bfloat16*>(scratch));
    DenseParams p{};
    p.a = static_cast<const __nv_bfloat16*>(scratch);
    p.a_kc_stride = m_pad * 8; p.a_plane = 72 * m_pad * 8;
    p.wpack = im->dense_w[TC_QC4]; p.bias = im->dense_b[TC_QC4];
    p.nrows = M; p.nprod = precision == DAI_PREC_BF16X1 ? 1 : 3; p.kchunks = 9; p.ntn = 1;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.out_kc_stride = rows_pad * 8; p.out_plane = 72 * rows_pad * 8;
    p.nk.training = 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (M + 127) / 128;
    launch_dep(k_tc_dense<64, EPI_CONV4>, dim3(ntiles < sms ? ntiles : sms), dim3(384), DenseCfg<64, EPI_CONV4>::SMEM, st, true, p);
    return 2;
}

int tc_fc4(const TcWeights& tw, const DevWeights& w, int precision, const void* h3b, size_t rows_pad, int row0,
           const uint32_t* mask, int nrows, void* act0, cudaStream_t st, std::string* err) {
    TcImpl* im = static_cast<TcImpl*>(tw.impl);
    if (!im) { *err = "tensor-core weights not packed"; return -1; }
    DenseParams p{};
    p.a = static_cast<const __nv_bfloat16*>(h3b) + (size_t)row0 * 8;
    p.a_kc_stride = rows_pad * 8; p.a_plane = 32 * rows_pad * 8;
    p.wpack = im->fc4_wpack; p.bias = im->fc4_bias; p.mask = mask; p.out = static_cast<__nv_bfloat16*>(act0);
    p.nrows = nrows; p.nprod = precision == DAI_PREC_BF16X1 ? 1 : 3; p.kchunks = 4; p.ntn = 64;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = ((nrows + 127) / 128) * 64;
    // (an A-resident variant — m-tile's A kept in shared memory, B streamed in K = 32 stages through a 3-slot ring — was
    // measured 11 % SLOWER, 3.46 vs 3.11 ms per rollout: with 96 KB instead of 192 KB of loads in flight the kernel is
    // bound by load latency x bytes in flight, not by the L2 traffic that variant saves)
    static const bool pairs = !(getenv("DAI_TC_FC4_PAIR") && atoi(getenv("DAI_TC_FC4_PAIR")) == 0);     // A/B switch (experiments)
    if (pairs) {
        // CTA pairs with A resident (k_tc_fc4_pair): needs rows_pad to cover whole 256-row blocks (run_decoder pads by 256)
        const int mpairs = ((nrows + 127) / 128 + 1) / 2;
        const long long total = (long long)mpairs * 64;
        const int npairs = (int)std::max<long long>(1, std::min<long long>(total, sms / 2));
        static long long* dbg_counters = nullptr;
        static const bool want_counters = getenv("DAI_TC_COUNTERS") != nullptr;                // experiments only
        if (want_counters) {
            if (!dbg_counters) cudaMalloc(&dbg_counters, 8 * 8 * 512);
            cudaMemsetAsync(dbg_counters, 0, 8 * 8 * 512, st);
            p.counters = dbg_counters;
        }
        // act0 (64 KB per row, written once, read once by ct1 after the whole chunk) goes out with an L2 evict_first hint: FC4
        // 2.69 -> 2.64 ms per step in three interleaved pairs (its 16.8 MB weight image keeps its place in L2); env DAI_TC_FC4_STREAM=0: off
        static const bool stream_out = !(getenv("DAI_TC_FC4_STREAM") && atoi(getenv("DAI_TC_FC4_STREAM")) == 0);
        p.layer = stream_out ? 1 : 0;
        if (want_counters) k_tc_fc4_pair<true><<<2 * npairs, 384, Fc4Pair::SMEM, st>>>(p);
        else k_tc_fc4_pair<false><<<2 * npairs, 384, Fc4Pair::SMEM, st>>>(p);
        if (want_counters) {
            static int printed = 0;
            cudaStreamSynchronize(st);
            long long hc[8 * 512];
            cudaMemcpy(hc, dbg_counters, sizeof(long long) * 8 * 2 * npairs, cudaMemcpyDeviceToHost);
            double a[6] = {0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 2 * npairs; i += 2) for (int j = 0; j < 6; ++j) a[j] += (double)hc[i * 8 + j] / npairs;
            if (printed++ % 23 == 3)
                fprintf(stderr, "[tc counters] fc4 pair kernel rows %d: per leader cycles: mma loop %.0f (wait acc_empty %.0f, own B %.0f, peer B %.0f, A %.0f) | "
                        "tiles %.1f => %.0f cycles per tile\n", nrows, a[0], a[1], a[2], a[3], a[4], a[5], a[5] > 0 ? a[0] / a[5] : 0.0);
        }
        return 1;
    }
    k_tc_dense<256, EPI_FC4><<<ntiles < sms ? ntiles : sms, 384, DenseCfg<256, EPI_FC4>::SMEM, st>>>(p);
    return 1;
}

// ct2 -> ct3 fused on CTA pairs: act1 (blocked planes) -> last deconv's row planes in act3; `scratch` holds 2 images per CTA
size_t tc_ct23_scratch_bytes(int nrows) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int npairs = std::max(1, std::min((nrows + 1) / 2, sms / 2));
    return (size_t)(4 * npairs) * F23::SCR_IMAGE;     // 2 CTAs x 2 images x (32*32*64 elements x hi/lo bf16)
}

int tc_ct23(const TcWeights& tw, const DevWeights& w, int precision, const void* act1, void* scratch, void* act3, int nrows,
            cudaStream_t st, std::string* err) {
    TcImpl* im = static_cast<TcImpl*>(tw.impl);
    if (!im || !im->ct2.wpair || !im->ct3.wpair) { *err = "pair weights not packed"; return -1; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int npairs = std::max(1, std::min((nrows + 1) / 2, sms / 2));
    const int grid = 2 * npairs, srows = 2 * grid;
    CUtensorMap mapIn, mapScr;
    if (make_map(im, act1, nrows, 16, 16, 8, F23::C2::HX, F23::C2::HY, &mapIn, err) != 0) return -1;
    if (make_map(im, scratch, srows, 32, 32, 8, F23::C3::HX, F23::C3::HY, &mapScr, err) != 0) return -1;
    FusedParams p{};
    p.wpack2 = im->ct2.wpair; p.wpack3 = im->ct3.wpair; p.bias2 = w.ct2_b; p.bias3 = w.ct3_b;
    p.scratch = scratch; p.out = static_cast<float*>(act3);
    p.nrows = nrows; p.nprod = precision == DAI_PREC_BF16X1 ? 1 : 3; p.srows = srows;
    for (int i = 0; i < 288; ++i) p.w4[i] = make_float2(im->w4[i], im->w4[i]);
    static long long* dbg_counters = nullptr;
    static const bool want_counters = getenv("DAI_TC_COUNTERS") != nullptr;                // experiments only
    if (want_counters) {
        if (!dbg_counters) cudaMalloc(&dbg_counters, 16 * 8 * 512);
        cudaMemsetAsync(dbg_counters, 0, 16 * 8 * 512, st);
        p.counters = dbg_counters;
    }
    // Optional (env DAI_TC_L2PERSIST=1, experiments): mark the scratch as an L2-persisting access window for this launch, so
    // its dirty lines are not written back to HBM between the ct2 epilogue and ct3's loads.
    static int l2_state = -1;
    static size_t l2_win_max = 0, l2_persist = 0;
    if (l2_state < 0) {
        const char* e = getenv("DAI_TC_L2PERSIST");
        int max_persist = 0, max_win = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        l2_state = 0;
        if (e && atoi(e) != 0 && max_persist > 0 && max_win > 0) {
            l2_persist = std::min<size_t>((size_t)max_persist, (size_t)96 << 20);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, l2_persist) == cudaSuccess) { l2_win_max = (size_t)max_win; l2_state = 1; }
            else cudaGetLastError();
            fprintf(stderr, "[dai_tc] L2 persistence for the ct2->ct3 scratch: %s (max persisting %d MB, max window %d MB)\n",
                    l2_state ? "on" : "unavailable", max_persist >> 20, max_win >> 20);
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(F23::THREADS); cfg.dynamicSmemBytes = F23::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (l2_state == 1) {
        const size_t bytes = (size_t)srows * F23::SCR_IMAGE;
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = scratch;
        attr[0].val.accessPolicyWindow.num_bytes = std::min(bytes, l2_win_max);
        attr[0].val.accessPolicyWindow.hitRatio = bytes <= l2_persist ? 1.0f : (float)((double)l2_persist / (double)bytes);
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr; cfg.numAttrs = 1;
    } else if (pdl_enabled() && !want_counters) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // see pdl_wait() in the kernel
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    if ((want_counters ? cudaLaunchKernelEx(&cfg, k_tc_ct23<true>, mapIn, mapScr, p)
                       : cudaLaunchKernelEx(&cfg, k_tc_ct23<false>, mapIn, mapScr, p)) != cudaSuccess) {
        *err = std::string("launch of the ct2->ct3 pair kernel failed: ") + cudaGetErrorString(cudaGetLastError());
        return -1;
    }
    if (want_counters) {
        static int printed = 0;
        cudaStreamSynchronize(st);
        long long h[8 * 512];
        cudaMemcpy(h, dbg_counters, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost);
        double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < grid; i += 2) {
            a[0] += (double)h[i * 8] / npairs; a[1] += (double)h[i * 8 + 1] / npairs; a[2] += (double)h[i * 8 + 2] / npairs; a[3] += (double)h[i * 8 + 5] / npairs;
            a[4] += (double)h[i * 8 + 3] / npairs; a[5] += (double)h[i * 8 + 4] / npairs; a[6] += (double)h[i * 8 + 6] / npairs; a[7] += (double)h[i * 8 + 7] / npairs;
        }
        if (printed++ % 23 == 3)
            fprintf(stderr, "[tc counters] ct2+ct3 pair kernel rows %d: per leader cycles: mma loop %.0f (wait acc_empty %.0f, wait a_full %.0f) | "
                    "ct3 epilogue: loop %.0f (wait acc_full %.0f) | ct2 epilogue busy %.0f | a_full wait on ct2 tiles %.0f | images %.1f => %.0f cycles per image\n",
                    nrows, a[0], a[1], a[2], a[4], a[5], a[6], a[7], a[3], a[3] > 0 ? a[0] / a[3] : 0.0);
        if (printed % 23 == 4) {
            static long long hh[512 * 36 / 2];
            cudaMemcpy(hh, dbg_counters + 8 * 512, sizeof(long long) * 36 * npairs, cudaMemcpyDeviceToHost);
            double w[36];
            for (int i = 0; i < 36; ++i) { w[i] = 0; for (int q = 0; q < npairs; ++q) w[i] += (double)hh[q * 36 + i] / npairs; }
            fprintf(stderr, "[tc counters] ct2+ct3 issuer waits per stage item (cycles per image; hi plane | lo plane | accumulator):");
            for (int k = 0; k < 12; ++k) fprintf(stderr, " k%d %.0f|%.0f|%.0f", k, w[k] / a[3], w[12 + k] / a[3], w[24 + k] / a[3]);
            fprintf(stderr, "\n");
        }
    }
    return 1;
}

int tc_decoder_chunk(const TcWeights& tw, const DevWeights& w, int precision, const void* h3b, size_t rows_pad, int row0,
                     const uint32_t* mask, int nrows, void* act0, void* act1, void* act2, void* act3, const Ct4Args& c4in,
                     cudaStream_t st, std::string* err, LayerTimer* timer) {
    LayerTimer none;
    LayerTimer& T = timer ? *timer : none;
    int n = 0, rc;
    T.begin(0, nrows, st);
    if ((rc = tc_fc4(tw, w, precision, h3b, rows_pad, row0, mask, nrows, act0, st, err)) < 0) return -1;
    T.end(st);
    n += rc;
    // ct2 -> ct3 run as ONE kernel (k_tc_ct23): the 256 KB/row activation between them stays in an L2-resident scratch (DRAM
    // traffic of the two layers 3.02 -> 1.37 GB per 4800 rows) and the pair is ~1 % faster end to end than the two kernels.
    // env DAI_TC_FUSE23=0 selects the two separate kernels (A/B, and the reference for tests/test_gpu_layers.py).
    static const bool fuse23 = !(getenv("DAI_TC_FUSE23") && atoi(getenv("DAI_TC_FUSE23")) == 0);
    const void* in[3] = {act0, act1, act2};
    void* out[3] = {act1, act2, act3};
    for (int layer = 1; layer <= 3; ++layer) {
        if (fuse23 && layer == 2) {
            // ct2 -> ct3 in one kernel (timed as layer 2; layer 3 then has no launches); act2 serves as its L2-resident scratch
            T.begin(2, nrows, st);
            if ((rc = tc_ct23(tw, w, precision, act1, act2, act3, nrows, st, err)) < 0) return -1;
            T.end(st);
            n += rc;
            break;
        }
        T.begin(layer, nrows, st);
        if ((rc = tc_layer(tw, w, precision, layer, in[layer - 1], out[layer - 1], nrows, st, err)) < 0) return -1;
        T.end(st);
        n += rc;
    }
    Ct4Args c4 = c4in;
    c4.act3 = static_cast<const float*>(act3);     // row planes + border terms [row][PROJ_ROW_FLOATS]
    T.begin(4, nrows, st);
    n += launch_ct4_gather(w, c4, st);
    T.end(st);
    return n;
}

}  // namespace dai
