// Host-side launch interface of the kernels (internal; the public boundary is include/dai_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "dai_common.cuh"

namespace dai {

// Launch `k` so that it may begin while the previous kernel of `st` drains (programmatic dependent launch, see
// pdl_wait() in dai_common.cuh); dep = false or env DAI_PDL=0: an ordinary launch.  Inside a stream capture the
// relaxed dependency becomes a programmatic edge of the step graph.
bool pdl_enabled();
template <class... P, class... A>
inline cudaError_t launch_dep(void (*k)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool dep, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (dep && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, static_cast<P>(args)...);
}

// Repacked fp32 weights (built once in dai_commit_weights).
//   *_t  : hidden FC layers transposed to [Kpad][N] (coalesced over outputs)
//   tail FCs keep torch's [N][K] (warp-per-output dot products)
//   convs: [tap = kh*3+kw][Cin][Cout]
struct DevWeights {
    // Ps: 14->512->512->512->20 (src/torchmodel.py:41-52)
    float *ps_w0t, *ps_b0, *ps_w1t, *ps_b1, *ps_w2t, *ps_b2, *ps_w3, *ps_b3;
    // Po FCs: 10->256->256->256->16384 (:107-118); FC4 columns permuted to NHWC (h,w,c)
    float *po_w0t, *po_b0, *po_w1t, *po_b1, *po_w2t, *po_b2, *po_w3t, *po_b3;
    // Po deconvs (:120-126)
    float *ct1_w, *ct1_b, *ct2_w, *ct2_b, *ct3_w, *ct3_b, *ct4_w, *ct4_b;
    // Qs convs (:85-92) and FCs (:94-103, FC1 rows permuted to the NHWC flatten)
    float *qc1_w, *qc1_b, *qc2_w, *qc2_b, *qc3_w, *qc3_b, *qc4_w, *qc4_b;
    float *qf0_t, *qf0_b, *qf1_t, *qf1_b, *qf2_t, *qf2_b, *qf3, *qf3_b;
    // Qpi: 10->128->128->4 (:19-25)
    float *pi_w0t, *pi_b0, *pi_w1t, *pi_b1, *pi_w2, *pi_b2;
};

// ---- weight repacking on the device --------------------------------------
// Every packed weight image is described by a gather map over its destination elements: map[i] = index of the fp32
// source element of the state_dict tensor `key` (REPACK_LO set: the bf16 lo part of it; REPACK_NONE: zero padding).
// bf16 = 1: the image is bf16 (hi = rn(x), lo = rn(x - hi)); 0: fp32.  alias: no image, *dst = the stored tensor itself.
constexpr uint32_t REPACK_LO = 0x80000000u, REPACK_NONE = 0xffffffffu;
struct RepackJob {
    std::string key;
    std::vector<uint32_t> map;
    int bf16;
    bool alias;
    void** dst;
};
int  launch_repack(const float* src, const uint32_t* map, size_t n, int bf16, void* dst, cudaStream_t st);

// ---- transition net ------------------------------------------------------------------
struct PsArgs {
    const float* pi;      // [B][4]
    const float* s0;      // [B][10]
    int32_t B;
    int32_t nA, nB;       // sample slots of the loop-2a set (site siteA) and loop-2b set (siteB)
    int32_t sample0;      // global sample of slot 0 (both sets)
    int32_t extra_slot;   // >= 0: loop-2a slot that stands for global sample `extra_sample`
    int32_t extra_sample;
    int32_t siteA, siteB;
    float *meanA, *logvarA, *sampA;   // [nA][B][10]
    float *meanB, *logvarB, *sampB;   // [nB][B][10] (any may be null)
    NoiseKey nk;
};
int  launch_ps(const DevWeights& w, const PsArgs& a, cudaStream_t st);

// ---- decoder -------------------------------------------------------------------------
struct PoFcArgs {
    RowMap map;
    const float* z[3];     // per set: latent rows [Sl][B][10] (mode 0)
    int32_t mode[3];       // 0: direct, 1: reparameterize(rp_mean[b], rp_logvar[b]) with site rp_site
    int32_t zbcast[3];     // 1: z is [B][10] shared by all slots
    int32_t zslots[3];     // time-batched maps (map.sps > 0): slots per horizon step in z[set] (row = (t*zslots + j)*B + b)
    const float *rp_mean, *rp_logvar;   // [B][10]
    size_t rp_tstride;     // time-batched maps: floats between the [B][10] slabs of consecutive steps
    int32_t rp_site;
    float* h3;             // [rows][256] fp32 (CUDA-core FC4), or null
    unsigned short* h3b;   // [plane hi|lo][kc 32][rows_pad][8] bf16 (tensor-core FC4), or null
    size_t rows_pad;
    NoiseKey nk;
};
int  launch_po_fc123(const DevWeights& w, const PoFcArgs& a, cudaStream_t st);

// permuted dropout mask bits of the 16384-wide FC4 output, [rows][512] words in the consumer's column order
// (tc_order 0: NHWC p*64+c for the CUDA-core FC4; 1: the tensor-core FC4's ((pg*8+kc)*4+pl)*8+ce)
int  launch_fc4_mask(const RowMap& map, const NoiseKey& nk, int row0, int nrows, uint32_t* mask, int tc_order, cudaStream_t st);

// fp32 SIMT layers over a chunk of decoder rows [row0, row0+nrows)
int  launch_fc4_simt(const DevWeights& w, const float* h3, const uint32_t* mask, int nrows, float* act0, cudaStream_t st);
int  launch_fc4_simt_blocked(const DevWeights& w, const float* h3, const uint32_t* mask, int nrows, void* act0, cudaStream_t st);
int  launch_ct1_simt(const DevWeights& w, const float* act0, int nrows, float* act1, cudaStream_t st);
int  launch_ct2_simt(const DevWeights& w, const float* act1, int nrows, float* act2, cudaStream_t st);
int  launch_ct3_simt(const DevWeights& w, const float* act2, int nrows, float* act3, cudaStream_t st);

// last deconv + sigmoid + EFE pixel terms.  Per row: hsum = sum_px H_bernoulli(p),
// reward = check_reward(p).  Rows of set 0 (r < img_rows) also write the image.
// Output of the tensor-core ct3 layer per image: 3 row planes e[kh][64][64] of the last deconv (its channel and kw sums
// done), then the tile-border terms [kh 3][oy 64][tile x 4][2] (dai_tc.cu, OUT_PROJ epilogue).
constexpr int PROJ_TILES_X = 4;
constexpr int PROJ_EDGE = 3 * 64 * PROJ_TILES_X * 2;
constexpr int PROJ_ROW_FLOATS = 3 * 4096 + PROJ_EDGE;

struct Ct4Args {
    const float* act3;     // [nrows][64][64][32]
    int32_t row0, nrows;   // global row offset of this chunk
    int32_t img_rows;      // global rows < img_rows write img[r]
    float* img;            // [img_rows][4096]
    float* hsum;           // [rows]
    float* reward;         // [rows]
};
int  launch_ct4_efe(const DevWeights& w, const Ct4Args& a, cudaStream_t st);
// same, from the row planes + border terms [nrows][PROJ_ROW_FLOATS] the tensor-core ct3 epilogue writes into act3
int  launch_ct4_gather(const DevWeights& w, const Ct4Args& a, cudaStream_t st);
// test hook: the finished row planes e[kh] (border terms added), [nrows][3][4096]
int  launch_proj_rows(const float* act3, int nrows, float* out, cudaStream_t st);

// ---- encoder -------------------------------------------------------------------------
struct QsArgs {
    const float* img;      // [rows][4096]
    int32_t rows;
    RowMap map;            // nsets = 1; noise site base map.site[0]
    float *c1, *c2, *c3, *c4;   // workspaces: [rows][31*31*32], [rows][15*15*32], [rows][7*7*64], [rows][576]
    float *mean, *logvar, *samp;   // [rows][10]; samp may be null
    NoiseKey nk;
};
int  launch_qs(const DevWeights& w, const QsArgs& a, cudaStream_t st);
// the same encoder in pieces, so conv2/conv3 can run on tensor cores (dai_tc.cu) between conv1 and the tail
int  launch_qs_conv1(const DevWeights& w, const float* img, int rows, float* c1, void* c1_parity, cudaStream_t st);
int  launch_qs_conv23_simt(const DevWeights& w, const float* c1, int rows, float* c2, float* c3, cudaStream_t st);
int  launch_qs_tail(const DevWeights& w, const QsArgs& a, cudaStream_t st);

// ---- pieces of the tensor-core MLP path (dai_tc.cu runs the hidden layers between them) ----------------
// K-blocked activations: bf16 hi/lo planes [plane][N/8][rows_pad][8]
NoiseRows ps_noise_rows(const PsArgs& a);
NoiseRows map_noise_rows(const RowMap& m);
int  launch_ps_l0(const DevWeights& w, const PsArgs& a, size_t rows_pad, void* out, cudaStream_t st);          // 14 -> 512
int  launch_ps_tail(const DevWeights& w, const PsArgs& a, size_t rows_pad, const void* in, cudaStream_t st);   // 512 -> 20 + reparam
int  launch_po_l0(const DevWeights& w, const PoFcArgs& a, size_t rows_pad, void* out, cudaStream_t st);        // 10 -> 256
int  launch_qs_tail20(const DevWeights& w, const QsArgs& a, size_t rows_pad, const void* in, cudaStream_t st); // 256 -> 20 (+ reparam)
int  launch_qs_conv4_kblocked(const DevWeights& w, const float* c3, int rows, size_t rows_pad, void* out, cudaStream_t st);

// ---- habit net -----------------------------------------------------------------------
int  launch_qpi(const DevWeights& w, const float* s, int B, float* logits, float* q, float* logq, cudaStream_t st);

// ---- scalar glue -----------------------------------------------------------------------
struct StepFinalizeArgs {
    int32_t B, Sl;
    int32_t T;             // horizon steps held by the buffers (0 or 1: one); steps are accumulated in order
    size_t lvA_tstride;    // floats between the steps' logvarA blocks
    const float* logvarA;  // [T][>=Sl][B][10] transition logvar of the loop-2a slots
    const float* qs_logvar;// [T*Sl][B][10]
    const float* reward;   // [3][T*Sl*B] (set 0 used)
    const float* hsum;     // [3][T*Sl*B] (sets 1,2 used)
    double* acc;           // [4][B] += sums of term0, term1, term2_1, term2_2
    const float* carry_src;// [B][10] or null
    float* carry_dst;      // [B][10]
};
int  launch_step_finalize(const StepFinalizeArgs& a, cudaStream_t st);
int  launch_combine(const double* sums, int B, int samples, float* G, float* t0, float* t1, float* t2, cudaStream_t st);
int  launch_select_actions(const float* G, int R, float temperature, const NoiseKey& nk, float* Ppi, float* logPpi,
                           int* choice, cudaStream_t st);
int  launch_reward_only(const float* o, int B, float* r, cudaStream_t st);
int  launch_traj_G(const float* reward, const float* hsum, const float* lv_traj, const float* qs_logvar,
                   int D, int depth, float* G, float* Gmean, cudaStream_t st);

// habit-policy rollout of mcts_step_simulate: `depth` sequential (Qpi -> categorical -> Ps) steps
struct SimArgs {
    const float* start;    // [K][10]
    int32_t K, depth, use_means;
    float *s0, *ps1, *mean, *logvar, *pi0;   // [K*depth][10|4], row = k*depth + t
    float* qpi;            // [K][4]
    NoiseKey nk;
};
int  launch_sim_rollout(const DevWeights& w, const SimArgs& a, cudaStream_t st);

// ---- device-resident search tree (dai_planner.cu; src/mcts.py:11-128) -----------------------
constexpr int PLAN_MAX_K = 32;        // leaves per batch
constexpr int PLAN_MAX_DEPTH = 64;    // edges on one root-to-leaf path
enum { PLAN_SIZE = 0, PLAN_DONE = 1, PLAN_STOP = 2, PLAN_ERR = 3, PLAN_LOGGED = 4, PLAN_NCTL = 8 };
struct PlanTree {
    float *W, *N, *Qpi;       // [cap][4]
    int32_t* child;           // [cap][4], -1 = none
    float* state;             // [cap][10]
    uint8_t* blocked;         // [cap] scratch of one selection round
    int32_t* ctl;             // [PLAN_NCTL]: nodes, expansions done, stop (threshold reached), error, paths logged
    int32_t* host_stop;       // mapped pinned flag the host polls between iterations
    int32_t cap, use_prior;
    float C;
    // log of every expansion's path (all_paths / all_paths_G of the reference's return tuple)
    int32_t* log_actions;     // [log_cap][PLAN_MAX_DEPTH]
    int32_t* log_len;         // [log_cap]
    float* log_G;             // [log_cap]
    int32_t log_cap;
};
struct PlanPicks {
    int32_t *leaf, *len, *nodes, *actions;   // [K], [K], [K][PLAN_MAX_DEPTH] x2 (edge d of pick j: (nodes, actions)[j][d])
    int32_t* count;                          // [1]
    int32_t* scratch;                        // [K * (PLAN_MAX_DEPTH + 1)]
};
int  launch_plan_init(const PlanTree& t, const float* qs0_mean, const float* qpi, cudaStream_t st);
int  launch_plan_select(const PlanTree& t, int k, float threshold, const PlanPicks& p, float* s_rows, float* starts, cudaStream_t st);
int  launch_plan_expand(const PlanTree& t, const PlanPicks& p, const float* G, const float* nxt, cudaStream_t st);
int  launch_plan_accumulate(const float* g, float* sims, int n, int first, cudaStream_t st);
int  launch_plan_backprop(const PlanTree& t, const PlanPicks& p, const float* sims, int nrep, const float* qpi, cudaStream_t st);
int  launch_plan_finish(const PlanTree& t, int* out, cudaStream_t st);

// ---- frame producer (dai_frames.cu; src/game_environment.py:39-66) --------------------------
struct FrameArgs {
    const float* s;        // [G][s_stride] latent classes as floats (Game.current_s), first 6 used
    int32_t s_stride;
    const float* last_r;   // [G]
    const uint32_t* bits;  // [count][128] bit-packed sprites
    long long count;
    long long base[6];     // index weights
    float* o;              // [G][4096]
    int32_t* n_bad;        // games whose index / reward is out of range (frame zeroed)
};
int  launch_pack_sprites(const uint8_t* px, long long count, uint32_t* bits, long long first, cudaStream_t st);
int  launch_render_frames(const FrameArgs& a, int G, cudaStream_t st);


}  // namespace dai
