/*
 * dai_b200.h — C ABI of the B200-native Monte-Carlo EFE rollout path.
 *
 * The reference (zfountas/deep-active-inference-mc @ d7e76d8) has no FFI: its boundary
 * is the Python attribute surface of ActiveInferenceModel and its three sub-modules
 * (src/torchmodel.py:149-393), used from src/mcts.py:71-85,158-164,188, src/util.py:62,
 * test_demo.py:56-57,136-137,150.  Each entry point below names the reference method it
 * replaces.  The Python mirror of that surface (deep-active-inference-mc_b200/
 * torchmodel.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success or a negative DAI_E_* code; nothing throws or
 *     aborts; dai_last_error() gives the message of the last failure on that handle;
 *   - all tensors are contiguous float32; images are NCHW (B,1,64,64) == (B,4096);
 *     unless a function says "host", pointers are device pointers on the handle's GPU and
 *     the work is enqueued on `stream` (a cudaStream_t passed as void*) and not waited for;
 *   - the caller owns every buffer it passes; the handle owns repacked weights and
 *     workspaces (freed in dai_destroy).  One handle per (process, device); a handle is
 *     not thread-safe; distinct handles are independent; no global mutable state;
 *   - noise: every compute call draws its MC-dropout masks and reparameterisation normals
 *     from Philox4x32-10 keyed by (seed + call_index) and addressed by
 *     (step, sample, site, row, element) — DESIGN.md "Noise".  Each entry point states
 *     how many call indices it consumes.
 */
#ifndef DAI_B200_H
#define DAI_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define DAI_API __attribute__((visibility("default")))
#else
#define DAI_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define DAI_OK              0
#define DAI_E_INVALID      -1   /* bad argument / shape / state            */
#define DAI_E_CUDA         -2   /* CUDA runtime or driver error            */
#define DAI_E_NOMEM        -3   /* device allocation failed                */
#define DAI_E_WEIGHTS      -4   /* weights missing or not committed        */
#define DAI_E_UNSUPPORTED  -5   /* configuration outside the built path    */

/* arithmetic of the contraction layers (dai_set_precision) */
#define DAI_PREC_FP32_SIMT   0  /* fp32 FMA on CUDA cores (debug / exact reference on device) */
#define DAI_PREC_BF16X3      1  /* tcgen05 kind::f16, hi/lo split, 3 products, fp32 accumulate (parity mode) */
#define DAI_PREC_BF16X1      2  /* tcgen05 single bf16 pass (fast mode, statistical validation only)        */

typedef struct dai_handle dai_handle;

typedef struct dai_config {
    int32_t s_dim;            /* 10  (ActiveInferenceModel.__init__, src/torchmodel.py:150) */
    int32_t pi_dim;           /* 4                                                        */
    int32_t resolution;       /* 64                                                       */
    int32_t colour_channels;  /* 1                                                        */
    int32_t precision;        /* DAI_PREC_*                                               */
    int32_t training;         /* 1: MC-dropout on (the reference's only mode), 0: identity */
} dai_config;

/* per-call counters filled by dai_get_stats */
typedef struct dai_stats {
    uint64_t kernel_launches; /* kernels of this library launched since the last reset */
    uint64_t calls;           /* compute entry points served                           */
    uint64_t workspace_bytes; /* device bytes currently held by the handle             */
    uint64_t repack_launches; /* gather kernels of the last dai_commit_weights         */
} dai_stats;

DAI_API int  dai_create(const dai_config* cfg, int device, dai_handle** out);
DAI_API int  dai_destroy(dai_handle* h);
DAI_API const char* dai_last_error(const dai_handle* h);
DAI_API const char* dai_version(void);

/* ---- weights: the 46 state_dict tensors of model_top / model_mid / model_down
 *      (src/torchmodel.py:167-177 checkpoint keys, e.g. "po_net.9.weight"); torch layouts
 *      (Linear (out,in); Conv2d (Cout,Cin,3,3); ConvTranspose2d (Cin,Cout,3,3)).
 *      `data` may be a host or a device pointer; the tensor is copied into the handle's own device
 *      copy (a device source never leaves the device) and marked dirty.  dai_set_weight waits for
 *      the copy; dai_set_weight_async enqueues it on `stream` (the source must stay valid until
 *      the stream reaches it; pass the stream that produced it, e.g. the optimizer's).
 *      dai_commit_weights re-packs — on the device, on `stream`, one gather kernel per packed
 *      image (transposes, NHWC permutations, tensor-core K-major bf16 hi/lo images) — ONLY the
 *      images of tensors set since the last commit, and must be called before the next compute
 *      call (on the same stream, or after synchronising).  Incremental: after an optimizer step
 *      that touched one tensor, set that tensor and commit (src/torchmodel.py:167-208, train.py:128-133). */
DAI_API int  dai_set_weight(dai_handle* h, const char* key, const float* data, const int64_t* shape, int ndim);
DAI_API int  dai_set_weight_async(dai_handle* h, const char* key, const float* data, const int64_t* shape, int ndim,
                                  void* stream);
DAI_API int  dai_commit_weights(dai_handle* h, void* stream);

DAI_API int  dai_set_rng(dai_handle* h, uint64_t seed, uint64_t call_index);
DAI_API int  dai_get_rng(const dai_handle* h, uint64_t* seed, uint64_t* call_index);
DAI_API int  dai_set_training(dai_handle* h, int training);      /* nn.Module.train()/eval() on the three nets */
DAI_API int  dai_set_precision(dai_handle* h, int precision);
DAI_API int  dai_get_stats(dai_handle* h, dai_stats* out, int reset);

/* ---- single-net forwards (1 call index each, dai_habit 0) -------------------------- */

/* ModelDown.encoder / encoder_with_sample (src/torchmodel.py:134-137,143-146).
 * o (B,4096) -> mean (B,10), logvar (B,10); sample (B,10) may be NULL. */
DAI_API int  dai_encode(dai_handle* h, const float* o, int B, float* mean, float* logvar, float* sample, void* stream);

/* ModelDown.decoder (src/torchmodel.py:139-141).  s (B,10) -> po (B,4096). */
DAI_API int  dai_decode(dai_handle* h, const float* s, int B, float* po, void* stream);

/* ModelMid.transition / transition_with_sample (src/torchmodel.py:58-66).
 * pi (B,4), s0 (B,10) -> mean, logvar (B,10); sample may be NULL. */
DAI_API int  dai_transition(dai_handle* h, const float* pi, const float* s0, int B,
                    float* mean, float* logvar, float* sample, void* stream);

/* ModelTop.encode_s (src/torchmodel.py:27-31).  s (B,10) -> logits, q, logq (B,4); any may be NULL. */
DAI_API int  dai_habit(dai_handle* h, const float* s, int B, float* logits, float* q, float* logq, void* stream);

/* ActiveInferenceModel.check_reward (src/torchmodel.py:210-212).  o (B,4096) -> r (B). */
DAI_API int  dai_check_reward(dai_handle* h, const float* o, int B, float* r, void* stream);

/* ---- EFE evaluators ---------------------------------------------------------------- */

/* calculate_G (src/torchmodel.py:270-300) for B (state, action) rows and `samples` MC
 * samples; 1 call index.  This rank evaluates samples [sample_begin, sample_end) (pass
 * 0, samples for everything).  Outputs, each may be NULL:
 *   sums (4,B) float64: sum over the local samples of term0, term1, term2_1, term2_2 (NOT
 *   divided) — the payload of the one all-reduce when samples are sharded over GPUs;
 *   G, t0, t1, t2 (B): the finished values — only meaningful when the rank holds all samples;
 *   ps1, ps1_mean, ps1_logvar (B,10), po1 (B,4096): from the globally LAST loop-2a sample
 *   (every rank recomputes that sample under the same noise key, so these agree on all ranks). */
DAI_API int  dai_calculate_G(dai_handle* h, const float* s0, const float* pi0, int B, int samples,
                     int sample_begin, int sample_end,
                     double* sums, float* G, float* t0, float* t1, float* t2,
                     float* ps1, float* ps1_mean, float* ps1_logvar, float* po1, void* stream);

/* calculate_G_mean (src/torchmodel.py:302-327); 1 call index. */
DAI_API int  dai_calculate_G_mean(dai_handle* h, const float* s0, const float* pi0, int B,
                          float* G, float* t0, float* t1, float* t2,
                          float* ps1_mean, float* po1, void* stream);

/* calculate_G_given_trajectory (src/torchmodel.py:329-352); 1 call index.  D rows. */
DAI_API int  dai_G_given_trajectory(dai_handle* h, const float* s0, const float* ps1, const float* ps1_mean,
                            const float* ps1_logvar, const float* pi0, int D, float* G, void* stream);

/* calculate_G_repeated (src/torchmodel.py:227-245) and calculate_G_4_repeated (:247-268);
 * 1 call index.  o (B,4096); pi (B,4) or NULL for eye(4) tiled over B/4 roots (row =
 * root*4 + action, src/util.py:57-60).  `four` selects the _4_ semantics (calculate_G_mean
 * when calc_mean).  T = steps sequential calculate_G evaluations with the s0 carry kept on
 * the device.  Sample sharding and outputs as in dai_calculate_G; sums accumulate over steps.
 * The steps are chained only through the transition net, so the evaluation order inside the call
 * is: the latent chain of all T steps, then the decoder / encoder / EFE work of groups of steps
 * batched over (step, sample) — same noise keys, same sums, same results as step by step.  A call
 * whose sizes and pointers repeat is replayed as one CUDA graph. */
DAI_API int  dai_rollout(dai_handle* h, const float* o, const float* pi, int B, int steps, int samples,
                 int calc_mean, int four, int sample_begin, int sample_end,
                 double* sums, float* G, float* t0, float* t1, float* t2, float* po1, void* stream);

/* Finish sharded sums after the all-reduce: sums (4,B) -> G, t0, t1, t2 (B). */
DAI_API int  dai_combine(dai_handle* h, const double* sums, int B, int samples,
                 float* G, float* t0, float* t1, float* t2, void* stream);

/* Same as dai_rollout with HOST buffers: copies o (and pi) host->device, runs, copies
 * G,t0,t1,t2 (and po1 if not NULL) device->host and waits.  This is the end-to-end call
 * bench.py times as `e2e`. */
DAI_API int  dai_rollout_host(dai_handle* h, const float* o_host, const float* pi_host, int B, int steps,
                      int samples, int calc_mean, int four,
                      float* G_host, float* t0_host, float* t1_host, float* t2_host, float* po1_host,
                      void* stream);

/* mcts_step_simulate (src/torchmodel.py:354-393); 2 call indices (rollout, trajectory).
 * starting_s (10) device; writes pi0 (depth,4), qpi (4) on the device and the mean G to
 * *G_host (waits for the stream). */
DAI_API int  dai_mcts_simulate(dai_handle* h, const float* starting_s, int depth, int use_means,
                       float* G_host, float* pi0, float* qpi, void* stream);

/* ---- next row: batched-leaf MCTS.  K independent mcts_step_simulate rollouts in ONE pass: the K
 * habit-policy rollouts run as K CTAs of one launch, the K*depth trajectory rows go through ONE
 * calculate_G_given_trajectory evaluation (row = k*depth + t) and the K per-trajectory means come back with one
 * stream wait.  starting_s (K,10) device; pi0 (K*depth,4), qpi (K,4) device; G_host (K) host.  Noise: rollout k draws
 * with row index k, trajectory row r with row index r, so K = 1 is exactly dai_mcts_simulate.  2 call indices. */
DAI_API int  dai_mcts_simulate_batch(dai_handle* h, const float* starting_s, int K, int depth, int use_means,
                                     float* G_host, float* pi0, float* qpi, void* stream);

/* One whole planning decision — active_inference_mcts (src/mcts.py:150-195) with `leaves` leaves per iteration — with
 * the search tree resident on the device: selection (the reference's argmax descent, src/mcts.py:39-62, made `leaves`
 * times with virtual visits so the leaves are distinct), expansion bookkeeping (:64-86), back-propagation (:88-96) and
 * the most-visited path (:98-106) are kernels between the EFE evaluations; the host enqueues the iterations and waits
 * ONCE, at the end (plus a poll of a mapped flag for the visit-count threshold of :176).  leaves = 1 is the
 * reference's search.  frame (4096) device, or qs0_mean (10) device if the caller already encoded it (then no encoder
 * call is made).  Results (host): path = the most-visited path before the opposite-action trimming of :108-126
 * (path_host holds 64 entries), all_paths (repeats x 64) / all_len / all_G = every expansion's action path and
 * simulated G (may be NULL).  Call indices: 1 (encoder, if frame) + 1 (root expansion) + per iteration 1 +
 * 2 * simulation_repeats — the same as the host-driven planner, so both make the same decisions. */
typedef struct dai_mcts_params {
    float   C;                             /* MCTS_Params.C (src/mcts.py:139)            */
    float   threshold;                     /* .threshold                                 */
    int32_t repeats;                       /* .repeats: expansions per decision          */
    int32_t simulation_repeats;            /* .simulation_repeats                        */
    int32_t simulation_depth;              /* .simulation_depth                          */
    int32_t use_means;                     /* .use_means: calculate_G_mean expansions    */
    int32_t using_prior_for_exploration;   /* .using_prior_for_exploration               */
    int32_t samples;                       /* MC samples per expansion (Node.expand)     */
    int32_t leaves;                        /* leaves per batch, 1..32                    */
} dai_mcts_params;
typedef struct dai_mcts_result {
    int32_t path_len;                      /* entries in path_host                       */
    int32_t repeats_done;                  /* expansions made before the search ended    */
    int32_t stopped;                       /* 1: the visit-count threshold ended it      */
    int32_t logged;                        /* entries in all_paths / all_len / all_G     */
} dai_mcts_result;
DAI_API int  dai_mcts_plan(dai_handle* h, const float* frame, const float* qs0_mean, const dai_mcts_params* prm,
                           dai_mcts_result* res, int32_t* path_host, int32_t* all_paths_host, int32_t* all_len_host,
                           float* all_G_host, void* stream);

/* ---- next row: batched many-roots action selection -------------------------
 * The action choice of make_batch_dsprites_active_inference (src/util.py:46-53,66-68) for R roots whose summed EFE
 * G (4R, row = root*4 + action) is already on the device: per root x = -G - max(-G), e = exp(x / temperature),
 * Ppi = e / sum(e), logPpi = x - log(sum(e) + 1e-20) (un-tempered, as softmax_multi_with_log returns it), and
 * choice ~ Categorical(Ppi) drawn from the keyed noise (site CAT, row = root).  1 call index. */
DAI_API int  dai_select_actions(dai_handle* h, const float* G, int R, float temperature,
                                float* Ppi, float* logPpi, int32_t* choice, void* stream);

/* ---- next row: the frame producer in front of the path ------------------------------------
 * Game.current_frame_all / s_to_o (src/game_environment.py:39-66) for G games in one kernel instead of a Python loop.
 * dai_frames_set_sprites uploads the dSprites `imgs` table (HOST uint8 (count,64,64), binary) once and keeps it
 * bit-packed in HBM (512 B per sprite); latents_sizes[6] = metadata['latents_sizes'] ([1,3,6,40,32,32]) must multiply
 * to count.  dai_frames_render: s (G, s_stride >= 6) = Game.current_s (latent classes as floats), last_r (G), both on
 * the device -> o (G,4096) float32: sprite[index] with the reward bar on rows 0..2 (r in [0,1]: columns 0..31 = r;
 * r in [-1,0): columns 32..63 = -r).  index = sum_i trunc(s_i) * base_i with base = the mixed-radix place values of
 * latents_sizes, or — reference_bases != 0 — the reference's s_bases as shipped ([1,3,6,40,32,32]).
 * Where the reference raises (index outside the table, reward outside [-1,1]) the frame is zeroed and counted: if
 * n_bad_host is not NULL the call waits for the stream and stores the count there.  No noise, 0 call indices. */
DAI_API int  dai_frames_set_sprites(dai_handle* h, const uint8_t* imgs_host, int64_t count, const int32_t* latents_sizes,
                                    void* stream);
DAI_API int  dai_frames_render(dai_handle* h, const float* s, int s_stride, const float* last_r, int G, int reference_bases,
                               float* o, int32_t* n_bad_host, void* stream);

/* ---- MC-sample sharding over GPUs (SURVEY.md §8 e; shards of src/torchmodel.py:273,287) ------------------------
 * One process per GPU, one handle per process.  Rank r of W evaluates the contiguous sample slice
 * [r*N/W ...) of every step (the first N % W ranks hold one more sample) with replicated weights; noise is keyed by
 * the GLOBAL sample index, the globally last loop-2a transition is recomputed on every rank, and the whole call
 * needs ONE all-reduce(sum) of the (4,B) float64 term sums — issued by the library itself (NCCL, resolved with
 * dlopen("libnccl.so.2") at run time: no link-time dependency) on the caller's stream.
 *   dai_comm_unique_id  rank 0: 128 bytes (ncclUniqueId) to hand to every rank out of band (MPI, a file, torch.distributed)
 *   dai_comm_init       collective over the W ranks; world = 1 detaches (no NCCL needed)
 *   dai_rollout_sharded / dai_calculate_G_sharded
 *                       dai_rollout / dai_calculate_G over this rank's slice + the all-reduce + the finish; every rank
 *                       receives the full result (G, terms, po1 / ps1 of the last sample).  All ranks must make the
 *                       same calls in the same order (same seed and call index => same noise keys). */
DAI_API int  dai_comm_unique_id(void* id128);
DAI_API int  dai_comm_init(dai_handle* h, const void* id128, int rank, int world);
DAI_API int  dai_comm_destroy(dai_handle* h);
DAI_API int  dai_comm_info(const dai_handle* h, int* rank, int* world);
DAI_API int  dai_rollout_sharded(dai_handle* h, const float* o, const float* pi, int B, int steps, int samples,
                         int calc_mean, int four, float* G, float* t0, float* t1, float* t2, float* po1, void* stream);
DAI_API int  dai_calculate_G_sharded(dai_handle* h, const float* s0, const float* pi0, int B, int samples,
                             float* G, float* t0, float* t1, float* t2,
                             float* ps1, float* ps1_mean, float* ps1_logvar, float* po1, void* stream);

/* ---- per-kernel timing (bench.py's roofline leg) -----------------------------------------
 * Between dai_profile_begin and dai_profile_end every decoder contraction kernel is bracketed by
 * CUDA events on its launch stream.  dai_profile_end waits for the stream and returns, per layer
 * (0: FC4, 1: ct1, 2: ct2, 3: ct3, 4: pixel terms), the summed device time in ms, the number of
 * launches and the decoder rows processed. */
DAI_API int  dai_profile_begin(dai_handle* h);
DAI_API int  dai_profile_end(dai_handle* h, float ms[5], int64_t launches[5], int64_t rows[5], void* stream);

/* ---- test hook ------------------------------------------------------------------------
 * One decoder contraction layer in isolation (layer 1: ConvT 64->64 s1 16x16; 2: ConvT 64->64 s2
 * 16x16->32x32; 3: ConvT 64->32 s2 32x32->64x64; src/torchmodel.py:120-124), bias + ReLU included,
 * fp32 NHWC in and out, in the given DAI_PREC_* arithmetic.  Used by tests to compare the tcgen05
 * kernels with the fp32 CUDA-core kernels layer by layer.  Exception: layer 3 in a tensor-core precision returns what
 * that kernel hands to the pixel kernel — the last deconv's channel and kw sums, out (nrows,3,4096):
 * e[kh][oy][ox] = sum_kw sum_c relu(ct3)[oy][ox+1-kw][c] * w4[c][kh][kw].
 * layer 23 (tensor-core precisions only) is the production pair kernel that runs layers 2 and 3 fused (the
 * 32x32x64 activation between them stays in an L2-resident scratch): in (nrows,256,64), out (nrows,3,4096) as for 3. */
DAI_API int  dai_debug_layer(dai_handle* h, int layer, int precision, const float* in, int nrows, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAI_B200_H */
