// C ABI (include/dai_b200.h): handle, weight repacking, workspaces and the orchestration of
// the kernels into the reference's EFE evaluators (src/torchmodel.py:210-393).
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/dai_b200.h"
#include "dai_kernels.h"
#include "dai_tc.h"

using namespace dai;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct WeightSpec {
    const char* key;
    int ndim;
    int64_t shape[4];
};

// state_dict of model_top / model_mid / model_down (src/torchmodel.py:19-25,41-52,84-128; D1 repaired)
const WeightSpec kSpecs[] = {
    {"qpi_net.0.weight", 2, {128, 10}},      {"qpi_net.0.bias", 1, {128}},
    {"qpi_net.2.weight", 2, {128, 128}},     {"qpi_net.2.bias", 1, {128}},
    {"qpi_net.4.weight", 2, {4, 128}},       {"qpi_net.4.bias", 1, {4}},
    {"ps_net.0.weight", 2, {512, 14}},       {"ps_net.0.bias", 1, {512}},
    {"ps_net.3.weight", 2, {512, 512}},      {"ps_net.3.bias", 1, {512}},
    {"ps_net.6.weight", 2, {512, 512}},      {"ps_net.6.bias", 1, {512}},
    {"ps_net.9.weight", 2, {20, 512}},       {"ps_net.9.bias", 1, {20}},
    {"qs_net.0.weight", 4, {32, 1, 3, 3}},   {"qs_net.0.bias", 1, {32}},
    {"qs_net.2.weight", 4, {32, 32, 3, 3}},  {"qs_net.2.bias", 1, {32}},
    {"qs_net.4.weight", 4, {64, 32, 3, 3}},  {"qs_net.4.bias", 1, {64}},
    {"qs_net.6.weight", 4, {64, 64, 3, 3}},  {"qs_net.6.bias", 1, {64}},
    {"qs_net.9.weight", 2, {256, 576}},      {"qs_net.9.bias", 1, {256}},
    {"qs_net.12.weight", 2, {256, 256}},     {"qs_net.12.bias", 1, {256}},
    {"qs_net.15.weight", 2, {256, 256}},     {"qs_net.15.bias", 1, {256}},
    {"qs_net.18.weight", 2, {20, 256}},      {"qs_net.18.bias", 1, {20}},
    {"po_net.0.weight", 2, {256, 10}},       {"po_net.0.bias", 1, {256}},
    {"po_net.3.weight", 2, {256, 256}},      {"po_net.3.bias", 1, {256}},
    {"po_net.6.weight", 2, {256, 256}},      {"po_net.6.bias", 1, {256}},
    {"po_net.9.weight", 2, {16384, 256}},    {"po_net.9.bias", 1, {16384}},
    {"po_net.13.weight", 4, {64, 64, 3, 3}}, {"po_net.13.bias", 1, {64}},
    {"po_net.15.weight", 4, {64, 64, 3, 3}}, {"po_net.15.bias", 1, {64}},
    {"po_net.17.weight", 4, {64, 32, 3, 3}}, {"po_net.17.bias", 1, {32}},
    {"po_net.19.weight", 4, {32, 1, 3, 3}},  {"po_net.19.bias", 1, {1}},
};
constexpr int kNumSpecs = sizeof(kSpecs) / sizeof(kSpecs[0]);

constexpr int kDecChunkDefault = 19200;  // decoder rows per activation chunk (182 KB of activations per row with the fused ct2->ct3 kernel).  Larger chunks amortise the per-launch ramp: 1024 -> 4800 rows = +7 %, 4800 -> 9600 +1 %, and at R = 32 (19200 rows) one chunk instead of four = +2.6 %
constexpr int kQsChunk = 4096;    // encoder rows per chunk (166 KB of conv features per row); chunks are equalised

}  // namespace

struct dai_handle {
    dai_config cfg{};
    int device = 0;
    std::string err;
    // weights: every state_dict tensor is stored on the device as given (raw_dev, allocated once);
    // every packed image is a RepackJob that a gather kernel rebuilds when its source tensor is dirty
    float* raw_dev[64] = {};           // by kSpecs index
    bool have[64] = {};
    bool dirty[64] = {};
    struct DevJob { int spec; uint32_t* map; size_t n; int bf16; void* dst; };
    std::vector<DevJob> jobs;
    bool planned = false;
    bool committed = false;
    DevWeights w{};
    TcWeights tcw{};
    std::vector<void*> wallocs;
    uint64_t repack_launches = 0;      // of the last commit
    uint64_t seed = 1234, call = 0;
    uint64_t launches = 0, calls = 0;
    int dec_chunk = kDecChunkDefault;   // env DAI_DEC_CHUNK (tuning only)
    // workspaces (grow-only)
    DevBuf mlpA, mlpB, ps, zB, h3, mask, act0, act1, act2, act3, img, hsum, reward, qc1, qc2, qc3, qc4, qs_out, acc, carry,
        pi_eye, traj, root, stage_in, stage_out, scratch;
    LayerTimer timer;
    // frame producer
    DevBuf sprites, sprite_stage, frame_flag;
    long long sprite_count = 0;
    long long place[6] = {0, 0, 0, 0, 0, 0}, sizes[6] = {0, 0, 0, 0, 0, 0};
    // device-resident planner
    DevBuf plan_tree, plan_picks, plan_rows, plan_out, plan_pi0;
    int32_t* plan_stop_host = nullptr;   // mapped pinned flag: the search's threshold test fired
    int32_t* plan_stop_dev = nullptr;
    cudaEvent_t plan_sel_ev = nullptr;   // recorded after every selection kernel: bounds the host's run-ahead to one batch
    // one EFE step as a replayed CUDA graph: cache of instantiated step graphs keyed by everything a
    // step's launches depend on except the noise key, which the kernels then read from `keybuf`
    struct StepGraph { uint64_t sig = 0; uint64_t gen = 0; cudaGraphExec_t exec = nullptr; uint64_t nlaunch = 0; int seen = 0; uint64_t stamp = 0; };
    std::vector<StepGraph> graphs;
    uint64_t alloc_gen = 0;            // bumped whenever a workspace is (re)allocated or a weight that travels as a kernel parameter
                                       // changes: invalidates every cached graph
    bool capturing = false, capture_failed = false;
    int graphs_enabled = 1;            // env DAI_GRAPHS=0 disables (A/B)
    uint32_t* keybuf = nullptr;        // device {k0, k1, step}
    // the legacy default stream cannot be captured: work a caller enqueues on it runs on this stream instead, fenced
    // with events on both sides (StreamSwap)
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    // side stream of a step: work that depends on nothing but the noise key (FC4's dropout bit planes) is forked onto it and
    // joined before its consumer — inside a captured step the fork/join become graph edges (env DAI_FORK=0 disables)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint64_t graph_stamp = 0, graph_replays = 0;
    // sample-shard communicator (NCCL, resolved at run time; SURVEY.md §8 e)
    void* comm = nullptr;      // ncclComm_t
    int comm_rank = 0, comm_world = 1;
    float* pinned = nullptr;   // small host result buffer
    size_t pinned_cap = 0;
};

namespace {

int fail(dai_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(h, DAI_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define RET(call)                 \
    do {                          \
        int rc__ = (call);        \
        if (rc__ != DAI_OK) return rc__; \
    } while (0)

// Graph capture needs a capturable stream.  If the caller hands in the legacy default stream (what torch's current stream is
// unless the user switched), the call's work is enqueued on the handle's own non-blocking stream, ordered after everything
// already on the caller's stream and followed by a wait on the caller's stream, so the caller-visible ordering is unchanged.
struct StreamSwap {
    dai_handle* h;
    cudaStream_t user, use;
    StreamSwap(dai_handle* h_, cudaStream_t st) : h(h_), user(st), use(st) {
        const bool legacy = st == nullptr || st == cudaStreamLegacy;
        if (!legacy || !h->graphs_enabled || !h->own_stream) return;
        if (cudaEventRecord(h->ev_in, user) != cudaSuccess || cudaStreamWaitEvent(h->own_stream, h->ev_in, 0) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        use = h->own_stream;
    }
    ~StreamSwap() {
        if (use == user) return;
        if (cudaEventRecord(h->ev_out, use) == cudaSuccess) cudaStreamWaitEvent(user, h->ev_out, 0);
        cudaGetLastError();
    }
    StreamSwap(const StreamSwap&) = delete;
    StreamSwap& operator=(const StreamSwap&) = delete;
};

int reserve(dai_handle* h, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return DAI_OK;
    if (h->capturing) {                 // a stream capture cannot allocate or synchronise: give up on this capture
        h->capture_failed = true;
        return fail(h, DAI_E_CUDA, "workspace growth during graph capture");
    }
    ++h->alloc_gen;
    if (b.p) {
        CK(cudaDeviceSynchronize());
        CK(cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
    }
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(h, DAI_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return DAI_OK;
}

template <class T>
T* ptr(DevBuf& b) { return static_cast<T*>(b.p); }

NoiseKey make_key(const dai_handle* h, uint64_t call, uint32_t step) {
    const uint64_t k = h->seed + call;
    NoiseKey nk{};
    nk.k0 = (uint32_t)(k & 0xffffffffu);
    nk.k1 = (uint32_t)(k >> 32);
    nk.step = step;
    nk.training = h->cfg.training;
    return nk;
}

int check_ready(dai_handle* h) {
    if (!h) return DAI_E_INVALID;
    if (!h->committed) return fail(h, DAI_E_WEIGHTS, "weights not committed (dai_set_weight x46, then dai_commit_weights)");
    CK(cudaSetDevice(h->device));
    return DAI_OK;
}

int post_launch(dai_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, DAI_E_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
    return DAI_OK;
}

// ---- weight repacking ----------------------------------------------------------------

int spec_index(const char* key) {
    for (int i = 0; i < kNumSpecs; ++i)
        if (strcmp(kSpecs[i].key, key) == 0) return i;
    return -1;
}

size_t spec_elems(int i) {
    size_t n = 1;
    for (int d = 0; d < kSpecs[i].ndim; ++d) n *= (size_t)kSpecs[i].shape[d];
    return n;
}

// gather maps of the CUDA-core (fp32) images; the tensor-core images are planned in dai_tc.cu (tc_plan_weights)
// torch Linear (N,K) -> [Kpad][N]
std::vector<uint32_t> map_transpose_pad(int N, int K, int Kpad) {
    std::vector<uint32_t> t((size_t)Kpad * N, REPACK_NONE);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) t[(size_t)k * N + n] = (uint32_t)((size_t)n * K + k);
    return t;
}

// ConvTranspose2d (Cin,Cout,3,3) -> [tap][Cin][Cout]
std::vector<uint32_t> map_convT(int Cin, int Cout) {
    std::vector<uint32_t> p((size_t)9 * Cin * Cout);
    for (int ci = 0; ci < Cin; ++ci)
        for (int co = 0; co < Cout; ++co)
            for (int t = 0; t < 9; ++t) p[((size_t)t * Cin + ci) * Cout + co] = (uint32_t)(((size_t)ci * Cout + co) * 9 + t);
    return p;
}

// Conv2d (Cout,Cin,3,3) -> [tap][Cin][Cout]
std::vector<uint32_t> map_conv(int Cout, int Cin) {
    std::vector<uint32_t> p((size_t)9 * Cin * Cout);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int t = 0; t < 9; ++t) p[((size_t)t * Cin + ci) * Cout + co] = (uint32_t)(((size_t)co * Cin + ci) * 9 + t);
    return p;
}

// Describe every packed image once per handle: device buffers + gather maps (uploaded), nothing packed yet.
int plan_weights(dai_handle* h) {
    std::vector<RepackJob> jobs;
    DevWeights& w = h->w;
    auto img = [&](const char* key, std::vector<uint32_t> map, float** dst) {
        jobs.push_back(RepackJob{key, std::move(map), 0, false, reinterpret_cast<void**>(dst)});
    };
    auto same = [&](const char* key, float** dst) {          // used as stored
        jobs.push_back(RepackJob{key, {}, 0, true, reinterpret_cast<void**>(dst)});
    };
    // Ps
    img("ps_net.0.weight", map_transpose_pad(512, 14, 16), &w.ps_w0t);   same("ps_net.0.bias", &w.ps_b0);
    img("ps_net.3.weight", map_transpose_pad(512, 512, 512), &w.ps_w1t); same("ps_net.3.bias", &w.ps_b1);
    img("ps_net.6.weight", map_transpose_pad(512, 512, 512), &w.ps_w2t); same("ps_net.6.bias", &w.ps_b2);
    same("ps_net.9.weight", &w.ps_w3);                                   same("ps_net.9.bias", &w.ps_b3);
    // Po FCs
    img("po_net.0.weight", map_transpose_pad(256, 10, 12), &w.po_w0t);   same("po_net.0.bias", &w.po_b0);
    img("po_net.3.weight", map_transpose_pad(256, 256, 256), &w.po_w1t); same("po_net.3.bias", &w.po_b1);
    img("po_net.6.weight", map_transpose_pad(256, 256, 256), &w.po_w2t); same("po_net.6.bias", &w.po_b2);
    {   // FC4: reference output index e = c*256 + p (Unflatten (64,16,16)) -> NHWC column n' = p*64 + c
        std::vector<uint32_t> t((size_t)256 * 16384), bp(16384);
        for (int c = 0; c < 64; ++c)
            for (int p = 0; p < 256; ++p) {
                const int e = c * 256 + p, n = p * 64 + c;
                bp[n] = (uint32_t)e;
                for (int k = 0; k < 256; ++k) t[(size_t)k * 16384 + n] = (uint32_t)((size_t)e * 256 + k);
            }
        img("po_net.9.weight", std::move(t), &w.po_w3t);
        img("po_net.9.bias", std::move(bp), &w.po_b3);
    }
    img("po_net.13.weight", map_convT(64, 64), &w.ct1_w); same("po_net.13.bias", &w.ct1_b);
    img("po_net.15.weight", map_convT(64, 64), &w.ct2_w); same("po_net.15.bias", &w.ct2_b);
    img("po_net.17.weight", map_convT(64, 32), &w.ct3_w); same("po_net.17.bias", &w.ct3_b);
    img("po_net.19.weight", map_convT(32, 1), &w.ct4_w);  same("po_net.19.bias", &w.ct4_b);
    // Qs
    img("qs_net.0.weight", map_conv(32, 1), &w.qc1_w);  same("qs_net.0.bias", &w.qc1_b);
    img("qs_net.2.weight", map_conv(32, 32), &w.qc2_w); same("qs_net.2.bias", &w.qc2_b);
    img("qs_net.4.weight", map_conv(64, 32), &w.qc3_w); same("qs_net.4.bias", &w.qc3_b);
    img("qs_net.6.weight", map_conv(64, 64), &w.qc4_w); same("qs_net.6.bias", &w.qc4_b);
    {   // FC1: reference flatten index f = c*9 + h*3 + w -> NHWC flatten f' = (h*3+w)*64 + c
        std::vector<uint32_t> t((size_t)576 * 256);
        for (int n = 0; n < 256; ++n)
            for (int c = 0; c < 64; ++c)
                for (int p = 0; p < 9; ++p) t[(size_t)(p * 64 + c) * 256 + n] = (uint32_t)((size_t)n * 576 + c * 9 + p);
        img("qs_net.9.weight", std::move(t), &w.qf0_t);
    }
    same("qs_net.9.bias", &w.qf0_b);
    img("qs_net.12.weight", map_transpose_pad(256, 256, 256), &w.qf1_t); same("qs_net.12.bias", &w.qf1_b);
    img("qs_net.15.weight", map_transpose_pad(256, 256, 256), &w.qf2_t); same("qs_net.15.bias", &w.qf2_b);
    same("qs_net.18.weight", &w.qf3);                                    same("qs_net.18.bias", &w.qf3_b);
    // Qpi
    img("qpi_net.0.weight", map_transpose_pad(128, 10, 12), &w.pi_w0t);   same("qpi_net.0.bias", &w.pi_b0);
    img("qpi_net.2.weight", map_transpose_pad(128, 128, 128), &w.pi_w1t); same("qpi_net.2.bias", &w.pi_b1);
    same("qpi_net.4.weight", &w.pi_w2);                                   same("qpi_net.4.bias", &w.pi_b2);
    // tensor-core operand images (bf16 hi/lo, K-major)
    std::string terr;
    if (tc_plan_weights(&h->tcw, &jobs, &terr) != 0) return fail(h, DAI_E_CUDA, "tensor-core weight planning failed: %s", terr.c_str());
    // device side: buffers + maps
    for (RepackJob& j : jobs) {
        const int si = spec_index(j.key.c_str());
        if (si < 0) return fail(h, DAI_E_INVALID, "internal: repack job for unknown key %s", j.key.c_str());
        if (j.alias) { *j.dst = h->raw_dev[si]; continue; }
        const size_t n = j.map.size();
        void *dmap = nullptr, *ddst = nullptr;
        if (cudaMalloc(&dmap, n * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&ddst, std::max<size_t>(n * (j.bf16 ? 2 : 4), 16)) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, DAI_E_NOMEM, "cudaMalloc(packed weights for %s) failed", j.key.c_str());
        }
        h->wallocs.push_back(dmap); h->wallocs.push_back(ddst);
        CK(cudaMemcpy(dmap, j.map.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        *j.dst = ddst;
        h->jobs.push_back(dai_handle::DevJob{si, static_cast<uint32_t*>(dmap), n, j.bf16, ddst});
    }
    h->planned = true;
    return DAI_OK;
}

// Re-pack the images of every dirty tensor on `st` (one gather kernel per image; nothing leaves the device except the
// 288 floats of po_net.19.weight and the 320 of qs_net.0, which travel as kernel parameters of ct3 / conv2).
int commit(dai_handle* h, cudaStream_t st) {
    for (int i = 0; i < kNumSpecs; ++i)
        if (!h->have[i]) return fail(h, DAI_E_WEIGHTS, "missing weight %s", kSpecs[i].key);
    if (!h->planned) {
        RET(plan_weights(h));
        for (int i = 0; i < kNumSpecs; ++i) h->dirty[i] = true;
    }
    h->repack_launches = 0;
    for (const dai_handle::DevJob& j : h->jobs)
        if (h->dirty[j.spec]) h->repack_launches += launch_repack(h->raw_dev[j.spec], j.map, j.n, j.bf16, j.dst, st);
    h->launches += h->repack_launches;
    RET(post_launch(h, "weight repack"));
    const int i19 = spec_index("po_net.19.weight");
    if (h->dirty[i19]) {
        float w19[288];
        CK(cudaMemcpyAsync(w19, h->raw_dev[i19], sizeof(w19), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        tc_set_w4(&h->tcw, w19);
        ++h->alloc_gen;          // these weights are kernel PARAMETERS: a graph captured with the old values must not be replayed
    }
    const int iq0 = spec_index("qs_net.0.weight"), iq0b = spec_index("qs_net.0.bias");
    if (h->dirty[iq0] || h->dirty[iq0b]) {
        // the encoder's first conv is computed inside conv2's kernel, its 320 parameters are constant-bank operands there
        float wq[288], bq[32];
        CK(cudaMemcpyAsync(wq, h->raw_dev[iq0], sizeof(wq), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(bq, h->raw_dev[iq0b], sizeof(bq), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        tc_set_conv1(&h->tcw, wq, bq);
        ++h->alloc_gen;          // kernel parameters as well (see above)
    }
    for (int i = 0; i < kNumSpecs; ++i) h->dirty[i] = false;
    h->committed = true;
    return DAI_OK;
}

// ---- MLPs: CUDA-core fused kernels (fp32_simt) or first layer / tcgen05 hidden layers / tail ----------

int run_ps(dai_handle* h, cudaStream_t st, const PsArgs& pa) {
    const int rows = (pa.nA + pa.nB) * pa.B;
    if (rows <= 0) return DAI_OK;
    if (h->cfg.precision == DAI_PREC_FP32_SIMT) {
        h->launches += launch_ps(h->w, pa, st);
        return post_launch(h, "transition");
    }
    const size_t rows_pad = ((size_t)rows + 127) / 128 * 128 + 128;
    RET(reserve(h, h->mlpA, rows_pad * 512 * 2 * sizeof(unsigned short)));
    RET(reserve(h, h->mlpB, rows_pad * 512 * 2 * sizeof(unsigned short)));
    const NoiseRows nr = ps_noise_rows(pa);
    std::string terr;
    h->launches += launch_ps_l0(h->w, pa, rows_pad, h->mlpA.p, st);
    int n1 = tc_dense_hidden(h->tcw, TC_PS1, h->cfg.precision, h->mlpA.p, h->mlpB.p, rows, rows_pad, pa.nk, nr, 1, st, &terr);
    int n2 = n1 < 0 ? -1 : tc_dense_hidden(h->tcw, TC_PS2, h->cfg.precision, h->mlpB.p, h->mlpA.p, rows, rows_pad, pa.nk, nr, 2, st, &terr);
    if (n1 < 0 || n2 < 0) return fail(h, DAI_E_CUDA, "tensor-core transition: %s", terr.c_str());
    h->launches += n1 + n2 + launch_ps_tail(h->w, pa, rows_pad, h->mlpA.p, st);
    return post_launch(h, "transition");
}

// ---- decoder over row sets -----------------------------------------------------------

// Runs Po on `fc` rows (sets x slots x B): FC1..3 fused, then per chunk FC4 -> ct1 -> ct2 -> ct3 -> pixel
// terms.  Rows < img_rows write their image into `img`.
int run_decoder(dai_handle* h, cudaStream_t st, PoFcArgs fc, int img_rows, float* img, float* hsum, float* reward) {
    const int rows = fc.map.rows();
    if (rows <= 0) return DAI_OK;
    const bool tc = h->cfg.precision != DAI_PREC_FP32_SIMT;
    const size_t rows_pad = ((size_t)rows + 127) / 128 * 128 + 256;      // the tensor-core FC4 reads whole 256-row blocks (CTA pairs) from any chunk start
    RET(reserve(h, h->h3, tc ? rows_pad * 256 * 2 * sizeof(unsigned short) : (size_t)rows * 256 * sizeof(float)));
    // equal chunks (a short last chunk leaves most SMs idle for a whole pass): ceil(rows / nchunks), 32-row granular
    const int nchunks = (rows + h->dec_chunk - 1) / h->dec_chunk;
    const int ch = std::min(rows, ((rows + nchunks - 1) / nchunks + 31) / 32 * 32);
    const size_t esz = sizeof(float);   // fp32 planes, or bf16 hi+lo planes: 4 bytes per element either way
    // FC4's dropout bit planes depend on the noise key alone: they are generated for ALL rows of the call on the side
    // stream, next to the (latency-bound) FC1..3 launches, instead of per chunk between the decoder's kernels
    const bool fork_mask = tc && fc.nk.training && h->side != nullptr && !h->timer.on && (size_t)rows * 2048 <= ((size_t)1 << 30);
    RET(reserve(h, h->mask, (size_t)(fork_mask ? rows : ch) * 512 * sizeof(uint32_t)));
    RET(reserve(h, h->act0, (size_t)ch * 16384 * esz));
    RET(reserve(h, h->act1, (size_t)ch * 16384 * esz));
    RET(reserve(h, h->act2, std::max((size_t)ch * 65536 * esz, tc ? tc_ct23_scratch_bytes(ch) : (size_t)0)));
    RET(reserve(h, h->act3, tc ? (size_t)ch * PROJ_ROW_FLOATS * sizeof(float) : (size_t)ch * 131072 * esz));
    fc.h3 = tc ? nullptr : ptr<float>(h->h3);
    fc.h3b = tc ? ptr<unsigned short>(h->h3) : nullptr;
    fc.rows_pad = rows_pad;
    if (fork_mask) {
        CK(cudaEventRecord(h->ev_fork, st));
        CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        h->launches += launch_fc4_mask(fc.map, fc.nk, 0, rows, ptr<uint32_t>(h->mask), 1, h->side);
        CK(cudaEventRecord(h->ev_join, h->side));
    }
    if (!tc) {
        h->launches += launch_po_fc123(h->w, fc, st);
    } else {
        // FC1 on CUDA cores (K = 10), FC2 / FC3 on tensor cores; FC3's K-blocked output is FC4's A operand
        RET(reserve(h, h->mlpA, rows_pad * 512 * 2 * sizeof(unsigned short)));
        RET(reserve(h, h->mlpB, rows_pad * 512 * 2 * sizeof(unsigned short)));
        const NoiseRows nr = map_noise_rows(fc.map);
        std::string terr;
        h->launches += launch_po_l0(h->w, fc, rows_pad, h->mlpA.p, st);
        int n1 = tc_dense_hidden(h->tcw, TC_PO1, h->cfg.precision, h->mlpA.p, h->mlpB.p, rows, rows_pad, fc.nk, nr, 1, st, &terr);
        int n2 = n1 < 0 ? -1 : tc_dense_hidden(h->tcw, TC_PO2, h->cfg.precision, h->mlpB.p, h->h3.p, rows, rows_pad, fc.nk, nr, 2, st, &terr);
        if (n1 < 0 || n2 < 0) return fail(h, DAI_E_CUDA, "tensor-core decoder FCs: %s", terr.c_str());
        h->launches += n1 + n2;
    }
    for (int r0 = 0; r0 < rows; r0 += ch) {
        const int n = std::min(ch, rows - r0);
        const uint32_t* mask = nullptr;
        if (fc.nk.training) {
            if (fork_mask && r0 == 0) CK(cudaStreamWaitEvent(st, h->ev_join, 0));
            if (!fork_mask) h->launches += launch_fc4_mask(fc.map, fc.nk, r0, n, ptr<uint32_t>(h->mask), tc ? 1 : 0, st);
            mask = ptr<uint32_t>(h->mask) + (fork_mask ? (size_t)r0 * 512 : 0);
        }
        const float* h3c = tc ? nullptr : fc.h3 + (size_t)r0 * 256;
        Ct4Args c4{};
        c4.row0 = r0; c4.nrows = n; c4.img_rows = img_rows; c4.img = img; c4.hsum = hsum; c4.reward = reward;
        if (!tc) {
            h->launches += launch_fc4_simt(h->w, h3c, mask, n, ptr<float>(h->act0), st);
            h->launches += launch_ct1_simt(h->w, ptr<float>(h->act0), n, ptr<float>(h->act1), st);
            h->launches += launch_ct2_simt(h->w, ptr<float>(h->act1), n, ptr<float>(h->act2), st);
            h->launches += launch_ct3_simt(h->w, ptr<float>(h->act2), n, ptr<float>(h->act3), st);
            c4.act3 = ptr<float>(h->act3);
            h->launches += launch_ct4_efe(h->w, c4, st);
        } else {
            std::string terr;
            const int nl = tc_decoder_chunk(h->tcw, h->w, h->cfg.precision, fc.h3b, rows_pad, r0, mask, n, h->act0.p,
                                            h->act1.p, h->act2.p, h->act3.p, c4, st, &terr, &h->timer);
            if (nl < 0) return fail(h, DAI_E_CUDA, "tensor-core decoder: %s", terr.c_str());
            h->launches += nl;
        }
    }
    return post_launch(h, "decoder");
}

// Runs Qs on `rows` images with noise map (B, Sl, sample0, site); outputs [rows][10] each.  Rows are ordered (slot, b);
// a launch covers either whole slots (B <= chunk) or a run of rows of ONE slot (B > chunk: noise row = b0 + local row).
int run_encoder(dai_handle* h, cudaStream_t st, const float* img, int B, int Sl, int sample0, int site,
                const NoiseKey& nk, float* mean, float* logvar, float* samp, int sps = 0) {
    const int rows = B * Sl;
    if (rows <= 0) return DAI_OK;
    // equal chunks rather than a full one and a short one
    int slots_per = Sl, rows_per = B;                // launch = slots_per whole slots, or rows_per rows of one slot
    if (B > kQsChunk) {
        const int nchunks = (B + kQsChunk - 1) / kQsChunk;
        slots_per = 1; rows_per = (B + nchunks - 1) / nchunks;
    } else if (rows > kQsChunk) {
        const int slots_max = kQsChunk / B;
        const int nchunks = (Sl + slots_max - 1) / slots_max;
        slots_per = (Sl + nchunks - 1) / nchunks;
    }
    const int chs = B > kQsChunk ? rows_per : slots_per * B;     // largest launch
    const bool tc = h->cfg.precision != DAI_PREC_FP32_SIMT;
    RET(reserve(h, h->qc1, (size_t)chs * 32768 * sizeof(float)));   // (31,31,32) fp32, or 2 x 4 parities x (16,16,32) bf16
    RET(reserve(h, h->qc2, (size_t)chs * 8192 * sizeof(float)));    // (15,15,32) fp32, or 2 x 4 parities x (8,8,32) bf16
    RET(reserve(h, h->qc3, (size_t)chs * 49 * 64 * sizeof(float)));
    RET(reserve(h, h->qc4, tc ? std::max((size_t)chs * 576 * sizeof(float), tc_qs_conv4_scratch_bytes(chs)) : (size_t)chs * 576 * sizeof(float)));
    const size_t rp_max = ((size_t)chs + 127) / 128 * 128 + 128;
    if (tc) {
        RET(reserve(h, h->mlpA, rp_max * 576 * 2 * sizeof(unsigned short)));
        RET(reserve(h, h->mlpB, rp_max * 576 * 2 * sizeof(unsigned short)));
    }
    for (int s0 = 0; s0 < Sl; s0 += slots_per) {
        const int ns = std::min(slots_per, Sl - s0);
        for (int b0 = 0; b0 < B; b0 += rows_per) {
            const int nb = std::min(rows_per, B - b0);
            const int r0 = s0 * B + b0;                 // first row of this launch
            const int n = ns * nb;                      // (rows_per == B) or (ns == 1)
            QsArgs a{};
            a.img = img + (size_t)r0 * IMG; a.rows = n;
            a.map.B = nb; a.map.Sl = ns; a.map.sample0 = sample0; a.map.slot0 = s0; a.map.sps = sps; a.map.nsets = 1; a.map.b0 = b0;
            a.map.site[0] = site; a.map.site[1] = site; a.map.site[2] = site;
            a.c1 = ptr<float>(h->qc1); a.c2 = ptr<float>(h->qc2); a.c3 = ptr<float>(h->qc3); a.c4 = ptr<float>(h->qc4);
            a.mean = mean + (size_t)r0 * S_DIM; a.logvar = logvar + (size_t)r0 * S_DIM;
            a.samp = samp ? samp + (size_t)r0 * S_DIM : nullptr;
            a.nk = nk;
            if (!tc) {
                h->launches += launch_qs(h->w, a, st);
                continue;
            }
            // conv1 inside conv2's kernel (env DAI_TC_FUSE_C1=0: the two-kernel path through a 128 KB/row HBM activation)
            static const bool fuse_c1 = !(getenv("DAI_TC_FUSE_C1") && atoi(getenv("DAI_TC_FUSE_C1")) == 0);
            if (!fuse_c1) h->launches += launch_qs_conv1(h->w, a.img, n, nullptr, h->qc1.p, st);
            std::string terr;
            const int nl = tc_qs_convs(h->tcw, h->w, h->cfg.precision, h->qc1.p, h->qc2.p, a.c3, n, st, &terr, fuse_c1 ? a.img : nullptr);
            if (nl < 0) return fail(h, DAI_E_CUDA, "tensor-core encoder convs: %s", terr.c_str());
            h->launches += nl;
            // conv4 as im2col + GEMM -> K-blocked operand; FC1..3 on tensor cores; tail (256 -> 20) on CUDA cores
            const size_t rp = ((size_t)n + 127) / 128 * 128 + 128;
            const NoiseRows nr = map_noise_rows(a.map);
            const int n4 = tc_qs_conv4(h->tcw, h->cfg.precision, a.c3, n, h->qc4.p, rp, h->mlpA.p, st, &terr);
            if (n4 < 0) return fail(h, DAI_E_CUDA, "tensor-core encoder conv4: %s", terr.c_str());
            h->launches += n4;
            int n0 = tc_dense_hidden(h->tcw, TC_QS0, h->cfg.precision, h->mlpA.p, h->mlpB.p, n, rp, nk, nr, 0, st, &terr);
            int n1 = n0 < 0 ? -1 : tc_dense_hidden(h->tcw, TC_QS1, h->cfg.precision, h->mlpB.p, h->mlpA.p, n, rp, nk, nr, 1, st, &terr);
            int n2 = n1 < 0 ? -1 : tc_dense_hidden(h->tcw, TC_QS2, h->cfg.precision, h->mlpA.p, h->mlpB.p, n, rp, nk, nr, 2, st, &terr);
            if (n0 < 0 || n1 < 0 || n2 < 0) return fail(h, DAI_E_CUDA, "tensor-core encoder FCs: %s", terr.c_str());
            h->launches += n0 + n1 + n2 + launch_qs_tail20(h->w, a, rp, h->mlpB.p, st);
        }
    }
    return post_launch(h, "encoder");
}

// ---- one EFE step (calculate_G / calculate_G_mean) -----------------------------------

struct StepSpec {
    const float* s0 = nullptr;    // [B][10]
    const float* pi = nullptr;    // [B][4]
    int B = 0;
    int samples = 1, j0 = 0, j1 = 1;
    int mean_variant = 0;         // 1: calculate_G_mean (src/torchmodel.py:302-327)
    NoiseKey nk{};
    double* acc = nullptr;        // [4][B]
    float* carry_dst = nullptr;   // [B][10] <- ps1 (or ps1_mean when carry_mean) of the last sample
    int carry_mean = 0;
    float *out_ps1 = nullptr, *out_mean = nullptr, *out_logvar = nullptr, *out_po1 = nullptr;
};

int run_step(dai_handle* h, cudaStream_t st, const StepSpec& sp) {
    const int B = sp.B;
    const int Sl = sp.mean_variant ? 1 : (sp.j1 - sp.j0);
    const int S = sp.mean_variant ? 1 : sp.samples;
    const int j0 = sp.mean_variant ? 0 : sp.j0;
    const bool own_last = (j0 + Sl == S) && Sl > 0;
    const int nA = Sl + (own_last ? 0 : 1);
    const int last_slot = nA - 1;
    const size_t slab = (size_t)B * S_DIM;          // one [B][10] slab
    // ps buffer: meanA, logvarA, sampA [nA] ; meanB, sampB [Sl]
    RET(reserve(h, h->ps, (3 * (size_t)nA + 2 * (size_t)Sl) * slab * sizeof(float)));
    float* meanA = ptr<float>(h->ps);
    float* logvarA = meanA + nA * slab;
    float* sampA = logvarA + nA * slab;
    float* meanB = sampA + nA * slab;
    float* sampB = meanB + Sl * slab;
    PsArgs pa{};
    pa.pi = sp.pi; pa.s0 = sp.s0; pa.B = B; pa.nA = nA; pa.nB = Sl; pa.sample0 = j0;
    pa.extra_slot = own_last ? -1 : Sl; pa.extra_sample = S - 1;
    pa.siteA = SITE_PS_A; pa.siteB = SITE_PS_B;
    pa.meanA = meanA; pa.logvarA = logvarA; pa.sampA = sampA;
    pa.meanB = meanB; pa.logvarB = nullptr; pa.sampB = sampB;
    pa.nk = sp.nk;
    RET(run_ps(h, st, pa));

    const int rows = 3 * Sl * B;
    RET(reserve(h, h->img, (size_t)std::max(Sl, 1) * B * IMG * sizeof(float)));
    RET(reserve(h, h->hsum, (size_t)std::max(rows, 1) * sizeof(float)));
    RET(reserve(h, h->reward, (size_t)std::max(rows, 1) * sizeof(float)));
    RET(reserve(h, h->qs_out, (size_t)std::max(Sl, 1) * slab * 2 * sizeof(float)));
    float* qs_mean = ptr<float>(h->qs_out);
    float* qs_logvar = qs_mean + (size_t)std::max(Sl, 1) * slab;
    if (Sl > 0) {
        PoFcArgs fc{};
        fc.map.B = B; fc.map.Sl = Sl; fc.map.sample0 = j0; fc.map.nsets = 3;
        fc.map.site[0] = SITE_PO_A; fc.map.site[1] = SITE_PO_B1; fc.map.site[2] = SITE_PO_B2;
        fc.z[0] = sp.mean_variant ? meanA : sampA;
        fc.z[1] = sp.mean_variant ? meanB : sampB;
        fc.z[2] = nullptr;
        fc.mode[0] = 0; fc.mode[1] = 0; fc.mode[2] = 1;
        fc.rp_mean = meanA + last_slot * slab; fc.rp_logvar = logvarA + last_slot * slab; fc.rp_site = SITE_RP_B;
        fc.nk = sp.nk;
        RET(run_decoder(h, st, fc, Sl * B, ptr<float>(h->img), ptr<float>(h->hsum), ptr<float>(h->reward)));
        RET(run_encoder(h, st, ptr<float>(h->img), B, Sl, j0, SITE_QS_A, sp.nk, qs_mean, qs_logvar, nullptr));
    }
    StepFinalizeArgs fa{};
    fa.B = B; fa.Sl = Sl; fa.logvarA = logvarA; fa.qs_logvar = qs_logvar;
    fa.reward = ptr<float>(h->reward); fa.hsum = ptr<float>(h->hsum); fa.acc = sp.acc;
    fa.carry_src = sp.carry_dst ? ((sp.carry_mean ? meanA : sampA) + last_slot * slab) : nullptr;
    fa.carry_dst = sp.carry_dst;
    // outputs of the last sample are read before the carry overwrites anything they alias
    if (sp.out_ps1) CK(cudaMemcpyAsync(sp.out_ps1, sampA + last_slot * slab, slab * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (sp.out_mean) CK(cudaMemcpyAsync(sp.out_mean, meanA + last_slot * slab, slab * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (sp.out_logvar) CK(cudaMemcpyAsync(sp.out_logvar, logvarA + last_slot * slab, slab * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (sp.out_po1) {
        if (own_last) {
            CK(cudaMemcpyAsync(sp.out_po1, ptr<float>(h->img) + (size_t)(Sl - 1) * B * IMG, (size_t)B * IMG * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
        } else {
            // this rank does not hold the last sample: decode it once more (same noise key => same image on every rank)
            PoFcArgs fc{};
            fc.map.B = B; fc.map.Sl = 1; fc.map.sample0 = S - 1; fc.map.nsets = 1;
            fc.map.site[0] = SITE_PO_A; fc.map.site[1] = SITE_PO_A; fc.map.site[2] = SITE_PO_A;
            fc.z[0] = sampA + last_slot * slab; fc.mode[0] = 0;
            fc.nk = sp.nk;
            RET(reserve(h, h->scratch, (size_t)2 * B * sizeof(float)));
            RET(run_decoder(h, st, fc, B, sp.out_po1, ptr<float>(h->scratch), ptr<float>(h->scratch) + B));
        }
    }
    h->launches += launch_step_finalize(fa, st);
    return post_launch(h, "step finalize");
}

__global__ void k_set_key(uint32_t* keybuf, uint32_t k0, uint32_t k1, uint32_t step) {
    keybuf[0] = k0; keybuf[1] = k1; keybuf[2] = step;
}

uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
}

// A cached CUDA graph per launch sequence: the first evaluation with a given signature runs eagerly (it sizes the
// workspaces), the second is captured (stream capture of the very same launch code) and instantiated, later ones are
// ONE cudaGraphLaunch — the kernels of a step (or of a whole time-batched rollout) without per-launch host work or
// launch gaps.  The noise key of the call and the step index are the only things that change between replays; the
// kernels read them from h->keybuf (NoiseKey::dyn), which a one-thread kernel writes ahead of every launch.
// `body(dyn)` enqueues the sequence on `st` with NoiseKey::dyn = dyn (null: eager, the key travels as a parameter).
template <class Body>
int run_cached(dai_handle* h, cudaStream_t st, uint64_t sig, const NoiseKey& nk, Body&& body) {
    if (!h->graphs_enabled || h->timer.on) return body(nullptr);
    dai_handle::StepGraph* g = nullptr;
    for (auto& e : h->graphs)
        if (e.sig == sig) { g = &e; break; }
    if (g && g->gen != h->alloc_gen) {          // a workspace moved since: the captured pointers are stale
        if (g->exec) cudaGraphExecDestroy(g->exec);
        g->exec = nullptr; g->seen = 0; g->gen = h->alloc_gen;
    }
    if (!g) {
        if (h->graphs.size() >= 24) {           // drop the least recently used entry
            size_t lru = 0;
            for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].stamp < h->graphs[lru].stamp) lru = i;
            if (h->graphs[lru].exec) cudaGraphExecDestroy(h->graphs[lru].exec);
            h->graphs.erase(h->graphs.begin() + (long)lru);
        }
        h->graphs.push_back(dai_handle::StepGraph{});
        g = &h->graphs.back();
        g->sig = sig; g->gen = h->alloc_gen;
    }
    g->stamp = ++h->graph_stamp;
    if (!h->keybuf) CK(cudaMalloc(&h->keybuf, 16));
    if (g->exec) {
        k_set_key<<<1, 1, 0, st>>>(h->keybuf, nk.k0, nk.k1, nk.step);
        CK(cudaGraphLaunch(g->exec, st));
        h->launches += g->nlaunch + 1;
        ++h->graph_replays;
        return post_launch(h, "graph replay");
    }
    if (g->seen++ == 0) {                        // first time: eager (grows the workspaces), and see whether anything moved
        const uint64_t gen0 = h->alloc_gen;
        RET(body(nullptr));
        g->gen = h->alloc_gen;
        if (h->alloc_gen != gen0) g->seen = 1;
        return DAI_OK;
    }
    // capture
    k_set_key<<<1, 1, 0, st>>>(h->keybuf, nk.k0, nk.k1, nk.step);
    const uint64_t l0 = h->launches;
    h->capturing = true; h->capture_failed = false;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        h->capturing = false;
        h->graphs_enabled = 0;                   // this stream cannot be captured (e.g. the legacy default stream): stay eager
        return body(nullptr);
    }
    const int rc = body(h->keybuf);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    h->capturing = false;
    if (rc != DAI_OK || ce != cudaSuccess || !graph || h->capture_failed) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        h->launches = l0;
        g->seen = 1;                             // try again on a later call
        if (h->capture_failed) { h->capture_failed = false; return body(nullptr); }
        if (rc != DAI_OK) return rc;
        return body(nullptr);
    }
    cudaGraphExec_t exec = nullptr;
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        cudaGraphDestroy(graph);
        h->launches = l0;
        h->graphs_enabled = 0;
        return body(nullptr);
    }
    cudaGraphDestroy(graph);
    g->exec = exec; g->nlaunch = h->launches - l0; g->gen = h->alloc_gen;
    CK(cudaGraphLaunch(g->exec, st));
    h->launches += 1;
    ++h->graph_replays;
    return post_launch(h, "graph (first launch)");
}

uint64_t sig_of(std::initializer_list<uint64_t> fields) {
    uint64_t sig = 1469598103934665603ull;
    for (uint64_t f : fields) sig = mix64(sig, f);
    return sig;
}

// run_step through a cached graph (keyed by everything its launches depend on except the noise key)
int run_step_cached(dai_handle* h, cudaStream_t st, const StepSpec& sp) {
    const uint64_t sig = sig_of({1, (uint64_t)(uintptr_t)sp.s0, (uint64_t)(uintptr_t)sp.pi, (uint64_t)sp.B, (uint64_t)sp.samples, (uint64_t)sp.j0,
                                 (uint64_t)sp.j1, (uint64_t)sp.mean_variant, (uint64_t)(uintptr_t)sp.acc, (uint64_t)(uintptr_t)sp.carry_dst,
                                 (uint64_t)sp.carry_mean, (uint64_t)(uintptr_t)sp.out_ps1, (uint64_t)(uintptr_t)sp.out_mean,
                                 (uint64_t)(uintptr_t)sp.out_logvar, (uint64_t)(uintptr_t)sp.out_po1, (uint64_t)h->cfg.precision,
                                 (uint64_t)h->cfg.training, (uint64_t)h->dec_chunk, (uint64_t)(uintptr_t)st});
    return run_cached(h, st, sig, sp.nk, [&](const uint32_t* dyn) {
        StepSpec cs = sp;
        cs.nk.dyn = dyn;
        if (dyn) cs.nk.step = 0;                 // the replay's base step comes from the key buffer
        return run_step(h, st, cs);
    });
}

// ---- the horizon of a rollout, time-batched -------------------------------------------------------------------------
// The steps of calculate_G_repeated (src/torchmodel.py:329-349) are chained only through the transition net: the state a
// step starts from is the previous step's ps1 (or its mean) of the last MC sample, which Ps alone produces; the decoder
// and encoder passes of a step feed nothing but that step's own EFE terms.  So the latent chain of ALL steps runs first
// (4 small launches per step), and the pixel work of a GROUP of steps then runs as one launch sequence whose slots are
// (step, sample) pairs (RowMap::sps): a single root's 600 decoder rows per step become 6,000-row launches that fill
// the machine, and the ~25 latency-bound launches of a step are paid once per group instead of once per step.
// Row sets stay set-major ([set][step][sample][b]), so the images of set 0 are the encoder's input as they lie; the
// noise of a row is keyed on (sample, step) exactly as in a launch of that step alone, and the steps' fp64 sums are
// accumulated in step order — the results are those of the step-by-step schedule.
int run_rollout_steps(dai_handle* h, cudaStream_t st, const float* s_root, const float* pi, int B, int T, int samples, int j0in, int j1,
                      int mean_variant, int calc_mean, NoiseKey nk, double* acc, float* out_po1) {
    const int Sl = mean_variant ? 1 : (j1 - j0in);
    const int S = mean_variant ? 1 : samples;
    const int j0 = mean_variant ? 0 : j0in;
    const bool own_last = (j0 + Sl == S) && Sl > 0;
    const int nA = Sl + (own_last ? 0 : 1);
    const int last_slot = nA - 1;
    const size_t slab = (size_t)B * S_DIM;
    // ps buffer: meanA, logvarA, sampA [T][nA] ; meanB, sampB [T][Sl]
    RET(reserve(h, h->ps, (size_t)T * (3 * (size_t)nA + 2 * (size_t)Sl) * slab * sizeof(float)));
    float* meanA = ptr<float>(h->ps);
    float* logvarA = meanA + (size_t)T * nA * slab;
    float* sampA = logvarA + (size_t)T * nA * slab;
    float* meanB = sampA + (size_t)T * nA * slab;
    float* sampB = meanB + (size_t)T * Sl * slab;
    // 1. the latent chain
    for (int t = 0; t < T; ++t) {
        PsArgs pa{};
        pa.pi = pi;
        pa.s0 = t == 0 ? s_root : (calc_mean ? meanA : sampA) + ((size_t)(t - 1) * nA + last_slot) * slab;
        pa.B = B; pa.nA = nA; pa.nB = Sl; pa.sample0 = j0;
        pa.extra_slot = own_last ? -1 : Sl; pa.extra_sample = S - 1;
        pa.siteA = SITE_PS_A; pa.siteB = SITE_PS_B;
        pa.meanA = meanA + (size_t)t * nA * slab; pa.logvarA = logvarA + (size_t)t * nA * slab; pa.sampA = sampA + (size_t)t * nA * slab;
        pa.meanB = meanB + (size_t)t * Sl * slab; pa.logvarB = nullptr; pa.sampB = sampB + (size_t)t * Sl * slab;
        pa.nk = nk; pa.nk.step = nk.step + (uint32_t)t;
        RET(run_ps(h, st, pa));
    }
    if (Sl > 0) {
        // 2. pixel work, in equal groups of steps of up to 2.5 decoder chunks (R = 16, T = 10: two groups of five steps)
        const int rows_step = 3 * Sl * B;
        const int tg_max = std::max(1, std::min(T, std::min(255, (5 * h->dec_chunk / 2) / std::max(rows_step, 1))));
        const int ngroups = (T + tg_max - 1) / tg_max;
        const int tg = (T + ngroups - 1) / ngroups;
        RET(reserve(h, h->img, (size_t)tg * Sl * B * IMG * sizeof(float)));
        RET(reserve(h, h->hsum, (size_t)tg * rows_step * sizeof(float)));
        RET(reserve(h, h->reward, (size_t)tg * rows_step * sizeof(float)));
        RET(reserve(h, h->qs_out, (size_t)tg * Sl * slab * 2 * sizeof(float)));
        for (int t0 = 0; t0 < T; t0 += tg) {
            const int n = std::min(tg, T - t0);
            float* qs_mean = ptr<float>(h->qs_out);
            float* qs_logvar = qs_mean + (size_t)n * Sl * slab;
            NoiseKey gk = nk;
            gk.step = nk.step + (uint32_t)t0;
            PoFcArgs fc{};
            fc.map.B = B; fc.map.Sl = n * Sl; fc.map.sps = Sl; fc.map.sample0 = j0; fc.map.nsets = 3;
            fc.map.site[0] = SITE_PO_A; fc.map.site[1] = SITE_PO_B1; fc.map.site[2] = SITE_PO_B2;
            fc.z[0] = (mean_variant ? meanA : sampA) + (size_t)t0 * nA * slab; fc.zslots[0] = nA;
            fc.z[1] = (mean_variant ? meanB : sampB) + (size_t)t0 * Sl * slab; fc.zslots[1] = Sl;
            fc.z[2] = nullptr;
            fc.mode[0] = 0; fc.mode[1] = 0; fc.mode[2] = 1;
            fc.rp_mean = meanA + ((size_t)t0 * nA + last_slot) * slab; fc.rp_logvar = logvarA + ((size_t)t0 * nA + last_slot) * slab;
            fc.rp_tstride = (size_t)nA * slab; fc.rp_site = SITE_RP_B;
            fc.nk = gk;
            RET(run_decoder(h, st, fc, n * Sl * B, ptr<float>(h->img), ptr<float>(h->hsum), ptr<float>(h->reward)));
            RET(run_encoder(h, st, ptr<float>(h->img), B, n * Sl, j0, SITE_QS_A, gk, qs_mean, qs_logvar, nullptr, Sl));
            StepFinalizeArgs fa{};
            fa.B = B; fa.Sl = Sl; fa.T = n; fa.lvA_tstride = (size_t)nA * slab;
            fa.logvarA = logvarA + (size_t)t0 * nA * slab; fa.qs_logvar = qs_logvar;
            fa.reward = ptr<float>(h->reward); fa.hsum = ptr<float>(h->hsum); fa.acc = acc;
            if (out_po1 && own_last && t0 + n == T)
                CK(cudaMemcpyAsync(out_po1, ptr<float>(h->img) + ((size_t)(n - 1) * Sl + (Sl - 1)) * B * IMG, (size_t)B * IMG * sizeof(float),
                                   cudaMemcpyDeviceToDevice, st));
            h->launches += launch_step_finalize(fa, st);
        }
    }
    if (out_po1 && !own_last) {
        // this rank does not hold the last sample: decode it once more (same noise key => same image on every rank)
        PoFcArgs fc{};
        fc.map.B = B; fc.map.Sl = 1; fc.map.sample0 = S - 1; fc.map.nsets = 1;
        fc.map.site[0] = SITE_PO_A; fc.map.site[1] = SITE_PO_A; fc.map.site[2] = SITE_PO_A;
        fc.z[0] = sampA + ((size_t)(T - 1) * nA + last_slot) * slab; fc.mode[0] = 0;
        fc.nk = nk; fc.nk.step = nk.step + (uint32_t)(T - 1);
        RET(reserve(h, h->scratch, (size_t)2 * B * sizeof(float)));
        RET(run_decoder(h, st, fc, B, out_po1, ptr<float>(h->scratch), ptr<float>(h->scratch) + B));
    }
    return post_launch(h, "rollout steps");
}

__global__ void k_fill_eye(float* pi, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 4) pi[i] = ((i >> 2) & 3) == (i & 3) ? 1.0f : 0.0f;
}

int finish_outputs(dai_handle* h, cudaStream_t st, int B, int samples, double* sums, float* G, float* t0, float* t1, float* t2) {
    double* acc = ptr<double>(h->acc);
    if (sums) CK(cudaMemcpyAsync(sums, acc, (size_t)4 * B * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (G || t0 || t1 || t2) h->launches += launch_combine(acc, B, samples, G, t0, t1, t2, st);
    return post_launch(h, "combine");
}

int rollout_impl(dai_handle* h, cudaStream_t st, const float* o, const float* pi, int B, int steps, int samples,
                 int calc_mean, int four, int j0, int j1, double* sums, float* G, float* t0, float* t1, float* t2,
                 float* po1) {
    if (B <= 0 || steps <= 0 || samples <= 0 || j0 < 0 || j1 > samples || j0 > j1)
        return fail(h, DAI_E_INVALID, "rollout: bad sizes B=%d steps=%d samples=%d range=[%d,%d)", B, steps, samples, j0, j1);
    if (!pi && (B % 4) != 0) return fail(h, DAI_E_INVALID, "rollout: pi == NULL needs B %% 4 == 0 (row = root*4 + action)");
    const uint64_t call = h->call++;
    ++h->calls;
    NoiseKey nk = make_key(h, call, 0);
    const size_t slab = (size_t)B * S_DIM;
    RET(reserve(h, h->root, 3 * slab * sizeof(float)));
    RET(reserve(h, h->carry, slab * sizeof(float)));
    RET(reserve(h, h->acc, (size_t)4 * B * sizeof(double)));
    float* m0 = ptr<float>(h->root);
    float* lv0 = m0 + slab;
    float* smp0 = lv0 + slab;
    const int mean_variant = (four && calc_mean) ? 1 : 0;
    // qs0 = encoder(o) (+ reparameterize), src/torchmodel.py:228-229 / :248-249; the carry starts from it
    auto begin = [&](const NoiseKey& k, const float*& pp) -> int {
        RET(run_encoder(h, st, o, B, 1, 0, SITE_QS_ROOT, k, m0, lv0, smp0));
        CK(cudaMemcpyAsync(h->carry.p, calc_mean ? m0 : smp0, slab * sizeof(float), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemsetAsync(h->acc.p, 0, (size_t)4 * B * sizeof(double), st));
        if (!pp) {
            RET(reserve(h, h->pi_eye, (size_t)B * 4 * sizeof(float)));
            k_fill_eye<<<(B * 4 + 255) / 256, 256, 0, st>>>(ptr<float>(h->pi_eye), B);
            ++h->launches;
            pp = ptr<float>(h->pi_eye);
        }
        return DAI_OK;
    };
    // time-batched horizon (run_rollout_steps) on the tensor-core path — the whole call, root encode and final combine
    // included, is ONE cached graph; env DAI_TBATCH=0 or the fp32 CUDA-core precision: one launch sequence per step
    static const bool tbatch_env = !(getenv("DAI_TBATCH") && atoi(getenv("DAI_TBATCH")) == 0);
    if (tbatch_env && h->cfg.precision != DAI_PREC_FP32_SIMT && steps > 1 && samples <= (int)SAMPLE_MASK) {
        const uint64_t sig = sig_of({2, (uint64_t)(uintptr_t)o, (uint64_t)(uintptr_t)pi, (uint64_t)B, (uint64_t)steps, (uint64_t)samples,
                                     (uint64_t)j0, (uint64_t)j1, (uint64_t)mean_variant, (uint64_t)calc_mean, (uint64_t)(uintptr_t)po1,
                                     (uint64_t)(uintptr_t)sums, (uint64_t)(uintptr_t)G, (uint64_t)(uintptr_t)t0, (uint64_t)(uintptr_t)t1,
                                     (uint64_t)(uintptr_t)t2, (uint64_t)h->cfg.precision, (uint64_t)h->cfg.training, (uint64_t)h->dec_chunk,
                                     (uint64_t)(uintptr_t)st});
        return run_cached(h, st, sig, nk, [&](const uint32_t* dyn) -> int {
            NoiseKey k = nk;
            k.dyn = dyn;
            k.step = 0;
            const float* pp = pi;
            RET(begin(k, pp));
            RET(run_rollout_steps(h, st, ptr<float>(h->carry), pp, B, steps, samples, j0, j1, mean_variant, calc_mean, k,
                                  ptr<double>(h->acc), po1));
            return finish_outputs(h, st, B, mean_variant ? 1 : samples, sums, G, t0, t1, t2);
        });
    }
    RET(begin(nk, pi));
    for (int t = 0; t < steps; ++t) {
        StepSpec sp;
        sp.s0 = ptr<float>(h->carry); sp.pi = pi; sp.B = B;
        sp.samples = samples; sp.j0 = j0; sp.j1 = j1; sp.mean_variant = mean_variant;
        sp.nk = nk; sp.nk.step = (uint32_t)t;
        sp.acc = ptr<double>(h->acc);
        sp.carry_dst = ptr<float>(h->carry); sp.carry_mean = calc_mean;
        sp.out_po1 = (t == steps - 1) ? po1 : nullptr;
        RET(run_step_cached(h, st, sp));
    }
    return finish_outputs(h, st, B, mean_variant ? 1 : samples, sums, G, t0, t1, t2);
}

}  // namespace

// =======================================================================================
// C ABI
// =======================================================================================
extern "C" {

int dai_comm_destroy(dai_handle* h);

const char* dai_version(void) { return "dai_b200 0.1 (sm_100a)"; }

int dai_create(const dai_config* cfg, int device, dai_handle** out) {
    if (!cfg || !out) return DAI_E_INVALID;
    *out = nullptr;
    if (cfg->s_dim != 10 || cfg->pi_dim != 4 || cfg->resolution != 64 || cfg->colour_channels != 1)
        return DAI_E_UNSUPPORTED;   // the 32-px / 3-action branch is dead in the reference
    if (cfg->precision < DAI_PREC_FP32_SIMT || cfg->precision > DAI_PREC_BF16X1) return DAI_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return DAI_E_CUDA;
    }
    dai_handle* h = new (std::nothrow) dai_handle();
    if (!h) return DAI_E_NOMEM;
    h->cfg = *cfg;
    h->device = device;
    if (const char* e = getenv("DAI_DEC_CHUNK")) {
        const int v = atoi(e);
        if (v >= 32 && v <= 32768) h->dec_chunk = v;
    }
    if (const char* e = getenv("DAI_GRAPHS")) h->graphs_enabled = atoi(e) != 0;
    if (getenv("DAI_TC_COUNTERS") || getenv("DAI_TC_DBG")) h->graphs_enabled = 0;      // the experiment hooks synchronise
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return DAI_E_CUDA; }
    if (cudaMallocHost(&h->pinned, 4096) != cudaSuccess) { delete h; return DAI_E_NOMEM; }
    h->pinned_cap = 4096;
    if (h->graphs_enabled) {
        if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            h->graphs_enabled = 0;
        }
    }
    const char* fk = getenv("DAI_FORK");
    if (!(fk && atoi(fk) == 0)) {
        if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            if (h->side) cudaStreamDestroy(h->side);
            h->side = nullptr;
        }
    }
    *out = h;
    return DAI_OK;
}

int dai_destroy(dai_handle* h) {
    if (!h) return DAI_E_INVALID;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&h->mlpA, &h->mlpB, &h->ps, &h->zB, &h->h3, &h->mask, &h->act0, &h->act1, &h->act2, &h->act3, &h->img, &h->hsum,
                      &h->reward, &h->qc1, &h->qc2, &h->qc3, &h->qc4, &h->qs_out, &h->acc, &h->carry, &h->pi_eye,
                      &h->traj, &h->root, &h->stage_in, &h->stage_out, &h->scratch, &h->sprites, &h->sprite_stage, &h->frame_flag, &h->plan_tree, &h->plan_picks, &h->plan_rows, &h->plan_out, &h->plan_pi0};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    for (void* p : h->wallocs) cudaFree(p);
    for (float* p : h->raw_dev) if (p) cudaFree(p);
    for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->keybuf) cudaFree(h->keybuf);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side) cudaStreamDestroy(h->side);
    tc_release(&h->tcw);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->plan_stop_host) cudaFreeHost(h->plan_stop_host);
    if (h->plan_sel_ev) cudaEventDestroy(h->plan_sel_ev);
    dai_comm_destroy(h);
    delete h;
    return DAI_OK;
}

const char* dai_last_error(const dai_handle* h) { return h ? h->err.c_str() : "null handle"; }

static int set_weight_impl(dai_handle* h, const char* key, const float* data, const int64_t* shape, int ndim, cudaStream_t st, bool wait) {
    if (!h || !key || !data || !shape) return DAI_E_INVALID;
    const int i = spec_index(key);
    if (i < 0) return fail(h, DAI_E_INVALID, "unknown weight key %s", key);
    if (ndim != kSpecs[i].ndim) return fail(h, DAI_E_INVALID, "%s: ndim %d, expected %d", key, ndim, kSpecs[i].ndim);
    for (int d = 0; d < ndim; ++d)
        if (shape[d] != kSpecs[i].shape[d])
            return fail(h, DAI_E_INVALID, "%s: dim %d is %lld, expected %lld", key, d, (long long)shape[d], (long long)kSpecs[i].shape[d]);
    CK(cudaSetDevice(h->device));
    const size_t n = spec_elems(i);
    if (!h->raw_dev[i]) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(float));
        if (e != cudaSuccess) { cudaGetLastError(); return fail(h, DAI_E_NOMEM, "cudaMalloc(%s) failed: %s", key, cudaGetErrorString(e)); }
        h->raw_dev[i] = static_cast<float*>(p);
    }
    // host or device source (UVA): a device tensor never leaves the device
    CK(cudaMemcpyAsync(h->raw_dev[i], data, n * sizeof(float), cudaMemcpyDefault, st));
    if (wait) CK(cudaStreamSynchronize(st));
    h->have[i] = true; h->dirty[i] = true;
    h->committed = false;
    return DAI_OK;
}

int dai_set_weight(dai_handle* h, const char* key, const float* data, const int64_t* shape, int ndim) {
    return set_weight_impl(h, key, data, shape, ndim, nullptr, true);
}

int dai_set_weight_async(dai_handle* h, const char* key, const float* data, const int64_t* shape, int ndim, void* stream) {
    return set_weight_impl(h, key, data, shape, ndim, (cudaStream_t)stream, false);
}

int dai_commit_weights(dai_handle* h, void* stream) {
    if (!h) return DAI_E_INVALID;
    CK(cudaSetDevice(h->device));
    return commit(h, (cudaStream_t)stream);
}

int dai_set_rng(dai_handle* h, uint64_t seed, uint64_t call_index) {
    if (!h) return DAI_E_INVALID;
    h->seed = seed; h->call = call_index;
    return DAI_OK;
}

int dai_get_rng(const dai_handle* h, uint64_t* seed, uint64_t* call_index) {
    if (!h) return DAI_E_INVALID;
    if (seed) *seed = h->seed;
    if (call_index) *call_index = h->call;
    return DAI_OK;
}

int dai_set_training(dai_handle* h, int training) {
    if (!h) return DAI_E_INVALID;
    h->cfg.training = training ? 1 : 0;
    return DAI_OK;
}

int dai_set_precision(dai_handle* h, int precision) {
    if (!h) return DAI_E_INVALID;
    if (precision < DAI_PREC_FP32_SIMT || precision > DAI_PREC_BF16X1) return fail(h, DAI_E_INVALID, "unknown precision %d", precision);
    h->cfg.precision = precision;
    return DAI_OK;
}

int dai_get_stats(dai_handle* h, dai_stats* out, int reset) {
    if (!h || !out) return DAI_E_INVALID;
    out->kernel_launches = h->launches;
    out->calls = h->calls;
    size_t total = 0;
    DevBuf* bufs[] = {&h->mlpA, &h->mlpB, &h->ps, &h->zB, &h->h3, &h->mask, &h->act0, &h->act1, &h->act2, &h->act3, &h->img, &h->hsum,
                      &h->reward, &h->qc1, &h->qc2, &h->qc3, &h->qc4, &h->qs_out, &h->acc, &h->carry, &h->pi_eye,
                      &h->traj, &h->root, &h->stage_in, &h->stage_out, &h->scratch, &h->sprites, &h->sprite_stage, &h->frame_flag, &h->plan_tree, &h->plan_picks, &h->plan_rows, &h->plan_out, &h->plan_pi0};
    for (DevBuf* b : bufs) total += b->cap;
    out->workspace_bytes = total;
    out->repack_launches = h->repack_launches;
    if (reset) { h->launches = 0; h->calls = 0; }
    return DAI_OK;
}

int dai_encode(dai_handle* h, const float* o, int B, float* mean, float* logvar, float* sample, void* stream) {
    RET(check_ready(h));
    if (!o || !mean || !logvar || B <= 0) return fail(h, DAI_E_INVALID, "encode: bad arguments");
    const NoiseKey nk = make_key(h, h->call++, 0);
    ++h->calls;
    return run_encoder(h, (cudaStream_t)stream, o, B, 1, 0, SITE_QS_ROOT, nk, mean, logvar, sample);
}

int dai_decode(dai_handle* h, const float* s, int B, float* po, void* stream) {
    RET(check_ready(h));
    if (!s || !po || B <= 0) return fail(h, DAI_E_INVALID, "decode: bad arguments");
    PoFcArgs fc{};
    fc.map.B = B; fc.map.Sl = 1; fc.map.sample0 = 0; fc.map.nsets = 1;
    fc.map.site[0] = fc.map.site[1] = fc.map.site[2] = SITE_PO_A;
    fc.z[0] = s; fc.mode[0] = 0;
    fc.nk = make_key(h, h->call++, 0);
    ++h->calls;
    RET(reserve(h, h->scratch, (size_t)2 * B * sizeof(float)));
    return run_decoder(h, (cudaStream_t)stream, fc, B, po, ptr<float>(h->scratch), ptr<float>(h->scratch) + B);
}

int dai_transition(dai_handle* h, const float* pi, const float* s0, int B, float* mean, float* logvar, float* sample,
                   void* stream) {
    RET(check_ready(h));
    if (!pi || !s0 || !mean || !logvar || B <= 0) return fail(h, DAI_E_INVALID, "transition: bad arguments");
    PsArgs pa{};
    pa.pi = pi; pa.s0 = s0; pa.B = B; pa.nA = 1; pa.nB = 0; pa.sample0 = 0; pa.extra_slot = -1;
    pa.siteA = SITE_PS_A; pa.siteB = SITE_PS_B;
    pa.meanA = mean; pa.logvarA = logvar; pa.sampA = sample;
    pa.nk = make_key(h, h->call++, 0);
    ++h->calls;
    return run_ps(h, (cudaStream_t)stream, pa);
}

int dai_habit(dai_handle* h, const float* s, int B, float* logits, float* q, float* logq, void* stream) {
    RET(check_ready(h));
    if (!s || B <= 0) return fail(h, DAI_E_INVALID, "habit: bad arguments");
    ++h->calls;
    h->launches += launch_qpi(h->w, s, B, logits, q, logq, (cudaStream_t)stream);
    return post_launch(h, "habit");
}

int dai_check_reward(dai_handle* h, const float* o, int B, float* r, void* stream) {
    if (!h) return DAI_E_INVALID;
    CK(cudaSetDevice(h->device));
    if (!o || !r || B <= 0) return fail(h, DAI_E_INVALID, "check_reward: bad arguments");
    ++h->calls;
    h->launches += launch_reward_only(o, B, r, (cudaStream_t)stream);
    return post_launch(h, "check_reward");
}

int dai_calculate_G(dai_handle* h, const float* s0, const float* pi0, int B, int samples, int sample_begin,
                    int sample_end, double* sums, float* G, float* t0, float* t1, float* t2, float* ps1,
                    float* ps1_mean, float* ps1_logvar, float* po1, void* stream) {
    RET(check_ready(h));
    if (!s0 || !pi0 || B <= 0 || samples <= 0 || sample_begin < 0 || sample_end > samples || sample_begin > sample_end)
        return fail(h, DAI_E_INVALID, "calculate_G: bad arguments");
    StreamSwap sw(h, (cudaStream_t)stream);
    cudaStream_t st = sw.use;
    RET(reserve(h, h->acc, (size_t)4 * B * sizeof(double)));
    CK(cudaMemsetAsync(h->acc.p, 0, (size_t)4 * B * sizeof(double), st));
    StepSpec sp;
    sp.s0 = s0; sp.pi = pi0; sp.B = B; sp.samples = samples; sp.j0 = sample_begin; sp.j1 = sample_end;
    sp.nk = make_key(h, h->call++, 0);
    ++h->calls;
    sp.acc = ptr<double>(h->acc);
    sp.out_ps1 = ps1; sp.out_mean = ps1_mean; sp.out_logvar = ps1_logvar; sp.out_po1 = po1;
    RET(run_step_cached(h, st, sp));
    return finish_outputs(h, st, B, samples, sums, G, t0, t1, t2);
}

int dai_calculate_G_mean(dai_handle* h, const float* s0, const float* pi0, int B, float* G, float* t0, float* t1,
                         float* t2, float* ps1_mean, float* po1, void* stream) {
    RET(check_ready(h));
    if (!s0 || !pi0 || B <= 0) return fail(h, DAI_E_INVALID, "calculate_G_mean: bad arguments");
    StreamSwap sw(h, (cudaStream_t)stream);
    cudaStream_t st = sw.use;
    RET(reserve(h, h->acc, (size_t)4 * B * sizeof(double)));
    CK(cudaMemsetAsync(h->acc.p, 0, (size_t)4 * B * sizeof(double), st));
    StepSpec sp;
    sp.s0 = s0; sp.pi = pi0; sp.B = B; sp.mean_variant = 1;
    sp.nk = make_key(h, h->call++, 0);
    ++h->calls;
    sp.acc = ptr<double>(h->acc);
    sp.out_mean = ps1_mean; sp.out_po1 = po1;
    RET(run_step_cached(h, st, sp));
    return finish_outputs(h, st, B, 1, nullptr, G, t0, t1, t2);
}

// calculate_G_given_trajectory over D rows = D / depth trajectories of `depth` rows; G (D) and Gmean (D / depth) device.
static int trajectory_impl(dai_handle* h, cudaStream_t st, const float* s0, const float* ps1, const float* ps1_mean,
                           const float* ps1_logvar, const float* pi0, int D, int depth, float* G, float* Gmean) {
    const NoiseKey nk = make_key(h, h->call++, 0);
    ++h->calls;
    const size_t slab = (size_t)D * S_DIM;
    RET(reserve(h, h->ps, slab * sizeof(float)));
    float* sampB = ptr<float>(h->ps);
    PsArgs pa{};
    pa.pi = pi0; pa.s0 = s0; pa.B = D; pa.nA = 0; pa.nB = 1; pa.sample0 = 0; pa.extra_slot = -1;
    pa.siteA = SITE_PS_A; pa.siteB = SITE_PS_B; pa.sampB = sampB; pa.nk = nk;
    RET(run_ps(h, st, pa));
    RET(reserve(h, h->img, (size_t)D * IMG * sizeof(float)));
    RET(reserve(h, h->hsum, (size_t)3 * D * sizeof(float)));
    RET(reserve(h, h->reward, (size_t)3 * D * sizeof(float)));
    RET(reserve(h, h->qs_out, 2 * slab * sizeof(float)));
    PoFcArgs fc{};
    fc.map.B = D; fc.map.Sl = 1; fc.map.sample0 = 0; fc.map.nsets = 3;
    fc.map.site[0] = SITE_PO_A; fc.map.site[1] = SITE_PO_B1; fc.map.site[2] = SITE_PO_B2;
    fc.z[0] = ps1; fc.z[1] = sampB; fc.mode[0] = 0; fc.mode[1] = 0; fc.mode[2] = 1;
    fc.rp_mean = ps1_mean; fc.rp_logvar = ps1_logvar; fc.rp_site = SITE_RP_B; fc.nk = nk;
    RET(run_decoder(h, st, fc, D, ptr<float>(h->img), ptr<float>(h->hsum), ptr<float>(h->reward)));
    float* qs_mean = ptr<float>(h->qs_out);
    float* qs_logvar = qs_mean + slab;
    RET(run_encoder(h, st, ptr<float>(h->img), D, 1, 0, SITE_QS_A, nk, qs_mean, qs_logvar, nullptr));
    h->launches += launch_traj_G(ptr<float>(h->reward), ptr<float>(h->hsum), ps1_logvar, qs_logvar, D, depth, G, Gmean, st);
    return post_launch(h, "trajectory G");
}

int dai_G_given_trajectory(dai_handle* h, const float* s0, const float* ps1, const float* ps1_mean,
                           const float* ps1_logvar, const float* pi0, int D, float* G, void* stream) {
    RET(check_ready(h));
    if (!s0 || !ps1 || !ps1_mean || !ps1_logvar || !pi0 || D <= 0 || D > 256)
        return fail(h, DAI_E_INVALID, "G_given_trajectory: bad arguments (1 <= depth <= 256)");
    RET(reserve(h, h->scratch, 16 * sizeof(float)));
    return trajectory_impl(h, (cudaStream_t)stream, s0, ps1, ps1_mean, ps1_logvar, pi0, D, D, G, ptr<float>(h->scratch));
}

int dai_rollout(dai_handle* h, const float* o, const float* pi, int B, int steps, int samples, int calc_mean, int four,
                int sample_begin, int sample_end, double* sums, float* G, float* t0, float* t1, float* t2, float* po1,
                void* stream) {
    RET(check_ready(h));
    if (!o) return fail(h, DAI_E_INVALID, "rollout: o is NULL");
    StreamSwap sw(h, (cudaStream_t)stream);
    return rollout_impl(h, sw.use, o, pi, B, steps, samples, calc_mean, four, sample_begin, sample_end,
                        sums, G, t0, t1, t2, po1);
}

int dai_combine(dai_handle* h, const double* sums, int B, int samples, float* G, float* t0, float* t1, float* t2,
                void* stream) {
    if (!h || !sums || B <= 0 || samples <= 0) return DAI_E_INVALID;
    CK(cudaSetDevice(h->device));
    h->launches += launch_combine(sums, B, samples, G, t0, t1, t2, (cudaStream_t)stream);
    return post_launch(h, "combine");
}

int dai_rollout_host(dai_handle* h, const float* o_host, const float* pi_host, int B, int steps, int samples,
                     int calc_mean, int four, float* G_host, float* t0_host, float* t1_host, float* t2_host,
                     float* po1_host, void* stream) {
    RET(check_ready(h));
    if (!o_host || B <= 0) return fail(h, DAI_E_INVALID, "rollout_host: bad arguments");
    StreamSwap sw(h, (cudaStream_t)stream);
    cudaStream_t st = sw.use;
    const size_t in_bytes = (size_t)B * IMG * sizeof(float) + (size_t)B * 4 * sizeof(float);
    const size_t out_f = (size_t)4 * B + (po1_host ? (size_t)B * IMG : 0);
    RET(reserve(h, h->stage_in, in_bytes));
    RET(reserve(h, h->stage_out, out_f * sizeof(float)));
    float* o_dev = ptr<float>(h->stage_in);
    float* pi_dev = o_dev + (size_t)B * IMG;
    CK(cudaMemcpyAsync(o_dev, o_host, (size_t)B * IMG * sizeof(float), cudaMemcpyHostToDevice, st));
    if (pi_host) CK(cudaMemcpyAsync(pi_dev, pi_host, (size_t)B * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
    float* Gd = ptr<float>(h->stage_out);
    float* po1d = po1_host ? Gd + 4 * (size_t)B : nullptr;
    RET(rollout_impl(h, st, o_dev, pi_host ? pi_dev : nullptr, B, steps, samples, calc_mean, four, 0, samples, nullptr,
                     Gd, Gd + B, Gd + 2 * B, Gd + 3 * B, po1d));
    if (G_host) CK(cudaMemcpyAsync(G_host, Gd, B * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (t0_host) CK(cudaMemcpyAsync(t0_host, Gd + B, B * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (t1_host) CK(cudaMemcpyAsync(t1_host, Gd + 2 * B, B * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (t2_host) CK(cudaMemcpyAsync(t2_host, Gd + 3 * B, B * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (po1_host) CK(cudaMemcpyAsync(po1_host, po1d, (size_t)B * IMG * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return DAI_OK;
}

// K rollouts + one trajectory evaluation, enqueued only: the K per-trajectory means are left in h->scratch (device)
static int simulate_batch_enqueue(dai_handle* h, cudaStream_t st, const float* starting_s, int K, int depth, int use_means,
                                  float* pi0, float* qpi) {
    const size_t slab = (size_t)K * depth * S_DIM;
    RET(reserve(h, h->traj, 4 * slab * sizeof(float)));
    RET(reserve(h, h->scratch, (size_t)std::max(K, 16) * sizeof(float)));
    float* s0 = ptr<float>(h->traj);
    SimArgs sa{};
    sa.start = starting_s; sa.K = K; sa.depth = depth; sa.use_means = use_means;
    sa.s0 = s0; sa.ps1 = s0 + slab; sa.mean = s0 + 2 * slab; sa.logvar = s0 + 3 * slab;
    sa.pi0 = pi0; sa.qpi = qpi;
    sa.nk = make_key(h, h->call++, 0);
    h->launches += launch_sim_rollout(h->w, sa, st);
    RET(post_launch(h, "simulate rollout"));
    // calculate_G_given_trajectory over the K*depth rows (next call index), per-trajectory mean (src/torchmodel.py:392)
    return trajectory_impl(h, st, sa.s0, sa.ps1, sa.mean, sa.logvar, pi0, K * depth, depth, nullptr, ptr<float>(h->scratch));
}

int dai_mcts_simulate_batch(dai_handle* h, const float* starting_s, int K, int depth, int use_means, float* G_host,
                            float* pi0, float* qpi, void* stream) {
    RET(check_ready(h));
    if (!starting_s || !G_host || !pi0 || !qpi || K <= 0 || K > 4096 || depth <= 0 || depth > 256)
        return fail(h, DAI_E_INVALID, "mcts_simulate: bad arguments (1 <= K <= 4096, 1 <= depth <= 256)");
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)K * sizeof(float) > h->pinned_cap) {
        float* np = nullptr;
        CK(cudaMallocHost(&np, (size_t)K * sizeof(float)));
        cudaFreeHost(h->pinned);
        h->pinned = np; h->pinned_cap = (size_t)K * sizeof(float);
    }
    RET(simulate_batch_enqueue(h, st, starting_s, K, depth, use_means, pi0, qpi));
    CK(cudaMemcpyAsync(h->pinned, h->scratch.p, (size_t)K * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int k = 0; k < K; ++k) G_host[k] = h->pinned[k];
    return DAI_OK;
}

int dai_mcts_simulate(dai_handle* h, const float* starting_s, int depth, int use_means, float* G_host, float* pi0,
                      float* qpi, void* stream) {
    return dai_mcts_simulate_batch(h, starting_s, 1, depth, use_means, G_host, pi0, qpi, stream);
}

int dai_mcts_plan(dai_handle* h, const float* frame, const float* qs0_mean_in, const dai_mcts_params* prm,
                  dai_mcts_result* res, int32_t* path_host, int32_t* all_paths_host, int32_t* all_len_host,
                  float* all_G_host, void* stream) {
    RET(check_ready(h));
    if ((!frame && !qs0_mean_in) || !prm || !res || !path_host)
        return fail(h, DAI_E_INVALID, "mcts_plan: bad arguments");
    const int K = prm->leaves, R = prm->repeats, depth = prm->simulation_depth, nrep = prm->simulation_repeats;
    if (K < 1 || K > PLAN_MAX_K || R < 0 || R > 100000 || depth < 1 || depth > 256 || nrep < 1 || prm->samples < 1)
        return fail(h, DAI_E_INVALID, "mcts_plan: 1 <= leaves <= %d, repeats >= 0, 1 <= simulation_depth <= 256, "
                    "simulation_repeats >= 1, samples >= 1", PLAN_MAX_K);
    StreamSwap sw(h, (cudaStream_t)stream);
    cudaStream_t st = sw.use;
    if (!h->plan_stop_host) {
        CK(cudaHostAlloc(&h->plan_stop_host, 64, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(&h->plan_stop_dev, h->plan_stop_host, 0));
    }
    if (!h->plan_sel_ev) CK(cudaEventCreateWithFlags(&h->plan_sel_ev, cudaEventDisableTiming));
    *(volatile int32_t*)h->plan_stop_host = 0;      // the previous decision ended with a stream wait: no kernel writes it now
    // ---- carve the tree, the picks and the row buffers
    const int cap = 1 + PI_DIM * (R + K + 2), log_cap = std::max(R, 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t oW = take((size_t)cap * PI_DIM * 4), oN = take((size_t)cap * PI_DIM * 4), oQ = take((size_t)cap * PI_DIM * 4),
                 oC = take((size_t)cap * PI_DIM * 4), oS = take((size_t)cap * S_DIM * 4), oB = take((size_t)cap),
                 oCtl = take(PLAN_NCTL * 4), oLa = take((size_t)log_cap * PLAN_MAX_DEPTH * 4), oLl = take((size_t)log_cap * 4),
                 oLg = take((size_t)log_cap * 4);
    RET(reserve(h, h->plan_tree, off));
    uint8_t* tb = ptr<uint8_t>(h->plan_tree);
    PlanTree t{};
    t.W = (float*)(tb + oW); t.N = (float*)(tb + oN); t.Qpi = (float*)(tb + oQ); t.child = (int32_t*)(tb + oC);
    t.state = (float*)(tb + oS); t.blocked = tb + oB; t.ctl = (int32_t*)(tb + oCtl); t.host_stop = h->plan_stop_dev;
    t.cap = cap; t.use_prior = prm->using_prior_for_exploration ? 1 : 0; t.C = prm->C;
    t.log_actions = (int32_t*)(tb + oLa); t.log_len = (int32_t*)(tb + oLl); t.log_G = (float*)(tb + oLg); t.log_cap = log_cap;
    off = 0;
    const size_t pLeaf = take(K * 4), pLen = take(K * 4), pNodes = take((size_t)K * PLAN_MAX_DEPTH * 4),
                 pActs = take((size_t)K * PLAN_MAX_DEPTH * 4), pCnt = take(4), pScr = take((size_t)K * (PLAN_MAX_DEPTH + 1) * 4);
    RET(reserve(h, h->plan_picks, off));
    uint8_t* pb = ptr<uint8_t>(h->plan_picks);
    PlanPicks pk{};
    pk.leaf = (int32_t*)(pb + pLeaf); pk.len = (int32_t*)(pb + pLen); pk.nodes = (int32_t*)(pb + pNodes);
    pk.actions = (int32_t*)(pb + pActs); pk.count = (int32_t*)(pb + pCnt); pk.scratch = (int32_t*)(pb + pScr);
    // rows: s_rows (4K,10), starts (K,10), G (4K), nxt (4K,10), sims (K), qpi (K,4), root mean/logvar (10 each), root qpi (4)
    off = 0;
    const size_t rS = take((size_t)4 * K * S_DIM * 4), rSt = take((size_t)K * S_DIM * 4), rG = take((size_t)4 * K * 4),
                 rNx = take((size_t)4 * K * S_DIM * 4), rSim = take((size_t)K * 4), rQ = take((size_t)K * PI_DIM * 4),
                 rM = take(S_DIM * 4), rLv = take(S_DIM * 4), rRq = take(PI_DIM * 4), rPi = take((size_t)4 * K * PI_DIM * 4);
    RET(reserve(h, h->plan_rows, off));
    uint8_t* rb = ptr<uint8_t>(h->plan_rows);
    float *s_rows = (float*)(rb + rS), *starts = (float*)(rb + rSt), *Gd = (float*)(rb + rG), *nxt = (float*)(rb + rNx),
          *sims = (float*)(rb + rSim), *qpi = (float*)(rb + rQ), *m0 = (float*)(rb + rM), *lv0 = (float*)(rb + rLv),
          *rootq = (float*)(rb + rRq), *pi_eye = (float*)(rb + rPi);
    RET(reserve(h, h->plan_pi0, (size_t)K * depth * PI_DIM * 4));
    RET(reserve(h, h->plan_out, (size_t)(1 + PLAN_MAX_DEPTH) * 4));
    k_fill_eye<<<(4 * K * 4 + 255) / 256, 256, 0, st>>>(pi_eye, 4 * K);
    ++h->launches;

    // ---- root: qs0 = encoder mean (src/mcts.py:158), habit prior (:164), first expansion (:172)
    const float* root_mean = qs0_mean_in;
    if (!root_mean) {
        RET(dai_encode(h, frame, 1, m0, lv0, nullptr, (void*)st));
        root_mean = m0;
    }
    RET(dai_habit(h, root_mean, 1, nullptr, rootq, nullptr, (void*)st));
    h->launches += launch_plan_init(t, root_mean, rootq, st);
    auto expand = [&](int kv) -> int {          // kv <= 0: the root itself
        const int rows = PI_DIM * std::max(kv, 1);
        h->launches += launch_plan_select(t, kv, prm->threshold, pk, s_rows, starts, st);
        CK(cudaEventRecord(h->plan_sel_ev, st));
        if (prm->use_means) RET(dai_calculate_G_mean(h, s_rows, pi_eye, rows, Gd, nullptr, nullptr, nullptr, nxt, nullptr, (void*)st));
        else RET(dai_calculate_G(h, s_rows, pi_eye, rows, prm->samples, 0, prm->samples, nullptr, Gd, nullptr, nullptr, nullptr,
                                 nxt, nullptr, nullptr, nullptr, (void*)st));
        h->launches += launch_plan_expand(t, pk, Gd, nxt, st);
        return post_launch(h, "planner expansion");
    };
    RET(expand(0));
    const uint64_t call_after_root = h->call;
    const uint64_t calls_per_batch = 1 + 2 * (uint64_t)nrep;    // one EFE evaluation + nrep x (rollout, trajectory)
    int done = 0, leaves_now = PI_DIM;
    while (done < R) {
        // The threshold test of batch i runs in its selection kernel (src/mcts.py:176).  Wait for the PREVIOUS selection
        // before enqueuing the next batch: the GPU still has that batch's evaluation queued, so it never starves, and the
        // host is never more than one batch ahead of a stop (the batch whose selection raised it is a no-op on the device).
        CK(cudaEventSynchronize(h->plan_sel_ev));
        if (*(volatile int32_t*)h->plan_stop_host) break;
        const int kv = std::min(std::min(K, leaves_now), R - done);
        RET(expand(kv));
        for (int r = 0; r < nrep; ++r) {
            RET(simulate_batch_enqueue(h, st, starts, kv, depth, 0, ptr<float>(h->plan_pi0), qpi));
            h->launches += launch_plan_accumulate(ptr<float>(h->scratch), sims, kv, r == 0, st);
        }
        h->launches += launch_plan_backprop(t, pk, sims, nrep, qpi, st);
        RET(post_launch(h, "planner back-propagation"));
        done += kv; leaves_now += 3 * kv;
    }
    // ---- result: one wait per decision
    h->launches += launch_plan_finish(t, ptr<int32_t>(h->plan_out), st);
    RET(post_launch(h, "planner finish"));
    int32_t out[1 + PLAN_MAX_DEPTH], ctl[PLAN_NCTL];
    CK(cudaMemcpyAsync(out, h->plan_out.p, sizeof(out), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctl, t.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ctl[PLAN_ERR]) return fail(h, DAI_E_UNSUPPORTED, "mcts_plan: %s", ctl[PLAN_ERR] == 1 ? "search path deeper than 64 edges" : "tree capacity exceeded");
    res->path_len = out[0]; res->repeats_done = ctl[PLAN_DONE]; res->stopped = ctl[PLAN_STOP]; res->logged = ctl[PLAN_LOGGED];
    {   // the call index after a decision counts only the batches the search used (a batch enqueued behind a stop drew
        // nothing that was kept), so it equals the host-driven planner's and does not depend on timing
        int d = 0, lv = PI_DIM;
        uint64_t nb = 0;
        while (d < ctl[PLAN_DONE]) { const int kv = std::min(std::min(K, lv), R - d); d += kv; lv += 3 * kv; ++nb; }
        h->call = call_after_root + nb * calls_per_batch;
    }
    if (out[0] > PLAN_MAX_DEPTH) return fail(h, DAI_E_UNSUPPORTED, "mcts_plan: decision path deeper than 64 edges");
    for (int i = 0; i < out[0]; ++i) path_host[i] = out[1 + i];
    const int nlog = std::min(ctl[PLAN_LOGGED], log_cap);
    if (nlog > 0 && all_paths_host && all_len_host && all_G_host) {
        CK(cudaMemcpyAsync(all_paths_host, t.log_actions, (size_t)nlog * PLAN_MAX_DEPTH * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(all_len_host, t.log_len, (size_t)nlog * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(all_G_host, t.log_G, (size_t)nlog * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return DAI_OK;
}

int dai_select_actions(dai_handle* h, const float* G, int R, float temperature, float* Ppi, float* logPpi,
                       int32_t* choice, void* stream) {
    if (!h) return DAI_E_INVALID;
    CK(cudaSetDevice(h->device));
    if (!G || R <= 0 || !(temperature > 0.0f)) return fail(h, DAI_E_INVALID, "select_actions: bad arguments");
    const NoiseKey nk = make_key(h, h->call++, 0);
    ++h->calls;
    h->launches += launch_select_actions(G, R, temperature, nk, Ppi, logPpi, choice, (cudaStream_t)stream);
    return post_launch(h, "select_actions");
}

int dai_frames_set_sprites(dai_handle* h, const uint8_t* imgs_host, int64_t count, const int32_t* latents_sizes, void* stream) {
    if (!h) return DAI_E_INVALID;
    CK(cudaSetDevice(h->device));
    if (!imgs_host || !latents_sizes || count <= 0) return fail(h, DAI_E_INVALID, "frames_set_sprites: bad arguments");
    long long prod = 1;
    for (int i = 0; i < 6; ++i) {
        if (latents_sizes[i] <= 0) return fail(h, DAI_E_INVALID, "frames_set_sprites: latents_sizes[%d] = %d", i, latents_sizes[i]);
        prod *= latents_sizes[i];
    }
    if (prod != count) return fail(h, DAI_E_INVALID, "frames_set_sprites: %lld sprites but latents_sizes multiply to %lld", (long long)count, prod);
    cudaStream_t st = (cudaStream_t)stream;
    RET(reserve(h, h->sprites, (size_t)count * 512));
    RET(reserve(h, h->frame_flag, sizeof(int32_t)));
    const long long chunk = 16384;                              // 64 MB of uint8 pixels per staging pass
    RET(reserve(h, h->sprite_stage, (size_t)std::min<long long>(chunk, count) * 4096));
    for (long long first = 0; first < count; first += chunk) {
        const long long n = std::min<long long>(chunk, count - first);
        CK(cudaMemcpyAsync(h->sprite_stage.p, imgs_host + first * 4096, (size_t)n * 4096, cudaMemcpyHostToDevice, st));
        h->launches += launch_pack_sprites(ptr<uint8_t>(h->sprite_stage), n, ptr<uint32_t>(h->sprites), first, st);
        RET(post_launch(h, "pack sprites"));
        CK(cudaStreamSynchronize(st));                          // the staging buffer (and pageable host memory) is reused
    }
    h->sprite_count = count;
    long long pv = 1;
    for (int i = 5; i >= 0; --i) { h->sizes[i] = latents_sizes[i]; h->place[i] = pv; pv *= latents_sizes[i]; }
    return DAI_OK;
}

int dai_frames_render(dai_handle* h, const float* s, int s_stride, const float* last_r, int G, int reference_bases,
                      float* o, int32_t* n_bad_host, void* stream) {
    if (!h) return DAI_E_INVALID;
    CK(cudaSetDevice(h->device));
    if (h->sprite_count <= 0) return fail(h, DAI_E_INVALID, "frames_render: no sprite table (dai_frames_set_sprites)");
    if (!s || !last_r || !o || G <= 0 || s_stride < 6) return fail(h, DAI_E_INVALID, "frames_render: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    FrameArgs a{};
    a.s = s; a.s_stride = s_stride; a.last_r = last_r; a.bits = ptr<uint32_t>(h->sprites); a.count = h->sprite_count;
    static const long long kReferenceBases[6] = {1, 3, 6, 40, 32, 32};     // Game.s_bases, src/game_environment.py:25
    for (int i = 0; i < 6; ++i) a.base[i] = reference_bases ? kReferenceBases[i] : h->place[i];
    a.o = o; a.n_bad = ptr<int32_t>(h->frame_flag);
    ++h->calls;
    CK(cudaMemsetAsync(h->frame_flag.p, 0, sizeof(int32_t), st));
    h->launches += launch_render_frames(a, G, st);
    RET(post_launch(h, "render frames"));
    if (n_bad_host) {
        CK(cudaMemcpyAsync(h->pinned, h->frame_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        *n_bad_host = *reinterpret_cast<int32_t*>(h->pinned);
    }
    return DAI_OK;
}

int dai_profile_begin(dai_handle* h) {
    if (!h) return DAI_E_INVALID;
    for (auto& r : h->timer.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    h->timer.recs.clear();
    h->timer.on = true;
    return DAI_OK;
}

int dai_profile_end(dai_handle* h, float ms[5], int64_t launches[5], int64_t rows[5], void* stream) {
    if (!h || !ms || !launches || !rows) return DAI_E_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    h->timer.on = false;
    for (int i = 0; i < 5; ++i) { ms[i] = 0.0f; launches[i] = 0; rows[i] = 0; }
    for (auto& r : h->timer.recs) {
        float t = 0.0f;
        if (r.b && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.layer >= 0 && r.layer < 5) {
            ms[r.layer] += t; launches[r.layer] += 1; rows[r.layer] += r.rows;
        }
        cudaEventDestroy(r.a);
        if (r.b) cudaEventDestroy(r.b);
    }
    h->timer.recs.clear();
    return DAI_OK;
}

int dai_debug_layer(dai_handle* h, int layer, int precision, const float* in, int nrows, float* out, void* stream) {
    RET(check_ready(h));
    if (!in || !out || nrows <= 0 || ((layer < 1 || layer > 3) && layer != 23)) return fail(h, DAI_E_INVALID, "debug_layer: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (layer == 23) {
        // ct2 -> ct3 fused pair kernel: (nrows,256,64) fp32 NHWC in, the last deconv's row planes (nrows,3,4096) out
        if (precision == DAI_PREC_FP32_SIMT) return fail(h, DAI_E_INVALID, "debug_layer 23 is the tensor-core pair kernel");
        RET(reserve(h, h->act0, (size_t)nrows * 256 * 64 * 4));
        RET(reserve(h, h->act2, tc_ct23_scratch_bytes(nrows)));
        RET(reserve(h, h->act3, (size_t)nrows * PROJ_ROW_FLOATS * sizeof(float)));
        h->launches += tc_to_blocked(in, nrows, 256, 64, h->act0.p, st);
        std::string terr;
        h->timer.begin(2, nrows, st);
        const int nl = tc_ct23(h->tcw, h->w, precision, h->act0.p, h->act2.p, h->act3.p, nrows, st, &terr);
        if (nl < 0) return fail(h, DAI_E_CUDA, "tensor-core pair kernel: %s", terr.c_str());
        h->timer.end(st);
        h->launches += nl + launch_proj_rows(ptr<float>(h->act3), nrows, out, st);
        return post_launch(h, "debug layer (ct2+ct3 pair)");
    }
    if (precision == DAI_PREC_FP32_SIMT) {
        if (layer == 1) h->launches += launch_ct1_simt(h->w, in, nrows, out, st);
        if (layer == 2) h->launches += launch_ct2_simt(h->w, in, nrows, out, st);
        if (layer == 3) h->launches += launch_ct3_simt(h->w, in, nrows, out, st);
        return post_launch(h, "debug layer (simt)");
    }
    const int hw_in = layer == 3 ? 1024 : 256, hw_out = layer == 1 ? 256 : (layer == 2 ? 1024 : 4096);
    RET(reserve(h, h->act0, (size_t)nrows * hw_in * 64 * 4));
    RET(reserve(h, h->act1, (size_t)nrows * hw_out * 64 * 4));
    h->launches += tc_to_blocked(in, nrows, hw_in, 64, h->act0.p, st);
    std::string terr;
    if (layer == 3) RET(reserve(h, h->act3, (size_t)nrows * PROJ_ROW_FLOATS * sizeof(float)));
    void* dst = layer == 3 ? h->act3.p : h->act1.p;
    h->timer.begin(layer, nrows, st);
    const int nl = tc_layer(h->tcw, h->w, precision, layer, h->act0.p, dst, nrows, st, &terr);
    if (nl < 0) return fail(h, DAI_E_CUDA, "tensor-core layer: %s", terr.c_str());
    h->timer.end(st);
    h->launches += nl;
    if (layer != 3) h->launches += tc_from_blocked(h->act1.p, nrows, hw_out, 64, out, st);
    else h->launches += launch_proj_rows(ptr<float>(h->act3), nrows, out, st);
    return post_launch(h, "debug layer (tc)");
}

}  // extern "C"

// =======================================================================================
// Sample-shard communicator: the ONE collective of a sharded rollout (SURVEY.md §8 e) lives behind the C ABI.
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy already loaded into the process — e.g. torch's — or
// the system one), so the library has no link-time dependency on it and single-GPU users never touch it.
// =======================================================================================
namespace {

struct NcclId { char internal[128]; };          // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { api.why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return api; }
    api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(lib, "ncclCommInitRank"));
    api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(lib, "ncclAllReduce"));
    api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(lib, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(lib, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    if (!api.ok) api.why = "libnccl.so.2 lacks a required symbol";
    return api;
}

constexpr int kNcclFloat64 = 8, kNcclSum = 0;   // nccl.h: ncclFloat64, ncclSum

void shard_of(int samples, int rank, int world, int* j0, int* j1) {
    // contiguous; the first samples % world ranks hold one more (sharding.shard_range)
    const int base = samples / world, rem = samples % world;
    *j0 = rank * base + std::min(rank, rem);
    *j1 = *j0 + base + (rank < rem ? 1 : 0);
}

int allreduce_sums(dai_handle* h, double* sums, int B, cudaStream_t st) {
    if (h->comm_world == 1) return DAI_OK;
    const int rc = nccl().AllReduce(sums, sums, (size_t)4 * B, kNcclFloat64, kNcclSum, h->comm, st);
    if (rc != 0) return fail(h, DAI_E_CUDA, "ncclAllReduce: %s", nccl().GetErrorString(rc));
    return DAI_OK;
}

}  // namespace

extern "C" {

int dai_comm_unique_id(void* id128) {
    if (!id128) return DAI_E_INVALID;
    if (!nccl().ok) return DAI_E_UNSUPPORTED;
    NcclId id;
    if (nccl().GetUniqueId(&id) != 0) return DAI_E_CUDA;
    memcpy(id128, &id, sizeof(id));
    return DAI_OK;
}

int dai_comm_init(dai_handle* h, const void* id128, int rank, int world) {
    if (!h || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) return fail(h, DAI_E_INVALID, "comm_init: bad arguments");
    CK(cudaSetDevice(h->device));
    if (h->comm) { nccl().CommDestroy(h->comm); h->comm = nullptr; }
    h->comm_rank = rank; h->comm_world = world;
    if (world == 1) return DAI_OK;
    if (!nccl().ok) return fail(h, DAI_E_UNSUPPORTED, "comm_init: NCCL unavailable (%s)", nccl().why.c_str());
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = nccl().CommInitRank(&h->comm, world, id, rank);
    if (rc != 0) { h->comm = nullptr; h->comm_world = 1; h->comm_rank = 0; return fail(h, DAI_E_CUDA, "ncclCommInitRank: %s", nccl().GetErrorString(rc)); }
    return DAI_OK;
}

int dai_comm_destroy(dai_handle* h) {
    if (!h) return DAI_E_INVALID;
    if (h->comm) { cudaSetDevice(h->device); nccl().CommDestroy(h->comm); h->comm = nullptr; }
    h->comm_rank = 0; h->comm_world = 1;
    return DAI_OK;
}

int dai_comm_info(const dai_handle* h, int* rank, int* world) {
    if (!h) return DAI_E_INVALID;
    if (rank) *rank = h->comm_rank;
    if (world) *world = h->comm_world;
    return DAI_OK;
}

int dai_rollout_sharded(dai_handle* h, const float* o, const float* pi, int B, int steps, int samples, int calc_mean, int four,
                        float* G, float* t0, float* t1, float* t2, float* po1, void* stream) {
    RET(check_ready(h));
    if (!o) return fail(h, DAI_E_INVALID, "rollout_sharded: o is NULL");
    StreamSwap sw(h, (cudaStream_t)stream);
    cudaStream_t st = sw.use;
    int j0 = 0, j1 = samples;
    const bool sharded = h->comm_world > 1 && !(four && calc_mean);    // calculate_G_mean steps have one sample: nothing to shard
    if (sharded) shard_of(samples, h->comm_rank, h->comm_world, &j0, &j1);
    if (!sharded) return rollout_impl(h, st, o, pi, B, steps, samples, calc_mean, four, 0, samples, nullptr, G, t0, t1, t2, po1);
    RET(rollout_impl(h, st, o, pi, B, steps, samples, calc_mean, four, j0, j1, nullptr, nullptr, nullptr, nullptr, nullptr, po1));
    RET(allreduce_sums(h, ptr<double>(h->acc), B, st));
    return finish_outputs(h, st, B, samples, nullptr, G, t0, t1, t2);
}

int dai_calculate_G_sharded(dai_handle* h, const float* s0, const float* pi0, int B, int samples, float* G, float* t0, float* t1,
                            float* t2, float* ps1, float* ps1_mean, float* ps1_logvar, float* po1, void* stream) {
    RET(check_ready(h));
    if (!s0 || !pi0 || B <= 0 || samples <= 0) return fail(h, DAI_E_INVALID, "calculate_G_sharded: bad arguments");
    StreamSwap sw(h, (cudaStream_t)stream);
    cudaStream_t st = sw.use;
    int j0 = 0, j1 = samples;
    if (h->comm_world > 1) shard_of(samples, h->comm_rank, h->comm_world, &j0, &j1);
    RET(reserve(h, h->acc, (size_t)4 * B * sizeof(double)));
    CK(cudaMemsetAsync(h->acc.p, 0, (size_t)4 * B * sizeof(double), st));
    StepSpec sp;
    sp.s0 = s0; sp.pi = pi0; sp.B = B; sp.samples = samples; sp.j0 = j0; sp.j1 = j1;
    sp.nk = make_key(h, h->call++, 0);
    ++h->calls;
    sp.acc = ptr<double>(h->acc);
    sp.out_ps1 = ps1; sp.out_mean = ps1_mean; sp.out_logvar = ps1_logvar; sp.out_po1 = po1;
    RET(run_step(h, st, sp));
    RET(allreduce_sums(h, ptr<double>(h->acc), B, st));
    return finish_outputs(h, st, B, samples, nullptr, G, t0, t1, t2);
}

}  // extern "C"
