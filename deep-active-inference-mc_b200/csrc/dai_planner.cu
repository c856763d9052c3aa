// Device-resident search tree of the batched-leaf planner (reference: src/mcts.py:11-128,150-195).
// The tree is a structure of arrays in HBM; selection (the reference's argmax descent, made K times with virtual
// visits so the K leaves are distinct), expansion bookkeeping, back-propagation and the final most-visited path are
// single-warp kernels between the EFE evaluations, so a decision needs no host round trip until its end.
// All arithmetic is fp32 in the reference's operation order (torch CPU semantics: NaN wins an argmax).
#include "dai_kernels.h"

namespace dai {

namespace {

__device__ __forceinline__ bool better(float v, float best) { return v > best || (isnan(v) && !isnan(best)); }

__device__ __forceinline__ int argmax4(const float* v) {          // torch.argmax: first maximum, NaN is a maximum
    int a = 0;
    for (int i = 1; i < PI_DIM; ++i) if (better(v[i], v[a])) a = i;
    return a;
}

// src/mcts.py:39-47: Q normalised to a distribution plus the C/N exploration bonus
__device__ void selection_scores(const PlanTree& t, int i, float* sc) {
    float q[PI_DIM], mn, sum = 0.0f;
    for (int a = 0; a < PI_DIM; ++a) q[a] = __fdiv_rn(t.W[i * PI_DIM + a], t.N[i * PI_DIM + a]);
    mn = q[0];
    for (int a = 1; a < PI_DIM; ++a) if (q[a] < mn || isnan(q[a])) mn = q[a];
    for (int a = 0; a < PI_DIM; ++a) { q[a] = __fsub_rn(q[a], mn); sum = __fadd_rn(sum, q[a]); }
    for (int a = 0; a < PI_DIM; ++a) {
        const float qn = __fdiv_rn(q[a], sum);
        const float n = t.N[i * PI_DIM + a];
        const float bonus = t.use_prior ? __fdiv_rn(__fmul_rn(__fmul_rn(t.C, t.Qpi[i * PI_DIM + a]), 1.0f), n)
                                        : __fdiv_rn(__fmul_rn(t.C, 1.0f), n);
        sc[a] = __fadd_rn(qn, bonus);
    }
}

__device__ __forceinline__ bool is_leaf(const PlanTree& t, int i) {
    for (int a = 0; a < PI_DIM; ++a) if (t.child[i * PI_DIM + a] < 0) return true;
    return false;
}

}  // namespace

// root node from the encoder mean and the habit prior; statistics cleared
__global__ void k_plan_init(PlanTree t, const float* qs0_mean, const float* qpi) {
    const int tid = threadIdx.x;
    for (int i = tid; i < t.cap * PI_DIM; i += blockDim.x) { t.W[i] = 0.0f; t.N[i] = 0.0f; t.Qpi[i] = 0.0f; t.child[i] = -1; }
    for (int i = tid; i < t.cap; i += blockDim.x) t.blocked[i] = 0;
    if (tid < S_DIM) t.state[tid] = qs0_mean[tid];
    __syncthreads();
    if (tid < PI_DIM) t.Qpi[tid] = qpi[tid];
    if (tid == 0) { t.ctl[PLAN_SIZE] = 1; t.ctl[PLAN_DONE] = 0; t.ctl[PLAN_STOP] = 0; t.ctl[PLAN_ERR] = 0; t.ctl[PLAN_LOGGED] = 0; }
}

// One selection round: the threshold test of the search loop (src/mcts.py:176), then up to k descents.  k <= 0 selects
// the root itself with an empty path (the root expansion, src/mcts.py:172).  Writes the picks and gathers the states
// the expansion (4 rows per leaf) and the simulations (1 row per leaf) will read.
__global__ void k_plan_select(PlanTree t, int k, float threshold, PlanPicks p, float* s_rows, float* starts) {
    if (threadIdx.x != 0) return;
    int npick = 0;
    if (k <= 0) {
        p.leaf[0] = 0; p.len[0] = 0;
        npick = 1;
    } else if (!t.ctl[PLAN_STOP]) {
        // calc_threshold(normalization(root.N)) > threshold  (src/mcts.py:131-135,176)
        float x[PI_DIM], s = 0.0f, mx, mean = 0.0f;
        for (int a = 0; a < PI_DIM; ++a) s = __fadd_rn(s, t.N[a]);
        for (int a = 0; a < PI_DIM; ++a) x[a] = __fdiv_rn(t.N[a], s);
        mx = x[0];
        for (int a = 1; a < PI_DIM; ++a) if (better(x[a], mx)) mx = x[a];
        for (int a = 0; a < PI_DIM; ++a) mean = __fadd_rn(mean, x[a]);
        mean = __fdiv_rn(mean, (float)PI_DIM);
        if (__fsub_rn(mx, mean) > threshold) {
            t.ctl[PLAN_STOP] = 1;
            *t.host_stop = 1;
        } else {
            // Tree.select_batch (deep-active-inference-mc_b200/mcts.py): virtual visits on a scratch copy of W / N
            int touched[PLAN_MAX_K * PLAN_MAX_DEPTH];
            float w0[PLAN_MAX_K * PLAN_MAX_DEPTH], n0[PLAN_MAX_K * PLAN_MAX_DEPTH];
            int ntouched = 0, nblocked = 0;
            int* blist = p.scratch;                       // nodes blocked in this round (cleared at the end)
            while (npick < k && !t.blocked[0]) {
                int cur = 0, len = 0;
                int* nodes = p.nodes + npick * PLAN_MAX_DEPTH;
                int* acts = p.actions + npick * PLAN_MAX_DEPTH;
                while (true) {
                    float sc[PI_DIM];
                    selection_scores(t, cur, sc);
                    for (int a = 0; a < PI_DIM; ++a) if (t.blocked[t.child[cur * PI_DIM + a]]) sc[a] = -INFINITY;
                    const int a = argmax4(sc);
                    if (len >= PLAN_MAX_DEPTH) { t.ctl[PLAN_ERR] = 1; break; }
                    nodes[len] = cur; acts[len] = a; ++len;           // the edge (cur, a); its child comes next
                    cur = t.child[cur * PI_DIM + a];
                    if (is_leaf(t, cur)) break;
                }
                p.leaf[npick] = cur; p.len[npick] = len;
                ++npick;
                t.blocked[cur] = 1; blist[nblocked++] = cur;
                for (int d = len - 1; d >= 0; --d) {                  // a node with no unblocked child is exhausted
                    const int up = nodes[d];
                    bool all = true;
                    for (int a = 0; a < PI_DIM; ++a) all = all && t.blocked[t.child[up * PI_DIM + a]];
                    if (!all) break;
                    t.blocked[up] = 1; blist[nblocked++] = up;
                }
                for (int d = 0; d < len; ++d) {                        // one virtual visit per edge, mean kept
                    const int e = nodes[d] * PI_DIM + acts[d];
                    touched[ntouched] = e; w0[ntouched] = t.W[e]; n0[ntouched] = t.N[e]; ++ntouched;
                    const float q = __fdiv_rn(t.W[e], t.N[e]);
                    t.N[e] = __fadd_rn(t.N[e], 1.0f);
                    t.W[e] = __fmul_rn(q, t.N[e]);
                }
            }
            for (int i = ntouched - 1; i >= 0; --i) { t.W[touched[i]] = w0[i]; t.N[touched[i]] = n0[i]; }
            for (int i = 0; i < nblocked; ++i) t.blocked[blist[i]] = 0;
        }
    }
    p.count[0] = npick;
    // rows for the model calls; slots without a pick read the root (their results are ignored)
    const int kk = k <= 0 ? 1 : k;
    for (int j = 0; j < kk; ++j) {
        const int leaf = j < npick ? p.leaf[j] : 0;
        for (int d = 0; d < S_DIM; ++d) {
            const float v = t.state[leaf * S_DIM + d];
            for (int a = 0; a < PI_DIM; ++a) s_rows[(j * PI_DIM + a) * S_DIM + d] = v;
            starts[j * S_DIM + d] = v;
        }
    }
}

// Tree.expand_batch bookkeeping: W -= G, N += 1, four children per leaf from the next states (src/mcts.py:82-85)
__global__ void k_plan_expand(PlanTree t, PlanPicks p, const float* G, const float* nxt) {
    if (threadIdx.x != 0) return;
    const int npick = p.count[0];
    int size = t.ctl[PLAN_SIZE];
    for (int j = 0; j < npick; ++j) {
        const int leaf = p.leaf[j];
        if (size + PI_DIM > t.cap) { t.ctl[PLAN_ERR] = 2; break; }
        for (int a = 0; a < PI_DIM; ++a) {
            const int e = leaf * PI_DIM + a;
            t.W[e] = __fsub_rn(t.W[e], G[j * PI_DIM + a]);
            t.N[e] = __fadd_rn(t.N[e], 1.0f);
            t.child[e] = size + a;
            for (int d = 0; d < S_DIM; ++d) t.state[(size + a) * S_DIM + d] = nxt[(j * PI_DIM + a) * S_DIM + d];
        }
        size += PI_DIM;
    }
    t.ctl[PLAN_SIZE] = size;
}

// sims[k] (+)= G[k] over the simulation repeats
__global__ void k_plan_accumulate(const float* g, float* sims, int n, int first) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < n) sims[i] = first ? __fadd_rn(0.0f, g[i]) : __fadd_rn(sims[i], g[i]);
}

// Qpi of the expanded leaves, back-propagation of the mean simulated G along each path (src/mcts.py:88-96,186-191)
__global__ void k_plan_backprop(PlanTree t, PlanPicks p, const float* sims, int nrep, const float* qpi) {
    if (threadIdx.x != 0) return;
    const int npick = p.count[0];
    for (int j = 0; j < npick; ++j) {
        const int leaf = p.leaf[j];
        for (int a = 0; a < PI_DIM; ++a) t.Qpi[leaf * PI_DIM + a] = qpi[j * PI_DIM + a];
        const float g = __fdiv_rn(sims[j], (float)nrep);
        const int len = p.len[j];
        for (int d = 0; d < len; ++d) {
            const int e = p.nodes[j * PLAN_MAX_DEPTH + d] * PI_DIM + p.actions[j * PLAN_MAX_DEPTH + d];
            t.W[e] = __fsub_rn(t.W[e], g);
            t.N[e] = __fadd_rn(t.N[e], 1.0f);
        }
        const int slot = t.ctl[PLAN_LOGGED];
        if (slot < t.log_cap) {
            t.log_len[slot] = len;
            t.log_G[slot] = g;
            for (int d = 0; d < len; ++d) t.log_actions[slot * PLAN_MAX_DEPTH + d] = p.actions[j * PLAN_MAX_DEPTH + d];
        }
        t.ctl[PLAN_LOGGED] = slot + 1;
    }
    t.ctl[PLAN_DONE] += npick;
}

// Tree.most_visited_path before trimming (src/mcts.py:98-106): argmax N down to a leaf
__global__ void k_plan_finish(PlanTree t, int* out /* [0] = len, [1..] = actions */) {
    if (threadIdx.x != 0) return;
    int cur = 0, len = 0;
    while (true) {
        const int a = argmax4(t.N + cur * PI_DIM);
        if (len < PLAN_MAX_DEPTH) out[1 + len] = a;
        ++len;
        cur = t.child[cur * PI_DIM + a];
        if (cur < 0 || is_leaf(t, cur)) break;
    }
    out[0] = len;
}

int launch_plan_init(const PlanTree& t, const float* qs0_mean, const float* qpi, cudaStream_t st) {
    k_plan_init<<<1, 256, 0, st>>>(t, qs0_mean, qpi);
    return 1;
}
int launch_plan_select(const PlanTree& t, int k, float threshold, const PlanPicks& p, float* s_rows, float* starts, cudaStream_t st) {
    k_plan_select<<<1, 32, 0, st>>>(t, k, threshold, p, s_rows, starts);
    return 1;
}
int launch_plan_expand(const PlanTree& t, const PlanPicks& p, const float* G, const float* nxt, cudaStream_t st) {
    k_plan_expand<<<1, 32, 0, st>>>(t, p, G, nxt);
    return 1;
}
int launch_plan_accumulate(const float* g, float* sims, int n, int first, cudaStream_t st) {
    k_plan_accumulate<<<(n + 127) / 128, 128, 0, st>>>(g, sims, n, first);
    return 1;
}
int launch_plan_backprop(const PlanTree& t, const PlanPicks& p, const float* sims, int nrep, const float* qpi, cudaStream_t st) {
    k_plan_backprop<<<1, 32, 0, st>>>(t, p, sims, nrep, qpi);
    return 1;
}
int launch_plan_finish(const PlanTree& t, int* out, cudaStream_t st) {
    k_plan_finish<<<1, 32, 0, st>>>(t, out);
    return 1;
}

}  // namespace dai
