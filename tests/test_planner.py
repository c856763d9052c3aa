"""The planner (dai_b200.mcts) against the reference's src/mcts.py (where /root/reference is mounted) and, on
the GPU, the CUDA model against the oracle under the planner."""
import os
import sys

import numpy as np
import pytest
import torch

import cases
from oracle import efe_oracle as O

REF = "/root/reference"


def _frame(seed=3):
    import dai_b200.synthetic as syn
    return torch.from_numpy(syn.make_frames(1, seed))[0, 0]          # (64, 64) like game.current_frame


def _params(mod, repeats, use_means, threshold, depth=2, samples=1):
    p = mod.MCTS_Params()
    p.repeats, p.use_means, p.threshold, p.simulation_depth = repeats, use_means, threshold, depth
    if hasattr(p, "samples"):
        p.samples = samples
    return p


def _ints(paths):
    return [[int(a) for a in pth] for pth in paths]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (authoring container only)")
@pytest.mark.parametrize("use_means,threshold", [(True, 2.0), (False, 2.0), (True, 0.5)])
def test_same_decision_as_reference_planner(use_means, threshold):
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.mcts as ref_mcts                    # the reference planner, unmodified
    from dai_b200 import mcts as my_mcts
    w = cases.weights_for("w0")
    a = O.OracleModel(w, seed=77)
    b = O.OracleModel(w, seed=77)
    ra = ref_mcts.active_inference_mcts(a, _frame(), _params(ref_mcts, 5, use_means, threshold), o_shape=(1, 64, 64))
    rb = my_mcts.active_inference_mcts(b, _frame(), _params(my_mcts, 5, use_means, threshold), o_shape=(1, 64, 64))
    assert [int(x) for x in ra[0]] == [int(x) for x in rb[0]]
    assert ra[1] == rb[1] and ra[2] == rb[2]
    assert _ints(ra[3]) == _ints(rb[3])
    assert np.allclose(ra[4], rb[4], rtol=0, atol=0)
    assert a.call == b.call                         # the model was called the same number of times, in the same order


def test_planner_structure_and_samples_param():
    from dai_b200 import mcts as my_mcts
    m = O.OracleModel(cases.weights_for("w0"), seed=5)
    p = _params(my_mcts, 3, False, 2.0, depth=2, samples=2)
    path, reps, explored, all_paths, all_G = my_mcts.active_inference_mcts(m, _frame(), p, o_shape=(1, 64, 64))
    assert reps == 3 and explored == 3 * 2 and len(all_paths) == 3 and len(all_G) == 3
    assert all(0 <= a < 4 for a in path)
    assert my_mcts.active_inference_mcts(m, [], p)[0] == [0]


@pytest.mark.gpu
@pytest.mark.parametrize("use_means", [True, False])
def test_cuda_model_plans_like_the_oracle(use_means):
    from dai_b200 import mcts as my_mcts
    from dai_b200.torchmodel import ActiveInferenceModel
    w = cases.weights_for("w0")
    gpu = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(w)
    gpu.set_rng(77, 0)
    ora = O.OracleModel(w, seed=77)
    pg = my_mcts.active_inference_mcts(gpu, _frame(), _params(my_mcts, 6, use_means, 2.0, depth=3, samples=2), o_shape=(1, 64, 64))
    po = my_mcts.active_inference_mcts(ora, _frame(), _params(my_mcts, 6, use_means, 2.0, depth=3, samples=2), o_shape=(1, 64, 64))
    assert pg[0] == po[0] and pg[1] == po[1] and pg[3] == po[3]
    assert np.allclose(pg[4], po[4], rtol=1e-4, atol=1e-3)


# ---- batched-leaf planner ------------------------------------------------------------------

@pytest.mark.parametrize("use_means", [True, False])
def test_batched_planner_with_one_leaf_is_the_sequential_planner(use_means):
    from dai_b200 import mcts as my_mcts
    w = cases.weights_for("w0")
    a, b = O.OracleModel(w, seed=77), O.OracleModel(w, seed=77)
    p = _params(my_mcts, 5, use_means, 2.0, depth=2, samples=2)
    ra = my_mcts.active_inference_mcts(a, _frame(), p, o_shape=(1, 64, 64))
    rb = my_mcts.active_inference_mcts_batched(b, _frame(), p, o_shape=(1, 64, 64), leaves=1)
    assert ra[0] == rb[0] and ra[1] == rb[1] and ra[2] == rb[2] and ra[3] == rb[3]
    assert np.allclose(ra[4], rb[4], rtol=0, atol=1e-6)
    assert a.call == b.call


def test_batched_planner_claims_distinct_leaves_and_counts_expansions():
    from dai_b200 import mcts as my_mcts
    m = O.OracleModel(cases.weights_for("w0"), seed=5)
    p = _params(my_mcts, 10, True, 2.0, depth=2)
    calls0 = m.call
    path, reps, explored, all_paths, all_G = my_mcts.active_inference_mcts_batched(m, _frame(), p, o_shape=(1, 64, 64), leaves=4)
    assert reps == 10 and explored == 10 * 2 and len(all_paths) == 10 and len(all_G) == 10
    for i in (0, 4):                                   # the leaves of one batch are distinct
        assert len({tuple(x) for x in all_paths[i:i + 4]}) == 4
    assert len({tuple(x) for x in all_paths[8:10]}) == 2
    assert all(0 <= a < 4 for a in path)
    # model calls: encoder, habit net (no call index), root expansion, then per batch 1 expansion + 2 (simulate) = 3 batches
    assert m.call - calls0 == 1 + 1 + 3 * 3


def test_select_batch_restores_statistics_and_exhausts_small_trees():
    from dai_b200.mcts import Tree
    t = Tree(4, 64, 1.0, False)
    root = t.add(torch.zeros(10))
    for a in range(4):
        t.child[root, a] = t.add(torch.zeros(10))
    t.W[root] = torch.tensor([-50.0, -52.0, -49.0, -51.0])
    t.N[root] = 1.0
    W0, N0 = t.W.clone(), t.N.clone()
    picks = t.select_batch(root, 6)                     # only 4 leaves exist
    assert sorted(a[0] for _, a in picks) == [0, 1, 2, 3]
    assert picks[0][1] == [2]                           # best Q first, as the sequential rule would
    assert torch.equal(t.W, W0) and torch.equal(t.N, N0)


def test_oracle_batched_simulation_rows_are_keyed_by_leaf():
    w = cases.weights_for("w0")
    a, b = O.OracleModel(w, seed=9), O.OracleModel(w, seed=9)
    s = torch.from_numpy(np.random.default_rng(2).normal(size=(3, 10)).astype(np.float32))
    G1, p1, q1 = a.mcts_step_simulate(s[0], 3)
    Gb, pb, qb = b.mcts_step_simulate_batch(s[:1], 3)
    assert abs(G1 - float(Gb[0])) < 1e-6 and torch.equal(p1, pb[0]) and torch.equal(q1, qb[0])
    G3, p3, q3 = b.mcts_step_simulate_batch(s, 3)
    assert G3.shape == (3,) and p3.shape == (3, 3, 4) and q3.shape == (3, 4)
    assert torch.equal(p3.sum(-1), torch.ones(3, 3))


@pytest.mark.gpu
def test_cuda_batched_simulation_matches_oracle():
    from dai_b200.torchmodel import ActiveInferenceModel
    w = cases.weights_for("w0")
    gpu = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(w)
    ora = O.OracleModel(w, seed=41)
    s = torch.from_numpy(np.random.default_rng(4).normal(size=(5, 10)).astype(np.float32))
    gpu.set_rng(41, 0)
    Gg, pg, qg = gpu.mcts_step_simulate_batch(s, 4)
    Go, po, qo = ora.mcts_step_simulate_batch(s, 4)
    assert torch.equal(pg.cpu(), po)
    assert torch.allclose(qg.cpu(), qo, rtol=1e-4, atol=1e-6)
    assert torch.allclose(Gg.cpu(), Go, rtol=1e-4)
    # K = 1 is the reference's call
    gpu.set_rng(41, 7); ora.set_rng(41, 7)
    G1, p1, q1 = gpu.mcts_step_simulate(s[2], 4)
    gpu.set_rng(41, 7)
    Gb, pb, qb = gpu.mcts_step_simulate_batch(s[2:3], 4)
    Go1, po1, qo1 = ora.mcts_step_simulate(s[2], 4)
    assert abs(G1 - Go1) <= 1e-4 * abs(Go1) and torch.equal(p1.cpu(), po1)
    assert G1 == float(Gb[0]) and torch.equal(p1, pb[0]) and torch.equal(q1, qb[0])


@pytest.mark.gpu
@pytest.mark.parametrize("use_means", [True, False])
def test_cuda_model_plans_like_the_oracle_with_batched_leaves(use_means):
    from dai_b200 import mcts as my_mcts
    from dai_b200.torchmodel import ActiveInferenceModel
    w = cases.weights_for("w0")
    gpu = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(w)
    gpu.set_rng(77, 0)
    ora = O.OracleModel(w, seed=77)
    p = _params(my_mcts, 9, use_means, 2.0, depth=3, samples=2)
    pg = my_mcts.active_inference_mcts_batched(gpu, _frame(), p, o_shape=(1, 64, 64), leaves=4)
    po = my_mcts.active_inference_mcts_batched(ora, _frame(), p, o_shape=(1, 64, 64), leaves=4)
    assert pg[0] == po[0] and pg[1] == po[1] and pg[3] == po[3]
    assert np.allclose(pg[4], po[4], rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("use_means,threshold,leaves,repeats", [(True, 2.0, 4, 9), (False, 2.0, 8, 20), (True, 0.5, 2, 40),
                                                                 (False, 2.0, 1, 5)])
def test_device_resident_planner_makes_the_host_planners_decisions(use_means, threshold, leaves, repeats):
    """dai_mcts_plan (tree on the device, one host wait per decision) against the host-driven batched planner on the
    same CUDA model and noise: same path, same expansions, same simulated G."""
    from dai_b200 import mcts as my_mcts
    from dai_b200.torchmodel import ActiveInferenceModel
    gpu = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(cases.weights_for("w0"))
    p = _params(my_mcts, repeats, use_means, threshold, depth=3, samples=2)
    gpu.set_rng(77, 0)
    host = my_mcts.active_inference_mcts_batched(gpu, _frame(), p, o_shape=(1, 64, 64), leaves=leaves)
    calls_host = gpu._engine.get_rng()[1]
    gpu.set_rng(77, 0)
    dev = my_mcts.active_inference_mcts_device(gpu, _frame(), p, o_shape=(1, 64, 64), leaves=leaves)
    assert dev[0] == host[0] and dev[1] == host[1] and dev[2] == host[2]
    assert dev[3] == _ints(host[3])
    assert np.allclose(dev[4], host[4], rtol=0, atol=1e-6)
    assert gpu._engine.get_rng()[1] == calls_host          # also after a stop by threshold: the call index is timing-independent
    if threshold > 1.0:
        assert dev[1] == repeats
    # leaves = 1 is the reference's sequential search
    if leaves == 1:
        gpu.set_rng(77, 0)
        seq = my_mcts.active_inference_mcts(gpu, _frame(), p, o_shape=(1, 64, 64))
        assert dev[0] == seq[0] and dev[3] == _ints(seq[3])


def test_select_batch_on_random_trees_claims_distinct_leaves_and_leaves_no_trace():
    """Property test of the virtual-visit selection on randomly grown trees: k distinct leaves (or all of them),
    every path ends in its leaf, statistics untouched; the first pick is the sequential rule's pick."""
    from dai_b200.mcts import Tree
    rng = np.random.default_rng(0)
    for trial in range(40):
        t = Tree(4, 400, float(rng.uniform(0.2, 2.0)), bool(trial % 2))
        root = t.add(torch.zeros(10))
        leaves = [root]
        for _ in range(int(rng.integers(1, 25))):                  # expand random leaves
            i = leaves.pop(int(rng.integers(0, len(leaves))))
            for a in range(4):
                c = t.add(torch.zeros(10))
                t.child[i, a] = c
                leaves.append(c)
            t.W[i] = torch.from_numpy(-rng.uniform(40, 60, 4).astype(np.float32))
            t.N[i] = torch.from_numpy(rng.integers(1, 6, 4).astype(np.float32))
            t.Qpi[i] = torch.softmax(torch.from_numpy(rng.normal(size=4).astype(np.float32)), 0)
        W0, N0 = t.W.clone(), t.N.clone()
        k = int(rng.integers(1, 12))
        picks = t.select_batch(root, k)
        ids = [nodes[-1] for nodes, _ in picks]
        assert len(ids) == min(k, len(leaves)) and len(set(ids)) == len(ids)
        assert all(t.is_leaf(i) for i in ids)
        for nodes, actions in picks:
            cur = root
            for n, a in zip(nodes, actions):
                assert int(t.child[cur, a]) == n
                cur = n
        assert torch.equal(t.W, W0) and torch.equal(t.N, N0)
        seq_nodes, seq_actions = t.select(root)
        assert picks[0][1] == seq_actions and picks[0][0] == seq_nodes
