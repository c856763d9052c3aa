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
