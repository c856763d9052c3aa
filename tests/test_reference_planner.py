"""the reference's UNMODIFIED planner (src/mcts.py:150-195) drives the CUDA drop-in model.
The reference is imported from /root/reference where mounted, else from the byte-for-byte copy that
__graft_entry__.build() stages under the git-ignored baseline/_ref/ (it travels to the GPU box)."""
import numpy as np
import pytest
import torch

import cases
from oracle import efe_oracle as O
from oracle import reference_model as RM

pytestmark = pytest.mark.skipif(not RM.available(), reason="reference neither mounted nor staged under baseline/_ref")


def _frame(seed=3):
    import dai_b200.synthetic as syn
    return torch.from_numpy(syn.make_frames(1, seed))[0, 0]


def _params(mod, repeats, use_means, threshold, depth):
    p = mod.MCTS_Params()
    p.repeats, p.use_means, p.threshold, p.simulation_depth = repeats, use_means, threshold, depth
    return p


def test_staged_reference_is_the_mounted_reference():
    """The staged copy is byte-identical to the mount (authoring container), and it imports with its shims."""
    import filecmp
    import os
    if os.path.isdir(os.path.join(RM.MOUNT, "src")) and os.path.isdir(os.path.join(RM.STAGED, "src")):
        for name in os.listdir(os.path.join(RM.MOUNT, "src")):
            if name.endswith(".py"):
                assert filecmp.cmp(os.path.join(RM.MOUNT, "src", name), os.path.join(RM.STAGED, "src", name), shallow=False), name
    m = RM.load(cases.weights_for("w0"))
    assert tuple(m.model_down.qs_net[9].weight.shape) == (256, 576) and m.pi_dim == 4


def test_reference_model_equals_oracle_under_torch_rng():
    """The real reference and the oracle port consume torch's generator identically (the pin, re-run wherever the
    reference is available — also on the GPU box from the staged copy)."""
    w = cases.weights_for("w0")
    ref = RM.load(w)
    W = O.to_torch(w)
    o = torch.from_numpy(__import__("dai_b200.synthetic", fromlist=["x"]).make_frames(1, 3)).repeat(4, 1, 1, 1)
    with torch.no_grad():
        torch.manual_seed(11)
        a = ref.calculate_G_4_repeated(o, steps=2, samples=3)
        torch.manual_seed(11)
        b = O.calculate_G_repeated(W, o, None, 2, False, 3, O.TorchStreamNoise(), four=True)
    assert torch.equal(a[0], b[0]) and all(torch.equal(x, y) for x, y in zip(a[1], b[1])) and torch.equal(a[2], b[2])


@pytest.mark.gpu
@pytest.mark.parametrize("use_means,threshold,repeats,depth", [(True, 2.0, 12, 3), (False, 2.0, 8, 4), (True, 0.5, 40, 2)])
def test_unmodified_reference_planner_drives_the_cuda_model(use_means, threshold, repeats, depth):
    """src/mcts.py runs unmodified over the CUDA model; every evaluation it requests is also made by the oracle on the
    same inputs and noise key and must agree (cases.TeacherForced explains why two independent long searches may
    legitimately part at a near-tie).  A short search is additionally required to make the oracle-driven search's
    decisions outright."""
    from dai_b200.torchmodel import ActiveInferenceModel
    _, ref_mcts, _ = RM.modules()
    w = cases.weights_for("w0")
    gpu = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(w)
    gpu.set_rng(77, 0)
    tf = cases.TeacherForced(gpu, O.OracleModel(w, seed=77), 77)
    p = _params(ref_mcts, repeats, use_means, threshold, depth)
    rg = ref_mcts.active_inference_mcts(tf, _frame(), p, o_shape=(1, 64, 64))
    assert rg[1] <= repeats and len(rg[3]) == rg[1] and rg[2] == rg[1] * depth
    assert tf.calls == 3 + 2 * rg[1]
    assert all(0 <= int(a) < 4 for a in rg[0])
    # independent searches, short horizon: same decisions
    gpu.set_rng(78, 0)
    ora = O.OracleModel(w, seed=78)
    q = _params(ref_mcts, 4, use_means, 2.0, depth)
    a = ref_mcts.active_inference_mcts(gpu, _frame(), q, o_shape=(1, 64, 64))
    b = ref_mcts.active_inference_mcts(ora, _frame(), q, o_shape=(1, 64, 64))
    ints = lambda paths: [[int(x) for x in pth] for pth in paths]
    assert [int(x) for x in a[0]] == [int(x) for x in b[0]] and a[1] == b[1] and ints(a[3]) == ints(b[3])
    assert np.allclose(a[4], b[4], rtol=1e-4, atol=0)
    assert gpu._engine.get_rng()[1] == ora.call
