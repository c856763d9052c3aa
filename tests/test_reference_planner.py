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
    from dai_b200.torchmodel import ActiveInferenceModel
    _, ref_mcts, _ = RM.modules()
    w = cases.weights_for("w0")
    gpu = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, device="cuda:0").load_numpy_weights(w)
    gpu.set_rng(77, 0)
    ora = O.OracleModel(w, seed=77)
    p = _params(ref_mcts, repeats, use_means, threshold, depth)
    rg = ref_mcts.active_inference_mcts(gpu, _frame(), p, o_shape=(1, 64, 64))
    ro = ref_mcts.active_inference_mcts(ora, _frame(), p, o_shape=(1, 64, 64))
    ints = lambda paths: [[int(a) for a in pth] for pth in paths]
    assert [int(a) for a in rg[0]] == [int(a) for a in ro[0]]
    assert rg[1] == ro[1] and rg[2] == ro[2]
    assert ints(rg[3]) == ints(ro[3])
    assert np.allclose(rg[4], ro[4], rtol=1e-4, atol=0)
    assert gpu._engine.get_rng()[1] == ora.call
