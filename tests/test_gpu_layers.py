"""tcgen05 contraction layers against the fp32 CUDA-core kernels, layer by layer (test hook
dai_debug_layer), on random non-negative activations like the ones ReLU+dropout produce."""
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from dai_b200.torchmodel import ActiveInferenceModel
    m = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, precision="bf16x3", device="cuda:0")
    m.load_numpy_weights(cases.weights_for("w0"))
    m._sync()
    return m._engine


@pytest.mark.parametrize("rows", [1, 3, 40])
@pytest.mark.parametrize("layer", [1, 2, 3])
def test_tc_layer_matches_fp32(eng, layer, rows):
    hw_in = 1024 if layer == 3 else 256
    g = torch.Generator(device="cuda").manual_seed(layer * 100 + rows)
    x = torch.rand(rows, hw_in, 64, device="cuda", generator=g)
    x = x * (torch.rand(rows, hw_in, 64, device="cuda", generator=g) < 0.5) * 2.0
    ref = eng.debug_layer(layer, "fp32_simt", x)
    if layer == 3:
        # the tensor-core ct3 epilogue emits the last deconv's channel and kw sums:
        # e[kh][oy][ox] = sum_kw <relu(out)[oy][ox+1-kw], w4[:, kh, kw]>
        w4 = torch.from_numpy(cases.weights_for("w0")["po_net.19.weight"]).cuda().reshape(32, 9)
        d = torch.einsum("npc,ct->ntp", ref.double(), w4.double()).reshape(rows, 3, 3, 64, 64)
        e = d[:, :, 1].clone()
        e[..., :-1] += d[:, :, 0][..., 1:]
        e[..., 1:] += d[:, :, 2][..., :-1]
        ref = e.reshape(rows, 3, 4096).float()
    got = eng.debug_layer(layer, "bf16x3", x)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-5 * max(scale, 1.0), "layer %d rows %d: max err %.3e (scale %.3e)" % (layer, rows, err, scale)
    fast = eng.debug_layer(layer, "bf16x1", x)
    ferr = (fast - ref).abs().max().item()
    assert ferr <= 2e-2 * max(scale, 1.0), "layer %d bf16x1: max err %.3e" % (layer, ferr)
