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


def _row_planes(ct3_out, rows):
    """(rows,4096,32) relu'd ct3 output -> the last deconv's channel and kw sums e[kh] (rows,3,4096)."""
    w4 = torch.from_numpy(cases.weights_for("w0")["po_net.19.weight"]).cuda().reshape(32, 9)
    d = torch.einsum("npc,ct->ntp", ct3_out.double(), w4.double()).reshape(rows, 3, 3, 64, 64)
    e = d[:, :, 1].clone()
    e[..., :-1] += d[:, :, 0][..., 1:]
    e[..., 1:] += d[:, :, 2][..., :-1]
    return e.reshape(rows, 3, 4096).float()


@pytest.mark.parametrize("rows", [1, 2, 3, 40, 333])
def test_fused_ct2_ct3_pair_kernel_matches_fp32(eng, rows):
    """k_tc_ct23 (CTA pairs, cta_group::2, act2 through the L2-resident scratch) against ct2 -> ct3 on the fp32
    CUDA-core kernels: 1 row (rank 1 idle), odd counts, one image pair per CTA pair, and 333 rows = up to three
    software-pipelined stages per pair with a partner-less last image."""
    g = torch.Generator(device="cuda").manual_seed(2300 + rows)
    x = torch.rand(rows, 256, 64, device="cuda", generator=g)
    x = x * (torch.rand(rows, 256, 64, device="cuda", generator=g) < 0.5) * 2.0
    mid = eng.debug_layer(2, "fp32_simt", x)
    ref = _row_planes(eng.debug_layer(3, "fp32_simt", mid), rows)
    got = eng.debug_layer(23, "bf16x3", x)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= 4e-5 * max(scale, 1.0), "rows %d: max err %.3e (scale %.3e)" % (rows, err, scale)
    # and the same numbers as the two separate tensor-core kernels up to fp32 accumulation order (same products; the pair
    # kernel issues them plane-major)
    sep = eng.debug_layer(3, "bf16x3", eng.debug_layer(2, "bf16x3", x))
    assert (got - sep).abs().max().item() <= 1e-5 * max(scale, 1.0), "rows %d: fused differs from ct2 then ct3 by %.3e" % (rows, (got - sep).abs().max().item())
    fast = eng.debug_layer(23, "bf16x1", x)
    assert (fast - ref).abs().max().item() <= 3e-2 * max(scale, 1.0)
