"""Time one decoder layer in isolation (dai_debug_layer), CUDA events, many rows."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dai_b200
from dai_b200 import synthetic
from dai_b200.torchmodel import ActiveInferenceModel
m = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0).load_numpy_weights(synthetic.make_weights(0)); m._sync()
eng = m._engine
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
names = {1: "ct1", 2: "ct2", 3: "ct3"}
for layer in (1, 2, 3):
    hw = 1024 if layer == 3 else 256
    x = torch.rand(rows, hw, 64, device="cuda")
    for prec in ("bf16x3", "bf16x1"):
        for _ in range(3): eng.debug_layer(layer, prec, x)
        eng.profile_begin()
        for _ in range(20): eng.debug_layer(layer, prec, x)
        ms, n, r = eng.profile_end()[names[layer]]
        print("layer %d %s rows %d: %.1f us/launch (kernel only)" % (layer, prec, rows, ms * 1000 / n))
