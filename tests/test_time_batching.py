"""The time-batched horizon (latent chain of all steps first, pixel work of a group of steps as one launch
sequence over (step, sample) slots, one CUDA graph per rollout) against the step-by-step schedule of the same library
(env DAI_TBATCH=0, read once per process — hence two subprocesses): every output of calculate_G_repeated /
calculate_G_4_repeated (src/torchmodel.py:329-393) must be BIT-EQUAL, for whole evaluations and for sample shards
(including the shard that does not hold the last sample and decodes it once more), repeated calls (eager, captured,
replayed), several group sizes (a small decoder chunk makes groups of 1, 2 and 3 steps) and eval mode."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import hashlib, json, sys
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import cases
from dai_b200.torchmodel import ActiveInferenceModel
m = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0, precision="bf16x3", device="cuda:0").load_numpy_weights(cases.weights_for("w0"))
m._sync()
eng = m._engine
rng = np.random.default_rng(11)
out = {}
def digest(d):
    h = hashlib.sha256()
    for k in sorted(d):
        if d[k] is not None:
            h.update(d[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()
o2 = torch.from_numpy(rng.random((2, 4096), dtype=np.float32)).cuda()
o5 = torch.from_numpy(rng.random((5, 4096), dtype=np.float32)).cuda()
pi5 = torch.eye(4, device="cuda")[torch.tensor([0, 3, 1, 2, 2])]
for rep in range(3):                      # eager, captured, replayed
    eng.set_rng(5, 0)
    out["four_T6_N7_rep%%d" %% rep] = digest(eng.rollout(o2.repeat_interleave(4, 0), None, 6, 7, four=True))
eng.set_rng(5, 0); out["ragged_T4_N3"] = digest(eng.rollout(o5, pi5, 4, 3))
eng.set_rng(5, 0); out["mean_T5"] = digest(eng.rollout(o2.repeat_interleave(4, 0), None, 5, 4, calc_mean=True, four=True))
eng.set_rng(5, 0); out["mean_notfour_T5"] = digest(eng.rollout(o5, pi5, 5, 4, calc_mean=True))
for j0, j1 in ((0, 3), (3, 7), (7, 7)):   # shards: without the last sample, with it, empty
    eng.set_rng(9, 2)
    out["shard_%%d_%%d" %% (j0, j1)] = digest(eng.rollout(o5, pi5, 7, 7, shard=(j0, j1)))
m.model_down.eval(); m.model_mid.eval(); m._sync()
eng.set_rng(5, 0); out["eval_T3_N2"] = digest(eng.rollout(o5, pi5, 3, 2))
print("DIGESTS " + json.dumps(out))
"""


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DIGESTS ")][-1]
    return json.loads(line[len("DIGESTS "):])


@pytest.mark.parametrize("chunk", ["19200", "96"])
def test_time_batched_rollout_is_bit_equal_to_the_stepwise_schedule(chunk):
    # chunk 96: 2.5 chunks = 240 rows per group -> groups of 1 .. 3 steps at these sizes, several decoder chunks per group
    stepwise = _run({"DAI_TBATCH": "0", "DAI_DEC_CHUNK": chunk})
    batched = _run({"DAI_TBATCH": "1", "DAI_DEC_CHUNK": chunk})
    assert sorted(stepwise) == sorted(batched)
    diff = [k for k in stepwise if stepwise[k] != batched[k]]
    assert diff == []
    # the three repetitions (eager / captured / replayed graph) of one call are the same evaluation
    assert batched["four_T6_N7_rep0"] == batched["four_T6_N7_rep1"] == batched["four_T6_N7_rep2"]
