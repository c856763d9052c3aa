"""mcts_step_simulate's habit-policy rollout (src/torchmodel.py:354-388) on a cluster of 16 CTAs per starting state
(k_sim_rollout_cluster: the two 512 x 512 transition layers resident in shared memory, column slices exchanged through
distributed shared memory) against the one-CTA kernel of the same library (env DAI_SIM_CLUSTER=0, read once per
process — hence two subprocesses): every output BIT-EQUAL (the FMA chains keep their order), for one and several
starting states, sampled and mean trajectories, train and eval mode."""
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
rng = np.random.default_rng(23)
out = {}
def digest(ts):
    h = hashlib.sha256()
    for t in ts:
        h.update(np.ascontiguousarray(torch.as_tensor(t).detach().cpu().numpy()).tobytes())
    return h.hexdigest()
s1 = torch.from_numpy(rng.standard_normal((1, 10)).astype(np.float32))
s19 = torch.from_numpy(rng.standard_normal((19, 10)).astype(np.float32))       # 19 clusters: more than one wave of 16-CTA clusters
for name, s, depth, means in (("k1", s1, 10, False), ("k1_means", s1, 7, True), ("k19", s19, 10, False), ("k19_d3", s19, 3, True)):
    m.set_rng(31, 4)
    out[name] = digest(m.mcts_step_simulate_batch(s, depth, use_means=means))
m.model_down.eval(); m.model_mid.eval(); m.model_top.eval()
m.set_rng(31, 4)
out["eval"] = digest(m.mcts_step_simulate_batch(s19, 5, use_means=False))
print("DIGESTS " + json.dumps(out))
"""


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DIGESTS ")][-1]
    return json.loads(line[len("DIGESTS "):])


def test_cluster_rollout_is_bit_equal_to_the_one_cta_kernel():
    one = _run({"DAI_SIM_CLUSTER": "0"})
    cluster = _run({"DAI_SIM_CLUSTER": "1"})
    assert sorted(one) == sorted(cluster)
    assert [k for k in one if one[k] != cluster[k]] == []
