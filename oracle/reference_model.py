"""TEST INFRASTRUCTURE — locate and load the UNMODIFIED reference (zfountas/deep-active-inference-mc).

The reference is pure Python with no build system, so "building oracle/_ref" means making its `src/`
package importable where /root/reference does not exist (the GPU box): `stage()` copies the reference's
`src/*.py` byte for byte into the git-ignored `baseline/_ref/` (it travels with the gpurun snapshot; it is
never committed).  `load()` imports `src.torchmodel` / `src.mcts` from the first location that has them and
applies the two runtime shims of SURVEY.md §0.1 (D1: encoder FC1 576->256, D2: `precision`) to the
INSTANCE — no source edits.

Only tests/, bench.py's `--impl reference` / `cpu_baseline` legs and __graft_entry__.build() (staging)
use this module; the product package never imports it.
"""
import filecmp
import importlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOUNT = "/root/reference"
STAGED = os.path.join(ROOT, "baseline", "_ref")


def stage(force=False):
    """Copy /root/reference/src (and LICENSE) to baseline/_ref when the mount exists.  Returns the staged
    path or None."""
    src = os.path.join(MOUNT, "src")
    if not os.path.isdir(src):
        return STAGED if os.path.isdir(os.path.join(STAGED, "src")) else None
    dst = os.path.join(STAGED, "src")
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        a, b = os.path.join(src, name), os.path.join(dst, name)
        if force or not os.path.exists(b) or not filecmp.cmp(a, b, shallow=False):
            shutil.copyfile(a, b)
            os.chmod(b, 0o644)
    lic = os.path.join(MOUNT, "LICENSE")
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(STAGED, "LICENSE"))
        os.chmod(os.path.join(STAGED, "LICENSE"), 0o644)
    return STAGED


def location():
    """Directory holding the reference's `src/` package: $DAI_REFERENCE_DIR, the mount, or the staged copy."""
    for d in (os.environ.get("DAI_REFERENCE_DIR"), MOUNT, STAGED):
        if d and os.path.isfile(os.path.join(d, "src", "torchmodel.py")):
            return d
    return None


def available():
    return location() is not None


def modules():
    """(src.torchmodel, src.mcts, src.util) of the reference, imported unmodified."""
    d = location()
    if d is None:
        raise ImportError("reference not available (neither %s nor %s)" % (MOUNT, STAGED))
    sys.dont_write_bytecode = True
    if d not in sys.path:
        sys.path.insert(0, d)
    return (importlib.import_module("src.torchmodel"), importlib.import_module("src.mcts"),
            importlib.import_module("src.util"))


def load(weights):
    """The reference's ActiveInferenceModel on CPU with SHIM-1 / SHIM-2 and the given state_dict arrays."""
    import torch
    tm, _, _ = modules()
    # the reference picks "cuda" whenever torch sees a GPU (src/torchmodel.py:151); its CPU path is what is wanted here
    # (SURVEY.md §0 fact 5), so the probe answers "no GPU" while the instance is constructed
    saved = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        m = tm.ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0)
    finally:
        torch.cuda.is_available = saved
    assert m.device.type == "cpu"
    m.model_down.qs_net[9] = torch.nn.Linear(576, 256)      # SHIM-1 (D1)
    m.precision = torch.float32                             # SHIM-2 (D2)
    for mod in (m.model_top, m.model_mid, m.model_down):
        mod.load_state_dict({k: torch.as_tensor(weights[k]) for k in mod.state_dict()})
    return m
