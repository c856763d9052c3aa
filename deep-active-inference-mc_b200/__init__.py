"""B200-native Monte-Carlo EFE rollout path behind the reference's call surface.

`ActiveInferenceModel` mirrors src/torchmodel.py:149-393 of
zfountas/deep-active-inference-mc; the compute lives in csrc/ (sm_100a CUDA)
behind the C ABI declared in include/dai_b200.h.
"""
from . import synthetic  # noqa: F401


def __getattr__(name):
    # engine / model pull in torch + the CUDA library: load on first use
    if name in ("ActiveInferenceModel", "ModelTop", "ModelMid", "ModelDown"):
        from . import torchmodel
        return getattr(torchmodel, name)
    if name in ("Engine", "load_library", "DaiError"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
