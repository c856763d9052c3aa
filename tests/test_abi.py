"""The C-ABI library loads and exports every symbol include/dai_b200.h declares (no compute
calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "dai_b200.h")).read()
    return sorted(set(re.findall(r"DAI_API\s+(?:const\s+char\*|int)\s+(dai_\w+)\s*\(", text)))


def _library():
    import dai_b200  # noqa: F401
    from dai_b200 import engine
    if not os.path.exists(engine.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    return engine


def test_header_declares_the_expected_surface():
    names = _declared()
    assert len(names) >= 20
    for must in ("dai_create", "dai_set_weight", "dai_commit_weights", "dai_calculate_G", "dai_calculate_G_mean",
                 "dai_G_given_trajectory", "dai_rollout", "dai_rollout_host", "dai_mcts_simulate", "dai_encode",
                 "dai_decode", "dai_transition", "dai_habit", "dai_combine"):
        assert must in names


def test_library_exports_every_declared_symbol():
    engine = _library()
    lib = ctypes.CDLL(engine.library_path())
    for name in _declared():
        assert getattr(lib, name) is not None
    assert set(engine.SIGNATURES) == set(_declared())
    lib.dai_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.dai_version()


def test_no_cpu_fallback_without_cuda():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    engine = _library()
    with pytest.raises(engine.DaiError):
        engine.Engine()


def test_create_rejects_unsupported_configs():
    engine = _library()
    lib = engine.load_library()
    h = ctypes.c_void_p()
    cfg = engine.DaiConfig(10, 3, 32, 1, 0, 1)      # the dead 32-px / 3-action branch
    assert lib.dai_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -5
    assert lib.dai_create(None, 0, ctypes.byref(h)) == -1


def test_shard_ranges_partition_the_samples():
    from dai_b200.sharding import shard_range
    for n in (1, 7, 50, 800):
        for w in (1, 2, 3, 8):
            cover = []
            for r in range(w):
                b, e = shard_range(n, r, w)
                cover += list(range(b, e))
            assert cover == list(range(n))


def test_ctypes_signatures_have_the_arity_the_header_declares():
    """Every binding in engine.SIGNATURES passes as many arguments as the C prototype takes (ABI drift shows up here,
    on the CPU, rather than as a crash on the GPU box)."""
    engine = _library()
    text = open(os.path.join(ROOT, "include", "dai_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"DAI_API\s+(?:const\s+char\*|int)\s+(dai_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S))
    assert set(protos) == set(engine.SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert n == len(engine.SIGNATURES[name][1]), (name, n, len(engine.SIGNATURES[name][1]))
    # the two parameter structs mirror the header field for field
    fields = re.search(r"typedef struct dai_mcts_params \{(.*?)\} dai_mcts_params;", text, flags=re.S).group(1)
    names = re.findall(r"(?:float|int32_t)\s+(\w+)\s*;", fields)
    assert names == [f[0] for f in engine.DaiMctsParams._fields_]
    fields = re.search(r"typedef struct dai_mcts_result \{(.*?)\} dai_mcts_result;", text, flags=re.S).group(1)
    assert re.findall(r"int32_t\s+(\w+)\s*;", fields) == [f[0] for f in engine.DaiMctsResult._fields_]
    fields = re.search(r"typedef struct dai_config \{(.*?)\} dai_config;", text, flags=re.S).group(1)
    assert re.findall(r"int32_t\s+(\w+)\s*;", fields) == [f[0] for f in engine.DaiConfig._fields_]
