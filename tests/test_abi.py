"""CPU: the C-ABI library builds, loads, and exports every symbol `include/qa_b200.h` declares, with
struct layouts that agree with the ctypes mirror.  No compute call is made (no GPU here)."""
import ctypes
import os
import re

from qa_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "qa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qa_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    lib = _abi.load()
    declared = _declared_symbols()
    assert "qa_post_physics_bbc" in declared and "qa_gae" in declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in qa_b200.h but not exported"
        assert name in _abi.SYMBOLS, f"{name} has no ctypes binding"
    assert sorted(_abi.SYMBOLS) == declared


def test_abi_version_and_layout_handshake():
    lib = _abi.load()
    assert lib.qa_version() == _abi.QA_ABI_VERSION
    assert b"sm_100a" in lib.qa_build_info()
    for which, st in enumerate(_abi.STRUCT_ORDER):
        assert lib.qa_struct_size(which) == ctypes.sizeof(st), st.__name__
    assert lib.qa_struct_size(999) == -1


def test_argument_validation_without_gpu():
    """Null / out-of-range arguments are rejected before any CUDA call."""
    lib = _abi.load()
    assert lib.qa_gae(None, None) == -1
    g = _abi.QaGaeArgs()
    assert lib.qa_gae(ctypes.byref(g), None) == -1
    a = _abi.QaActionPushArgs(num_envs=4, delay=9, clip=1.0, actions_in=8, action_history_buf=8, actions_out=8)
    assert lib.qa_action_push(ctypes.byref(a), None) == -2
    c, s = _abi.QaBbcConst(), _abi.QaBbcStepArgs()
    c.num_bodies = 40
    assert lib.qa_post_physics_bbc(ctypes.byref(c), ctypes.byref(s), None) == -2


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from qa_b200 import ops
    x = torch.zeros(4, 12)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.action_push(x, torch.zeros(4, 8, 12), x.clone(), 0, 400.0)
