"""CPU: the C-ABI library loads and exports every symbol include/bode_b200.h declares; compute calls fail
loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bode_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bode_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import bayesian_ode_b200 as bode
    lib = bode._lib.load()
    names = _declared_symbols()
    assert len(names) >= 7
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in bode._lib.SYMBOLS, f"{n} has no ctypes prototype"
    assert lib.bode_version() >= 100


def test_struct_layout_matches_header():
    import bayesian_ode_b200 as bode
    # 4 int32 + 66 doubles + 4 pointers + int64 + AT pointer; 3 x 4 bytes (+pad) + 4 pointers
    assert ctypes.sizeof(bode._lib.NpdeFieldStruct) == 16 + 66 * 8 + 6 * 8
    assert ctypes.sizeof(bode._lib.GridStruct) == 16 + 4 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    import bayesian_ode_b200 as bode
    with pytest.raises(bode._lib.BodeError):
        bode.NPDEField(torch.zeros(25, 2), torch.zeros(25, 2), 1.0, 0.75, 0.1)
    assert bode._lib.load().bode_device_sm_count() < 0        # CUDA error surfaces as a status, not a crash
    assert b"CUDA" in bode._lib.load().bode_last_error()
