"""
The C ABI: the shared library loads on a CPU-only box and exports every symbol that
include/torchpme_b200.h declares (no compute calls are made here).
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "torchpme_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tpme_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared_symbols()
    for must in ("tpme_spread", "tpme_gather", "tpme_gather_vjp", "tpme_kfilter_apply", "tpme_green_multiply",
                 "tpme_pair_forward", "tpme_pair_backward", "tpme_fft_plan_create", "tpme_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from torchpme_b200 import _native

    lib = ctypes.CDLL(_native.library_path())
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.tpme_abi_version() == 1


def test_binding_table_matches_header():
    from torchpme_b200 import _native

    assert sorted(_native.SIGNATURES) == _declared_symbols()
    _native.load()


def test_argument_errors_are_reported_without_a_gpu():
    """Host-side argument validation of the ABI runs before any CUDA call."""
    from torchpme_b200 import _native

    lib = _native.load()
    green = _native.make_green(2, 1.0, [0.0] * 9, exponent=9)
    rc = lib.tpme_green_multiply(0, None, 1, 4, 4, 4, ctypes.byref(green), None, None)
    assert rc != 0
    assert b"Unsupported exponent" in lib.tpme_last_error()
    rc = lib.tpme_spread(7, None, None, 1, 1, (ctypes.c_double * 9)(), 4, 4, 4, 4, 0, None, 0, None)
    assert rc != 0 and b"dtype" in lib.tpme_last_error()


def test_cpu_tensors_fail_loudly():
    """There is no CPU fallback: CPU tensors raise instead of silently computing elsewhere."""
    import torch

    import torchpme_b200 as tp

    calc = tp.PMECalculator(tp.CoulombPotential(smearing=0.5), mesh_spacing=0.25)
    with pytest.raises(tp.NativeLibraryError, match="CUDA-only"):
        calc(torch.ones(2, 1), torch.eye(3), torch.rand(2, 3), torch.tensor([[0, 1]]), torch.tensor([0.5]))
