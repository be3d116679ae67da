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


def test_kernel_bindings_refuse_cpu_tensors():
    """
    Device dispatch happens in the Python classes (CPU tensors -> torchpme_b200._cpu, the torch
    formulation); the kernel bindings themselves never compute anything for a CPU tensor and never
    fall back: they raise.
    """
    import torch

    from torchpme_b200 import _native

    with pytest.raises(_native.NativeLibraryError, match="CUDA-only"):
        _native.spread(torch.rand(2, 3), torch.ones(2, 1), [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0], (8, 8, 8), 4, 0)
    with pytest.raises(_native.NativeLibraryError, match="CUDA-only"):
        _native.pair_forward(torch.ones(2, 1), torch.tensor([[0, 1]]), torch.tensor([0.5]), None, None, False,
                             _native.make_pair_potential(1, 0.5))


def test_slab_and_peer_entry_points_validate_arguments_without_a_gpu():
    """host-side argument checks of the multi-GPU entry points run before any CUDA call"""
    from torchpme_b200 import _native

    lib = _native.load()
    err = lambda: lib.tpme_last_error().decode()  # noqa: E731
    r2u = (ctypes.c_double * 9)()
    # slab outside the mesh
    assert lib.tpme_spread_slab(0, None, None, 1, 1, r2u, 8, 8, 8, 6, 4, None, None, 4, 0, None, 0, None) != 0
    assert "slab" in err()
    # a point list without its length
    dummy = ctypes.c_void_p(16)
    assert lib.tpme_gather_slab(0, dummy, dummy, 1, 1, r2u, 8, 8, 8, 0, 8, dummy, None, 4, 0, dummy, None, None,
                                None) != 0
    assert "go together" in err()
    # non-power-of-two meshes are refused by the slab FFT pieces
    assert lib.tpme_slab_fft_yz(0, 1, dummy, dummy, 4, 12, 16, None) != 0 and "power-of-two" in err()
    green = _native.make_green(1, 1.0, [0.0] * 9)
    assert lib.tpme_slab_fft_x_green(0, dummy, 1, 16, 16, 16, 12, 8, ctypes.byref(green), None) != 0
    assert "y slab" in err()
    # exchange copy: element size and rank count
    ptrs = (ctypes.c_void_p * 1)(16)
    assert lib.tpme_slab_exchange_copy(4, dummy, ptrs, 1, 1, 1, 8, 0, 0, 0, 0, 0, None) != 0 and "complex" in err()
    assert lib.tpme_slab_exchange_copy(8, dummy, ptrs, 1, 99, 1, 8, 0, 0, 0, 0, 0, None) != 0 and "destinations" in err()
    # peers: world size must divide the mesh, buffers must be present
    peers = _native.make_slab_peers(0, [16, 32, 48], [16, 32, 48])
    assert lib.tpme_slab_fft_yz_push(0, dummy, 1, 16, 16, 16, ctypes.byref(peers), None) != 0 and "divide" in err()
    peers = _native.make_slab_peers(0, [16, 0], [16, 32])
    assert lib.tpme_slab_fft_yz_push(0, dummy, 1, 16, 16, 16, ctypes.byref(peers), None) != 0 and "null peer" in err()
    # all-reduce: alignment and rank layout
    a = (ctypes.c_void_p * 2)(16, 40)
    assert lib.tpme_peer_allreduce(0, a, a, 2, 0, 64, None) != 0 and "aligned" in err()
    assert lib.tpme_peer_allreduce(0, a, a, 2, 5, 64, None) != 0 and "rank layout" in err()
    assert lib.tpme_peer_barrier(a, 0, 0, dummy, 1.0, dummy, None) != 0 and "rank layout" in err()
    # neighbor search parameters
    search = _native._NeighborSearch()
    assert lib.tpme_nl_sort(0, dummy, 4, ctypes.byref(search), dummy, dummy, dummy, dummy, None) != 0
    assert "cutoff" in err()
    search.cutoff = 1.0
    for a in range(3):
        search.n_bins[a] = 1
    assert lib.tpme_nl_pairs(0, dummy, dummy, dummy, 4, ctypes.byref(search), 8, 0, dummy, dummy, dummy, dummy, None) != 0
    assert "singular" in err()
