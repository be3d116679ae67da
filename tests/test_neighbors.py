"""
Neighbor list: the per-atom code of ``torch-pme_b200/csrc/neighbors_core.h`` (wrap + bin, search loop:
what each CUDA thread executes) is compiled for the host (tests/native/nl_host.cpp, g++) and driven
through the package's own layout logic, then compared with the brute-force oracle on periodic /
non-periodic, cubic / triclinic, large / smaller-than-cutoff cells, half and full lists, fp64 and fp32.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import pme_oracle as oracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_search(tmp_path_factory):
    """the per-atom code of csrc/neighbors_core.h built for the host; passed to neighbor_list(_host_library=...)"""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("nl") / "nl_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "native", "nl_host.cpp")], check=True)
    return ctypes.CDLL(so)


def _canonical(idx, d, shifts):
    """unordered pairs as sorted rows (i, j, Sx, Sy, Sz) with i < j or (i == j and S > 0), plus distances"""
    idx, shifts = np.asarray(idx, dtype=np.int64), np.asarray(shifts, dtype=np.int64)
    rows = np.concatenate([idx, shifts], axis=1)
    neg = (shifts[:, 0] < 0) | ((shifts[:, 0] == 0) & ((shifts[:, 1] < 0) | ((shifts[:, 1] == 0) & (shifts[:, 2] < 0))))
    flip = (idx[:, 0] > idx[:, 1]) | ((idx[:, 0] == idx[:, 1]) & neg)
    rows[flip] = np.concatenate([idx[flip][:, ::-1], -shifts[flip]], axis=1)
    key = np.lexsort(rows.T[::-1])
    return rows[key], np.asarray(d, dtype=np.float64)[key]


CELLS = {
    "cubic_large": np.eye(3) * 14.0,
    "cubic_small": np.eye(3) * 2.3,                        # smaller than the cutoff: several images
    "triclinic": np.array([[9.0, 0.0, 0.0], [2.5, 8.0, 0.0], [-1.5, 2.0, 7.0]]),
    "flat": np.array([[12.0, 0.0, 0.0], [0.0, 3.1, 0.0], [0.5, 0.2, 6.5]]),
}


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("full", [False, True])
@pytest.mark.parametrize("name", sorted(CELLS))
def test_search_loop_matches_bruteforce_oracle(host_search, name, full, dtype):
    from torchpme_b200.neighbors import neighbor_list

    rng = np.random.default_rng(len(name) + full)
    cell = CELLS[name]
    n = 40 if name != "cubic_small" else 5
    pos = (rng.random((n, 3)) * 1.6 - 0.3) @ cell          # some atoms outside the cell
    cutoff = 3.7
    ref_idx, ref_d, ref_s = oracle.neighbor_list(pos, cell, cutoff, full=full)
    idx, d, s = neighbor_list(torch.tensor(pos, dtype=dtype), torch.tensor(cell, dtype=dtype), cutoff,
                              full_neighbor_list=full, _host_library=host_search)
    assert idx.dtype == torch.int64 and s.dtype == torch.int32 and d.dtype == dtype
    if full:
        # both directions present: compare as multisets of directed pairs
        a = np.concatenate([idx.numpy(), s.numpy()], axis=1)
        b = np.concatenate([ref_idx, ref_s], axis=1)
        ka, kb = np.lexsort(a.T[::-1]), np.lexsort(b.T[::-1])
        assert a.shape == b.shape
        np.testing.assert_array_equal(a[ka], b[kb])
        np.testing.assert_allclose(d.numpy()[ka], ref_d[kb], rtol=2e-5 if dtype == torch.float32 else 1e-12)
    else:
        rows, dist = _canonical(idx.numpy(), d.numpy(), s.numpy())
        ref_rows, ref_dist = _canonical(ref_idx, ref_d, ref_s)
        assert rows.shape == ref_rows.shape
        np.testing.assert_array_equal(rows, ref_rows)
        np.testing.assert_allclose(dist, ref_dist, rtol=2e-5 if dtype == torch.float32 else 1e-12)
        assert len(np.unique(rows, axis=0)) == len(rows)          # every pair exactly once


def test_non_periodic_all_pairs(host_search):
    from torchpme_b200.neighbors import distances_from, neighbor_list

    rng = np.random.default_rng(2)
    pos = torch.tensor(rng.random((9, 3)) * 4.0)
    cell = torch.eye(3, dtype=torch.float64)
    idx, d, s = neighbor_list(pos, cell, 100.0, periodic=(False, False, False), _host_library=host_search)
    assert idx.shape[0] == 9 * 8 // 2 and int(s.abs().sum()) == 0
    assert bool((idx[:, 0] < idx[:, 1]).all())
    np.testing.assert_allclose(d.numpy(), distances_from(pos, cell, idx, s).numpy(), rtol=1e-13)
    # slab geometry: periodic in x and y only
    cell = torch.tensor([[3.0, 0, 0], [0, 3.5, 0], [0, 0, 50.0]], dtype=torch.float64)
    idx, d, s = neighbor_list(pos, cell, 2.5, periodic=(True, True, False), _host_library=host_search)
    assert int(s[:, 2].abs().sum()) == 0 and idx.shape[0] > 0
    np.testing.assert_allclose(d.numpy(), distances_from(pos, cell, idx, s).numpy(), rtol=1e-12)
    assert float(d.max()) < 2.5
    # completeness against a brute-force count over the x / y images
    p, c = pos.numpy(), cell.numpy()
    expected = 0
    for sx in range(-3, 4):
        for sy in range(-3, 4):
            delta = p[None, :, :] + (sx * c[0] + sy * c[1])[None, None, :] - p[:, None, :]
            close = np.linalg.norm(delta, axis=-1) < 2.5
            for i in range(9):
                for j in range(9):
                    if close[i, j] and (i < j or (i == j and (sx, sy) > (0, 0))):
                        expected += 1
    assert idx.shape[0] == expected


def test_search_layout():
    from torchpme_b200.neighbors import search_layout

    assert search_layout(np.eye(3) * 90.24, 6.0) == ([30, 30, 30], [2, 2, 2])
    assert search_layout(np.eye(3) * 2.3, 3.7) == ([1, 1, 1], [2, 2, 2])
    n_bins, reach = search_layout(np.eye(3) * 10.0, 3.0, periodic=(True, False, True))
    assert n_bins == [6, 1, 6] and reach == [2, 0, 2]
    n_bins, _ = search_layout(np.eye(3) * 1e4, 1.0)
    assert n_bins[0] * n_bins[1] * n_bins[2] <= 1 << 21


def test_struct_layout_matches_header():
    import ctypes as c

    from torchpme_b200 import _native

    ns = _native._NeighborSearch
    assert c.sizeof(ns) == 120            # 9 doubles, 10 ints, 1 double (include/torchpme_b200.h)
    assert ns.n_bins.offset == 72 and ns.reach.offset == 84 and ns.periodic.offset == 96
    assert ns.full_list.offset == 108 and ns.cutoff.offset == 112


def test_product_path_refuses_cpu_tensors():
    import torchpme_b200 as tp
    from torchpme_b200.neighbors import neighbor_list

    with pytest.raises(tp.NativeLibraryError, match="CUDA-only"):
        neighbor_list(torch.rand(4, 3), torch.eye(3) * 5, 2.0)
