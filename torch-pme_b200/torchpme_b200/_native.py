"""
ctypes binding of ``libtorchpme_b200.so`` (C ABI declared in ``include/torchpme_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call fails the
error is raised to the caller.  PyTorch is used for device memory and streams only; every
function here takes torch tensors, checks that they are contiguous CUDA tensors and hands
raw device pointers to the library on torch's current stream.
"""

from __future__ import annotations

import ctypes
import os
import threading
import weakref

import numpy as np
import torch

_LIB_NAME = "libtorchpme_b200.so"
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)

P3M, LAGRANGE = 0, 1
METHOD_ID = {"P3M": P3M, "Lagrange": LAGRANGE}
GREEN_TABLE, GREEN_COULOMB, GREEN_IPL = 0, 1, 2

#: TPME_DEBUG_BOUNDS=1: validate the pair list against the number of atoms before every pair sum (host sync)
DEBUG_BOUNDS = os.environ.get("TPME_DEBUG_BOUNDS", "0") == "1"

# number of kernel launches issued through this binding (bench.py reports it)
launch_counter = 0


class NativeLibraryError(RuntimeError):
    pass


class _Green(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int),
        ("exponent", ctypes.c_int),
        ("p3m_nodes", ctypes.c_int),
        ("p3m_mode", ctypes.c_int),
        ("smearing", ctypes.c_double),
        ("prefactor", ctypes.c_double),
        ("scale", ctypes.c_double),
        ("recip", ctypes.c_double * 9),
        ("spacing", ctypes.c_double * 3),
        ("table", ctypes.c_void_p),
    ]


class _PairPotential(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int),
        ("exponent", ctypes.c_int),
        ("exclusion_degree", ctypes.c_int),
        ("reserved", ctypes.c_int),
        ("smearing", ctypes.c_double),
        ("prefactor", ctypes.c_double),
        ("exclusion_radius", ctypes.c_double),
    ]


class _PointEpilogue(ctypes.Structure):
    _fields_ = [
        ("add_coef", ctypes.c_void_p),
        ("dc", ctypes.c_void_p),
        ("scale", ctypes.c_double),
        ("self_half", ctypes.c_double),
        ("background", ctypes.c_double),
        ("coef2", ctypes.c_void_p),
        ("dvalues2", ctypes.c_void_p),
        ("vjp_scale", ctypes.c_double),
    ]


class _NeighborSearch(ctypes.Structure):
    _fields_ = [
        ("cell", ctypes.c_double * 9),
        ("n_bins", ctypes.c_int * 3),
        ("reach", ctypes.c_int * 3),
        ("periodic", ctypes.c_int * 3),
        ("full_list", ctypes.c_int),
        ("cutoff", ctypes.c_double),
    ]


class _TilePlan(ctypes.Structure):
    _fields_ = [(name, ctypes.c_int) for name in (
        "nx", "ny", "nz", "nodes", "tx", "ty", "zw", "npx", "npy", "nzc", "nzt", "gather_nzt", "n_bins", "row_stride",
        "plane_stride", "smem_bytes", "spread_threads", "gather_threads", "spread_batch", "spread_tiled")]


class _SlabPeers(ctypes.Structure):
    _fields_ = [
        ("n_ranks", ctypes.c_int),
        ("rank", ctypes.c_int),
        ("hat", ctypes.c_void_p * 16),
        ("hat_t", ctypes.c_void_p * 16),
    ]


_vp, _i, _i64, _dp = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_double)

#: every symbol the header declares, with its argument types (used by the loader and by
#: tests/test_abi.py to check the export table)
SIGNATURES = {
    "tpme_abi_version": ([], _i),
    "tpme_last_error": ([], ctypes.c_char_p),
    "tpme_spread": ([_i, _vp, _vp, _i64, _i, _dp, _i, _i, _i, _i, _i, _vp, _i, _vp], _i),
    "tpme_gather": ([_i, _vp, _vp, _i64, _i, _dp, _i, _i, _i, _i, _i, _vp, _vp,
                     ctypes.POINTER(_PointEpilogue), _vp], _i),
    "tpme_gather_vjp": ([_i, _vp, _vp, _vp, _i64, _i, _dp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp,
                         ctypes.POINTER(_PointEpilogue), _vp], _i),
    "tpme_spread_slab": ([_i, _vp, _vp, _i64, _i, _dp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _vp], _i),
    "tpme_gather_slab": ([_i, _vp, _vp, _i64, _i, _dp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp,
                          ctypes.POINTER(_PointEpilogue), _vp], _i),
    "tpme_gather_vjp_slab": ([_i, _vp, _vp, _vp, _i64, _i, _dp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp,
                              _i, _vp, ctypes.POINTER(_PointEpilogue), _vp], _i),
    "tpme_slab_select_points": ([_i, _vp, _i64, _dp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "tpme_tile_plan_make": ([_i, _i, _i, _i, _i, _i, _i64, ctypes.POINTER(_TilePlan)], _i),
    "tpme_tile_bin_count_ints": ([ctypes.POINTER(_TilePlan)], _i64),
    "tpme_tile_sort": ([_i, ctypes.POINTER(_TilePlan), _vp, _i64, _dp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "tpme_tile_spread": ([_i, ctypes.POINTER(_TilePlan), _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _i, _vp], _i),
    "tpme_tile_gather": ([_i, ctypes.POINTER(_TilePlan), _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _dp, _i, _vp, _vp,
                          _vp, _i, _vp, ctypes.POINTER(_PointEpilogue), _vp], _i),
    "tpme_slab_fft_yz": ([_i, _i, _vp, _vp, _i, _i, _i, _vp], _i),
    "tpme_slab_fft_x_green": ([_i, _vp, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_Green), _vp], _i),
    "tpme_slab_exchange_copy": ([_i, _vp, ctypes.POINTER(_vp), _i, _i, _i, _i64, _i64, _i64, _i64, _i64,
                                 _i64, _vp], _i),
    "tpme_slab_fft_yz_push": ([_i, _vp, _i, _i, _i, _i, ctypes.POINTER(_SlabPeers), _vp], _i),
    "tpme_slab_fft_x_green_push": ([_i, _i, _i, _i, _i, ctypes.POINTER(_Green), ctypes.POINTER(_SlabPeers), _vp], _i),
    "tpme_peer_buffer_create": ([_i64, ctypes.POINTER(_vp), ctypes.c_char_p], _i),
    "tpme_peer_buffer_open": ([ctypes.c_char_p, ctypes.POINTER(_vp)], _i),
    "tpme_peer_buffer_close": ([_vp], _i),
    "tpme_peer_buffer_destroy": ([_vp], _i),
    "tpme_peer_barrier": ([ctypes.POINTER(_vp), _i, _i, _vp, ctypes.c_double, _vp, _vp], _i),
    "tpme_nl_scratch_ints": ([_i64, ctypes.POINTER(_NeighborSearch)], _i64),
    "tpme_nl_sort": ([_i, _vp, _i64, ctypes.POINTER(_NeighborSearch), _vp, _vp, _vp, _vp, _vp], _i),
    "tpme_nl_pairs": ([_i, _vp, _vp, _vp, _i64, ctypes.POINTER(_NeighborSearch), _i64, _i, _vp, _vp, _vp, _vp,
                       _vp], _i),
    "tpme_multimem_allreduce": ([_i, _vp, _i, _i, _i64, _vp], _i),
    "tpme_peer_allreduce": ([_i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i, _i64, _vp], _i),
    "tpme_fft_plan_create": ([ctypes.POINTER(_vp), _i, _i, _i, _i, _i], _i),
    "tpme_fft_plan_destroy": ([_vp], _i),
    "tpme_fft_plan_uses_own_fft": ([_vp], _i),
    "tpme_rfft3": ([_vp, _vp, _vp, _vp], _i),
    "tpme_irfft3": ([_vp, _vp, _vp, _vp], _i),
    "tpme_green_multiply": ([_i, _vp, _i, _i, _i, _i, ctypes.POINTER(_Green), _vp, _vp], _i),
    "tpme_kfilter_apply": ([_vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_Green), _vp, _vp], _i),
    "tpme_green_table": ([_i, _vp, _i, _i, _i, ctypes.POINTER(_Green), _vp], _i),
    "tpme_green_table_vjp": ([_i, _vp, _vp, _i, _i, _i, _i, ctypes.c_double, _vp, _vp], _i),
    "tpme_pair_forward": ([_i, _vp, _vp, _i, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i,
                           ctypes.POINTER(_PairPotential), _vp, _vp], _i),
    "tpme_pair_backward": ([_i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i,
                            ctypes.POINTER(_PairPotential), _vp, _vp, _vp], _i),
    "tpme_pair_distances": ([_i, _vp, _dp, _vp, _i, _vp, _i64, _vp, _vp, _vp], _i),
    "tpme_pair_distances_backward": ([_i, _vp, _dp, _vp, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp], _i),
}

_lib = None
_lib_lock = threading.Lock()


def library_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises NativeLibraryError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise NativeLibraryError(
                f"{_LIB_NAME} not found at {_LIB_PATH}. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C torch-pme_b200/csrc`. There is no CPU / PyTorch fallback."
            )
        lib = ctypes.CDLL(_LIB_PATH, mode=ctypes.RTLD_LOCAL)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        if lib.tpme_abi_version() != 1:
            raise NativeLibraryError("ABI version mismatch between header and library")
        _lib = lib
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().tpme_last_error().decode(errors="replace")
        raise NativeLibraryError(f"{what} failed (code {rc}): {msg}")


def _dtype_id(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.float64:
        return 1
    raise TypeError(f"torchpme_b200 kernels support float32/float64, got {t.dtype}")


def _dev(t: torch.Tensor | None, name: str):
    """device pointer of a contiguous CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeLibraryError(
            f"`{name}` lives on {t.device}; torchpme_b200 is a CUDA-only implementation "
            "(no CPU fallback). Move the inputs to a CUDA device."
        )
    if not t.is_contiguous():
        raise NativeLibraryError(f"`{name}` must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def _on(t: torch.Tensor, name: str):
    """device guard for the launch; refuses non-CUDA tensors loudly (no CPU fallback)"""
    if not t.is_cuda:
        raise NativeLibraryError(
            f"`{name}` lives on {t.device}; torchpme_b200 is a CUDA-only implementation "
            "(no CPU fallback). Move the inputs to a CUDA device."
        )
    return torch.cuda.device(t.device)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _mat9(values) -> ctypes.Array:
    return (ctypes.c_double * 9)(*[float(v) for v in values])


def _count(n=1):
    global launch_counter
    launch_counter += n


# --------------------------------------------------------------------------------------
# mesh interpolation
# --------------------------------------------------------------------------------------
def _list_args(point_list):
    if point_list is None:
        return None, None
    lst, count = point_list
    return _dev(lst, "point_list"), _dev(count, "list_count")


def slab_select_points(positions, r2u, ns, nodes: int, slab):
    """(list int32 (N,), count int32 (1,)) of the points whose stencil reaches into `slab` = (x0, nx_local)"""
    lib = load()
    n = positions.shape[0]
    nx, ny, nz = ns
    lst = torch.empty(max(n, 1), dtype=torch.int32, device=positions.device)
    count = torch.empty(1, dtype=torch.int32, device=positions.device)
    with _on(positions, "positions"):
        _check(lib.tpme_slab_select_points(_dtype_id(positions), _dev(positions, "positions"), n, _mat9(r2u),
                                           nx, ny, nz, slab[0], slab[1], nodes, _dev(lst, "list"),
                                           _dev(count, "count"), _stream()), "tpme_slab_select_points")
    _count()
    return lst, count


# --------------------------------------------------------------------------------------
# tiled mesh interpolation (cell-sorted atoms, shared-memory pencils, TMA bulk copies)
# --------------------------------------------------------------------------------------
#: "auto": use the tiled kernels whenever the mesh / stencil is covered and the system is large
#: enough to pay for the sort; "off": always the direct kernels; "on": tiled whenever covered
TILE_MODE = os.environ.get("TPME_TILES", "auto")
TILE_MIN_POINTS = int(os.environ.get("TPME_TILE_MIN_POINTS", "65536"))
#: "auto" | "on" | "off": spread through the shared-memory pencils or the direct kernel (part of the tile
#: plan: the library reads TPME_TILE_SPREAD itself; this copy only keys the plan cache)
TILE_SPREAD = os.environ.get("TPME_TILE_SPREAD", "auto")


class TileSort:
    """
    The points of one spread / gather sequence binned by mesh tile (`tpme_tile_sort`): per-point
    records in bin order, the permutation and the bin starts.  Built once per set of positions and
    shared by the forward and backward launches of a step.
    """

    __slots__ = ("plan", "bin_start", "rec", "idx", "n_points", "r2u", "dtype", "positions", "count",
                 "spread_tiled")

    def __init__(self, plan, positions, r2u):
        lib = load()
        n = positions.shape[0]
        dev = positions.device
        self.plan, self.n_points, self.r2u, self.dtype = plan, n, r2u, positions.dtype
        self.positions = positions
        counts = torch.empty(int(lib.tpme_tile_bin_count_ints(ctypes.byref(plan))), dtype=torch.int32, device=dev)
        self.bin_start = torch.empty(plan.n_bins + 1, dtype=torch.int32, device=dev)
        key_rank = torch.empty((max(n, 1), 2), dtype=torch.int32, device=dev)
        self.rec = torch.empty((max(n, 1), 4), dtype=positions.dtype, device=dev)
        self.idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        # where the spread goes is part of the plan (tpme_tile_plan_make: tiled in fp64 and for dense
        # meshes, the direct kernel otherwise; TPME_TILE_SPREAD = on | off overrides)
        self.spread_tiled = bool(plan.spread_tiled)
        self.count = None
        with _on(positions, "positions"):
            _check(lib.tpme_tile_sort(_dtype_id(positions), ctypes.byref(plan), _dev(positions, "positions"), n,
                                      _mat9(r2u), _dev(counts, "bin_count"),
                                      _dev(self.bin_start, "bin_start"), _dev(key_rank, "key_rank"),
                                      _dev(self.rec, "sorted_rec"), _dev(self.idx, "sorted_idx"), _stream()),
                   "tpme_tile_sort")
        _count(3)


_tile_plans: dict = {}


def tile_plan(dtype, ns, nodes: int, method: int, n_points: int):
    """tiling of the mesh for the tiled kernels, or None when they do not cover this case"""
    if TILE_MODE == "off" or n_points <= 0 or (TILE_MODE != "on" and n_points < TILE_MIN_POINTS):
        return None
    # the plan depends on the point count only through the block sizes: bucket it
    key = (dtype, tuple(int(v) for v in ns), int(nodes), int(method), int(n_points).bit_length(), TILE_SPREAD)
    if key in _tile_plans:
        return _tile_plans[key]
    lib = load()
    plan = _TilePlan()
    if TILE_SPREAD in ("on", "off"):
        os.environ["TPME_TILE_SPREAD"] = TILE_SPREAD      # the library reads the switch when it makes the plan
    else:
        os.environ.pop("TPME_TILE_SPREAD", None)
    rc = lib.tpme_tile_plan_make(0 if dtype == torch.float32 else 1, int(ns[0]), int(ns[1]), int(ns[2]),
                                 int(nodes), int(method), int(n_points), ctypes.byref(plan))
    if rc not in (0, 3):
        _check(rc, "tpme_tile_plan_make")
    _tile_plans[key] = plan if rc == 0 else None
    return _tile_plans[key]


def tile_sort(positions, r2u, ns, nodes: int, method: int):
    """TileSort of `positions`, or None when the direct kernels should be used"""
    if not positions.is_cuda or positions.dtype not in (torch.float32, torch.float64):
        return None
    plan = tile_plan(positions.dtype, ns, nodes, method, positions.shape[0])
    return TileSort(plan, positions, r2u) if plan is not None else None


def tile_spread(tiles: TileSort, weights, method: int, out=None, accumulate=False):
    """`accumulate`: `out` already holds what the spread is added to (e.g. zeros written off the critical path)"""
    lib = load()
    n, c = weights.shape
    plan = tiles.plan
    if out is None:
        out = torch.empty((c, plan.nx, plan.ny, plan.nz), dtype=weights.dtype, device=weights.device)
    with _on(weights, "particle_weights"):
        _check(lib.tpme_tile_spread(_dtype_id(weights), ctypes.byref(plan), _dev(tiles.rec, "sorted_rec"),
                                    _dev(tiles.idx, "sorted_idx"), _dev(tiles.bin_start, "bin_start"),
                                    _dev(weights, "particle_weights"), n, c, method, _dev(out, "mesh"),
                                    int(accumulate and out is not None), _stream()), "tpme_tile_spread")
    _count()
    return out


def tile_gather(tiles: TileSort, mesh, method: int, values=None, dvalues=None, grad_positions=None,
                coef=None, accumulate=False, grad_r2u=None, epilogue=None):
    lib = load()
    c = mesh.shape[0]
    with _on(mesh, "mesh"):
        _check(lib.tpme_tile_gather(_dtype_id(mesh), ctypes.byref(tiles.plan), _dev(mesh, "mesh"),
                                    _dev(tiles.rec, "sorted_rec"), _dev(tiles.idx, "sorted_idx"),
                                    _dev(tiles.bin_start, "bin_start"), _dev(tiles.positions, "positions"),
                                    _dev(coef, "coef"), tiles.n_points, c, _mat9(tiles.r2u), method,
                                    _dev(values, "values"), _dev(dvalues, "dvalues"),
                                    _dev(grad_positions, "grad_positions"), int(accumulate),
                                    _dev(grad_r2u, "grad_r2u"),
                                    ctypes.byref(epilogue) if epilogue is not None else None, _stream()),
               "tpme_tile_gather")
    _count()


def spread(positions, weights, r2u, ns, nodes: int, method: int, out=None, slab=None, point_list=None,
           tiles: TileSort | None = None, accumulate=False):
    """
    `accumulate` (with `out`): add to `out` instead of zero-filling it first;
    `slab` = (x0, nx_local): spread into the local x slab (C, nx_local, ny, nz) only;
    `point_list` = result of :func:`slab_select_points` for that slab;
    `tiles` = :class:`TileSort` of `positions`: use the tiled kernel
    """
    if tiles is not None and slab is None:
        if tiles.spread_tiled:
            return tile_spread(tiles, weights, method, out=out, accumulate=accumulate)
    lib = load()
    n, c = weights.shape
    nx, ny, nz = ns
    x0, nxl = (0, nx) if slab is None else slab
    if out is None:
        out = torch.empty((c, nxl, ny, nz), dtype=positions.dtype, device=positions.device)
    with _on(positions, "positions"):
        _check(lib.tpme_spread_slab(_dtype_id(positions), _dev(positions, "positions"),
                                    _dev(weights, "particle_weights"), n, c, _mat9(r2u), nx, ny, nz,
                                    x0, nxl, *_list_args(point_list), nodes, method, _dev(out, "mesh"),
                                    int(accumulate), _stream()),
               "tpme_spread")
    _count()
    return out


def make_epilogue(add_coef, dc, scale, self_half, background, coef2=None, dvalues2=None,
                  vjp_scale=0.0) -> _PointEpilogue:
    e = _PointEpilogue()
    e.add_coef, e.dc = add_coef.data_ptr(), dc.data_ptr()
    e.scale, e.self_half, e.background = float(scale), float(self_half), float(background)
    e.coef2 = coef2.data_ptr() if coef2 is not None else None
    e.dvalues2 = dvalues2.data_ptr() if dvalues2 is not None else None
    e.vjp_scale = float(vjp_scale)
    e._keepalive = (add_coef, dc, coef2, dvalues2)
    return e


def _slab_of(mesh, slab):
    """(c, nx, ny, nz, x0, nx_local) of a full mesh or of a local slab `slab` = (x0, nx_global)"""
    c, nxl, ny, nz = mesh.shape
    if slab is None:
        return c, nxl, ny, nz, 0, nxl
    x0, nx = slab
    return c, nx, ny, nz, x0, nxl


def gather(mesh, positions, r2u, nodes: int, method: int, want_values=True, want_grad=False,
           values_out=None, epilogue: _PointEpilogue | None = None, slab=None, point_list=None,
           tiles: TileSort | None = None):
    """
    plain gather, or (with `epilogue` and `values_out`) the fused accumulate form.
    `slab` = (x0, nx_global): `mesh` is the local x slab and the results are partial sums.
    """
    lib = load()
    c, nx, ny, nz, x0, nxl = _slab_of(mesh, slab)
    n = positions.shape[0]
    values = values_out
    if values is None and want_values:
        values = torch.empty((n, c), dtype=mesh.dtype, device=mesh.device)
    dvalues = torch.empty((n, c, 3), dtype=mesh.dtype, device=mesh.device) if want_grad else None
    if tiles is not None and slab is None and n > 0:
        tile_gather(tiles, mesh, method, values=values, dvalues=dvalues, epilogue=epilogue)
        return values, dvalues
    with _on(mesh, "mesh"):
        _check(lib.tpme_gather_slab(_dtype_id(mesh), _dev(mesh, "mesh"), _dev(positions, "positions"), n,
                                    c, _mat9(r2u), nx, ny, nz, x0, nxl, *_list_args(point_list), nodes, method,
                                    _dev(values, "values"), _dev(dvalues, "dvalues"),
                                    ctypes.byref(epilogue) if epilogue is not None else None, _stream()),
               "tpme_gather")
    _count()
    return values, dvalues


def gather_vjp(mesh, positions, coef, r2u, nodes: int, method: int, grad_positions=None,
               want_values=False, want_grad_r2u=False, values_out=None,
               epilogue: _PointEpilogue | None = None, slab=None, point_list=None,
               tiles: TileSort | None = None):
    """returns (grad_positions, values | None, grad_r2u (3,3) | None); `slab` as in :func:`gather`"""
    lib = load()
    c, nx, ny, nz, x0, nxl = _slab_of(mesh, slab)
    n = positions.shape[0]
    accumulate = grad_positions is not None
    if grad_positions is None:
        grad_positions = torch.empty((n, 3), dtype=mesh.dtype, device=mesh.device)
    values = values_out
    if values is None and want_values:
        values = torch.empty((n, c), dtype=mesh.dtype, device=mesh.device)
    grad_r2u = torch.zeros((3, 3), dtype=mesh.dtype, device=mesh.device) if want_grad_r2u else None
    if tiles is not None and slab is None and n > 0 and c > 0:
        tile_gather(tiles, mesh, method, values=values, grad_positions=grad_positions, coef=coef,
                    accumulate=accumulate, grad_r2u=grad_r2u, epilogue=epilogue)
        return grad_positions, values, grad_r2u
    with _on(mesh, "mesh"):
        _check(lib.tpme_gather_vjp_slab(_dtype_id(mesh), _dev(mesh, "mesh"), _dev(positions, "positions"),
                                        _dev(coef, "coef"), n, c, _mat9(r2u), nx, ny, nz, x0, nxl,
                                        *_list_args(point_list), nodes, method,
                                        _dev(grad_positions, "grad_positions"),
                                        _dev(values, "values"), int(accumulate), _dev(grad_r2u, "grad_r2u"),
                                        ctypes.byref(epilogue) if epilogue is not None else None,
                                        _stream()), "tpme_gather_vjp")
    _count()
    return grad_positions, values, grad_r2u


# --------------------------------------------------------------------------------------
# reciprocal space
# --------------------------------------------------------------------------------------
class FFTPlan:
    """cuFFT R2C/C2R plan pair for a (batch, nx, ny, nz) real mesh."""

    def __init__(self, dtype: torch.dtype, ns, batch: int, device):
        lib = load()
        self.ns = tuple(int(v) for v in ns)
        self.batch = int(batch)
        self.dtype = dtype
        self.device = torch.device(device)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _check(lib.tpme_fft_plan_create(ctypes.byref(handle), 0 if dtype == torch.float32 else 1,
                                            *self.ns, self.batch), "tpme_fft_plan_create")
        self.handle = handle
        self.own_fft = int(lib.tpme_fft_plan_uses_own_fft(handle))  # 0 (cuFFT), 3 or 5 kernels

    def __del__(self):
        try:
            if getattr(self, "handle", None) and _lib is not None:
                _lib.tpme_fft_plan_destroy(self.handle)
        except Exception:
            pass


_plan_cache: dict = {}


def get_plan(dtype, ns, batch, device) -> FFTPlan:
    key = (dtype, tuple(int(v) for v in ns), int(batch), torch.device(device).index)
    plan = _plan_cache.get(key)
    if plan is None:
        if len(_plan_cache) > 32:
            _plan_cache.clear()
        plan = _plan_cache[key] = FFTPlan(dtype, ns, batch, device)
    return plan


def make_green(kind, scale, recip, spacing=(0.0, 0.0, 0.0), smearing=1.0, prefactor=1.0,
               exponent=1, p3m_nodes=0, table=None, p3m_mode=0, differential_order=2) -> _Green:
    g = _Green()
    g.kind, g.exponent, g.p3m_nodes = int(kind), int(exponent), int(p3m_nodes)
    g.p3m_mode = int(p3m_mode) | (int(differential_order) << 8 if p3m_mode else 0)
    g.smearing, g.prefactor, g.scale = float(smearing), float(prefactor), float(scale)
    for k in range(9):
        g.recip[k] = float(recip[k])
    for k in range(3):
        g.spacing[k] = float(spacing[k])
    g.table = table.data_ptr() if table is not None else None
    g._keepalive = table
    return g


def half_complex_shape(mesh_shape):
    c, nx, ny, nz = mesh_shape
    return (c, nx, ny, nz // 2 + 1, 2)


def kfilter_apply(mesh, green: _Green, keep_hat=False, want_dc=False):
    """irfft3(G * rfft3(mesh)); returns (filtered mesh, rfft3(mesh) | None[, sum of the mesh (C,)])."""
    lib = load()
    c, nx, ny, nz = mesh.shape
    plan = get_plan(mesh.dtype, (nx, ny, nz), c, mesh.device)
    out = torch.empty_like(mesh)
    work = torch.empty(half_complex_shape(mesh.shape), dtype=mesh.dtype, device=mesh.device)
    kept = torch.empty_like(work) if keep_hat else None
    dc = torch.empty(c, dtype=mesh.dtype, device=mesh.device) if want_dc else None
    if plan.own_fft and not keep_hat and (green.kind >= 3 or (green.p3m_nodes > 0 and green.p3m_mode & 255)):
        # spline kernels / P3M modes 1-3: the table kernel evaluates them per k-point (tpme_green_table), the
        # fused x pass of the hand-written FFT multiplies by that table (its register budget is tuned for the
        # closed forms)
        table = green_table(mesh.dtype, (nx, ny, nz), green, mesh.device)
        lean = _Green()
        lean.kind, lean.scale, lean.table = GREEN_TABLE, 1.0, table.data_ptr()
        for k in range(9):
            lean.recip[k] = green.recip[k]
        lean._keepalive = table
        green = lean
    with _on(mesh, "mesh"):
        _check(lib.tpme_kfilter_apply(plan.handle, _dev(mesh, "mesh_values"), _dev(out, "out"),
                                      _dev(work, "work"), _dev(kept, "keep"), ctypes.byref(green),
                                      _dev(dc, "dc"), _stream()), "tpme_kfilter_apply")
    # hand-written path: rows R2C, y lines, x lines . G, y lines, rows C2R (5 of our kernels);
    # cuFFT path: our Green multiply only (the cuFFT kernels are library launches)
    _count(plan.own_fft if plan.own_fft and not keep_hat else 1)
    if want_dc:
        return out, kept, dc
    return out, kept


def rfft3(mesh):
    lib = load()
    c, nx, ny, nz = mesh.shape
    plan = get_plan(mesh.dtype, (nx, ny, nz), c, mesh.device)
    hat = torch.empty(half_complex_shape(mesh.shape), dtype=mesh.dtype, device=mesh.device)
    with _on(mesh, "mesh"):
        _check(lib.tpme_rfft3(plan.handle, _dev(mesh, "mesh"), _dev(hat, "hat"), _stream()), "tpme_rfft3")
    return hat


def green_table(dtype, ns, green: _Green, device):
    lib = load()
    nx, ny, nz = ns
    out = torch.empty((nx, ny, nz // 2 + 1), dtype=dtype, device=device)
    with torch.cuda.device(device):
        _check(lib.tpme_green_table(0 if dtype == torch.float32 else 1, _dev(out, "table"), nx, ny, nz,
                                    ctypes.byref(green), _stream()), "tpme_green_table")
    _count()
    return out


def green_table_vjp(x_hat, y_hat, ns, scale: float):
    lib = load()
    nx, ny, nz = ns
    c = x_hat.shape[0]
    out = torch.empty((nx, ny, nz // 2 + 1), dtype=x_hat.dtype, device=x_hat.device)
    with _on(x_hat, "x_hat"):
        _check(lib.tpme_green_table_vjp(_dtype_id(x_hat), _dev(x_hat, "x_hat"), _dev(y_hat, "y_hat"), c,
                                        nx, ny, nz, float(scale), _dev(out, "grad_table"), _stream()),
               "tpme_green_table_vjp")
    _count()
    return out


# --------------------------------------------------------------------------------------
# slab-decomposed reciprocal space (multi-GPU)
# --------------------------------------------------------------------------------------
MAX_RANKS = 16
IPC_HANDLE_BYTES = 64


def slab_fft_yz(forward: bool, real_mesh, mesh_hat):
    """(y,z) passes of the local planes: real (C, nxl, ny, nz) <-> half-complex (C, nxl, ny, nz/2+1, 2)"""
    lib = load()
    c, nxl, ny, nz = real_mesh.shape
    with _on(real_mesh, "mesh"):
        _check(lib.tpme_slab_fft_yz(_dtype_id(real_mesh), int(forward), _dev(real_mesh, "mesh"),
                                    _dev(mesh_hat, "mesh_hat"), c * nxl, ny, nz, _stream()), "tpme_slab_fft_yz")
    # fused (y,z) plane kernel: 1 launch; otherwise rows + lines: 2
    _count(1 if (ny <= 128 and nz <= 128) else 2)


def slab_fft_x_green(mesh_hat_t, ns, y0: int, green: _Green):
    """in place on (C, nx, ny_local, nz/2+1, 2): x transform, multiply by G, inverse x transform"""
    lib = load()
    nx, ny, nz = ns
    c, nx_, nyl, nzh, _ = mesh_hat_t.shape
    assert nx_ == nx and nzh == nz // 2 + 1
    with _on(mesh_hat_t, "mesh_hat"):
        _check(lib.tpme_slab_fft_x_green(_dtype_id(mesh_hat_t), _dev(mesh_hat_t, "mesh_hat"), c, nx, ny, nz,
                                         y0, nyl, ctypes.byref(green), _stream()), "tpme_slab_fft_x_green")
    _count()


def slab_exchange_copy(src, dst_ptrs, n_c, n_p, n_a, run, src_strides, dst_strides):
    """
    dst[p][c * dst_c + a * dst_a + i] = src[c * src_c + p * src_p + a * src_a + i] in complex
    elements; `dst_ptrs` are raw device addresses (local blocks or mapped peer buffers).
    """
    lib = load()
    elem = 8 if src.dtype == torch.float32 else 16
    arr = (_vp * n_p)(*[int(p) for p in dst_ptrs])
    with _on(src, "src"):
        _check(lib.tpme_slab_exchange_copy(elem, _dev(src, "src"), arr, n_c, n_p, n_a, run,
                                           src_strides[0], src_strides[1], src_strides[2],
                                           dst_strides[0], dst_strides[1], _stream()),
               "tpme_slab_exchange_copy")
    _count()


def make_slab_peers(rank: int, hat_ptrs, hat_t_ptrs) -> _SlabPeers:
    peers = _SlabPeers()
    peers.n_ranks, peers.rank = len(hat_ptrs), int(rank)
    for p, (a, b) in enumerate(zip(hat_ptrs, hat_t_ptrs)):
        peers.hat[p], peers.hat_t[p] = int(a), int(b)
    return peers


def slab_fft_yz_push(real_mesh, ns, peers: _SlabPeers):
    """forward (y,z) passes of the local planes, results stored into the y-slab arrays of all ranks"""
    lib = load()
    nx, ny, nz = ns
    c = real_mesh.shape[0]
    with _on(real_mesh, "mesh"):
        _check(lib.tpme_slab_fft_yz_push(_dtype_id(real_mesh), _dev(real_mesh, "mesh"), c, nx, ny, nz,
                                         ctypes.byref(peers), _stream()), "tpme_slab_fft_yz_push")
    _count(2)


def slab_fft_x_green_push(dtype, device, n_channels: int, ns, green: _Green, peers: _SlabPeers):
    """x pass . G . inverse x pass of the local y rows, results stored into the x-slab arrays of all ranks"""
    lib = load()
    nx, ny, nz = ns
    with torch.cuda.device(device):
        _check(lib.tpme_slab_fft_x_green_push(0 if dtype == torch.float32 else 1, n_channels, nx, ny, nz,
                                              ctypes.byref(green), ctypes.byref(peers), _stream()),
               "tpme_slab_fft_x_green_push")
    _count()


class PeerBuffer:
    """cudaMalloc'ed exchange buffer with a CUDA IPC handle (not managed by torch's allocator)"""

    def __init__(self, n_bytes: int, device):
        lib = load()
        self.device = torch.device(device)
        self.n_bytes = int(n_bytes)
        ptr = _vp()
        handle = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            _check(lib.tpme_peer_buffer_create(self.n_bytes, ctypes.byref(ptr), handle), "tpme_peer_buffer_create")
        self.ptr = int(ptr.value)
        self.handle = handle.raw
        self._opened = []

    def open_peer(self, handle: bytes) -> int:
        lib = load()
        ptr = _vp()
        with torch.cuda.device(self.device):
            _check(lib.tpme_peer_buffer_open(handle, ctypes.byref(ptr)), "tpme_peer_buffer_open")
        self._opened.append(int(ptr.value))
        return int(ptr.value)

    def as_tensor(self, offset_bytes: int, shape, dtype) -> torch.Tensor:
        """torch view of a region of the buffer (through the CUDA array interface)"""
        n = 1
        for v in shape:
            n *= int(v)
        itemsize = torch.empty((), dtype=dtype).element_size()
        assert offset_bytes + n * itemsize <= self.n_bytes

        class _Region:
            pass

        region = _Region()
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        region.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": typestr,
                                           "data": (self.ptr + offset_bytes, False), "version": 2}
        region._owner = self
        with torch.cuda.device(self.device):
            return torch.as_tensor(region, device=self.device)

    def close(self):
        if _lib is None:
            return
        with torch.cuda.device(self.device):
            for p in self._opened:
                _lib.tpme_peer_buffer_close(_vp(p))
            self._opened = []
            if self.ptr:
                _lib.tpme_peer_buffer_destroy(_vp(self.ptr))
                self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def peer_barrier(flag_ptrs, rank: int, epoch, error_flag, timeout_seconds: float = 60.0):
    lib = load()
    n = len(flag_ptrs)
    arr = (_vp * n)(*[int(p) for p in flag_ptrs])
    with _on(epoch, "epoch"):
        _check(lib.tpme_peer_barrier(arr, n, rank, _dev(epoch, "epoch"), float(timeout_seconds),
                                     _dev(error_flag, "error_flag"), _stream()), "tpme_peer_barrier")
    _count()


def peer_allreduce(dtype, device, in_ptrs, out_ptrs, rank: int, n: int):
    """sum all-reduce of n reals between peer-mapped regions (call between two peer barriers)"""
    lib = load()
    w = len(in_ptrs)
    a = (_vp * w)(*[int(p) for p in in_ptrs])
    b = (_vp * w)(*[int(p) for p in out_ptrs])
    with torch.cuda.device(device):
        _check(lib.tpme_peer_allreduce(0 if dtype == torch.float32 else 1, a, b, w, rank, n, _stream()),
               "tpme_peer_allreduce")
    _count()


def multimem_allreduce(dtype, device, multicast_ptr: int, world: int, rank: int, n: int):
    """in-switch (NVLS) sum all-reduce of n reals behind a multicast address (call between two peer barriers)"""
    lib = load()
    with torch.cuda.device(device):
        _check(lib.tpme_multimem_allreduce(0 if dtype == torch.float32 else 1, _vp(int(multicast_ptr)), world, rank, n,
                                           _stream()), "tpme_multimem_allreduce")
    _count()


# --------------------------------------------------------------------------------------
# real space
# --------------------------------------------------------------------------------------
def make_pair_potential(kind, smearing=1.0, prefactor=1.0, exponent=1, exclusion_radius=None,
                        exclusion_degree=1) -> _PairPotential:
    p = _PairPotential()
    p.kind, p.exponent, p.exclusion_degree = int(kind), int(exponent), int(exclusion_degree)
    p.smearing, p.prefactor = float(smearing), float(prefactor)
    p.exclusion_radius = float(exclusion_radius) if exclusion_radius is not None else -1.0
    return p


def _index_args(idx):
    if idx.dtype == torch.int64:
        return 1
    if idx.dtype == torch.int32:
        return 0
    raise TypeError(f"neighbor_indices must be int32 or int64, got {idx.dtype}")


#: number of valid pairs (0-dim int64 device tensor) of index buffers that hold a list built on the device
#: without a host synchronisation (neighbors.DeviceNeighborList), keyed by the buffer's address
_pair_counts: dict = {}


def register_pair_count(indices: torch.Tensor, count: torch.Tensor) -> None:
    """the pairs ``p >= count`` of ``indices`` (and of everything indexed like it) are padding"""
    if count.dtype != torch.int64 or count.numel() != 1 or count.device != indices.device:
        raise ValueError("`count` must be a one-element int64 tensor on the device of `indices`")
    ref = weakref.ref(indices, lambda _, key=indices.data_ptr(): _pair_counts.pop(key, None))
    _pair_counts[indices.data_ptr()] = (ref, count)


def pair_count_of(indices: torch.Tensor):
    hit = _pair_counts.get(indices.data_ptr())
    return None if hit is None else hit[1]


def pair_forward(charges, idx, dist, pair_values, mask_u8, full_list: bool, pot: _PairPotential,
                 out=None):
    """accumulates into `out` (allocated and zeroed here when None)"""
    lib = load()
    n, c = charges.shape
    if DEBUG_BOUNDS and idx.numel():
        # the kernels trust the pair list (like the reference's index_add_ on CUDA, which device-asserts):
        # TPME_DEBUG_BOUNDS=1 checks it on the host first (one synchronisation)
        lo, hi = int(idx.min()), int(idx.max())
        if lo < 0 or hi >= n:
            raise IndexError(f"neighbor_indices must lie in [0, {n}), got values in [{lo}, {hi}]")
    if out is None:
        out = torch.zeros_like(charges)
    with _on(charges, "charges"):
        _check(lib.tpme_pair_forward(_dtype_id(charges), _dev(charges, "charges"),
                                     _dev(idx, "neighbor_indices"), _index_args(idx),
                                     _dev(dist, "neighbor_distances"), _dev(pair_values, "pair_values"),
                                     _dev(mask_u8, "pair_mask"), idx.shape[0], _dev(pair_count_of(idx), "pair count"),
                                     n, c, int(full_list),
                                     ctypes.byref(pot), _dev(out, "out"), _stream()), "tpme_pair_forward")
    _count()
    return out


def pair_backward(charges, idx, dist, pair_values, mask_u8, grad_out, full_list: bool,
                  pot: _PairPotential, want_charges=True, want_pairs=True, grad_charges_out=None,
                  grad_pairs_out=None):
    lib = load()
    n, c = charges.shape
    g_q = grad_charges_out
    if g_q is None and want_charges:
        g_q = torch.zeros_like(charges)
    g_p = grad_pairs_out
    count = pair_count_of(idx)
    if g_p is None and want_pairs:    # padding entries of a counted list are not written: zeros there
        g_p = (torch.empty if count is None else torch.zeros)(idx.shape[0], dtype=charges.dtype, device=charges.device)
    with _on(charges, "charges"):
        _check(lib.tpme_pair_backward(_dtype_id(charges), _dev(charges, "charges"),
                                      _dev(idx, "neighbor_indices"), _index_args(idx),
                                      _dev(dist, "neighbor_distances"), _dev(pair_values, "pair_values"),
                                      _dev(mask_u8, "pair_mask"), _dev(grad_out, "grad_out"),
                                      idx.shape[0], _dev(count, "pair count"), n, c, int(full_list),
                                      ctypes.byref(pot),
                                      _dev(g_q, "grad_charges"), _dev(g_p, "grad_pairs"), _stream()),
               "tpme_pair_backward")
    _count()
    return g_q, g_p


def pair_distances(positions, cell_host, idx, shifts, count, out):
    """out[p] = |r_j + S_p . cell - r_i| (tpme_pair_distances); `count`: device pair count or None"""
    lib = load()
    with _on(positions, "positions"):
        _check(lib.tpme_pair_distances(_dtype_id(positions), _dev(positions, "positions"),
                                       _mat9(np.asarray(cell_host).reshape(-1)), _dev(idx, "neighbor_indices"),
                                       _index_args(idx), _dev(shifts, "shifts"), idx.shape[0],
                                       _dev(count, "pair count"), _dev(out, "distances"), _stream()),
               "tpme_pair_distances")
    _count()
    return out


def pair_distances_backward(positions, cell_host, idx, shifts, grad_d, count, grad_positions, grad_cell):
    lib = load()
    with _on(positions, "positions"):
        _check(lib.tpme_pair_distances_backward(_dtype_id(positions), _dev(positions, "positions"),
                                                _mat9(np.asarray(cell_host).reshape(-1)),
                                                _dev(idx, "neighbor_indices"), _index_args(idx),
                                                _dev(shifts, "shifts"), _dev(grad_d, "grad_distances"),
                                                idx.shape[0], _dev(count, "pair count"),
                                                _dev(grad_positions, "grad_positions"),
                                                _dev(grad_cell, "grad_cell"), _stream()),
               "tpme_pair_distances_backward")
    _count()
