"""
Mesh building blocks on top of the CUDA kernels: ``MeshInterpolator`` (spread / gather),
``KSpaceFilter`` / ``P3MKSpaceFilter`` (FFT * G * iFFT) and the mesh-size / k-vector helpers.

Host-side mirror of ``src/torchpme/lib/{mesh_interpolator,kspace_filter,kvectors}.py``:
same class names, constructor arguments, method names and error texts.  What differs is
underneath: no ``(n^3, N)`` index / weight tensors and no k-vector tensors are ever
materialised -- ``compute_weights`` only records the points, and each stage is one kernel
launch of ``libtorchpme_b200.so`` wrapped in a ``torch.autograd.Function`` carrying the
analytic backward.
"""

from __future__ import annotations

import math
import weakref

import numpy as np
import torch
from torch.autograd.function import once_differentiable

from . import _cpu, _native

_NORMS = ("ortho", "forward", "backward")


# --------------------------------------------------------------------------------------
# host-side cell geometry (one device->host read per distinct cell tensor)
# --------------------------------------------------------------------------------------
class CellGeometry:
    """Numbers derived from a (3, 3) cell that the kernels take as by-value arguments."""

    def __init__(self, cell: torch.Tensor):
        c = cell.detach().to("cpu", torch.float64).numpy().copy()
        self.cell = c
        self.inverse = np.linalg.inv(c)
        self.volume = float(abs(np.linalg.det(c)))
        self.norms = np.linalg.norm(c, axis=1)
        self.recip = (2.0 * math.pi * self.inverse.T).reshape(-1).tolist()

    def r2u(self, ns) -> list:
        """row-major 3x3 with u = r @ r2u = ns * (r @ cell^-1)"""
        return (self.inverse * np.asarray(ns, dtype=np.float64)[None, :]).reshape(-1).tolist()

    def spacing(self, ns) -> list:
        return (self.norms / np.asarray(ns, dtype=np.float64)).tolist()

    def ns_mesh(self, mesh_spacing: float) -> tuple:
        """2^ceil(log2(2 |a_i| / h + 1))   (reference: lib/kvectors.py:5-21)"""
        return tuple(int(2 ** math.ceil(math.log2(2.0 * n / mesh_spacing + 1.0))) for n in self.norms)


_geometry_cache: dict = {}


def geometry_of(cell: torch.Tensor) -> CellGeometry:
    """Cached per tensor *object* and in-place version; a new tensor costs one sync."""
    key = id(cell)
    hit = _geometry_cache.get(key)
    if hit is not None:
        ref, version, geom = hit
        if ref() is cell and version == cell._version:
            return geom
    if len(_geometry_cache) > 64:
        _geometry_cache.clear()
    geom = CellGeometry(cell)
    _geometry_cache[key] = (weakref.ref(cell), cell._version, geom)
    return geom


def device_cell(cell, device, dtype: torch.dtype | None = None) -> torch.Tensor:
    """
    A (3, 3) cell that lives on the host (numpy array, nested list or CPU tensor) as a tensor on `device`
    whose geometry is already in the cache, computed from the host values: the calculators then take their
    by-value launch parameters (inverse cell, reciprocal cell, volume, mesh size) without the device -> host
    read, i.e. without a stream synchronisation.  For drivers whose box changes every step (NPT molecular
    dynamics): ``cell = torchpme_b200.device_cell(box, "cuda")`` instead of ``torch.tensor(box, device="cuda")``.
    """
    host = torch.as_tensor(cell).detach().to("cpu")
    if host.shape != (3, 3):
        raise ValueError(f"cell of shape {list(host.shape)} should be of shape (3, 3)")
    if dtype is None:
        dtype = host.dtype if host.dtype.is_floating_point else torch.get_default_dtype()
    host = host.to(dtype).contiguous()
    on_device = host.to(device=device, non_blocking=True) if torch.device(device).type != "cpu" else host.clone()
    if len(_geometry_cache) > 64:
        _geometry_cache.clear()
    _geometry_cache[id(on_device)] = (weakref.ref(on_device), on_device._version, CellGeometry(host))
    return on_device


def _host_ints(t: torch.Tensor) -> tuple:
    return tuple(int(v) for v in t.detach().to("cpu").tolist())


def get_ns_mesh(cell: torch.Tensor, mesh_spacing: float) -> torch.Tensor:
    """Mesh size (power of two per axis) for a target spacing, as an int64 tensor on ``cell.device``."""
    ns = geometry_of(cell).ns_mesh(mesh_spacing)
    return torch.tensor(ns, dtype=torch.int64, device=cell.device)


def generate_kvectors_for_mesh(cell: torch.Tensor, ns: torch.Tensor) -> torch.Tensor:
    """
    Reciprocal vectors of the rFFT layout, ``(nx, ny, nz//2+1, 3)``, as differentiable torch ops
    (reference: lib/kvectors.py:24-102).  Only the table route of the filter needs them.
    """
    if cell.shape != (3, 3):
        raise ValueError(f"cell of shape {list(cell.shape)} should be of shape (3, 3)")
    if ns.shape != (3,):
        raise ValueError(f"ns of shape {list(ns.shape)} should be of shape (3, )")
    if ns.device != cell.device:
        raise ValueError(
            f"`ns` and `cell` are not on the same device, got {ns.device} and {cell.device}."
        )
    nx, ny, nz = _host_ints(ns)
    return _kvectors(cell, (nx, ny, nz))


def generate_kvectors_for_ewald(cell: torch.Tensor, ns: torch.Tensor) -> torch.Tensor:
    """
    All reciprocal vectors of the ``(nx, ny, nz)`` grid as an ``(nx ny nz, 3)`` tensor (full frequency
    range along z; reference: lib/kvectors.py:24-74,104-130).  Utility only: the Ewald calculator
    that consumes it is not part of this package.
    """
    if cell.shape != (3, 3):
        raise ValueError(f"cell of shape {list(cell.shape)} should be of shape (3, 3)")
    if ns.shape != (3,):
        raise ValueError(f"ns of shape {list(ns.shape)} should be of shape (3, )")
    if ns.device != cell.device:
        raise ValueError(
            f"`ns` and `cell` are not on the same device, got {ns.device} and {cell.device}."
        )
    return _kvectors(cell, _host_ints(ns), half_z=False).reshape(-1, 3)


def _kvectors(cell, ns, half_z: bool = True):
    nx, ny, nz = ns
    inv = torch.linalg.inv_ex(cell)[0] if cell.is_cuda else torch.linalg.inv(cell)
    recip = 2 * torch.pi * inv.T
    opts = dict(device=cell.device, dtype=cell.dtype)
    fx = torch.fft.fftfreq(nx, **opts) * nx
    fy = torch.fft.fftfreq(ny, **opts) * ny
    fz = (torch.fft.rfftfreq(nz, **opts) if half_z else torch.fft.fftfreq(nz, **opts)) * nz
    return (fx[:, None, None, None] * recip[0]
            + fy[None, :, None, None] * recip[1]
            + fz[None, None, :, None] * recip[2])


# --------------------------------------------------------------------------------------
# autograd nodes
# --------------------------------------------------------------------------------------
class _StencilConfig:
    __slots__ = ("r2u", "ns", "nodes", "method", "tiles")

    def __init__(self, r2u, ns, nodes, method, tiles=None):
        self.r2u, self.ns, self.nodes, self.method = r2u, tuple(ns), int(nodes), int(method)
        self.tiles = tiles   # _native.TileSort of the registered points (None: direct kernels)


def _r2u_tensor(cell: torch.Tensor, ns) -> torch.Tensor:
    """differentiable device copy of r2u, only built when the cell needs a gradient"""
    ns_t = torch.tensor(ns, dtype=cell.dtype, device=cell.device)
    return torch.linalg.inv(cell) * ns_t[None, :]


class _Spread(torch.autograd.Function):
    """mesh[c, m] = sum_i w[i, c] W_i(m);  backward = gather of the mesh gradient."""

    @staticmethod
    def forward(ctx, positions, weights, r2u_t, cfg: _StencilConfig):
        pos = positions.detach().contiguous()
        w = weights.detach().contiguous()
        ctx.cfg = cfg
        ctx.save_for_backward(pos, w)
        return _native.spread(pos, w, cfg.r2u, cfg.ns, cfg.nodes, cfg.method, tiles=cfg.tiles)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_mesh):
        pos, w = ctx.saved_tensors
        cfg = ctx.cfg
        need_pos, need_w, need_r2u = ctx.needs_input_grad[:3]
        grad_mesh = grad_mesh.contiguous()
        g_pos = g_w = g_r2u = None
        if need_pos or need_r2u:
            g_pos, g_w, g_r2u = _native.gather_vjp(
                grad_mesh, pos, w, cfg.r2u, cfg.nodes, cfg.method, want_values=need_w,
                want_grad_r2u=need_r2u, tiles=cfg.tiles)
        elif need_w:
            g_w, _ = _native.gather(grad_mesh, pos, cfg.r2u, cfg.nodes, cfg.method, tiles=cfg.tiles)
        return (g_pos if need_pos else None), g_w, g_r2u, None


class _Gather(torch.autograd.Function):
    """values[i, c] = sum_m mesh[c, m] W_i(m);  backward = spread of the value gradient."""

    @staticmethod
    def forward(ctx, mesh, positions, r2u_t, cfg: _StencilConfig):
        mesh_c = mesh.detach().contiguous()
        pos = positions.detach().contiguous()
        need_pos = ctx.needs_input_grad[1]
        need_r2u = ctx.needs_input_grad[2]
        values, dvalues = _native.gather(mesh_c, pos, cfg.r2u, cfg.nodes, cfg.method,
                                         want_values=True, want_grad=need_pos and not need_r2u,
                                         tiles=cfg.tiles)
        ctx.cfg = cfg
        ctx.mesh_shape = mesh_c.shape
        ctx.save_for_backward(pos, dvalues, mesh_c if need_r2u else None)
        return values

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_values):
        pos, dvalues, mesh_c = ctx.saved_tensors
        cfg = ctx.cfg
        need_mesh, need_pos, need_r2u = ctx.needs_input_grad[:3]
        g = grad_values.contiguous()
        g_mesh = g_pos = g_r2u = None
        if need_mesh:
            g_mesh = _native.spread(pos, g, cfg.r2u, cfg.ns, cfg.nodes, cfg.method, tiles=cfg.tiles)
        if need_r2u:
            g_pos, _, g_r2u = _native.gather_vjp(mesh_c, pos, g, cfg.r2u, cfg.nodes, cfg.method,
                                                 want_grad_r2u=True, tiles=cfg.tiles)
        elif need_pos:
            g_pos = torch.einsum("ic,icd->id", g, dvalues)
        return g_mesh, (g_pos if need_pos else None), g_r2u, None


class _FilterConfig:
    __slots__ = ("green_args", "scale")

    def __init__(self, green_args: dict, scale: float):
        self.green_args, self.scale = green_args, scale


def _green_args(cfg: _FilterConfig, table):
    """the filter table (kind 0) is an autograd input; spline kinds carry their own knot table in the config"""
    return cfg.green_args if table is None else dict(cfg.green_args, table=table)


class _KFilter(torch.autograd.Function):
    """out = scale * irfft3(G * rfft3(mesh)) with unnormalised transforms (self-adjoint)."""

    @staticmethod
    def forward(ctx, mesh, table, cfg: _FilterConfig):
        mesh_c = mesh.detach().contiguous()
        need_table = table is not None and ctx.needs_input_grad[1]
        table_c = table.detach().contiguous() if table is not None else None
        green = _native.make_green(scale=cfg.scale, **_green_args(cfg, table_c))
        out, x_hat = _native.kfilter_apply(mesh_c, green, keep_hat=need_table)
        ctx.cfg = cfg
        ctx.save_for_backward(table_c, x_hat)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        table_c, x_hat = ctx.saved_tensors
        cfg = ctx.cfg
        need_mesh, need_table = ctx.needs_input_grad[:2]
        g = grad_out.contiguous()
        g_mesh = g_table = None
        if need_mesh:
            green = _native.make_green(scale=cfg.scale, **_green_args(cfg, table_c))
            g_mesh, g_hat = _native.kfilter_apply(g, green, keep_hat=need_table)
        elif need_table:
            g_hat = _native.rfft3(g)
        if need_table:
            ns = tuple(g.shape[1:])
            g_table = _native.green_table_vjp(x_hat, g_hat, ns, cfg.scale)
        return g_mesh, g_table, None


# --------------------------------------------------------------------------------------
# public blocks
# --------------------------------------------------------------------------------------
class MeshInterpolator(torch.nn.Module):
    """
    Spread per-point weights onto a mesh (:meth:`points_to_mesh`) and interpolate mesh values
    back to points (:meth:`mesh_to_points`) with P3M (1..5 nodes) or Lagrange (3..7 nodes)
    stencils (reference: ``lib/mesh_interpolator.py:4-457``).
    """

    # TorchScript (a calculator that owns an interpolator can be scripted, see calculators.py): nothing of
    # this class is compiled -- it has no forward, and its properties are host-side / inspection code
    __jit_unused_properties__ = ["inverse_cell", "interpolation_weights", "x_shifts", "y_shifts", "z_shifts",
                                 "x_indices", "y_indices", "z_indices"]

    def __init__(self, cell: torch.Tensor, ns_mesh: torch.Tensor, interpolation_nodes: int, method: str):
        super().__init__()
        allowed = {"Lagrange": (3, 4, 5, 6, 7), "P3M": (1, 2, 3, 4, 5)}
        if method not in allowed:
            raise ValueError(f"method '{method}' is not supported. Choose from 'Lagrange' or 'P3M'")
        if interpolation_nodes not in allowed[method]:
            lo, hi = allowed[method][0], allowed[method][-1]
            raise ValueError(
                f"`interpolation_nodes` is {interpolation_nodes} but only values from {lo} to {hi} "
                f"for method '{method}' are allowed"
            )
        self.method = method
        self.interpolation_nodes = interpolation_nodes
        self._ns_host = None
        self._points = None
        self._tiles = None
        self._torch_stencil = None
        self.cell = None
        self.ns_mesh = None
        self.update(cell, ns_mesh)

    def _apply(self, fn, *args, **kwargs):
        """`.to()` / `.cuda()` also move the geometry tensors this module holds as plain attributes"""
        super()._apply(fn, *args, **kwargs)
        if self.cell is not None:
            self.cell = fn(self.cell)
            self._dtype, self._device = self.cell.dtype, self.cell.device
        if self.ns_mesh is not None:
            self.ns_mesh = fn(self.ns_mesh)
        self._points = self._tiles = self._torch_stencil = None
        return self

    # the calculators already know ns on the host and skip the tensor round trip
    def _update_host(self, cell: torch.Tensor, ns_host: tuple):
        self.cell = cell
        self._dtype, self._device = cell.dtype, cell.device
        self._ns_host = tuple(ns_host)
        self.ns_mesh = None
        self._tiles = None
        self._torch_stencil = None

    def update(self, cell: torch.Tensor | None = None, ns_mesh: torch.Tensor | None = None) -> None:
        self._tiles = None
        self._torch_stencil = None
        if cell is not None:
            if cell.shape != (3, 3):
                raise ValueError(f"cell of shape {list(cell.shape)} should be of shape (3, 3)")
            self.cell = cell
            self._dtype, self._device = cell.dtype, cell.device
        if ns_mesh is not None:
            if ns_mesh.shape != (3,):
                raise ValueError(f"shape {list(ns_mesh.shape)} of `ns_mesh` has to be (3,)")
            self.ns_mesh = ns_mesh
            self._ns_host = None
        if self.ns_mesh is not None and self.cell.device != self.ns_mesh.device:
            raise ValueError(
                "`cell` and `ns_mesh` are on different devices, got "
                f"{self.cell.device} and {self.ns_mesh.device}"
            )

    @property
    def inverse_cell(self) -> torch.Tensor:
        return torch.linalg.inv(self.cell)

    def _ns(self) -> tuple:
        if self._ns_host is None:
            self._ns_host = _host_ints(self.ns_mesh)
        return self._ns_host

    def get_mesh_xyz(self) -> torch.Tensor:
        """Cartesian positions of the mesh points, ``(nx, ny, nz, 3)``."""
        axes = [torch.arange(n, dtype=self._dtype, device=self._device) / n for n in self._ns()]
        frac = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
        return frac @ self.cell

    def compute_weights(self, positions: torch.Tensor):
        """Register the points used by the next spread / gather calls (no tensors are built)."""
        if positions.device != self._device:
            raise ValueError(
                f"`positions` device {positions.device} is not the same as instance "
                f"device {self._device}"
            )
        if positions.dim() != 2 or positions.shape[1] != 3:
            raise ValueError(f"shape {list(positions.shape)} of `positions` has to be (N, 3)")
        self._points = positions
        self._tiles = None   # (dtype, TileSort | None) of these points, built at the first use
        self._torch_stencil = None   # (flat index, weight) of the torch formulation (CPU tensors, inspection)

    def _tile_sort(self, dtype, r2u, ns):
        """the registered points binned by mesh tile (the counterpart of the reference's index tensors)"""
        if self._tiles is None or self._tiles[0] != dtype:
            pos = self._points.detach().to(dtype).contiguous()
            tiles = _native.tile_sort(pos, r2u, ns, self.interpolation_nodes, _native.METHOD_ID[self.method]) \
                if pos.is_cuda else None
            self._tiles = (dtype, tiles)
        return self._tiles[1]

    def _stencil(self, dtype=None):
        if self._points is None:
            raise ValueError("`compute_weights` has to be called before interpolating")
        ns = self._ns()
        r2u = geometry_of(self.cell).r2u(ns)
        cfg = _StencilConfig(r2u, ns, self.interpolation_nodes, _native.METHOD_ID[self.method],
                             self._tile_sort(dtype or self._dtype, r2u, ns))
        r2u_t = _r2u_tensor(self.cell, ns) if self.cell.requires_grad else None
        return cfg, r2u_t

    # ---- torch formulation: CPU tensors, and the reference's inspection attributes ----------------
    def _stencil_torch(self):
        if self._points is None:
            raise ValueError("`compute_weights` has to be called before interpolating")
        if self._torch_stencil is None:
            self._torch_stencil = _cpu.stencil(self._points.to(self._dtype), self.cell, self._ns(),
                                               self.interpolation_nodes, self.method)
        return self._torch_stencil

    def _node_table(self):
        """1-D weights (n, N, 3) and node indices (n, N, 3) as the reference stores them (:65-79, 326-359)"""
        n = self.interpolation_nodes
        ns = self._ns()
        pos = self._points.to(self._dtype)
        ns_t = torch.tensor(ns, dtype=self._dtype, device=self._device)
        u = (pos @ torch.linalg.inv(self.cell)) * ns_t
        base = torch.floor(u.detach()) if n % 2 == 0 else torch.round(u.detach())
        x = u - (base + 0.5) if n % 2 == 0 else u - base
        w = _cpu._weights_1d(x, n, self.method).permute(2, 0, 1)                       # (n, N, 3)
        first = base.to(torch.int64) + (1 - (n + 1) // 2)
        offs = torch.arange(n, device=self._device)
        idx = (first[None] + offs[:, None, None]) % torch.tensor(ns, device=self._device)   # (n, N, 3)
        return w, idx

    @property
    def interpolation_weights(self) -> torch.Tensor:
        """1-D weights of the registered points, ``(n, N, 3)`` (computed on demand; the kernels never build it)"""
        return self._node_table()[0]

    def _shifts(self, axis: int) -> torch.Tensor:
        n = self.interpolation_nodes
        grids = torch.meshgrid(*[torch.arange(n, device=self._device)] * 3, indexing="ij")
        return grids[axis].reshape(-1)

    x_shifts = property(lambda self: self._shifts(0))
    y_shifts = property(lambda self: self._shifts(1))
    z_shifts = property(lambda self: self._shifts(2))

    def _indices(self, axis: int) -> torch.Tensor:
        return self._node_table()[1][self._shifts(axis), :, axis]                      # (n^3, N)

    x_indices = property(lambda self: self._indices(0))
    y_indices = property(lambda self: self._indices(1))
    z_indices = property(lambda self: self._indices(2))

    def points_to_mesh(self, particle_weights: torch.Tensor) -> torch.Tensor:
        if particle_weights.device != self._device:
            raise ValueError(
                f"`particle_weights` device {particle_weights.device} is not the same "
                f"as instance device {self._device}"
            )
        if particle_weights.dim() != 2:
            raise ValueError(
                f"`particle_weights` of dimension {particle_weights.dim()} has to be of dimension 2"
            )
        if not particle_weights.is_cuda:      # device dispatch: CPU tensors use the torch formulation
            flat, weight = self._stencil_torch()
            return _cpu.spread(flat, weight, particle_weights.to(self._dtype), self._ns())
        cfg, r2u_t = self._stencil()
        pos = self._points.to(self._dtype)
        return _Spread.apply(pos, particle_weights.to(self._dtype), r2u_t, cfg)

    def mesh_to_points(self, mesh_vals: torch.Tensor) -> torch.Tensor:
        if mesh_vals.dim() != 4:
            raise ValueError(f"`mesh_vals` of dimension {mesh_vals.dim()} has to be of dimension 4")
        if tuple(mesh_vals.shape[1:]) != tuple(self._ns()):
            raise ValueError(
                f"`mesh_vals` of shape {list(mesh_vals.shape)} does not match the mesh {list(self._ns())}"
            )
        if not mesh_vals.is_cuda:
            flat, weight = self._stencil_torch()
            return _cpu.gather(flat, weight.to(mesh_vals.dtype), mesh_vals)
        cfg, r2u_t = self._stencil(mesh_vals.dtype)
        return _Gather.apply(mesh_vals, self._points.to(mesh_vals.dtype), r2u_t, cfg)


class KSpaceKernel(torch.nn.Module):
    """Interface of a user-defined reciprocal-space kernel (reference: kspace_filter.py:7-34)."""

    def kernel_from_k_sq(self, kvectors: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            f"kernel_from_k_sq is not implemented for '{self.__class__.__name__}'"
        )


_nan_check = True


def set_nan_check(enabled: bool) -> None:
    """
    Toggle the eager NaN guard of :meth:`KSpaceFilter.forward` (it costs a device->host
    synchronisation, like the reference's, lib/kspace_filter.py:189-195).  It is skipped
    automatically while a CUDA graph is being captured.
    """
    global _nan_check
    _nan_check = bool(enabled)


def check_filter_result(result: torch.Tensor) -> None:
    """the reference's guard on the filtered mesh (lib/kspace_filter.py:189-195), same text; one host sync;
    skipped when switched off (:func:`set_nan_check`) and while a CUDA graph is being captured"""
    if not _nan_check or (result.is_cuda and torch.cuda.is_current_stream_capturing()):
        return
    if torch.isnan(result).any():
        raise ValueError(
            "NaNs detected in the k-space filter result. This are probably caused "
            "by an unsuitable `mesh_spacing`, resulting in a problematic grid of "
            f"shape: {list(result.shape)}. Try adjsuting the grid by using a "
            "different `mesh_spacing` value."
        )


class KSpaceFilter(torch.nn.Module):
    """
    ``irfftn(rfftn(mesh) * kernel(k^2))`` on a real-space mesh (reference:
    ``lib/kspace_filter.py:37-222``).  Kernels that are exactly a ``CoulombPotential`` or
    ``InversePowerLawPotential`` are evaluated inside the multiply kernel from the reciprocal
    cell; any other kernel (or a cell that requires grad) goes through a filter table built
    with differentiable torch ops.
    """

    __jit_unused_properties__ = ["_kvectors", "_k_sq", "_kfilter"]

    def __init__(self, cell, ns_mesh, kernel, fft_norm: str = "ortho", ifft_norm: str = "ortho"):
        super().__init__()
        if fft_norm not in _NORMS:
            raise ValueError(f"Invalid option '{fft_norm}' for the `fft_norm` parameter.")
        if ifft_norm not in _NORMS:
            raise ValueError(f"Invalid option '{ifft_norm}' for the `ifft_norm` parameter.")
        self._fft_norm, self._ifft_norm = fft_norm, ifft_norm
        self.kernel = kernel
        self._ns_host = None
        self.ns_mesh = None
        self._table = None
        self.update(cell, ns_mesh)

    def _apply(self, fn, *args, **kwargs):
        """`.to()` / `.cuda()` also move the geometry tensors this module holds as plain attributes"""
        super()._apply(fn, *args, **kwargs)
        if getattr(self, "cell", None) is not None:
            self.cell = fn(self.cell)
        if self.ns_mesh is not None:
            self.ns_mesh = fn(self.ns_mesh)
        self._table = None
        return self

    # ---- geometry -------------------------------------------------------------------------
    def _update_host(self, cell: torch.Tensor, ns_host: tuple):
        self.cell = cell
        self._ns_host = tuple(ns_host)
        self.ns_mesh = None
        self._table = None

    def update(self, cell: torch.Tensor | None = None, ns_mesh: torch.Tensor | None = None) -> None:
        if cell is not None:
            if cell.shape != (3, 3):
                raise ValueError(f"cell of shape {list(cell.shape)} should be of shape (3, 3)")
            self.cell = cell
        if ns_mesh is not None:
            if ns_mesh.shape != (3,):
                raise ValueError(f"shape {list(ns_mesh.shape)} of `ns_mesh` has to be (3,)")
            self.ns_mesh = ns_mesh
            self._ns_host = None
        if self.ns_mesh is not None and self.cell.device != self.ns_mesh.device:
            raise ValueError(
                "`cell` and `ns_mesh` are on different devices, got "
                f"{self.cell.device} and {self.ns_mesh.device}"
            )
        self._table = None  # the filter is re-derived lazily: kernel parameters may have changed

    def _ns(self) -> tuple:
        if self._ns_host is None:
            self._ns_host = _host_ints(self.ns_mesh)
        return self._ns_host

    # ---- filter tables (torch ops; only for generic kernels / cell gradients / inspection) --
    @property
    def _kvectors(self) -> torch.Tensor:
        return _kvectors(self.cell, self._ns())

    @property
    def _k_sq(self) -> torch.Tensor:
        return torch.linalg.norm(self._kvectors, dim=3) ** 2

    def _influence(self, kvectors):
        return None

    @property
    def _kfilter(self) -> torch.Tensor:
        if self._table is None:
            kv = self._kvectors
            table = self.kernel.kernel_from_k_sq(torch.linalg.norm(kv, dim=3) ** 2)
            infl = self._influence(kv)
            self._table = table if infl is None else infl * table
        return self._table

    # ---- application ----------------------------------------------------------------------
    def _scale(self, ns) -> float:
        n = float(ns[0] * ns[1] * ns[2])
        fwd = {"backward": 0.0, "ortho": 0.5, "forward": 1.0}[self._fft_norm]
        inv = {"backward": 1.0, "ortho": 0.5, "forward": 0.0}[self._ifft_norm]
        return n ** (-(fwd + inv))

    def _p3m_nodes(self) -> int:
        return 0

    def _p3m_mode(self) -> dict:
        return {}

    def _wants_table(self) -> bool:
        if self.cell.requires_grad:
            return True
        if any(t.requires_grad for t in list(self.kernel.parameters()) + list(self.kernel.buffers())):
            return True
        return getattr(self.kernel, "_native_filter", lambda: None)() is None

    def forward(self, mesh_values: torch.Tensor) -> torch.Tensor:
        if torch.jit.is_scripting():
            raise RuntimeError("KSpaceFilter is not available under TorchScript: script the calculator that owns it")
        return self._forward_impl(mesh_values)

    @torch.jit.unused
    def _forward_impl(self, mesh_values: torch.Tensor) -> torch.Tensor:
        if mesh_values.dim() != 4:
            raise ValueError(
                f"`mesh_values` needs to be a 4 dimensional tensor, got {mesh_values.dim()}"
            )
        if mesh_values.device != self.cell.device:
            raise ValueError(
                "`mesh_values` and the k-space filter are on different devices, got "
                f"{mesh_values.device} and {self.cell.device}"
            )
        ns = self._ns()
        if tuple(mesh_values.shape[-3:]) != ns:
            raise ValueError("The real-space mesh is inconsistent with the k-space grid.")
        geom = geometry_of(self.cell)
        scale = self._scale(ns)
        if not mesh_values.is_cuda:           # device dispatch: CPU tensors use torch.fft with the filter table
            result = _cpu.kfilter(mesh_values, self._kfilter.to(mesh_values.dtype), scale)
            check_filter_result(result)
            return result
        if self._wants_table():
            table = self._kfilter.to(mesh_values.dtype)
            cfg = _FilterConfig(dict(kind=_native.GREEN_TABLE, recip=geom.recip), scale)
            result = _KFilter.apply(mesh_values, table, cfg)
        else:
            cfg = _FilterConfig(dict(recip=geom.recip, spacing=geom.spacing(ns), p3m_nodes=self._p3m_nodes(),
                                     **self._p3m_mode(), **self.kernel._native_filter()), scale)
            result = _KFilter.apply(mesh_values, None, cfg)
        check_filter_result(result)
        return result


class P3MKSpaceFilter(KSpaceFilter):
    """
    Filter with the P3M influence function folded in (reference: ``lib/kspace_filter.py:225-363``).
    All modes (0: point-charge potential; 1-3 with the finite-difference operator of order 1-6) are evaluated
    per k-point inside the filter kernel (``csrc/green.cuh``); the torch table ``_influence`` below serves
    the differentiable route (cell / parameter gradients) and the CPU path.
    """

    _DIFF_COEFF = (
        (1.0, 0.0, 0.0, 0.0, 0.0, 0.0),
        (4 / 3, -1 / 3, 0.0, 0.0, 0.0, 0.0),
        (3 / 2, -3 / 5, 1 / 10, 0.0, 0.0, 0.0),
        (8 / 5, -4 / 5, 8 / 35, -1 / 35, 0.0, 0.0),
        (5 / 3, -20 / 21, 5 / 14, -5 / 63, 1 / 126, 0.0),
        (12 / 7, -15 / 14, 10 / 21, -1 / 7, 2 / 77, -1 / 465),
    )

    def __init__(self, cell, ns_mesh, interpolation_nodes: int, kernel, fft_norm: str = "ortho",
                 ifft_norm: str = "ortho", mode: int = 0, differential_order: int = 2):
        if mode not in (0, 1, 2, 3):
            raise ValueError(f"`mode` should be one of [0, 1, 2, 3], but got {mode}")
        if differential_order not in (1, 2, 3, 4, 5, 6):
            raise ValueError(
                f"`differential_order` should be one between 1 and 6, but got {differential_order}"
            )
        torch.nn.Module.__init__(self)
        self.interpolation_nodes = interpolation_nodes
        self.mode = mode
        self.differential_order = differential_order
        KSpaceFilter.__init__(self, cell, ns_mesh, kernel, fft_norm, ifft_norm)

    def _p3m_nodes(self) -> int:
        return self.interpolation_nodes

    def _p3m_mode(self) -> dict:
        return dict(p3m_mode=self.mode, differential_order=self.differential_order)

    def _influence(self, kvectors):
        ns = self._ns()
        ns_t = torch.tensor(ns, dtype=self.cell.dtype, device=self.cell.device)
        h = (torch.linalg.norm(self.cell, dim=1) / ns_t).reshape(1, 1, 1, 3)
        kh = kvectors * h
        u2 = torch.prod(torch.sinc(kh / (2 * torch.pi)), dim=-1) ** (2 * self.interpolation_nodes)
        if self.mode == 0:
            return torch.where(u2 == 0, 0.0, 1.0 / torch.where(u2 == 0, 1.0, u2))
        # finite-difference operator, eq. 30 of doi:10.1063/1.3000389 (kspace_filter.py:318-347)
        d_op = torch.zeros_like(kh)
        for order, coef in enumerate(self._DIFF_COEFF[self.differential_order - 1][: self.differential_order]):
            d_op = d_op + (coef / (order + 1)) * torch.sin(kh * (order + 1))
        d_op = d_op / h
        denom = u2 * torch.linalg.norm(d_op, dim=-1) ** (4 * self.mode)
        numer = torch.sum(kvectors * d_op, dim=-1) ** self.mode
        return torch.where(denom == 0, 0.0, numer / torch.where(denom == 0, 1.0, denom))
