"""
Pair potentials understood by the B200 calculators.

Host-side mirror of the reference potential interface
(``src/torchpme/potentials/{potential,coulomb,inversepowerlaw}.py``): same class names,
constructor arguments, method names and error texts, so user code that builds
``CoulombPotential(smearing=...)`` and hands it to a calculator keeps working.

The torch implementations of the methods below are *interface* code: they serve users
who call a potential directly and the differentiable "table" route of the k-space filter
(cell gradients, custom kernels).  On the calculators' fast path none of them runs -- the
short-range kernel and the Green's function are evaluated inside the CUDA kernels from
the scalar description returned by :meth:`Potential._native_descriptor`.
"""

from __future__ import annotations

import math

import torch

from . import _native

_SQRT2 = math.sqrt(2.0)


class _Exp1(torch.autograd.Function):
    """E1(x) with the series / continued-fraction split of ``lib/math.py:16-60``."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        out = torch.full_like(x, torch.inf)
        lo = (x > 0) & (x <= 1)
        hi = x > 1
        if bool(lo.any()):
            xs = x[lo]
            term = torch.ones_like(xs)
            total = torch.ones_like(xs)
            for k in range(1, 26):
                term = -term * k * xs / (k + 1.0) ** 2
                total = total + term
                if bool(torch.all(term.abs() <= total.abs() * 1e-15)):
                    break
            out[lo] = -0.5772156649015329 - torch.log(xs) + xs * total
        if bool(hi.any()):
            xl = x[hi]
            depth = int((20 + (80.0 / xl).to(torch.int32)).max())
            levels = 20 + (80.0 / xl).to(torch.int32)
            frac = torch.zeros_like(xl)
            for k in range(depth, 0, -1):
                # elements whose own depth is smaller than k start later (frac stays 0)
                nxt = k / (1.0 + k / (xl + frac))
                frac = torch.where(levels >= k, nxt, frac)
            out[hi] = torch.exp(-xl) / (xl + frac)
        return out

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        return -grad * torch.exp(-x) / x


def exp1(x: torch.Tensor) -> torch.Tensor:
    """Exponential integral E1 for x > 0."""
    return _Exp1.apply(x)


def gamma(x: torch.Tensor) -> torch.Tensor:
    """Complete gamma function through ``lgamma`` (``lib/math.py:5-13``)."""
    return torch.exp(torch.special.gammaln(x))


def gammaincc_over_powerlaw(exponent, z: torch.Tensor) -> torch.Tensor:
    """Gamma((3-p)/2, z) / z^((3-p)/2) for p = 1..6 (``lib/math.py:79-104``)."""
    p = int(exponent)
    if p == 1:
        return torch.exp(-z) / z
    if p == 2:
        return torch.sqrt(torch.pi / z) * torch.erfc(torch.sqrt(z))
    if p == 3:
        return exp1(z)
    if p == 4:
        return 2 * (torch.exp(-z) - torch.sqrt(torch.pi * z) * torch.erfc(torch.sqrt(z)))
    if p == 5:
        return torch.exp(-z) - z * exp1(z)
    if p == 6:
        root = torch.sqrt(torch.pi * z**3)
        return ((2 - 4 * z) * torch.exp(-z) + 4 * root * torch.erfc(torch.sqrt(z))) / 3
    raise ValueError(f"Unsupported exponent: {exponent}")


def _slab_correction(periodic, positions, cell, charges):
    """2-D periodic (slab) term of the 1/r potential (``potentials/coulomb.py:6-40``)."""
    if periodic is None:
        return torch.zeros_like(charges)
    flags = periodic.to(torch.bool)
    is_slab = flags.sum() == 2
    axis = torch.argmax((~flags).to(torch.int64) * is_slab.to(torch.int64))
    z = positions.index_select(1, axis.reshape(1))
    length = torch.linalg.norm(cell, dim=-1).index_select(0, axis.reshape(1))
    volume = torch.abs(torch.linalg.det(cell))
    q_tot = charges.sum(dim=0)
    m1 = (charges * z).sum(dim=0)
    m2 = (charges * z * z).sum(dim=0)
    slab = (4.0 * torch.pi / volume) * (z * m1 - 0.5 * (m2 + q_tot * z * z) - q_tot / 12.0 * length**2)
    return torch.where(is_slab, slab, torch.zeros_like(slab))


class Potential(torch.nn.Module):
    """
    Base class: range-separated pair potential V = V_SR + V_LR with smearing ``smearing``,
    optional inner exclusion zone and a global ``prefactor``
    (reference: ``potentials/potential.py:4-212``).
    """

    def __init__(self, smearing=None, exclusion_radius=None, exclusion_degree: int = 1,
                 prefactor: float = 1.0):
        super().__init__()
        if smearing is None:
            self.smearing = None
        else:
            self.register_buffer("smearing", torch.tensor(smearing, dtype=torch.float64))
        self.exclusion_radius = exclusion_radius
        self.exclusion_degree = exclusion_degree
        self.register_buffer("prefactor", torch.tensor(prefactor, dtype=torch.float64))
        self._host_scalars = None

    # -- scalar description for the CUDA kernels ------------------------------------------
    def _scalars(self):
        """(smearing, prefactor) as Python floats, cached (reading a CUDA buffer syncs once)."""
        key = (None if self.smearing is None else (self.smearing.data_ptr(), self.smearing._version),
               self.prefactor.data_ptr(), self.prefactor._version)
        if self._host_scalars is None or self._host_scalars[0] != key:
            s = None if self.smearing is None else float(self.smearing)
            self._host_scalars = (key, s, float(self.prefactor))
        return self._host_scalars[1], self._host_scalars[2]

    def _native_descriptor(self):
        """``None`` (generic torch route) or ``(green_kind, exponent)`` for in-kernel evaluation."""
        return None

    def _native_filter(self):
        """
        ``None`` (the filter is a table built from :meth:`lr_from_k_sq`) or the keyword arguments that
        describe ``G(k^2)`` to the filter kernel (``_native.make_green``: kind, exponent, smearing, prefactor
        [, table]).  Defaults to the in-kernel potential of :meth:`_native_descriptor`.
        """
        descriptor = self._native_descriptor()
        if descriptor is None:
            return None
        smearing, prefactor = self._scalars()
        return dict(kind=descriptor[0], exponent=descriptor[1], smearing=smearing, prefactor=prefactor)

    # -- interface -------------------------------------------------------------------------
    def f_cutoff(self, dist, pair_mask=None):
        if self.exclusion_radius is None:
            raise ValueError("Cannot compute cutoff function when `exclusion_radius` is not set")
        bump = (0.5 * (1 - torch.cos(torch.pi * (dist / self.exclusion_radius)))) ** self.exclusion_degree
        out = torch.where(dist < self.exclusion_radius, 1 - bump, 0.0)
        return out if pair_mask is None else out * pair_mask

    def from_dist(self, dist, pair_mask=None):
        raise NotImplementedError(f"from_dist is not implemented for {self.__class__.__name__}")

    def lr_from_dist(self, dist, pair_mask=None):
        raise NotImplementedError(f"lr_from_dist is not implemented for {self.__class__.__name__}")

    def lr_from_k_sq(self, k_sq):
        raise NotImplementedError(f"lr_from_k_sq is not implemented for {self.__class__.__name__}")

    def sr_from_dist(self, dist, pair_mask=None):
        if self.smearing is None:
            raise ValueError(
                "Cannot compute range-separated potential when `smearing` is not specified."
            )
        long_range = self.lr_from_dist(dist, pair_mask=pair_mask)
        if self.exclusion_radius is None:
            return self.from_dist(dist, pair_mask=pair_mask) - long_range
        return -long_range * self.f_cutoff(dist, pair_mask=pair_mask)

    def kernel_from_k_sq(self, k_sq):
        return self.lr_from_k_sq(k_sq)

    def self_contribution(self):
        raise NotImplementedError(f"self_contribution is not implemented for {self.__class__.__name__}")

    def background_correction(self):
        raise NotImplementedError(
            f"background_correction is not implemented for {self.__class__.__name__}"
        )

    def pbc_correction(self, periodic, positions, cell, charges):
        return self.prefactor * torch.zeros_like(charges)

    def _need_smearing(self, what: str):
        if self.smearing is None:
            raise ValueError(f"Cannot compute {what} without specifying `smearing`.")


def _masked(values, pair_mask):
    return values if pair_mask is None else values * pair_mask


class CoulombPotential(Potential):
    """Smoothed 1/r potential (reference: ``potentials/coulomb.py:43-171``)."""

    def _native_descriptor(self):
        return (_native.GREEN_COULOMB, 1) if type(self) is CoulombPotential else None

    def from_dist(self, dist, pair_mask=None):
        return self.prefactor * _masked(1.0 / dist.clamp(min=1e-15), pair_mask)

    def lr_from_dist(self, dist, pair_mask=None):
        self._need_smearing("long-range contribution")
        val = torch.erf(dist / self.smearing / _SQRT2) / dist.clamp(min=1e-12)
        return self.prefactor * _masked(val, pair_mask)

    def lr_from_k_sq(self, k_sq):
        self._need_smearing("long-range kernel")
        at_zero = k_sq == 0
        safe = torch.where(at_zero, 1.0, k_sq)  # keeps the backward NaN free
        val = 4 * torch.pi * torch.exp(-0.5 * self.smearing**2 * safe) / safe
        return self.prefactor * torch.where(at_zero, 0.0, val)

    def self_contribution(self):
        self._need_smearing("self contribution")
        return self.prefactor * math.sqrt(2 / math.pi) / self.smearing

    def background_correction(self):
        self._need_smearing("background correction")
        return self.prefactor * torch.pi * self.smearing**2

    def pbc_correction(self, periodic, positions, cell, charges):
        return self.prefactor * _slab_correction(periodic, positions, cell, charges)


class InversePowerLawPotential(Potential):
    """1/r^p potentials, p = 1..6 (reference: ``potentials/inversepowerlaw.py:10-173``)."""

    def __init__(self, exponent: int, smearing=None, exclusion_radius=None,
                 exclusion_degree: int = 1, prefactor: float = 1.0):
        super().__init__(smearing, exclusion_radius, exclusion_degree, prefactor)
        gammaincc_over_powerlaw(exponent, torch.tensor(1.0))  # validates the exponent
        self.register_buffer("exponent", torch.tensor(exponent, dtype=torch.float64))
        self._p = int(exponent)

    def _native_descriptor(self):
        return (_native.GREEN_IPL, self._p) if type(self) is InversePowerLawPotential else None

    def from_dist(self, dist, pair_mask=None):
        return self.prefactor * _masked(torch.pow(dist.clamp(min=1e-15), -self.exponent), pair_mask)

    def lr_from_dist(self, dist, pair_mask=None):
        self._need_smearing("long-range contribution")
        half_p = self.exponent / 2
        x = (0.5 * dist**2 / self.smearing**2).clamp(min=1e-15)
        scale = 1.0 / (2 * self.smearing**2) ** half_p
        val = scale * torch.special.gammainc(half_p, x) / x**half_p
        return self.prefactor * _masked(val, pair_mask)

    def lr_from_k_sq(self, k_sq):
        self._need_smearing("long-range kernel")
        p_eff = (3 - self.exponent) / 2
        scale = torch.pi**1.5 / gamma(self.exponent / 2) * (2 * self.smearing**2) ** p_eff
        z = 0.5 * self.smearing**2 * k_sq
        safe = torch.where(z == 0, 1.0, z)
        # k = 0: divergent for p <= 3 (dropped: neutralising background), finite for p > 3
        at_zero = -scale / p_eff if self._p > 3 else 0.0
        body = scale * gammaincc_over_powerlaw(self._p, safe)
        return self.prefactor * torch.where(k_sq == 0, at_zero, body)

    def self_contribution(self):
        self._need_smearing("self contribution")
        half_p = self.exponent / 2
        return self.prefactor / gamma(half_p + 1) / (2 * self.smearing**2) ** half_p

    def background_correction(self):
        self._need_smearing("background correction")
        if self._p >= 3:
            return torch.zeros_like(self.smearing)
        num = torch.pi**1.5 * (2 * self.smearing**2) ** ((3 - self.exponent) / 2)
        return self.prefactor * num / ((3 - self.exponent) * gamma(self.exponent / 2))

    def pbc_correction(self, periodic, positions, cell, charges):
        if self._p == 1:
            return self.prefactor * _slab_correction(periodic, positions, cell, charges)
        return super().pbc_correction(periodic, positions, cell, charges)


class SplinePotential(Potential):
    """
    Tabulated potential: a cubic spline of ``(r_grid, y_grid)`` in real space and of
    ``(k_grid^2, yhat_grid)`` in reciprocal space, the latter computed from the former when not
    given (reference: ``potentials/spline.py:12-169``).  Purely long-ranged by default
    (``sr_from_dist`` is zero); on the calculators it is served by the generic routes (per-pair
    values and a filter table built from these torch functions).
    """

    def __init__(self, r_grid: torch.Tensor, y_grid: torch.Tensor, k_grid=None, yhat_grid=None,
                 reciprocal: bool = False, y_at_zero=None, yhat_at_zero=None, smearing=None,
                 exclusion_radius=None, exclusion_degree: int = 1, prefactor: float = 1.0):
        super().__init__(smearing, exclusion_radius, exclusion_degree, prefactor)
        from .splines import CubicSpline, CubicSplineReciprocal, compute_second_derivatives, compute_spline_ft

        if len(y_grid) != len(r_grid):
            raise ValueError("Length of radial grid and value array mismatch.")
        if reciprocal and torch.min(r_grid) <= 0.0:
            raise ValueError("Positive-valued radial grid is needed for reciprocal axis spline.")
        self.register_buffer("r_grid", r_grid)
        self.register_buffer("y_grid", y_grid)
        self._reciprocal = bool(reciprocal)
        if k_grid is None:
            # 2 pi / r (ascending) on a reciprocal axis, the radial grid itself otherwise
            k_grid = 2 * torch.pi / r_grid.flip(0) if reciprocal else r_grid.detach().clone()
        self.register_buffer("k_grid", k_grid)
        if yhat_grid is None:
            yhat_grid = compute_spline_ft(k_grid, r_grid, y_grid, compute_second_derivatives(r_grid, y_grid))
        self.register_buffer("yhat_grid", yhat_grid)
        self._y_at_zero_arg, self._yhat_at_zero_arg = y_at_zero, yhat_at_zero
        self._build = (CubicSpline, CubicSplineReciprocal)
        self._splines_for = None

    def _splines(self):
        """the spline objects, rebuilt when the buffers moved (``.to(device / dtype)``)"""
        key = (self.r_grid.device, self.r_grid.dtype, self.r_grid.data_ptr(), self.yhat_grid.data_ptr())
        if self._splines_for != key:
            plain, recip = self._build
            if self._reciprocal:
                self._spline = recip(self.r_grid, self.y_grid, y_at_zero=self._y_at_zero_arg)
                self._krn_spline = recip(self.k_grid**2, self.yhat_grid, y_at_zero=self._yhat_at_zero_arg)
            else:
                self._spline = plain(self.r_grid, self.y_grid)
                self._krn_spline = plain(self.k_grid**2, self.yhat_grid)
            zero = torch.zeros(1, dtype=self.r_grid.dtype, device=self.r_grid.device)
            self._y_at_zero = self._spline(zero) if self._y_at_zero_arg is None else \
                torch.as_tensor(self._y_at_zero_arg, dtype=zero.dtype, device=zero.device)
            self._yhat_at_zero = self._krn_spline(zero) if self._yhat_at_zero_arg is None else \
                torch.as_tensor(self._yhat_at_zero_arg, dtype=zero.dtype, device=zero.device)
            self._splines_for = key
        return self._spline, self._krn_spline

    def _native_filter(self):
        """G(k^2) = prefactor * spline(k^2) evaluated per k-point inside the filter kernel (green kinds 3 / 4):
        the knots, values and second derivatives travel as one small float64 table"""
        _, krn = self._splines()
        if not self.k_grid.is_cuda:
            return None
        key = self._splines_for
        if getattr(self, "_filter_table_for", None) != key:
            if self._reciprocal:
                inv, head = krn._inverse_axis, krn._zero_spline
                parts = [inv.x_points, inv.y_points, inv.d2y_points, head.x_points, head.y_points, head.d2y_points]
                n = inv.x_points.shape[0]
            else:
                parts = [krn.x_points, krn.y_points, krn.d2y_points]
                n = krn.x_points.shape[0]
            self._filter_table = torch.cat([p.detach().to(torch.float64).reshape(-1) for p in parts]).contiguous()
            self._filter_knots = int(n)
            self._filter_table_for = key
        return dict(kind=4 if self._reciprocal else 3, exponent=self._filter_knots, smearing=1.0,
                    prefactor=self._scalars()[1], table=self._filter_table)

    def from_dist(self, dist, pair_mask=None):
        # as in the reference: the prefactor multiplies the (already scaled) long-range part again
        return self.prefactor * (self.lr_from_dist(dist, pair_mask) + self.sr_from_dist(dist, pair_mask))

    def sr_from_dist(self, dist, pair_mask=None):
        return 0.0 * dist

    def lr_from_dist(self, dist, pair_mask=None):
        return self.prefactor * self._splines()[0](dist)

    def lr_from_k_sq(self, k_sq):
        return self.prefactor * self._splines()[1](k_sq)

    def self_contribution(self):
        self._splines()
        return self.prefactor * self._y_at_zero

    def background_correction(self):
        # (1,) tensor of the default dtype, like the reference's ``prefactor * torch.zeros(1)``
        return self.prefactor * torch.zeros(1, device=self.prefactor.device)


class CombinedPotential(Potential):
    """
    Weighted sum of potentials, with fixed or trainable weights (reference:
    ``potentials/combined.py:6-124``).  All members must be range separated (then ``smearing`` is
    required) or all direct (then it must be omitted).
    """

    def __init__(self, potentials, initial_weights=None, learnable_weights: bool = True, smearing=None,
                 exclusion_radius=None, exclusion_degree: int = 1):
        super().__init__(smearing=smearing, exclusion_radius=exclusion_radius, exclusion_degree=exclusion_degree)
        has_smearing = [pot.smearing is not None and bool(pot.smearing) for pot in potentials]
        if any(has_smearing) and not all(has_smearing):
            raise ValueError(
                "Cannot combine direct (`smearing=None`) and range-separated (`smearing=float`) potentials.")
        if all(has_smearing) and not self.smearing:
            raise ValueError(
                "You should specify a `smearing` when combining range-separated (`smearing=float`) potentials.")
        if not any(has_smearing) and self.smearing:
            raise ValueError("Cannot specify `smearing` when combining direct (`smearing=None`) potentials.")
        if initial_weights is None:
            initial_weights = torch.ones(len(potentials))
        elif len(initial_weights) != len(potentials):
            raise ValueError("The number of initial weights must match the number of potentials being combined")
        self.potentials = torch.nn.ModuleList(potentials)
        if learnable_weights:
            self.weights = torch.nn.Parameter(initial_weights)
        else:
            self.register_buffer("weights", initial_weights)

    def _mix(self, parts):
        stacked = torch.stack([p.to(self.weights.dtype) if p.dtype != self.weights.dtype else p for p in parts], dim=-1)
        return torch.inner(self.weights.to(stacked.dtype), stacked)

    def from_dist(self, dist, pair_mask=None):
        return self._mix([pot.from_dist(dist, pair_mask) for pot in self.potentials])

    def sr_from_dist(self, dist, pair_mask=None):
        return self._mix([pot.sr_from_dist(dist, pair_mask) for pot in self.potentials])

    def lr_from_dist(self, dist, pair_mask=None):
        return self._mix([pot.lr_from_dist(dist, pair_mask) for pot in self.potentials])

    def lr_from_k_sq(self, k_sq):
        return self._mix([pot.lr_from_k_sq(k_sq) for pot in self.potentials])

    def self_contribution(self):
        return self._mix([pot.self_contribution().reshape(()) for pot in self.potentials])

    def background_correction(self):
        return self._mix([pot.background_correction().reshape(()) for pot in self.potentials])
