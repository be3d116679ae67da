"""
Cubic splines for tabulated potentials (host-side interface code, plain torch ops).

Mirror of the reference's ``lib/splines.py`` interface -- ``CubicSpline``,
``CubicSplineReciprocal``, ``compute_second_derivatives``, ``compute_spline_ft`` -- written from the
textbook definitions:

* natural cubic spline: second derivatives ``M`` from the tridiagonal continuity system
  ``h_{i-1}/6 M_{i-1} + (h_{i-1}+h_i)/3 M_i + h_i/6 M_{i+1} = dy_i/h_i - dy_{i-1}/h_{i-1}``,
  ``M_0 = M_{n-1} = 0``; evaluation ``A y_i + B y_{i+1} + ((A^3-A) M_i + (B^3-B) M_{i+1}) h^2/6``;
* radial Fourier transform ``f^(k) = 4 pi int sin(kr)/k r f(r) dr`` of the spline: on every interval
  ``r f(r)`` is a quartic ``P``, integrated exactly by repeated integration by parts
  (``int P sin = -P cos/k + P' sin/k^2 + P'' cos/k^3 - P''' sin/k^4 - P'''' cos/k^5``), with a
  Gauss-Legendre rule where ``k dr`` is small (the closed form cancels catastrophically there);
  beyond the last point the function continues as the natural spline in ``1/r`` through
  ``(0, 0), (1/r_n, y_n), (1/r_{n-1}, y_{n-1})``, i.e. ``c1/r + c3/r^3``, whose transform is
  ``4 pi/k [c1 cos(k r_n)/k + c3 (sin(k r_n)/r_n - k Ci(k r_n))]``.

These run where the potential's tensors live; on the calculators' path they feed the generic
routes (per-pair values / filter table), not hand-written kernels.
"""

from __future__ import annotations

import numpy as np
import torch


def compute_second_derivatives(x_points: torch.Tensor, y_points: torch.Tensor) -> torch.Tensor:
    """Second derivatives of the natural cubic spline through ``(x_points, y_points)``."""
    n = x_points.shape[0]
    if n < 3:
        return torch.zeros_like(y_points)
    h = x_points[1:] - x_points[:-1]
    slope = (y_points[1:] - y_points[:-1]) / h
    # interior equations as a dense tridiagonal system (grids have at most a few thousand points;
    # keeps the solve differentiable in both x and y)
    main = (h[:-1] + h[1:]) / 3
    system = torch.diag(main) + torch.diag(h[1:-1] / 6, 1) + torch.diag(h[1:-1] / 6, -1)
    rhs = slope[1:] - slope[:-1]
    inner = torch.linalg.solve(system, rhs)
    zero = torch.zeros(1, dtype=y_points.dtype, device=y_points.device)
    return torch.cat([zero, inner, zero])


class CubicSpline(torch.nn.Module):
    """Natural cubic spline of a real function; outside the grid the end intervals are continued."""

    def __init__(self, x_points: torch.Tensor, y_points: torch.Tensor):
        super().__init__()
        self.x_points = x_points
        self.y_points = y_points
        self.d2y_points = compute_second_derivatives(x_points, y_points)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xs, ys, ms = self.x_points, self.y_points, self.d2y_points
        lo = (torch.searchsorted(xs, x, right=True) - 1).clamp(0, xs.shape[0] - 2)
        h = xs[lo + 1] - xs[lo]
        a = (xs[lo + 1] - x) / h
        b = 1 - a
        curvature = (a * (a * a - 1)) * ms[lo] + (b * (b * b - 1)) * ms[lo + 1]
        return a * ys[lo] + b * ys[lo + 1] + curvature * (h * h / 6)


class CubicSplineReciprocal(torch.nn.Module):
    """
    Spline on a ``1/x`` axis that decays smoothly to zero for ``x -> infinity`` (the point
    ``(1/x = 0, y = 0)`` is added); below the first grid point a small direct spline through
    ``(0, y_at_zero), (x_0, y_0), (x_1, y_1)`` takes over.  ``x_points`` must be positive.
    """

    def __init__(self, x_points: torch.Tensor, y_points: torch.Tensor, y_at_zero=None):
        super().__init__()
        zero = torch.zeros(1, dtype=x_points.dtype, device=x_points.device)
        self._inverse_axis = CubicSpline(torch.cat([zero, 1.0 / x_points.flip(0)]),
                                         torch.cat([zero.to(y_points.dtype), y_points.flip(0)]))
        if y_at_zero is None:
            y_at_zero = y_points[0]
        self._y_at_zero = y_at_zero
        head_x = torch.stack([zero[0], x_points[0], x_points[1]])
        head_y = torch.stack([torch.as_tensor(y_at_zero, dtype=y_points.dtype, device=y_points.device),
                              y_points[0], y_points[1]])
        self._zero_spline = CubicSpline(head_x, head_y)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        first = self._zero_spline.x_points[1]
        below = x < first
        safe = torch.where(below, first, x)
        return torch.where(below, self._zero_spline(x), self._inverse_axis(1.0 / safe))


_GL_NODES, _GL_WEIGHTS = np.polynomial.legendre.leggauss(24)


def compute_spline_ft(k_points: torch.Tensor, x_points: torch.Tensor, y_points: torch.Tensor,
                      d2y_points: torch.Tensor) -> torch.Tensor:
    """
    Radial Fourier transform ``4 pi int sin(kr)/k r f(r) dr`` of the cubic spline
    ``(x_points, y_points, d2y_points)`` at ``k_points``, tail beyond the last point included (see the
    module docstring).  Evaluated once at construction time, in float64 numpy (needs ``scipy`` for
    the cosine integral) whatever the dtype of the grids -- the reference evaluates in the grids'
    dtype and loses accuracy for float32 grids (its tests/lib/test_splines.py:60-95 expects that).
    """
    try:
        from scipy.special import sici
    except ImportError as err:  # pragma: no cover
        raise ImportError("Computing the Fourier-domain kernel based on a spline requires scipy") from err
    k = k_points.detach().cpu().numpy().astype(np.float64).reshape(-1)
    r = x_points.detach().cpu().numpy().astype(np.float64)
    y = y_points.detach().cpu().numpy().astype(np.float64)
    m = d2y_points.detach().cpu().numpy().astype(np.float64)
    r0, r1 = r[:-1], r[1:]
    h = r1 - r0
    # cubic on [r0, r1] in powers of t = r - r0:  y0 + c1 t + c2 t^2 + c3 t^3
    c1 = (y[1:] - y[:-1]) / h - h * (2 * m[:-1] + m[1:]) / 6
    c2 = m[:-1] / 2
    c3 = (m[1:] - m[:-1]) / (6 * h)
    # quartic P(t) = (r0 + t) * cubic(t), coefficients p0..p4
    p = np.stack([r0 * y[:-1], y[:-1] + r0 * c1, c1 + r0 * c2, c2 + r0 * c3, c3])          # (5, n-1)

    def poly_derivs(t):
        """P, P', P'', P''', P'''' at t (per interval)"""
        return (p[0] + t * (p[1] + t * (p[2] + t * (p[3] + t * p[4]))),
                p[1] + t * (2 * p[2] + t * (3 * p[3] + t * 4 * p[4])),
                2 * p[2] + t * (6 * p[3] + t * 12 * p[4]),
                6 * p[3] + t * 24 * p[4],
                24 * p[4] + 0 * t)

    nodes = 0.5 * (_GL_NODES + 1)                              # Gauss-Legendre on [0, 1]
    out = np.zeros_like(k)
    # tail: natural spline in u = 1/r through (0,0), (1/r_n, y_n), (1/r_{n-1}, y_{n-1})
    u0, u1 = 1.0 / r[-1], 1.0 / r[-2]
    slope0, slope1 = y[-1] / u0, (y[-2] - y[-1]) / (u1 - u0)
    m_tail = (slope1 - slope0) / (u1 / 3)                      # single interior equation, M(0) = M(u1) = 0
    t1 = y[-1] * r[-1] - m_tail / (6 * r[-1])                  # coefficient of 1/r
    t3 = m_tail * r[-1] / 6                                    # coefficient of 1/r^3
    for idx, kk in enumerate(k):
        if kk == 0.0:
            # int r^2 f(r) dr over the splined range (the tail's k -> 0 limit diverges like the
            # potential itself and is dropped, as in the reference)
            tt = h[None, :] * nodes[:, None]
            vals = p[0] + tt * (p[1] + tt * (p[2] + tt * (p[3] + tt * p[4])))          # r f(r)
            integral = (0.5 * _GL_WEIGHTS[:, None] * vals * (r0[None, :] + tt)).sum(0).dot(h)
            out[idx] = 4 * np.pi * integral
            continue
        small = kk * h < 0.5
        total = 0.0
        if np.any(~small):
            sel = ~small
            acc = 0.0
            for t, sign in ((h, 1.0), (np.zeros_like(h), -1.0)):
                d0, d1, d2, d3, d4 = poly_derivs(t)
                arg = kk * (r0 + t)
                s, c = np.sin(arg), np.cos(arg)
                anti = -d0 * c / kk + d1 * s / kk**2 + d2 * c / kk**3 - d3 * s / kk**4 - d4 * c / kk**5
                acc = acc + sign * anti
            total += acc[sel].sum()
        if np.any(small):
            sel = small
            tt = h[sel][None, :] * nodes[:, None]                                 # (q, n_small)
            ps = p[:, sel]
            vals = ps[0] + tt * (ps[1] + tt * (ps[2] + tt * (ps[3] + tt * ps[4])))
            integrand = vals * np.sin(kk * (r0[sel][None, :] + tt))
            total += (0.5 * _GL_WEIGHTS[:, None] * integrand).sum(0).dot(h[sel])
        ci = sici(kk * r[-1])[1]
        tail = t1 * np.cos(kk * r[-1]) / kk + t3 * (np.sin(kk * r[-1]) / r[-1] - kk * ci)
        out[idx] = 4 * np.pi / kk * (total + tail)
    return torch.as_tensor(out.reshape(tuple(k_points.shape)), dtype=k_points.dtype, device=k_points.device)
