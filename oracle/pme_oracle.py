"""
CPU oracle for the PME / P3M hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a numpy restatement of the algorithm that the reference
(lab-cosmo/torch-pme @ e29fa56) runs for ``PMECalculator/P3MCalculator.forward``.
It exists only so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` have an independent
checker.  Nothing under ``torch-pme_b200/`` may import it.

Parity status: PINNED.  ``tests/test_oracle_pinning.py`` checks this file against
(i) the reference's own known answers (Madelung constants, ``tests/helpers.py:19-139``;
GROMACS SPME energies/forces, ``examples/coulomb_test_frames.xyz``;
closed forms of ``tests/test_potentials.py:86-182``) and (ii) golden tensors generated
by importing the reference itself in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

Every function cites the reference lines (relative to ``/root/reference/src/torchpme``)
whose arithmetic it restates.  Conventions: ``cell[i]`` is the i-th lattice vector,
``charges`` is ``(N, C)``, meshes are ``(C, nx, ny, nz)``, k-space meshes are
``(C, nx, ny, nz//2+1)``.
"""

from __future__ import annotations

import math

import numpy as np

import scipy.fft as _sfft  # multi-threaded pocketfft (workers=-1)
import scipy.special as _sp  # erf / erfc / regularised incomplete gamma

# --------------------------------------------------------------------------------------
# interpolation weight polynomials
# --------------------------------------------------------------------------------------
# Integer numerators of the 1-D weight polynomials in ascending powers of x, x in
# [-1/2, 1/2], and their common denominators.
#   P3M      : lib/mesh_interpolator.py:171-209
#   Lagrange : lib/mesh_interpolator.py:228-300
_WEIGHT_TABLE = {
    ("P3M", 1): (1, [[1]]),
    ("P3M", 2): (2, [[1, -2], [1, 2]]),
    ("P3M", 3): (8, [[1, -4, 4], [6, 0, -8], [1, 4, 4]]),
    ("P3M", 4): (
        48,
        [[1, -6, 12, -8], [23, -30, -12, 24], [23, 30, -12, -24], [1, 6, 12, 8]],
    ),
    ("P3M", 5): (
        384,
        [
            [1, -8, 24, -32, 16],
            [76, -176, 96, 64, -64],
            [230, 0, -240, 0, 96],
            [76, 176, 96, -64, -64],
            [1, 8, 24, 32, 16],
        ],
    ),
    ("Lagrange", 3): (2, [[0, -1, 1], [2, 0, -2], [0, 1, 1]]),
    ("Lagrange", 4): (
        48,
        [[-3, 2, 12, -8], [27, -54, -12, 24], [27, 54, -12, -24], [-3, -2, 12, 8]],
    ),
    ("Lagrange", 5): (
        24,
        [
            [0, 2, -1, -2, 1],
            [0, -16, 16, 4, -4],
            [24, 0, -30, 0, 6],
            [0, 16, 16, -4, -4],
            [0, -2, -1, 2, 1],
        ],
    ),
    ("Lagrange", 6): (
        3840,
        [
            [45, -18, -200, 80, 80, -32],
            [-375, 250, 1560, -1040, -240, 160],
            [2250, -4500, -1360, 2720, 160, -320],
            [2250, 4500, -1360, -2720, 160, 320],
            [-375, -250, 1560, 1040, -240, -160],
            [45, 18, -200, -80, 80, 32],
        ],
    ),
    ("Lagrange", 7): (
        720,
        [
            [0, -12, 4, 15, -5, -3, 1],
            [0, 108, -54, -120, 60, 12, -6],
            [0, -540, 540, 195, -195, -15, 15],
            [720, 0, -980, 0, 280, 0, -20],
            [0, 540, 540, -195, -195, 15, 15],
            [0, -108, -54, 120, 60, -12, -6],
            [0, 12, 4, -15, -5, 3, 1],
        ],
    ),
}


def weight_coefficients(method: str, nodes: int) -> np.ndarray:
    """(nodes, nodes) float64 polynomial coefficients, ascending powers of x."""
    if (method, nodes) not in _WEIGHT_TABLE:
        raise ValueError(f"unsupported interpolation ({method}, {nodes})")
    den, num = _WEIGHT_TABLE[(method, nodes)]
    return np.asarray(num, dtype=np.float64) / den


def weights_1d(x: np.ndarray, nodes: int, method: str, deriv: bool = False):
    """
    1-D interpolation weights w_a(x), a = 0..nodes-1 (lib/mesh_interpolator.py:149-301),
    or their derivative dw_a/dx.  ``x`` has any shape; result has shape (nodes, *x.shape).
    """
    coef = weight_coefficients(method, nodes).astype(x.dtype)
    out = np.zeros((nodes,) + x.shape, dtype=x.dtype)
    for a in range(nodes):
        c = coef[a]
        if deriv:
            c = c[1:] * np.arange(1, nodes, dtype=x.dtype)
        acc = np.zeros_like(x)
        for ck in c[::-1]:  # Horner
            acc = acc * x + ck
        out[a] = acc
    return out


def mesh_coordinates(positions: np.ndarray, cell: np.ndarray, ns) -> np.ndarray:
    """u = ns * (positions @ cell^-1)   (lib/mesh_interpolator.py:104-112,326)."""
    inv = np.linalg.inv(cell)
    return np.asarray(ns, dtype=positions.dtype) * (positions @ inv)


def stencil(u: np.ndarray, nodes: int):
    """
    Base index and offset of every point (lib/mesh_interpolator.py:329-341): even node
    counts use floor and the offset from the cell midpoint, odd ones round-half-to-even
    (``torch.round``) and the offset from the nearest mesh point.  Returns
    ``(i0 (N,3) int64, x (N,3))``; node ``a`` lives at index ``i0 + a + 1 - (nodes+1)//2``
    (lib/mesh_interpolator.py:350-359), wrapped with a non-negative modulo.
    """
    if nodes % 2 == 0:
        i0 = np.floor(u)
        x = u - (i0 + 0.5)
    else:
        i0 = np.rint(u)
        x = u - i0
    return i0.astype(np.int64), x.astype(u.dtype)


def _node_indices(i0: np.ndarray, nodes: int, ns) -> np.ndarray:
    """(nodes, N, 3) wrapped mesh indices of the stencil."""
    first = 1 - (nodes + 1) // 2
    ns = np.asarray(ns, dtype=np.int64)
    return np.stack([(i0 + (first + a)) % ns for a in range(nodes)], axis=0)


def points_to_mesh(
    weights: np.ndarray, positions: np.ndarray, cell: np.ndarray, ns, nodes: int, method: str
) -> np.ndarray:
    """
    Charge assignment rho_c[m] = sum_i q_ic wx wy wz  (lib/mesh_interpolator.py:379-426).
    """
    nx, ny, nz = (int(v) for v in ns)
    u = mesh_coordinates(positions, cell, ns)
    i0, x = stencil(u, nodes)
    w = weights_1d(x, nodes, method)  # (n, N, 3)
    idx = _node_indices(i0, nodes, ns)  # (n, N, 3)
    n_ch = weights.shape[1]
    mesh = np.zeros((n_ch, nx * ny * nz), dtype=positions.dtype)
    for a in range(nodes):
        # flat index / weight of all (b, c) partners of x-node a
        flat = (
            (idx[a, None, None, :, 0] * ny + idx[None, :, None, :, 1]) * nz
            + idx[None, None, :, :, 2]
        ).reshape(-1)
        wprod = (
            w[a, None, None, :, 0] * w[None, :, None, :, 1] * w[None, None, :, :, 2]
        )  # (1->n, n, N) broadcast to (n, n, N) with the leading axis = b
        wprod = wprod.reshape(nodes * nodes, -1)
        for ch in range(n_ch):
            vals = (wprod * weights[None, :, ch]).reshape(-1)
            mesh[ch] += np.bincount(flat, weights=vals, minlength=nx * ny * nz).astype(
                positions.dtype
            )
    return mesh.reshape(n_ch, nx, ny, nz)


def mesh_to_points(
    mesh: np.ndarray,
    positions: np.ndarray,
    cell: np.ndarray,
    nodes: int,
    method: str,
    gradient: bool = False,
):
    """
    Back-interpolation V_ic = sum_m mesh_c[m] wx wy wz  (lib/mesh_interpolator.py:428-457).

    With ``gradient=True`` additionally returns dV_ic/dr_i, shape (N, C, 3): the quantity
    PyTorch's tape produces by differentiating the weight polynomials through
    ``u = ns * r @ cell^-1`` (the integer base index is not differentiable,
    lib/mesh_interpolator.py:334,340).
    """
    n_ch, nx, ny, nz = mesh.shape
    ns = (nx, ny, nz)
    u = mesh_coordinates(positions, cell, ns)
    i0, x = stencil(u, nodes)
    w = weights_1d(x, nodes, method)
    idx = _node_indices(i0, nodes, ns)
    n_pts = positions.shape[0]
    out = np.zeros((n_pts, n_ch), dtype=mesh.dtype)
    if gradient:
        dw = weights_1d(x, nodes, method, deriv=True)
        dout_du = np.zeros((n_pts, n_ch, 3), dtype=mesh.dtype)
    flat_mesh = mesh.reshape(n_ch, -1)
    for a in range(nodes):
        for b in range(nodes):
            base = (idx[a, :, 0] * ny + idx[b, :, 1]) * nz
            for c in range(nodes):
                vals = flat_mesh[:, base + idx[c, :, 2]].T  # (N, C)
                wxy = w[a, :, 0] * w[b, :, 1]
                out += vals * (wxy * w[c, :, 2])[:, None]
                if gradient:
                    dout_du[:, :, 0] += vals * (dw[a, :, 0] * w[b, :, 1] * w[c, :, 2])[:, None]
                    dout_du[:, :, 1] += vals * (w[a, :, 0] * dw[b, :, 1] * w[c, :, 2])[:, None]
                    dout_du[:, :, 2] += vals * (wxy * dw[c, :, 2])[:, None]
    if not gradient:
        return out
    # du_a/dr_b = ns_a * inv[b, a]
    jac = np.linalg.inv(cell) * np.asarray(ns, dtype=mesh.dtype)[None, :]  # (b, a)
    dout_dr = np.einsum("ica,ba->icb", dout_du, jac)
    return out, dout_dr


# --------------------------------------------------------------------------------------
# mesh size, k-vectors, Green's functions
# --------------------------------------------------------------------------------------
def get_ns_mesh(cell: np.ndarray, mesh_spacing: float) -> np.ndarray:
    """ns_a = 2^ceil(log2(2|cell_a|/h + 1))   (lib/kvectors.py:5-21)."""
    norms = np.linalg.norm(cell, axis=1)
    return (2 ** np.ceil(np.log2(2.0 * norms / mesh_spacing + 1.0))).astype(np.int64)


def kvectors_for_mesh(cell: np.ndarray, ns) -> np.ndarray:
    """
    Reciprocal vectors in rFFT layout, (nx, ny, nz//2+1, 3)  (lib/kvectors.py:24-102):
    k = fx*B[0] + fy*B[1] + fz*B[2], B = 2 pi (cell^-1)^T, f = fftfreq*n / rfftfreq*n.
    """
    nx, ny, nz = (int(v) for v in ns)
    recip = 2.0 * np.pi * np.linalg.inv(cell).T
    fx = np.fft.fftfreq(nx) * nx
    fy = np.fft.fftfreq(ny) * ny
    fz = np.fft.rfftfreq(nz) * nz
    k = (
        fx[:, None, None, None] * recip[0]
        + fy[None, :, None, None] * recip[1]
        + fz[None, None, :, None] * recip[2]
    )
    return k.astype(cell.dtype)


_EULER = 0.577215664901532860606512090082402431


def exp1(x: np.ndarray) -> np.ndarray:
    """
    Exponential integral E1 (lib/math.py:16-60): power series for 0 < x <= 1 (<= 25
    terms), continued fraction with m = 20 + floor(80/x) levels for x > 1, inf at 0.
    """
    x = np.asarray(x)
    out = np.full(x.shape, np.inf, dtype=x.dtype)
    small = (x > 0) & (x <= 1)
    if small.any():
        xs = x[small]
        e1 = np.ones_like(xs)
        r = np.ones_like(xs)
        for k in range(1, 26):
            r = -r * k * xs / (k + 1.0) ** 2
            e1 = e1 + r
            if np.all(np.abs(r) <= np.abs(e1) * 1e-15):
                break
        out[small] = -_EULER - np.log(xs) + xs * e1
    large = x > 1
    if large.any():
        xl = x[large]
        m = 20 + (80.0 / xl).astype(np.int32)
        t0 = np.zeros_like(xl)
        for k in range(int(m.max()), 0, -1):
            t0 = k / (1.0 + k / (xl + t0))
        out[large] = np.exp(-xl) / (xl + t0)
    return out


def _erfc(x: np.ndarray) -> np.ndarray:
    return _sp.erfc(x)


def gammaincc_over_powerlaw(exponent: int, z: np.ndarray) -> np.ndarray:
    """Gamma((3-p)/2, z)/z^((3-p)/2) closed forms for p = 1..6 (lib/math.py:79-104)."""
    if exponent == 1:
        return np.exp(-z) / z
    if exponent == 2:
        return np.sqrt(np.pi / z) * _erfc(np.sqrt(z))
    if exponent == 3:
        return exp1(z)
    if exponent == 4:
        return 2 * (np.exp(-z) - np.sqrt(np.pi * z) * _erfc(np.sqrt(z)))
    if exponent == 5:
        return np.exp(-z) - z * exp1(z)
    if exponent == 6:
        return ((2 - 4 * z) * np.exp(-z) + 4 * np.sqrt(np.pi * z**3) * _erfc(np.sqrt(z))) / 3
    raise ValueError(f"Unsupported exponent: {exponent}")


class PotentialSpec:
    """
    Plain description of the two potentials on the path.

    ``kind='coulomb'``  -> potentials/coulomb.py:43-171
    ``kind='ipl'``      -> potentials/inversepowerlaw.py:10-173  (1/r^p, p = 1..6)
    """

    def __init__(
        self,
        kind: str,
        smearing: float,
        exponent: int = 1,
        prefactor: float = 1.0,
        exclusion_radius: float | None = None,
        exclusion_degree: int = 1,
    ):
        if kind not in ("coulomb", "ipl"):
            raise ValueError(kind)
        self.kind = kind
        # None: no range separation (direct calculator, calculators/calculator.py:53-61)
        self.smearing = None if smearing is None else float(smearing)
        self.exponent = int(exponent) if kind == "ipl" else 1
        self.prefactor = float(prefactor)
        self.exclusion_radius = exclusion_radius
        self.exclusion_degree = exclusion_degree

    # ---- real space ------------------------------------------------------------
    def from_dist(self, d):
        """coulomb.py:80-96 / inversepowerlaw.py:54-71"""
        if self.kind == "coulomb":
            return self.prefactor / np.maximum(d, 1e-15)
        return self.prefactor * np.maximum(d, 1e-15) ** (-float(self.exponent))

    def lr_from_dist(self, d):
        """coulomb.py:98-120 / inversepowerlaw.py:73-106"""
        s = self.smearing
        if self.kind == "coulomb":
            return self.prefactor * _sp.erf(d / s / 2.0**0.5) / np.maximum(d, 1e-12)
        x = np.maximum(0.5 * d**2 / s**2, 1e-15)
        peff = self.exponent / 2.0
        pre = 1.0 / (2 * s**2) ** peff
        return self.prefactor * pre * _sp.gammainc(peff, x) / x**peff

    def f_cutoff(self, d):
        """potential.py:58-88"""
        rc = self.exclusion_radius
        inner = 1 - ((1 - np.cos(np.pi * (d / rc))) * 0.5) ** self.exclusion_degree
        return np.where(d < rc, inner, 0.0)

    def sr_from_dist(self, d):
        """potential.py:106-138"""
        if self.exclusion_radius is None:
            return self.from_dist(d) - self.lr_from_dist(d)
        return -self.lr_from_dist(d) * self.f_cutoff(d)

    def sr_from_dist_closed(self, d, deriv: bool = False):
        """
        Cancellation-free closed form of ``from_dist - lr_from_dist``:
        Q(p/2, d^2/2s^2)/d^p (asserted for p=1,2,3 by the
        reference in tests/test_potentials.py:103-111), and its derivative in d.
        """
        s = self.smearing
        p = float(self.exponent)
        x = 0.5 * d**2 / s**2
        # Q(p/2, x) in elementary functions (p = 1, 2, 3 are the forms the
        # reference asserts in tests/test_potentials.py:103-111), generic incomplete gamma otherwise
        p_int = int(self.exponent)
        if p_int == 1:
            q = _sp.erfc(np.sqrt(x))
        elif p_int == 2:
            q = np.exp(-x)
        elif p_int == 3:
            q = _sp.erfc(np.sqrt(x)) + 2 * np.sqrt(x / np.pi) * np.exp(-x)
        elif p_int == 4:
            q = np.exp(-x) * (1 + x)
        elif p_int == 5:
            q = _sp.erfc(np.sqrt(x)) + 2 * np.sqrt(x / np.pi) * np.exp(-x) * (1 + 2 * x / 3)
        elif p_int == 6:
            q = np.exp(-x) * (1 + x + 0.5 * x * x)
        else:
            q = _sp.gammaincc(p / 2.0, x)
        v = self.prefactor * q / d**p
        if not deriv:
            return v
        dq = -(x ** (p / 2.0 - 1.0)) * np.exp(-x) / math.gamma(p / 2.0) * (d / s**2)
        dv = self.prefactor * (dq / d**p - p * q / d ** (p + 1.0))
        return v, dv

    # ---- reciprocal space -----------------------------------------------------
    def lr_from_k_sq(self, k_sq):
        """coulomb.py:122-142 / inversepowerlaw.py:108-141"""
        s = self.smearing
        zero = k_sq == 0
        if self.kind == "coulomb":
            m = np.where(zero, 1.0, k_sq)
            return self.prefactor * np.where(zero, 0.0, 4 * np.pi * np.exp(-0.5 * s**2 * m) / m)
        p = self.exponent
        peff = (3 - p) / 2.0
        pre = np.pi**1.5 / math.gamma(p / 2.0) * (2 * s**2) ** peff
        x = 0.5 * s**2 * k_sq
        m = np.where(x == 0, 1.0, x)
        k0 = -pre / peff if p > 3 else 0.0
        return self.prefactor * np.where(zero, k0, pre * gammaincc_over_powerlaw(p, m))

    # ---- scalar corrections ---------------------------------------------------
    def self_contribution(self):
        """coulomb.py:144-150 / inversepowerlaw.py:143-150"""
        s = self.smearing
        if self.kind == "coulomb":
            return self.prefactor * (2 / np.pi) ** 0.5 / s
        ph = self.exponent / 2.0
        return self.prefactor / math.gamma(ph + 1) / (2 * s**2) ** ph

    def background_correction(self):
        """coulomb.py:152-158 / inversepowerlaw.py:152-164"""
        s = self.smearing
        if self.kind == "coulomb":
            return self.prefactor * np.pi * s**2
        p = self.exponent
        if p >= 3:
            return 0.0
        pre = np.pi**1.5 * (2 * s**2) ** ((3 - p) / 2.0)
        return self.prefactor * pre / ((3 - p) * math.gamma(p / 2.0))

    def has_slab_correction(self):
        """coulomb.py:160-167 / inversepowerlaw.py:166-169"""
        return self.kind == "coulomb" or self.exponent == 1


def pbc_correction(periodic, positions, cell, charges) -> np.ndarray:
    """2-D slab term, non-zero only with exactly two periodic axes (coulomb.py:6-40)."""
    if periodic is None or int(np.sum(periodic)) != 2:
        return np.zeros_like(charges)
    axis = int(np.argmax(~np.asarray(periodic, dtype=bool)))
    z = positions[:, axis : axis + 1]
    blen = np.linalg.norm(cell, axis=-1)[axis]
    vol = abs(np.linalg.det(cell))
    qtot = charges.sum(0)
    m1 = (charges * z).sum(0)
    m2 = (charges * z**2).sum(0)
    return (4.0 * np.pi / vol) * (z * m1 - 0.5 * (m2 + qtot * z**2) - qtot / 12.0 * blen**2)


def p3m_influence(kvectors: np.ndarray, cell: np.ndarray, ns, nodes: int) -> np.ndarray:
    """
    Mode-0 influence function 1/U^2, U^2 = [prod_a sinc(k_a h_a / 2 pi)]^(2 nodes) with
    the reference's Cartesian-component convention h_a = |cell_a|/n_a
    (lib/kspace_filter.py:307-316,349-361); 0 where U^2 == 0.
    """
    h = np.linalg.norm(cell, axis=1) / np.asarray(ns, dtype=cell.dtype)
    kh = kvectors * h.reshape(1, 1, 1, 3)
    u2 = np.prod(np.sinc(kh / (2 * np.pi)), axis=-1) ** (2 * nodes)
    m = np.where(u2 == 0, 1.0, u2)
    return np.where(u2 == 0, 0.0, 1.0 / m)


def _rfftn(a):
    return _sfft.rfftn(a, axes=(1, 2, 3), workers=-1)


def _irfftn(a, shape):
    return _sfft.irfftn(a, s=shape, axes=(1, 2, 3), workers=-1)


_NORM_EXP = {"backward": (0.0, 1.0), "ortho": (0.5, 0.5), "forward": (1.0, 0.0)}


def kspace_filter(mesh: np.ndarray, kfilter: np.ndarray, fft_norm="ortho", ifft_norm="ortho"):
    """
    irfftn(rfftn(mesh) * kfilter) over the last three axes with torch's ``norm`` keywords
    (lib/kspace_filter.py:122-197).  The calculators use ("backward", "forward"), i.e. no
    scaling either way (calculators/pme.py:72-78).
    """
    n_mesh = float(np.prod(mesh.shape[1:]))
    scale = n_mesh ** (-_NORM_EXP[fft_norm][0]) * n_mesh ** (1.0 - _NORM_EXP[ifft_norm][1])
    # numpy/scipy inverse carries 1/n ("backward"); fold every combination into one factor
    hat = _rfftn(mesh) * kfilter
    out = _irfftn(hat, mesh.shape[1:]) * scale
    return out.astype(mesh.dtype)


# --------------------------------------------------------------------------------------
# calculators
# --------------------------------------------------------------------------------------
def compute_rspace(pot: PotentialSpec, charges, neighbor_indices, neighbor_distances,
                   full_neighbor_list=False, closed_form=False, pair_mask=None):
    """
    V_i = 1/2 sum_(i,j) q_j v(d_ij)   (calculators/calculator.py:43-87): v = v_SR with a smearing,
    the bare potential (times 1 - f_cut if an exclusion radius is set) without one (:53-61);
    masked pairs contribute nothing (potential.py:106-138, `* pair_mask`).
    """
    d = neighbor_distances
    if pot.smearing is None:
        v = pot.from_dist(d)
        if pot.exclusion_radius is not None:
            v = v * (1 - pot.f_cutoff(d))
    else:
        v = pot.sr_from_dist_closed(d) if closed_form else pot.sr_from_dist(d)
    if pair_mask is not None:
        v = v * pair_mask
    ii = neighbor_indices[:, 0]
    jj = neighbor_indices[:, 1]
    out = np.zeros_like(charges)
    n = charges.shape[0]
    for c in range(charges.shape[1]):
        out[:, c] += np.bincount(ii, weights=charges[jj, c] * v, minlength=n)
        if not full_neighbor_list:
            out[:, c] += np.bincount(jj, weights=charges[ii, c] * v, minlength=n)
    return (out / 2).astype(charges.dtype)


def kfilter_for(pot: PotentialSpec, cell, ns, method: str, nodes: int) -> np.ndarray:
    """G(k) on the half mesh; P3M multiplies the influence function (kspace_filter.py:293-305)."""
    kv = kvectors_for_mesh(cell, ns)
    k_sq = np.linalg.norm(kv, axis=3) ** 2
    g = pot.lr_from_k_sq(k_sq)
    if method == "P3M":
        g = p3m_influence(kv, cell, ns, nodes) * g
    return g.astype(cell.dtype)


def compute_kspace(pot: PotentialSpec, charges, cell, positions, mesh_spacing, nodes, method,
                   periodic=None, ns=None):
    """Long-range part incl. self/background/slab terms (calculators/pme.py:88-143)."""
    if ns is None:
        ns = get_ns_mesh(cell, mesh_spacing)
    g = kfilter_for(pot, cell, ns, method, nodes)
    rho = points_to_mesh(charges, positions, cell, ns, nodes, method)
    phi = kspace_filter(rho, g, "backward", "forward")
    ivol = 1.0 / abs(np.linalg.det(cell))
    v = mesh_to_points(phi, positions, cell, nodes, method) * ivol
    v = v - charges * pot.self_contribution()
    v = v - 2 * pot.background_correction() * charges.sum(0) * ivol
    if pot.has_slab_correction():
        v = v + pot.prefactor * pbc_correction(periodic, positions, cell, charges)
    return (v / 2).astype(charges.dtype)


def calculator_forward(pot, charges, cell, positions, neighbor_indices, neighbor_distances,
                       mesh_spacing, nodes=4, method="Lagrange", full_neighbor_list=False,
                       periodic=None, ns=None):
    """``Calculator.forward`` for PME ('Lagrange') / P3M ('P3M') (calculators/calculator.py:103-189)."""
    sr = compute_rspace(pot, charges, neighbor_indices, neighbor_distances, full_neighbor_list)
    lr = compute_kspace(pot, charges, cell, positions, mesh_spacing, nodes, method, periodic, ns)
    return sr + lr


def calculator_step(pot, charges, cell, positions, neighbor_indices, neighbor_distances,
                    mesh_spacing, nodes=4, method="Lagrange", full_neighbor_list=False,
                    grad_out=None, ns=None):
    """
    One benchmark *step*: forward, then the analytic backward of L = sum(grad_out * V)
    (default grad_out = charges, i.e. L = sum_i q_i V_i) with respect to positions
    (k-space path), charges and neighbor distances.  This restates what PyTorch's tape
    derives from the forward ops: one more spread (of grad_out),
    one more filter pass (the filter is self-adjoint for real even G) and two
    derivative-weight gathers.  3-D periodic only (no slab term).

    Returns ``dict(V, dpos, dq, dd)``.
    """
    if grad_out is None:
        grad_out = charges
    if ns is None:
        ns = get_ns_mesh(cell, mesh_spacing)
    dt = charges.dtype
    n = charges.shape[0]
    ii = neighbor_indices[:, 0]
    jj = neighbor_indices[:, 1]
    half = not full_neighbor_list

    # ---- forward --------------------------------------------------------------
    if pot.exclusion_radius is None:
        v, dv = pot.sr_from_dist_closed(neighbor_distances, deriv=True)
    else:  # numerical derivative only needed for the exclusion variant
        v = pot.sr_from_dist(neighbor_distances)
        eps = 1e-6
        dv = (pot.sr_from_dist(neighbor_distances + eps) - pot.sr_from_dist(neighbor_distances - eps)) / (2 * eps)
    v_sr = np.zeros_like(charges)
    for c in range(charges.shape[1]):
        v_sr[:, c] += np.bincount(ii, weights=charges[jj, c] * v, minlength=n)
        if half:
            v_sr[:, c] += np.bincount(jj, weights=charges[ii, c] * v, minlength=n)
    v_sr /= 2

    g = kfilter_for(pot, cell, ns, method, nodes)
    ivol = 1.0 / abs(np.linalg.det(cell))
    rho = points_to_mesh(charges, positions, cell, ns, nodes, method)
    phi = kspace_filter(rho, g, "backward", "forward")
    v_g, dv_g = mesh_to_points(phi, positions, cell, nodes, method, gradient=True)
    v_lr = v_g * ivol - charges * pot.self_contribution()
    v_lr = v_lr - 2 * pot.background_correction() * charges.sum(0) * ivol
    v_lr /= 2
    V = (v_sr + v_lr).astype(dt)

    # ---- backward -------------------------------------------------------------
    gq = grad_out
    # real space
    pair_w = (gq[ii] * charges[jj]).sum(1)
    if half:
        pair_w = pair_w + (gq[jj] * charges[ii]).sum(1)
    dd = 0.5 * dv * pair_w
    dq = np.zeros_like(charges)
    for c in range(charges.shape[1]):
        dq[:, c] += 0.5 * np.bincount(jj, weights=gq[ii, c] * v, minlength=n)
        if half:
            dq[:, c] += 0.5 * np.bincount(ii, weights=gq[jj, c] * v, minlength=n)
    # reciprocal space
    rho_g = points_to_mesh(gq, positions, cell, ns, nodes, method)
    psi = kspace_filter(rho_g, g, "backward", "forward")
    u_g, du_g = mesh_to_points(psi, positions, cell, nodes, method, gradient=True)
    dq += 0.5 * (ivol * u_g - gq * pot.self_contribution()
                 - 2 * pot.background_correction() * ivol * gq.sum(0))
    dpos = 0.5 * ivol * (
        (gq[:, :, None] * dv_g).sum(1) + (charges[:, :, None] * du_g).sum(1)
    )
    return dict(V=V, dpos=dpos.astype(dt), dq=dq.astype(dt), dd=dd.astype(dt))


# --------------------------------------------------------------------------------------
# test utilities (not part of the restated path): neighbor lists, synthetic crystals
# --------------------------------------------------------------------------------------
def neighbor_list(positions, cell, cutoff, full=False):
    """
    Brute-force periodic neighbor list (substitute for the external ``vesin`` package used
    by the reference's tests, tests/helpers.py:240-275): every pair of (atom, periodic
    image) within ``cutoff`` exactly once for half lists and twice for full lists,
    including images of an atom with itself.  Returns (idx (P,2) int64, d (P,), S (P,3)).
    """
    positions = np.asarray(positions, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64)
    n = len(positions)
    inv = np.linalg.inv(cell)
    # number of images needed along each lattice direction
    heights = 1.0 / np.linalg.norm(inv, axis=0)
    reps = np.ceil(cutoff / heights).astype(int) + 1
    rng = [np.arange(-r, r + 1) for r in reps]
    shifts = np.stack(np.meshgrid(*rng, indexing="ij"), -1).reshape(-1, 3)
    out_i, out_j, out_d, out_s = [], [], [], []
    for s in shifts:
        delta = positions[None, :, :] + (s @ cell)[None, None, :] - positions[:, None, :]
        dist = np.linalg.norm(delta, axis=-1)
        mask = dist < cutoff
        if not s.any():
            mask &= ~np.eye(n, dtype=bool)
        i, j = np.nonzero(mask)
        if not full:
            # keep each unordered (i, j, S) / (j, i, -S) pair once
            key = tuple(s)
            if key < (0, 0, 0):
                continue
            if key == (0, 0, 0):
                keep = i < j
            else:
                keep = np.ones(len(i), dtype=bool)
            i, j = i[keep], j[keep]
        out_i.append(i)
        out_j.append(j)
        out_d.append(dist[i, j])
        out_s.append(np.broadcast_to(s, (len(i), 3)))
    idx = np.stack([np.concatenate(out_i), np.concatenate(out_j)], 1).astype(np.int64)
    return idx, np.concatenate(out_d), np.concatenate(out_s).astype(np.int64)
