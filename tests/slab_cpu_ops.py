"""
TEST INFRASTRUCTURE ONLY -- a CPU emulation of the ``torchpme_b200._native`` function set that
the slab-decomposed calculators call, built on the numpy oracle.

It exists so that the host-side logic of ``torchpme_b200.distributed`` (slab layout, the stride
arithmetic of the exchange copies, all-to-all / all-reduce plumbing, pair-list sharding,
gradient assembly) can be exercised with ``gloo`` and world_size 2 on a machine without a GPU.
The product never selects it: the calculators default to the CUDA library and raise when it
is missing; only tests pass ``_ops=CpuOps()``.
"""
import ctypes
from types import SimpleNamespace

import numpy as np
import torch

from oracle import pme_oracle as oracle

_METHOD = {0: "P3M", 1: "Lagrange"}


def _np(t):
    return t.detach().cpu().numpy()


def _cell_from_r2u(r2u, ns):
    inv = np.asarray(r2u, dtype=np.float64).reshape(3, 3) / np.asarray(ns, dtype=np.float64)[None, :]
    return np.linalg.inv(inv)


class CpuOps:
    # ---- descriptors ----------------------------------------------------------------------
    @staticmethod
    def make_green(kind, scale, recip, spacing=(0.0, 0.0, 0.0), smearing=1.0, prefactor=1.0, exponent=1,
                   p3m_nodes=0, table=None):
        return SimpleNamespace(kind=kind, scale=scale, recip=np.asarray(recip, dtype=np.float64).reshape(3, 3),
                               smearing=smearing, prefactor=prefactor, exponent=exponent, p3m_nodes=p3m_nodes)

    @staticmethod
    def make_pair_potential(kind, smearing=1.0, prefactor=1.0, exponent=1, exclusion_radius=None,
                            exclusion_degree=1):
        name = {1: "coulomb", 2: "ipl"}[kind]
        return SimpleNamespace(kind=kind, spec=oracle.PotentialSpec(name, smearing, exponent, prefactor,
                                                                    exclusion_radius, exclusion_degree))

    @staticmethod
    def make_epilogue(add_coef, dc, scale, self_half, background, coef2=None, dvalues2=None, vjp_scale=0.0):
        return SimpleNamespace(add_coef=add_coef, dc=dc, scale=scale, self_half=self_half,
                               background=background, coef2=coef2, dvalues2=dvalues2, vjp_scale=vjp_scale)

    # ---- mesh interpolation -----------------------------------------------------------------
    @staticmethod
    def slab_select_points(positions, r2u, ns, nodes, slab):
        """the emulation visits every point (foreign points contribute exact zeros)"""
        return None

    @staticmethod
    def spread(positions, weights, r2u, ns, nodes, method, out=None, slab=None, point_list=None):
        cell = _cell_from_r2u(r2u, ns)
        full = oracle.points_to_mesh(_np(weights), _np(positions), cell, ns, nodes, _METHOD[method])
        x0, nxl = (0, ns[0]) if slab is None else slab
        return torch.from_numpy(np.ascontiguousarray(full[:, x0:x0 + nxl])).to(positions.dtype)

    @staticmethod
    def _gather_parts(mesh, positions, r2u, nodes, method, slab):
        c, nxl, ny, nz = mesh.shape
        x0, nx = (0, nxl) if slab is None else slab
        full = np.zeros((c, nx, ny, nz), dtype=np.float64)
        full[:, x0:x0 + nxl] = _np(mesh)
        cell = _cell_from_r2u(r2u, (nx, ny, nz))
        return oracle.mesh_to_points(full, _np(positions).astype(np.float64), cell, nodes, _METHOD[method],
                                     gradient=True)

    @classmethod
    def gather(cls, mesh, positions, r2u, nodes, method, want_values=True, want_grad=False, values_out=None,
               epilogue=None, slab=None, point_list=None):
        vals, dvals = cls._gather_parts(mesh, positions, r2u, nodes, method, slab)
        vals_t = torch.from_numpy(vals).to(mesh.dtype)
        if epilogue is not None:
            e = epilogue
            values_out.copy_(values_out + e.scale * vals_t - e.add_coef * e.self_half - e.background * e.dc)
            vals_t = values_out
        elif values_out is not None:
            values_out.copy_(vals_t)
            vals_t = values_out
        return (vals_t if (want_values or values_out is not None) else None,
                torch.from_numpy(dvals).to(mesh.dtype) if want_grad else None)

    @classmethod
    def gather_vjp(cls, mesh, positions, coef, r2u, nodes, method, grad_positions=None, want_values=False,
                   want_grad_r2u=False, values_out=None, epilogue=None, slab=None, point_list=None):
        vals, dvals = cls._gather_parts(mesh, positions, r2u, nodes, method, slab)
        vals_t = torch.from_numpy(vals).to(mesh.dtype)
        g = torch.einsum("ic,icd->id", coef.double(), torch.from_numpy(dvals))
        if epilogue is not None and epilogue.coef2 is not None:
            g = (g + torch.einsum("ic,icd->id", epilogue.coef2.double(), epilogue.dvalues2.double())) * epilogue.vjp_scale
        g = g.to(mesh.dtype)
        if grad_positions is None:
            grad_positions = g
        else:
            grad_positions.add_(g)
        if values_out is not None:
            e = epilogue
            if e is not None:
                values_out.copy_(values_out + e.scale * vals_t - e.add_coef * e.self_half - e.background * e.dc)
            else:
                values_out.copy_(vals_t)
        return grad_positions, values_out, None

    # ---- slab FFT pieces ----------------------------------------------------------------------
    @staticmethod
    def slab_fft_yz(forward, real_mesh, mesh_hat):
        c, nxl, ny, nz = real_mesh.shape
        if forward:
            h = np.fft.rfft2(_np(real_mesh).astype(np.float64), axes=(2, 3))
            mesh_hat.copy_(torch.view_as_real(torch.from_numpy(h)).to(mesh_hat.dtype))
        else:
            h = torch.view_as_complex(mesh_hat.double().contiguous()).numpy()
            r = np.fft.irfft2(h, s=(ny, nz), axes=(2, 3)) * (ny * nz)
            real_mesh.copy_(torch.from_numpy(r).to(real_mesh.dtype))

    @staticmethod
    def slab_fft_x_green(mesh_hat_t, ns, y0, green):
        nx, ny, nz = ns
        c, _, nyl, nzh, _ = mesh_hat_t.shape
        cell = np.linalg.inv(green.recip.T / (2 * np.pi))
        name = {1: "coulomb", 2: "ipl"}[green.kind]
        spec = oracle.PotentialSpec(name, green.smearing, green.exponent, green.prefactor)
        g = oracle.kfilter_for(spec, cell, ns, "P3M" if green.p3m_nodes > 0 else "Lagrange", green.p3m_nodes)
        g = g[:, y0:y0 + nyl, :] * green.scale
        h = torch.view_as_complex(mesh_hat_t.double().contiguous()).numpy()
        h = np.fft.fft(h, axis=1) * g[None]
        h = np.fft.ifft(h, axis=1) * nx
        mesh_hat_t.copy_(torch.view_as_real(torch.from_numpy(h)).to(mesh_hat_t.dtype))

    @staticmethod
    def slab_exchange_copy(src, dst_ptrs, n_c, n_p, n_a, run, src_strides, dst_strides):
        """the same index arithmetic as the CUDA kernel, on host pointers"""
        elem = 2 * src.element_size()
        base = src.data_ptr()
        s_c, s_p, s_a = src_strides
        d_c, d_a = dst_strides
        for c in range(n_c):
            for p in range(n_p):
                for a in range(n_a):
                    ctypes.memmove(int(dst_ptrs[p]) + (c * d_c + a * d_a) * elem,
                                   base + (c * s_c + p * s_p + a * s_a) * elem, run * elem)

    # ---- real space -----------------------------------------------------------------------------
    @staticmethod
    def pair_forward(charges, idx, dist, pair_values, mask_u8, full_list, pot, out=None):
        v = oracle.compute_rspace(pot.spec, _np(charges).astype(np.float64), _np(idx),
                                  _np(dist).astype(np.float64), full_list, closed_form=pot.spec.exclusion_radius is None)
        v = torch.from_numpy(v).to(charges.dtype)
        if out is None:
            return v
        out.add_(v)
        return out

    @staticmethod
    def pair_backward(charges, idx, dist, pair_values, mask_u8, grad_out, full_list, pot, want_charges=True,
                      want_pairs=True, grad_charges_out=None, grad_pairs_out=None):
        q, g = _np(charges).astype(np.float64), _np(grad_out).astype(np.float64)
        d = _np(dist).astype(np.float64)
        ii, jj = _np(idx)[:, 0], _np(idx)[:, 1]
        v, dv = pot.spec.sr_from_dist_closed(d, deriv=True)
        n = q.shape[0]
        half = not full_list
        if want_pairs:
            w = (g[ii] * q[jj]).sum(1)
            if half:
                w = w + (g[jj] * q[ii]).sum(1)
            grad_pairs_out.copy_(torch.from_numpy(0.5 * dv * w).to(charges.dtype))
        if want_charges:
            dq = np.zeros_like(q)
            for c in range(q.shape[1]):
                dq[:, c] += 0.5 * np.bincount(jj, weights=g[ii, c] * v, minlength=n)
                if half:
                    dq[:, c] += 0.5 * np.bincount(ii, weights=g[jj, c] * v, minlength=n)
            grad_charges_out.add_(torch.from_numpy(dq).to(charges.dtype))
        return grad_charges_out, grad_pairs_out
