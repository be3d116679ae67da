"""
Slab-decomposed PME / P3M across the GPUs of one node.

One process per GPU (``torch.distributed``).  Every rank passes the *same* replicated
``charges / cell / positions`` (16 B per atom) and the same neighbor list to
``forward`` -- the reference call signature is unchanged -- and gets the same full ``(N, C)``
potentials back.  Inside:

* the mesh is cut into x slabs, rank ``g`` owns the planes ``[g nx/W, (g+1) nx/W)``;
* **spread**: every rank scans all atoms and keeps only the stencil planes that fall into its
  slab (``tpme_spread_slab``) -- no halo, no reduction;
* **FFT . G . iFFT**: local (y,z) passes, exchange x slabs -> y slabs, x pass fused with the
  Green's function on the local y rows, exchange back, inverse (y,z) passes.  The exchange is
  * ``transport="p2p"``: fused into the FFT kernels -- the y pass and the x pass store their
    results straight into the peers' arrays over NVLink (CUDA IPC mappings) while other thread
    blocks still compute, followed by a device-side flag barrier; no copy kernel, no library call;
  * ``transport="p2p-copy"``: one copy kernel that packs, transfers (peer stores) and unpacks;
  * ``transport="nccl"``: ``all_to_all_single`` between a packing and an unpacking copy kernel;
* **gather**: partial sums over the local planes (``tpme_gather_slab``); they are summed over
  the ranks by the same all-reduce that combines the real-space pair sum, whose pair list is
  cut into ``W`` contiguous chunks;
* backward mirrors this (the filter is self-adjoint) and ends with one all-reduce of the
  ``(N, 3 + C)`` position / charge gradients.  With ``shard_pairs=True`` (replicated pair list)
  the gradient w.r.t. ``neighbor_distances`` is all-reduced as well, so every returned gradient is
  the full, replicated one; with ``shard_pairs=False`` every rank passes its own chunk of the pair
  list and gets the gradient of that chunk.

Collectives per step (forward + backward): 4 exchanges of the half-complex mesh
(``C nx ny (nz/2+1)`` complex numbers / W per rank each) and 2 all-reduces.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist
from torch.autograd.function import once_differentiable

from . import _native
from ._checks import validate_parameters
from .calculators import P3MCalculator, PMECalculator, _side_stream
from .mesh import geometry_of


class SlabLayout:
    """Host-side partition of an ``(nx, ny, nz)`` mesh and of a pair list over ``world`` ranks."""

    def __init__(self, ns, world: int, rank: int):
        nx, ny, nz = (int(v) for v in ns)
        if world < 1 or not 0 <= rank < world:
            raise ValueError(f"invalid rank {rank} for world size {world}")
        if world > _native.MAX_RANKS:
            raise ValueError(f"at most {_native.MAX_RANKS} ranks are supported, got {world}")
        if nx % world or ny % world:
            raise ValueError(
                f"mesh {nx} x {ny} x {nz} cannot be cut into {world} x slabs and {world} y slabs; "
                "the world size has to divide nx and ny"
            )
        self.ns, self.world, self.rank = (nx, ny, nz), world, rank
        self.nxl, self.nyl, self.nzh = nx // world, ny // world, nz // 2 + 1
        self.x0, self.y0 = rank * self.nxl, rank * self.nyl
        #: complex elements of one exchanged block (one rank's x planes x one rank's y rows)
        self.block = self.nxl * self.nyl * self.nzh

    def pair_range(self, n_pairs: int):
        """contiguous chunk of the pair list owned by this rank"""
        per = -(-n_pairs // self.world)
        lo = min(self.rank * per, n_pairs)
        return lo, min(lo + per, n_pairs)


#: seconds a rank waits in a device-side peer barrier before it gives up.  A time-out sets the error
#: flag of the exchange / reducer; the next all-reduce then returns NaN for EVERY element (so a
#: stalled peer can never produce silently wrong potentials or forces) and ``check()`` raises.
PEER_TIMEOUT_SECONDS = float(os.environ.get("TPME_PEER_TIMEOUT", "60"))


def _exchange_handles(handle: bytes, device, group, world: int):
    """all-gather of the CUDA IPC handles (device tensors under NCCL, host tensors under gloo)"""
    on = device if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=on)
    handles = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(handles, mine, group=group)
    return [bytes(h.cpu().tolist()) for h in handles]


def _ptr(t: torch.Tensor, offset_elems: int = 0) -> int:
    return t.data_ptr() + offset_elems * t.element_size()


class NcclExchange:
    """x slabs <-> y slabs through ``all_to_all_single`` with packing kernels on both sides."""

    name = "nccl"

    def __init__(self, layout: SlabLayout, channels: int, dtype, device, group, ops):
        self.layout, self.c, self.group, self.ops = layout, channels, group, ops
        w, blk = layout.world, layout.block
        shape = (w * channels * blk, 2)
        self.buf_a = torch.empty(shape, dtype=dtype, device=device)
        self.buf_b = torch.empty(shape, dtype=dtype, device=device)
        nx, ny, _ = layout.ns
        self.hat = torch.empty((channels, layout.nxl, ny, layout.nzh, 2), dtype=dtype, device=device)
        self.hat_t = torch.empty((channels, nx, layout.nyl, layout.nzh, 2), dtype=dtype, device=device)

    def x_to_y(self):
        """self.hat (C, nxl, ny, nzh) -> self.hat_t (C, nx, nyl, nzh) of every rank"""
        lay, c, ops = self.layout, self.c, self.ops
        w, blk, run = lay.world, lay.block, lay.nyl * lay.nzh
        ny = lay.ns[1]
        # pack: block p = hat[:, :, p nyl:(p+1) nyl, :] as (C, nxl, nyl, nzh)
        ops.slab_exchange_copy(self.hat, [_ptr(self.buf_a, 2 * p * c * blk) for p in range(w)],
                               c, w, lay.nxl, run, (lay.nxl * ny * lay.nzh, run, ny * lay.nzh), (blk, run))
        recv = self.hat_t.view(-1, 2) if c == 1 else self.buf_b
        dist.all_to_all_single(recv, self.buf_a, group=self.group)
        if c > 1:
            # (W, C, blk) -> (C, W, blk)
            ops.slab_exchange_copy(self.buf_b, [_ptr(self.hat_t, 2 * p * blk) for p in range(w)],
                                   c, w, 1, blk, (blk, c * blk, 0), (w * blk, 0))
        return self.hat_t

    def y_to_x(self):
        """self.hat_t (C, nx, nyl, nzh) -> self.hat (C, nxl, ny, nzh) of every rank"""
        lay, c, ops = self.layout, self.c, self.ops
        w, blk, run = lay.world, lay.block, lay.nyl * lay.nzh
        ny = lay.ns[1]
        if c == 1:
            send = self.hat_t.view(-1, 2)
        else:
            send = self.buf_a
            ops.slab_exchange_copy(self.hat_t, [_ptr(self.buf_a, 2 * p * c * blk) for p in range(w)],
                                   c, w, 1, blk, (w * blk, blk, 0), (blk, 0))
        dist.all_to_all_single(self.buf_b, send, group=self.group)
        # block from rank p holds the y rows of p: (W, C, nxl, nyl, nzh) -> (C, nxl, W, nyl, nzh)
        ops.slab_exchange_copy(self.buf_b, [_ptr(self.hat, 2 * p * run) for p in range(w)],
                               c, w, lay.nxl, run, (blk, c * blk, run), (lay.nxl * ny * lay.nzh, ny * lay.nzh))
        return self.hat


class PeerExchange:
    """
    x slabs <-> y slabs by storing straight into the peers' buffers over NVLink: one copy kernel
    packs, transfers and unpacks; a flag barrier in peer memory orders it (CUDA IPC, one node).
    """

    name = "p2p"
    _FLAG_BYTES = 256

    def __init__(self, layout: SlabLayout, channels: int, dtype, device, group, ops=None):
        if device.type != "cuda":
            raise ValueError("the peer-memory exchange needs CUDA devices")
        self.layout, self.c, self.group, self.ops = layout, channels, group, ops or _native
        w = layout.world
        nx, ny, _ = layout.ns
        esize = 8 if dtype == torch.float32 else 16
        half = channels * layout.nxl * ny * layout.nzh * esize   # == C * nx * nyl * nzh * esize
        half = (half + 255) // 256 * 256
        self._off_x, self._off_t = self._FLAG_BYTES, self._FLAG_BYTES + half
        self.buffer = _native.PeerBuffer(self._FLAG_BYTES + 2 * half, device)
        # exchange the IPC handles and map the peers' buffers
        handles = _exchange_handles(self.buffer.handle, device, group, w)
        self.peer_base = []
        for p in range(w):
            if p == layout.rank:
                self.peer_base.append(self.buffer.ptr)
            else:
                self.peer_base.append(self.buffer.open_peer(handles[p]))
        self.esize = esize
        self.hat = self.buffer.as_tensor(self._off_x, (channels, layout.nxl, ny, layout.nzh, 2), dtype)
        self.hat_t = self.buffer.as_tensor(self._off_t, (channels, nx, layout.nyl, layout.nzh, 2), dtype)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        self.dtype, self.device = dtype, device
        self.peers = _native.make_slab_peers(layout.rank, [b + self._off_x for b in self.peer_base],
                                             [b + self._off_t for b in self.peer_base])
        dist.barrier(group=group)   # every rank has mapped every buffer before the first store

    # fused compute + exchange: the FFT passes store their results into the peers' arrays
    def push_yz(self, rho_local):
        _native.slab_fft_yz_push(rho_local, self.layout.ns, self.peers)
        self._barrier()
        return self.hat_t

    def push_x_green(self, green):
        _native.slab_fft_x_green_push(self.dtype, self.device, self.c, self.layout.ns, green, self.peers)
        self._barrier()
        return self.hat

    def _barrier(self):
        _native.peer_barrier(self.peer_base, self.layout.rank, self.epoch, self.error, PEER_TIMEOUT_SECONDS)

    def check(self):
        """host-side check of the barrier time-out flag (a device->host read)"""
        code = int(self.error.item())
        if code:
            raise RuntimeError(f"peer barrier timed out waiting for rank {code - 1}")

    def x_to_y(self):
        lay, c = self.layout, self.c
        w, blk, run = lay.world, lay.block, lay.nyl * lay.nzh
        ny = lay.ns[1]
        # my x planes go to rows [rank nxl, (rank+1) nxl) of every peer's (C, nx, nyl, nzh) array
        dst = [self.peer_base[p] + self._off_t + lay.rank * blk * self.esize for p in range(w)]
        self.ops.slab_exchange_copy(self.hat, dst, c, w, lay.nxl, run,
                                    (lay.nxl * ny * lay.nzh, run, ny * lay.nzh), (w * blk, run))
        self._barrier()
        return self.hat_t

    def y_to_x(self):
        lay, c = self.layout, self.c
        w, blk, run = lay.world, lay.block, lay.nyl * lay.nzh
        ny = lay.ns[1]
        # my y rows go to rows [rank nyl, (rank+1) nyl) of every peer's (C, nxl, ny, nzh) array
        dst = [self.peer_base[p] + self._off_x + lay.rank * run * self.esize for p in range(w)]
        self.ops.slab_exchange_copy(self.hat_t, dst, c, w, lay.nxl, run,
                                    (w * blk, blk, run), (lay.nxl * ny * lay.nzh, ny * lay.nzh))
        self._barrier()
        return self.hat


class TorchReducer:
    """partial results summed with ``torch.distributed.all_reduce`` (NCCL, or gloo in the CPU tests)"""

    def __init__(self, group, world):
        self.group, self.world = group, world

    def zeros(self, n: int, dtype, device) -> torch.Tensor:
        return torch.zeros(n, dtype=dtype, device=device)

    def all_reduce(self, flat: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(flat, group=self.group)
        return flat


class PeerReducer:
    """
    Sum all-reduce over NVLink peer memory (``tpme_peer_allreduce``): the kernels accumulate their
    partial results straight into this rank's input region; after a flag barrier every rank pulls
    its slice from all peers, sums and pushes the result to all output regions.  One kernel and two
    barriers instead of a ring all-reduce (2 (W - 1) latency-bound steps at these sizes).
    """

    _FLAG_BYTES = 256

    def __init__(self, n_max: int, dtype, device, group, world: int, rank: int):
        self.world, self.rank, self.dtype, self.device = world, rank, dtype, device
        esize = 4 if dtype == torch.float32 else 8
        self.n_max = n_max
        region = (n_max * esize + 16 * world + 255) // 256 * 256
        self._off_in, self._off_out = self._FLAG_BYTES, self._FLAG_BYTES + region
        self.buffer = _native.PeerBuffer(self._FLAG_BYTES + 2 * region, device)
        handles = _exchange_handles(self.buffer.handle, device, group, world)
        self.peer_base = [self.buffer.ptr if p == rank else self.buffer.open_peer(handles[p])
                          for p in range(world)]
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        self._in = self.buffer.as_tensor(self._off_in, (n_max,), dtype)
        self._out = self.buffer.as_tensor(self._off_out, (n_max,), dtype)
        self._watched = [self.error]
        self._nan = torch.full((), float("nan"), dtype=dtype, device=device)
        dist.barrier(group=group)

    def watch(self, error_flag: torch.Tensor) -> None:
        """also poison the results when this (exchange) barrier flag is set"""
        if all(error_flag is not f for f in self._watched):
            self._watched.append(error_flag)

    def zeros(self, n: int, dtype, device) -> torch.Tensor:
        assert n <= self.n_max and dtype == self.dtype
        return self._in[:n].zero_()

    def all_reduce(self, flat: torch.Tensor) -> torch.Tensor:
        """`flat` is the view handed out by :meth:`zeros`; returns a fresh tensor with the sums"""
        n = flat.numel()
        _native.peer_barrier(self.peer_base, self.rank, self.epoch, self.error, PEER_TIMEOUT_SECONDS)
        _native.peer_allreduce(self.dtype, self.device, [b + self._off_in for b in self.peer_base],
                               [b + self._off_out for b in self.peer_base], self.rank, n)
        _native.peer_barrier(self.peer_base, self.rank, self.epoch, self.error, PEER_TIMEOUT_SECONDS)
        # a barrier that timed out anywhere in this step means partly written buffers: the copy out
        # of the exchange region doubles as the poisoning (NaN everywhere), no host sync needed
        failed = self._watched[0] if len(self._watched) == 1 else torch.stack(self._watched).sum()
        return torch.where(failed != 0, self._nan, self._out[:n])

    def check(self):
        code = int(self.error.item())
        if code:
            raise RuntimeError(f"peer barrier timed out waiting for rank {code - 1}")


class MulticastReducer:
    """
    Sum all-reduce through the NVSwitch (NVLS, ``tpme_multimem_allreduce``): the partial results of all
    ranks sit at the same offset of a symmetric allocation that is bound to ONE multicast address
    (``torch.distributed._symmetric_memory`` provides allocation and rendezvous -- plumbing); rank r
    reduces slice r with ``multimem.ld_reduce`` (the switch returns the sum over all GPUs) and broadcasts
    it in place with ``multimem.st``.  Same interface and barrier protocol as :class:`PeerReducer`.
    Raises ``RuntimeError`` when the node has no multicast support (the caller falls back).
    """

    _FLAG_BYTES = 256

    def __init__(self, n_max: int, dtype, device, group, world: int, rank: int):
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank, self.dtype, self.device = world, rank, dtype, device
        esize = 4 if dtype == torch.float32 else 8
        self.n_max = n_max
        self._flag_elems = self._FLAG_BYTES // esize
        region = (n_max * esize + 16 * world + 255) // 256 * 256
        self._buf = symm_mem.empty(self._flag_elems + region // esize, dtype=dtype, device=device)
        pg = group if group is not None else dist.group.WORLD
        self._handle = symm_mem.rendezvous(self._buf, pg.group_name)
        mc = int(getattr(self._handle, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("no multicast (NVLS) support on this node")
        self._mc_data = mc + self._FLAG_BYTES
        self.peer_base = [int(p) for p in self._handle.buffer_ptrs]
        self._buf.zero_()
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        self._watched = [self.error]
        self._nan = torch.full((), float("nan"), dtype=dtype, device=device)
        self._data = self._buf[self._flag_elems:]
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def watch(self, error_flag: torch.Tensor) -> None:
        if all(error_flag is not f for f in self._watched):
            self._watched.append(error_flag)

    def zeros(self, n: int, dtype, device) -> torch.Tensor:
        assert n <= self.n_max and dtype == self.dtype
        return self._data[:n].zero_()

    def all_reduce(self, flat: torch.Tensor) -> torch.Tensor:
        n = flat.numel()
        _native.peer_barrier(self.peer_base, self.rank, self.epoch, self.error, PEER_TIMEOUT_SECONDS)
        _native.multimem_allreduce(self.dtype, self.device, self._mc_data, self.world, self.rank, n)
        _native.peer_barrier(self.peer_base, self.rank, self.epoch, self.error, PEER_TIMEOUT_SECONDS)
        failed = self._watched[0] if len(self._watched) == 1 else torch.stack(self._watched).sum()
        return torch.where(failed != 0, self._nan, self._data[:n])

    def check(self):
        code = int(self.error.item())
        if code:
            raise RuntimeError(f"peer barrier timed out waiting for rank {code - 1}")


def make_reducer(n_max: int, dtype, device, group, world: int, rank: int, kind: str = "auto"):
    """
    peer-memory reducer of the p2p transports: NVLS multicast when the node supports it (all ranks
    agree through one tiny all-reduce), plain peer pointers otherwise.  TPME_SLAB_REDUCER = auto |
    multimem | peer overrides.
    """
    kind = os.environ.get("TPME_SLAB_REDUCER", kind)
    reducer = None
    if kind in ("auto", "multimem") and dist.get_backend(group) == "nccl":
        try:
            reducer = MulticastReducer(n_max, dtype, device, group, world, rank)
        except Exception:
            if kind == "multimem":
                raise
            reducer = None
        ok = torch.tensor([1 if reducer is not None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok) == 0:
            reducer = None
    if reducer is None:
        reducer = PeerReducer(n_max, dtype, device, group, world, rank)
    return reducer


class SlabFilter:
    """``irfft3(G * rfft3(.))`` of a mesh distributed as x slabs (buffers reused across calls)."""

    def __init__(self, layout: SlabLayout, channels: int, dtype, device, group, transport="nccl",
                 ops=None):
        self.layout, self.ops = layout, ops or _native
        self.fused = False
        if transport in ("p2p", "p2p-copy"):
            self.exchange = PeerExchange(layout, channels, dtype, device, group, self.ops)
            self.fused = transport == "p2p"
        elif transport == "nccl":
            self.exchange = NcclExchange(layout, channels, dtype, device, group, self.ops)
        else:
            raise ValueError(f"unknown transport '{transport}' (choose 'nccl', 'p2p' or 'p2p-copy')")

    def apply(self, rho_local: torch.Tensor, green) -> torch.Tensor:
        """``rho_local`` (C, nxl, ny, nz) -> filtered slab of the same shape"""
        ex, lay, ops = self.exchange, self.layout, self.ops
        if self.fused:
            # transport "p2p": the y pass and the x pass push their results to the peers themselves
            ex.push_yz(rho_local)
            hat = ex.push_x_green(green)
        else:
            ops.slab_fft_yz(True, rho_local, ex.hat)
            hat_t = ex.x_to_y()
            ops.slab_fft_x_green(hat_t, lay.ns, lay.y0, green)
            hat = ex.y_to_x()
        out = torch.empty_like(rho_local)
        ops.slab_fft_yz(False, out, hat)
        return out


class _SlabStepConfig:
    __slots__ = ("r2u", "ns", "nodes", "method", "green_args", "pair_pot", "full_list", "half_ivolume",
                 "self_half", "background_ivolume", "layout", "filter", "group", "ops", "shard_pairs",
                 "reducer", "n_atoms")


class _SlabMeshPotential(torch.autograd.Function):
    """
    Distributed counterpart of ``calculators._FusedMeshPotential``: one autograd node per
    calculator forward; see the module docstring for the data flow.
    """

    @staticmethod
    def forward(ctx, charges, positions, distances, neighbor_indices, mask_u8, cfg):
        ops, lay = cfg.ops, cfg.layout
        q = charges.detach().contiguous()
        pos = positions.detach().contiguous()
        need_pos = ctx.needs_input_grad[1]
        n_pairs = neighbor_indices.shape[0]
        lo, hi = lay.pair_range(n_pairs) if cfg.shard_pairs else (0, n_pairs)
        idx = neighbor_indices[lo:hi].contiguous()
        d = distances.detach()[lo:hi].contiguous()
        mask = None if mask_u8 is None else mask_u8[lo:hi].contiguous()
        out = cfg.reducer.zeros(q.numel(), q.dtype, q.device).view(q.shape)
        cuda = q.is_cuda
        if cuda:
            main = torch.cuda.current_stream(q.device)
            side = _side_stream(q.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ops.pair_forward(q, idx, d, None, mask, cfg.full_list, cfg.pair_pot, out=out)
        else:
            ops.pair_forward(q, idx, d, None, mask, cfg.full_list, cfg.pair_pot, out=out)
        # the atoms whose stencil reaches into this rank's slab (~N/W of them): every mesh kernel of
        # the step strides over this list
        plist = ops.slab_select_points(pos, cfg.r2u, cfg.ns, cfg.nodes, (lay.x0, lay.nxl)) if lay.world > 1 else None
        rho = ops.spread(pos, q, cfg.r2u, cfg.ns, cfg.nodes, cfg.method, slab=(lay.x0, lay.nxl), point_list=plist)
        green = ops.make_green(scale=1.0, **cfg.green_args)
        phi = cfg.filter.apply(rho, green)
        if cuda:
            main.wait_stream(side)
        # out += 1/(2 Vol) * (partial gather); the self / background terms are added once, after
        # the all-reduce
        zero_dc = torch.zeros(q.shape[1], dtype=q.dtype, device=q.device)
        epi = ops.make_epilogue(q, zero_dc, cfg.half_ivolume, 0.0, 0.0)
        _, dvalues = ops.gather(phi, pos, cfg.r2u, cfg.nodes, cfg.method, want_grad=need_pos,
                                values_out=out, epilogue=epi, slab=(lay.x0, lay.ns[0]), point_list=plist)
        out = cfg.reducer.all_reduce(out.view(-1)).view(q.shape)
        out = out - q * cfg.self_half - cfg.background_ivolume * q.sum(dim=0)
        ctx.cfg, ctx.pair_range, ctx.n_pairs, ctx.plist = cfg, (lo, hi), n_pairs, plist
        ctx.save_for_backward(q, pos, d, idx, mask, dvalues)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        q, pos, d, idx, mask, dvalues = ctx.saved_tensors
        cfg = ctx.cfg
        ops, lay = cfg.ops, cfg.layout
        need_q, need_pos, need_d = ctx.needs_input_grad[:3]
        g = grad_out.contiguous()
        n, c = q.shape
        # one flat buffer so that a single all-reduce returns dL/dpositions (and dL/dcharges)
        flat = cfg.reducer.zeros(n * (3 + c) if need_q else 3 * n, q.dtype, q.device)
        g_pos = flat[: 3 * n].view(n, 3)
        g_q = flat[3 * n:].view(n, c) if need_q else None
        g_d = None
        cuda = q.is_cuda
        forked = False
        if need_q or need_d:
            lo, hi = ctx.pair_range
            if need_d:
                g_d = torch.zeros(ctx.n_pairs, dtype=q.dtype, device=q.device) if (hi - lo) != ctx.n_pairs \
                    else torch.empty(ctx.n_pairs, dtype=q.dtype, device=q.device)
            g_d_local = g_d[lo:hi] if need_d else None
            if cuda:
                main = torch.cuda.current_stream(q.device)
                side = _side_stream(q.device)
                side.wait_stream(main)
                forked = True
                with torch.cuda.stream(side):
                    ops.pair_backward(q, idx, d, None, mask, g, cfg.full_list, cfg.pair_pot,
                                      want_charges=need_q, want_pairs=need_d,
                                      grad_charges_out=g_q if need_q else None, grad_pairs_out=g_d_local)
            else:
                ops.pair_backward(q, idx, d, None, mask, g, cfg.full_list, cfg.pair_pot,
                                  want_charges=need_q, want_pairs=need_d,
                                  grad_charges_out=g_q if need_q else None, grad_pairs_out=g_d_local)
        if need_q or need_pos:
            plist = ctx.plist
            rho_g = ops.spread(pos, g, cfg.r2u, cfg.ns, cfg.nodes, cfg.method, slab=(lay.x0, lay.nxl),
                               point_list=plist)
            green = ops.make_green(scale=1.0, **cfg.green_args)
            psi = cfg.filter.apply(rho_g, green)
            if forked:
                torch.cuda.current_stream(q.device).wait_stream(_side_stream(q.device))
                forked = False
            zero_dc = torch.zeros(c, dtype=q.dtype, device=q.device)
            epi = ops.make_epilogue(g, zero_dc, cfg.half_ivolume, 0.0, 0.0,
                                    coef2=g if need_pos else None, dvalues2=dvalues,
                                    vjp_scale=cfg.half_ivolume)
            slab = (lay.x0, lay.ns[0])
            if need_pos:
                ops.gather_vjp(psi, pos, q, cfg.r2u, cfg.nodes, cfg.method, grad_positions=g_pos,
                               values_out=g_q if need_q else None, epilogue=epi, slab=slab, point_list=plist)
            else:
                ops.gather(psi, pos, cfg.r2u, cfg.nodes, cfg.method, values_out=g_q, epilogue=epi, slab=slab,
                           point_list=plist)
        if forked:
            torch.cuda.current_stream(q.device).wait_stream(_side_stream(q.device))
        flat = cfg.reducer.all_reduce(flat)
        if need_d and cfg.shard_pairs and lay.world > 1:
            # the ranks hold the same replicated pair list and each filled its own chunk: sum the
            # chunks so that every returned gradient has the same replicated meaning (positions.grad
            # through differentiable distances is then the full force on every rank).  Callers who
            # want to avoid this (P,) all-reduce hand each rank its own chunk with shard_pairs=False.
            dist.all_reduce(g_d, group=cfg.group)
        g_pos = flat[: 3 * n].view(n, 3)
        if need_q:
            g_q = flat[3 * n:].view(n, c)
            g_q = g_q - g * cfg.self_half - cfg.background_ivolume * g.sum(dim=0)
        return (g_q if need_q else None), (g_pos if need_pos else None), g_d, None, None, None


class _SlabMixin:
    """
    Adds ``process_group`` / ``transport`` / ``shard_pairs`` to a mesh calculator and routes
    ``forward`` through the slab-decomposed pipeline.  Restrictions (checked): in-kernel
    potentials (Coulomb, inverse power law), 3-D periodic, no cell / potential-parameter
    gradients, power-of-two meshes whose x and y sizes are multiples of the world size.
    """

    def _init_slab(self, process_group, transport, shard_pairs, ops):
        self.process_group = process_group
        self.transport = transport
        self.shard_pairs = shard_pairs
        self._ops = ops or _native
        self._slab_cfg = None
        self._slab_cfgs = {}

    def _world(self):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("torch.distributed has to be initialised for the slab-decomposed calculators")
        return dist.get_world_size(self.process_group), dist.get_rank(self.process_group)

    @torch.jit.unused
    def _forward_impl(self, charges, cell, positions, neighbor_indices, neighbor_distances,
                      periodic=None, node_mask=None, pair_mask=None, kvectors=None):
        if node_mask is not None or kvectors is not None:
            raise NotImplementedError("Batching not implemented for mesh-based calculators")
        if periodic is not None:
            raise NotImplementedError("the slab-decomposed calculators support 3-D periodic systems only")
        validate_parameters(charges, cell, positions, neighbor_indices, neighbor_distances,
                            periodic, pair_mask, node_mask, kvectors)
        pot = self.potential
        if pot._native_descriptor() is None:
            raise NotImplementedError(
                "the slab-decomposed calculators need an in-kernel potential "
                "(CoulombPotential or InversePowerLawPotential)")
        if cell.requires_grad or any(t.requires_grad for t in list(pot.parameters()) + list(pot.buffers())):
            raise NotImplementedError(
                "cell / potential-parameter gradients are not available in the slab-decomposed path")
        world, rank = self._world()
        if self._ops is _native and not positions.is_cuda:
            raise _native.NativeLibraryError(
                f"`positions` lives on {positions.device}; torchpme_b200 is a CUDA-only implementation "
                "(no CPU fallback). Move the inputs to a CUDA device.")
        geom = geometry_of(cell)
        ns = geom.ns_mesh(self.mesh_spacing)
        kind, exponent = pot._native_descriptor()
        smearing, prefactor = pot._scalars()
        n_channels = charges.shape[1]
        # keyed by the cell *values*: a clone of the same cell (e.g. the static buffer of a captured
        # CUDA graph) must find the same exchange buffers again
        key = (geom.cell.tobytes(), ns, kind, exponent, smearing, prefactor, pot.exclusion_radius,
               pot.exclusion_degree, self.full_neighbor_list, world, rank, n_channels, charges.dtype,
               charges.device, charges.shape[0])
        cfg = self._slab_cfgs.get(key)
        if cfg is None:
            ops = self._ops
            cfg = _SlabStepConfig()
            cfg.ops, cfg.group, cfg.shard_pairs = ops, self.process_group, self.shard_pairs
            cfg.layout = SlabLayout(ns, world, rank)
            cfg.r2u, cfg.ns = geom.r2u(ns), ns
            cfg.nodes, cfg.method = self.interpolation_nodes, _native.METHOD_ID[self._method]
            cfg.green_args = dict(kind=kind, exponent=exponent, smearing=smearing, prefactor=prefactor,
                                  recip=geom.recip, spacing=geom.spacing(ns),
                                  p3m_nodes=self.interpolation_nodes if self._method == "P3M" else 0)
            cfg.pair_pot = ops.make_pair_potential(kind, smearing, prefactor, exponent,
                                                   pot.exclusion_radius, pot.exclusion_degree)
            cfg.full_list = self.full_neighbor_list
            ivolume = 1.0 / geom.volume
            cfg.half_ivolume = 0.5 * ivolume
            cfg.self_half = 0.5 * float(pot.self_contribution())
            cfg.background_ivolume = float(pot.background_correction()) * ivolume
            cfg.filter = SlabFilter(cfg.layout, n_channels, charges.dtype, charges.device,
                                    self.process_group, self.transport, ops)
            cfg.n_atoms = charges.shape[0]
            if self.transport.startswith("p2p") and world > 1:
                cfg.reducer = make_reducer(cfg.n_atoms * (3 + n_channels), charges.dtype, charges.device,
                                           self.process_group, world, rank)
                cfg.reducer.watch(cfg.filter.exchange.error)
            else:
                cfg.reducer = TorchReducer(self.process_group, world)
            if len(self._slab_cfgs) >= 4:   # a captured graph keeps its own reference (GraphedStep)
                self._slab_cfgs.pop(next(iter(self._slab_cfgs)))
            self._slab_cfgs[key] = cfg
        self._slab_cfg = cfg
        mask_u8 = None if pair_mask is None else pair_mask.contiguous().view(torch.uint8)
        return _SlabMeshPotential.apply(charges, positions, neighbor_distances, neighbor_indices,
                                        mask_u8, cfg)


class SlabPMECalculator(_SlabMixin, PMECalculator):
    """:class:`PMECalculator` with the mesh slab-decomposed over the ranks of ``process_group``."""

    def __init__(self, potential, mesh_spacing: float, interpolation_nodes: int = 4,
                 full_neighbor_list: bool = False, process_group=None, transport: str = "nccl",
                 shard_pairs: bool = True, _ops=None):
        PMECalculator.__init__(self, potential, mesh_spacing, interpolation_nodes, full_neighbor_list)
        self._init_slab(process_group, transport, shard_pairs, _ops)


class SlabP3MCalculator(_SlabMixin, P3MCalculator):
    """:class:`P3MCalculator` with the mesh slab-decomposed over the ranks of ``process_group``."""

    def __init__(self, potential, mesh_spacing: float, interpolation_nodes: int = 4,
                 full_neighbor_list: bool = False, process_group=None, transport: str = "nccl",
                 shard_pairs: bool = True, _ops=None):
        P3MCalculator.__init__(self, potential, mesh_spacing, interpolation_nodes, full_neighbor_list)
        self._init_slab(process_group, transport, shard_pairs, _ops)
