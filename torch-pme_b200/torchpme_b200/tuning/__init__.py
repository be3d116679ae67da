"""
Timing harness of the tuning layer (reference: ``src/torchpme/tuning/tuner.py:283-373``).

The reference ranks hyper-parameter candidates (smearing / mesh spacing / cutoff) by the wall-clock
time of a few forward + backward calls, measured with ``time.monotonic`` and WITHOUT a device
synchronisation: with a hot path of 70 - 900 us per step on a B200 that measures the Python launch
overhead, not the kernels, and the ranking becomes noise.
:class:`TuningTimings` keeps the reference's constructor and ``forward(calculator) -> seconds``
contract and times on the device instead: CUDA events on the current stream around every repeat
(synchronised before and after), the L2 optionally flushed between repeats; CPU tensors are timed
with ``time.perf_counter`` like the reference.  The error-bound formulas and the grid search that
consume these timings are host-side code outside the hot path and are not reproduced here.
"""

from __future__ import annotations

import time

import torch

from .._checks import validate_parameters

__all__ = ["TuningTimings"]


class TuningTimings(torch.nn.Module):
    """
    Average execution time of ``calculator.forward`` (+ backward of ``result.sum()``) on one structure.

    Same arguments as the reference class; ``flush_l2`` (extra) writes a buffer larger than the L2
    before every timed repeat so that candidates are compared cold, as an MD step would see them.
    """

    def __init__(self, charges: torch.Tensor, cell: torch.Tensor, positions: torch.Tensor,
                 neighbor_indices: torch.Tensor, neighbor_distances: torch.Tensor, n_repeat: int = 4,
                 n_warmup: int = 4, run_backward: bool | None = True, flush_l2: bool = False):
        super().__init__()
        validate_parameters(charges, cell, positions, neighbor_indices, neighbor_distances, None, None, None, None)
        self.charges = charges
        self.cell = cell
        self.positions = positions
        self.n_repeat = n_repeat
        self.n_warmup = n_warmup
        self.run_backward = run_backward
        self.neighbor_indices = neighbor_indices
        self.neighbor_distances = neighbor_distances
        self._flush = None
        if flush_l2 and positions.is_cuda:
            self._flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=positions.device)

    def _one(self, calculator):
        positions = self.positions.clone()
        cell = self.cell.clone()
        charges = self.charges.clone()
        if self.run_backward:       # like the reference: no gradient w.r.t. the distances
            positions.requires_grad_(True)
            cell.requires_grad_(True)
            charges.requires_grad_(True)
        return positions, cell, charges

    def forward(self, calculator: torch.nn.Module) -> float:
        """average seconds per call (device time for CUDA tensors)"""
        on_gpu = self.positions.is_cuda
        total = 0.0
        for it in range(self.n_repeat + self.n_warmup):
            if it == self.n_warmup:
                total = 0.0
            positions, cell, charges = self._one(calculator)
            if on_gpu:
                if self._flush is not None:
                    self._flush.zero_()
                start = torch.cuda.Event(enable_timing=True)
                stop = torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(self.positions.device)
                start.record()
            else:
                t0 = time.perf_counter()
            result = calculator.forward(positions=positions, charges=charges, cell=cell,
                                        neighbor_indices=self.neighbor_indices,
                                        neighbor_distances=self.neighbor_distances)
            value = result.sum()
            if self.run_backward:
                value.backward(retain_graph=True)
            if on_gpu:
                stop.record()
                torch.cuda.synchronize(self.positions.device)
                total += start.elapsed_time(stop) * 1e-3
            else:
                total += time.perf_counter() - t0
        return total / self.n_repeat
