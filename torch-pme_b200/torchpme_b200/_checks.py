"""
Input validation for ``Calculator.forward``.

The reference raises ``ValueError`` / ``TypeError`` with specific texts
(``src/torchpme/_utils.py:4-170``) that downstream tests match with regular expressions,
so the texts are part of the interface and are reproduced here; the checks themselves are
organised per argument instead of as one long function.
"""

from __future__ import annotations

import torch

_REF = "the `positions` class"


def _same_device(name, tensor, device):
    if tensor.device != device:
        raise ValueError(
            f"device of `{name}` ({tensor.device}) must be same as that of {_REF} ({device})"
        )


def _same_dtype(name, tensor, dtype):
    if tensor.dtype != dtype:
        raise TypeError(f"type of `{name}` ({tensor.dtype}) must be same as that of {_REF} ({dtype})")


def _bool_mask(name, mask, expected_shape, device, shape_message):
    if mask.shape != expected_shape:
        raise ValueError(shape_message)
    if mask.device != device:
        raise ValueError(
            f"device of `{name}` ({mask.device}) must be same as that of {_REF} ({device})"
        )
    if mask.dtype != torch.bool:
        raise TypeError(f"type of `{name}` ({mask.dtype}) must be torch.bool")


def validate_parameters(charges, cell, positions, neighbor_indices, neighbor_distances,
                        periodic=None, pair_mask=None, node_mask=None, kvectors=None) -> None:
    dtype, device = positions.dtype, positions.device
    n_atoms = positions.shape[-2]

    if list(positions.shape) != [n_atoms, 3]:
        raise ValueError(
            "`positions` must be a tensor with shape [n_atoms, 3], got tensor with shape "
            f"{list(positions.shape)}"
        )

    if list(cell.shape) != [3, 3]:
        raise ValueError(
            f"`cell` must be a tensor with shape [3, 3], got tensor with shape {list(cell.shape)}"
        )
    _same_dtype("cell", cell, dtype)
    _same_device("cell", cell, device)

    if charges.dim() != 2:
        raise ValueError(
            f"`charges` must be a 2-dimensional tensor, got tensor with {charges.dim()} "
            f"dimension(s) and shape {list(charges.shape)}"
        )
    if charges.shape[0] != n_atoms:
        raise ValueError(
            "`charges` must be a tensor with shape [n_atoms, n_channels], with `n_atoms` being "
            f"the same as the variable `positions`. Got tensor with shape {list(charges.shape)} "
            f"where positions contains {len(positions)} atoms"
        )
    _same_dtype("charges", charges, dtype)
    _same_device("charges", charges, device)

    if neighbor_indices.shape[1] != 2:
        raise ValueError(
            "neighbor_indices is expected to have shape [num_neighbors, 2], but got "
            f"{list(neighbor_indices.shape)} for one structure"
        )
    _same_device("neighbor_indices", neighbor_indices, device)
    if neighbor_distances.shape != neighbor_indices[:, 0].shape:
        raise ValueError(
            "`neighbor_indices` and `neighbor_distances` need to have shapes [num_neighbors, 2] "
            f"and [num_neighbors], but got {list(neighbor_indices.shape)} and "
            f"{list(neighbor_distances.shape)}"
        )
    _same_device("neighbor_distances", neighbor_distances, device)
    if neighbor_distances.dtype != dtype:
        raise TypeError(
            f"type of `neighbor_distances` ({neighbor_distances.dtype}) must be same as that of "
            f"{_REF} ({dtype})"
        )

    if periodic is not None:
        if periodic.shape != (3,):
            raise ValueError(
                f"`periodic` must be a tensor of shape (3,), got tensor with shape {list(periodic.shape)}"
            )
        _same_device("periodic", periodic, device)

    if pair_mask is not None:
        _bool_mask(
            "pair_mask", pair_mask, neighbor_indices[:, 0].shape, device,
            "`pair_mask` must have the same shape as the number of neighbors, got tensor with "
            f"shape {list(pair_mask.shape)} while the number of neighbors is {neighbor_indices.shape[0]}",
        )

    if node_mask is not None:
        _bool_mask(
            "node_mask", node_mask, (n_atoms,), device,
            f"`node_mask` must have shape [n_atoms], got tensor with shape {list(node_mask.shape)} "
            f"where n_atoms is {n_atoms}",
        )

    if kvectors is not None:
        if kvectors.shape[1] != 3:
            raise ValueError(
                f"`kvectors` must be a tensor of shape [n_kvecs, 3], got tensor with shape {list(kvectors.shape)}"
            )
        _same_device("kvectors", kvectors, device)
        _same_dtype("kvectors", kvectors, dtype)
