"""
torchpme_b200 -- B200-native PME / P3M long-range calculators behind the torch-pme API.

    import torchpme_b200 as torchpme
    calc = torchpme.P3MCalculator(torchpme.CoulombPotential(smearing=1.2), mesh_spacing=0.6)
    V = calc(charges, cell, positions, neighbor_indices, neighbor_distances)   # CUDA tensors

Only the hot path of the reference is provided: PME / P3M calculators,
Coulomb and inverse-power-law potentials, the mesh interpolator and k-space filter blocks.
``torchpme_b200.distributed`` holds the slab-decomposed (one system over several GPUs) variants,
``torchpme_b200.GraphedStep`` the CUDA-graph capture of an energy + forces step.
"""

from . import calculators, graphs, lib, mesh, potentials, prefactors, tuning  # noqa: F401
from ._native import NativeLibraryError, library_path  # noqa: F401
from .calculators import Calculator, P3MCalculator, PMECalculator
from .graphs import GraphedPositionsStep, GraphedStep  # noqa: F401
from .mesh import device_cell, set_nan_check  # noqa: F401
from .potentials import (CombinedPotential, CoulombPotential, InversePowerLawPotential, Potential,
                         SplinePotential)

__version__ = "0.1.0"
__all__ = [
    "Calculator",
    "P3MCalculator",
    "PMECalculator",
    "CoulombPotential",
    "InversePowerLawPotential",
    "SplinePotential",
    "CombinedPotential",
    "Potential",
]
