"""``torchpme.lib.kspace_filter``-compatible module path."""
from ..mesh import KSpaceFilter, KSpaceKernel, P3MKSpaceFilter  # noqa: F401
