"""``torchpme.lib.mesh_interpolator``-compatible module path."""
from ..mesh import MeshInterpolator  # noqa: F401
