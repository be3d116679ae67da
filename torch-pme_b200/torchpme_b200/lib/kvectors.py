"""``torchpme.lib.kvectors``-compatible module path."""
from ..mesh import generate_kvectors_for_ewald, generate_kvectors_for_mesh, get_ns_mesh  # noqa: F401
