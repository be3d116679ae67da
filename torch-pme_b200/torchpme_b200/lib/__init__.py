"""``torchpme.lib``-compatible namespace for the mesh building blocks."""
from ..mesh import (  # noqa: F401
    KSpaceFilter,
    KSpaceKernel,
    MeshInterpolator,
    P3MKSpaceFilter,
    generate_kvectors_for_ewald,
    generate_kvectors_for_mesh,
    get_ns_mesh,
)
from ..potentials import exp1, gamma, gammaincc_over_powerlaw  # noqa: F401
from ..splines import (  # noqa: F401
    CubicSpline,
    CubicSplineReciprocal,
    compute_second_derivatives,
    compute_spline_ft,
)
from . import kspace_filter, kvectors, math, mesh_interpolator, splines  # noqa: F401,E402
