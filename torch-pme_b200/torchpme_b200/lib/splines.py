"""``torchpme.lib.splines``-compatible module path."""
from ..splines import (  # noqa: F401
    CubicSpline,
    CubicSplineReciprocal,
    compute_second_derivatives,
    compute_spline_ft,
)
