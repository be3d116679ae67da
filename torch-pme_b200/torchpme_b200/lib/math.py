"""``torchpme.lib.math``-compatible module path."""
from ..potentials import exp1, gamma, gammaincc_over_powerlaw  # noqa: F401
