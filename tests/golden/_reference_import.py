"""
Import the *unmodified* reference (read-only at /root/reference) in the build container.

The only thing the source tree lacks is the setuptools-scm generated
``torchpme/_version.py``; we pre-seed ``sys.modules`` with a stub for it instead of
writing into the tree.  Used by make_golden.py and by tests that are skipped when
/root/reference is absent (it is absent on the GPU box).
"""
import os
import sys
import types

# the read-only tree in the build container, else the copy `oracle/make_ref.sh` made of it (which
# travels to the GPU box; see __graft_entry__.build)
_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = ("/root/reference/src", os.path.join(os.path.dirname(os.path.dirname(_HERE)), "oracle", "_ref"))
REFERENCE_SRC = next((c for c in _CANDIDATES if os.path.isdir(os.path.join(c, "torchpme"))), _CANDIDATES[0])


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "torchpme"))


def import_reference():
    if "torchpme" in sys.modules and getattr(sys.modules["torchpme"], "__graft_ref__", False):
        return sys.modules["torchpme"]
    if not available():
        raise ImportError("reference tree not present")
    stub = types.ModuleType("torchpme._version")
    stub.__version__ = "0.0.0+ref"
    stub.__version_tuple__ = (0, 0, 0)
    sys.modules["torchpme._version"] = stub
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import torchpme  # noqa

    torchpme.__graft_ref__ = True
    return torchpme
