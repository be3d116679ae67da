#!/bin/bash
# Makes the UNMODIFIED reference importable as  oracle/_ref/torchpme  (SURVEY.md appendix B):
# a plain copy of /root/reference/src/torchpme plus the two-line _version.py that its build system
# (setuptools-scm) would generate.  oracle/_ref/ is git-ignored (no reference sources in history) but
# NOT gpurun-ignored, so the copy travels to the GPU box, where bench.py --impl reference and the
# cpu_baseline / reference_cuda legs time it.  Test / bench infrastructure only: nothing under
# torch-pme_b200/ imports it.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
src="${1:-/root/reference/src/torchpme}"
if [ ! -d "$src" ]; then
  echo "make_ref: $src not found (GPU box: the prebuilt copy is used)"; exit 0
fi
rm -rf "$here/_ref"
mkdir -p "$here/_ref"
cp -r "$src" "$here/_ref/torchpme"
find "$here/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
printf '__version__ = "0.0.0+ref"\n__version_tuple__ = (0, 0, 0)\n' > "$here/_ref/torchpme/_version.py"
echo "make_ref: $(find "$here/_ref/torchpme" -name '*.py' | wc -l) files -> $here/_ref/torchpme"
