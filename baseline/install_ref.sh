#!/bin/bash
# Offline install of the UNMODIFIED reference's `models` package into baseline/_ref (git-ignored, travels to the GPU
# box with gpurun) so that `bench.py --impl reference` times the reference's own modules.  The reference ships no
# setup.py / pyproject, and /root/reference is read-only, so the install runs from a copy under /tmp that only ADDS a
# 5-line setup.py naming the package (no reference file is edited).
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${1:-/root/reference}
TMP=$(mktemp -d /tmp/effdet_ref.XXXXXX)
cp -r "$SRC/models" "$TMP/models"
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup
setup(name='efficientdet-pytorch-reference', version='0+fbe56e5', packages=['models'],
      description='toandaominh1997/EfficientDet.Pytorch models/ (unmodified), packaged for the reference bench arm')
PY
rm -rf "$REPO/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$REPO/baseline/_ref" "$TMP" 2>&1 | tail -3
rm -rf "$TMP"
ls "$REPO/baseline/_ref"
