#!/usr/bin/env bash
# Installs the UNMODIFIED reference (raphaelsty/mkb, mounted read-only at /root/reference) into the
# git-ignored baseline/_ref/ so that `bench.py --impl reference` and tests/golden/make_golden.py can
# import it, here and on the GPU box (baseline/_ref travels with gpurun; it never enters git history).
#   bash baseline/install_ref.sh
# `river` (reference dependency, absent from the wheelhouse) is used by the reference for two
# accumulators only; baseline/river_shim/ (ours, committed) provides them.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference}
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/reference"            # setuptools writes build/ and egg-info into the source tree
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/reference" 2>&1 | tail -3
cp -r "$HERE/river_shim/river" "$HERE/_ref/river"
rm -rf "$TMP"
python - <<PY
import sys; sys.path.insert(0, "$HERE/_ref")
import mkb, os
print("installed mkb", mkb.__version__, "->", os.path.dirname(mkb.__file__))
PY
