#!/usr/bin/env bash
# Installs the UNMODIFIED reference package (whisper_sae) into baseline/_ref for `bench.py --impl
# reference` (git-ignored; travels to the GPU box with the gpurun snapshot).  The reference builds
# with hatchling, which is neither installed nor in /opt/wheelhouse, and there is no network; so the
# install runs from a copy under /tmp whose [build-system] table alone is switched to setuptools
# (present in the image).  Every file under src/whisper_sae is installed byte for byte - the diff at
# the end proves it.  Dependencies are the image's own (torch, pydantic, pyyaml, rich): --no-deps.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${1:-/root/reference}
TMP=$(mktemp -d /tmp/wsae_ref_XXXX)
cp -r "$REF"/src "$REF"/pyproject.toml "$REF"/README.md "$TMP"/
python - "$TMP/pyproject.toml" <<'PY'
import re, sys
p = sys.argv[1]
s = open(p).read()
s = re.sub(r"\[build-system\].*?(?=\n\[)", '[build-system]\nrequires = ["setuptools"]\nbuild-backend = "setuptools.build_meta"\n', s, flags=re.S)
s = re.sub(r"\[tool\.hatch\.build\.targets\.wheel\].*?(?=\n\[)", '[tool.setuptools.packages.find]\nwhere = ["src"]\n', s, flags=re.S)
open(p, "w").write(s)
PY
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP" 2>&1 | tail -3
diff -r -x __pycache__ "$REF/src/whisper_sae" "$ROOT/baseline/_ref/whisper_sae" && echo "baseline/_ref/whisper_sae == reference sources (byte-identical)"
rm -rf "$TMP"
