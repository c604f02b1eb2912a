#!/bin/bash
# Install the UNMODIFIED reference package for the bench's comparator legs (cpu_baseline_reference / eager_gpu).
# The reference is pure Python (no setup.py / pyproject: `pip install /root/reference` has nothing to build), so the
# install is a copy of its package directory under a name that does not shadow ours.  baseline/_ref/ is git-ignored
# (never committed) but NOT gpurun-ignored: it travels to the GPU box with the snapshot.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}/kvcompress"
[ -d "$SRC" ] || { echo "install_ref: $SRC not found (the GPU box has no /root/reference; it uses the prebuilt copy)"; exit 0; }
mkdir -p "$ROOT/baseline/_ref"
rm -rf "$ROOT/baseline/_ref/kvcompress_ref"
cp -r "$SRC" "$ROOT/baseline/_ref/kvcompress_ref"
find "$ROOT/baseline/_ref" -name __pycache__ -type d -exec rm -rf {} +
echo "installed $(find "$ROOT/baseline/_ref/kvcompress_ref" -name '*.py' | wc -l) reference files into baseline/_ref/kvcompress_ref"
