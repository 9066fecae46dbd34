#!/bin/bash
# ORACLE recipe: install the UNMODIFIED reference (pure Python + Numba) from /root/reference into the git-ignored
# oracle/_ref/ so that bench.py's `--impl reference` / cpu_baseline legs can time the reference's own CPU path on
# the GPU box (where /root/reference does not exist).  Nothing under oracle/_ref is tracked or imported by the
# product.  The reference tree is read-only: pip builds from a copy under /tmp.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/ramba" ] || { echo "no reference at $REF: keeping whatever oracle/_ref holds"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
echo "installed the reference into $HERE/_ref"
