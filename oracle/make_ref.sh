#!/usr/bin/env bash
# Test/measurement infrastructure, not product code.
# Stages the UNMODIFIED Python reference (ashispati/InpaintNet) for the CPU arm of bench.py and the oracle checks:
# the reference is pure Python, so "building" it is copying its *.py files, byte for byte, from the read-only
# reference tree into oracle/_ref/ (git-ignored: never enters the history; NOT gpurun-ignored: it travels to the
# GPU box next to the built .so, where /root/reference does not exist).  Nothing under inpaintnet_b200/ imports it.
#   usage: oracle/make_ref.sh [reference_root]       (default /root/reference)
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
[ -d "$SRC/MeasureVAE" ] || { echo "make_ref: no reference tree at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST"
(cd "$SRC" && find . -name '*.py' -not -path './.git/*' -print0 | xargs -0 -I{} cp --parents {} "$DST/")
(cd "$SRC" && find . -name '*.py' -not -path './.git/*' -print0 | sort -z | xargs -0 sha256sum) > "$DST/SHA256SUMS"
echo "make_ref: $(find "$DST" -name '*.py' | wc -l) reference files staged in $DST"
