#!/bin/sh
# Builds oracle/_ref: the UNMODIFIED reference package, so that `bench.py --impl reference` and the
# CPU baseline can time the reference's own Sequential.fit() on the GPU box (where /root/reference
# does not exist).  The reference is pure Python (no setup.py / pyproject.toml, nothing to compile):
# "building" it is copying its .py files, byte for byte, into a package directory.
#   * test infrastructure only: nothing under shinnosuke_b200/ imports it;
#   * oracle/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored,
#     so it travels to the GPU box like the built .so files.
# Usage: sh oracle/make_ref.sh [reference_dir]     (default /root/reference)
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref/shinnosuke_ref"
[ -f "$SRC/models.py" ] || { echo "make_ref: no reference at $SRC" >&2; exit 1; }
rm -rf "$HERE/_ref"
mkdir -p "$DST"
(cd "$SRC" && find . -name '*.py' -not -path './docs/*' | while read -r f; do
    mkdir -p "$DST/$(dirname "$f")"
    cp "$f" "$DST/$f"
done)
(cd "$SRC" && find . -name '*.py' -not -path './docs/*' | sort | xargs sha256sum) > "$HERE/_ref/SHA256SUMS"
echo "make_ref: $(wc -l < "$HERE/_ref/SHA256SUMS") files -> $DST"
