#!/bin/bash
# Closes the V2 parity pin wherever the UPSTREAM tools exist (they do not in the build image: no network, no binaries).
# Needs `draco_decoder` (google/draco >= 1.4.3) and `basisu` (BinomialLLC/basis_universal >= 1.16) on $PATH.  Writes, for every committed
# fixture under tests/golden/liam/:
#   tests/golden/upstream/<name>.obj            draco_decoder's OBJ of the .drc   (positions / uvs / normals per face corner)
#   tests/golden/upstream/<name>_layerNNNN.png  basisu's RGBA32 unpack of every layer of the .ktx2
# tests/test_upstream_golden.py compares the oracle's (CPU) and the library's (GPU) outputs with these files when they are present.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"; out="$here/upstream"; mkdir -p "$out"
command -v draco_decoder >/dev/null || { echo "draco_decoder not on PATH"; exit 2; }
command -v basisu >/dev/null || { echo "basisu not on PATH"; exit 2; }
for f in "$here"/liam/*.drc; do
    n="$(basename "$f" .drc)"
    draco_decoder -i "$f" -o "$out/$n.obj" >/dev/null
done
tmp="$(mktemp -d)"
for f in "$here"/liam/*.ktx2; do
    n="$(basename "$f" .ktx2)"
    ( cd "$tmp" && rm -f ./* && basisu -unpack -no_ktx -format_only 13 -file "$f" >/dev/null )      # 13 = cTFRGBA32
    i=0
    for p in $(ls "$tmp"/*RGBA32*.png 2>/dev/null | sort); do
        cp "$p" "$out/${n}_layer$(printf '%04d' $i).png"; i=$((i + 1))
    done
    [ "$i" -gt 0 ] || { echo "basisu wrote no RGBA32 images for $f (file naming differs in this basisu version: copy them by hand)"; exit 3; }
done
rm -rf "$tmp"
echo "upstream goldens written to $out"
