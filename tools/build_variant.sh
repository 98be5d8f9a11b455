#!/bin/bash
# Builds the library from a git revision's csrc into lako_b200/_lib/liblako_b200_<tag>.so for same-box A/B timing:
#   tools/build_variant.sh HEAD base ; LAKO_B200_LIB=lako_b200/_lib/liblako_b200_base.so python bench.py ...
set -e
rev=${1:-HEAD}; tag=${2:-base}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$rev" lako_b200/csrc include | tar -x -C "$tmp"
objs=()
for f in "$tmp"/lako_b200/csrc/*.cu; do
  o="$tmp/$(basename "${f%.cu}").o"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I "$tmp/include" -c "$f" -o "$o" &
  objs+=("$o")
done
wait
nvcc -shared -arch=sm_100a -o "$root/lako_b200/_lib/liblako_b200_$tag.so" "${objs[@]}" -lcudart
rm -rf "$tmp"
echo "$root/lako_b200/_lib/liblako_b200_$tag.so"
