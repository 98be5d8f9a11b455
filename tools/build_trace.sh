#!/bin/bash
# current tree + -DLAKO_TRACE (clock64 phase stamps in the attention backward) -> _lib/liblako_b200_trace.so
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
objs=()
for f in "$root"/lako_b200/csrc/*.cu; do
  o="$tmp/$(basename "${f%.cu}").o"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${LAKO_TRACE_DEF--DLAKO_TRACE} $LAKO_EXTRA_DEFS -I "$root/include" -c "$f" -o "$o" &
  objs+=("$o")
done
wait
nvcc -shared -arch=sm_100a -o "$root/lako_b200/_lib/${LAKO_OUT:-liblako_b200_trace.so}" "${objs[@]}" -lcudart
rm -rf "$tmp"
