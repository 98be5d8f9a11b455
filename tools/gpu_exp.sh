#!/bin/bash
for f in old "" old "" old ""; do
lib=lako_b200/_lib/liblako_b200${f:+_$f}.so
echo "== $lib"
LAKO_B200_LIB=$PWD/$lib timeout 600 python tools/dropout_cost.py 2>&1 | tail -2 | head -1 | cut -c1-135
done
