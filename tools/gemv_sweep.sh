#!/bin/bash
# development sweep of the GeMV launch parameters (CTAs per SM, rows per group, x staging)
for cfg in "5,4,0" "6,4,0" "4,8,0" "5,8,0" "6,8,0" "8,4,0"; do
  HISPMV_GEMV=$cfg python tools/sweep.py --configs g3a,g8192,gtall --out /tmp/g.json 2>&1 | grep "gemv" | sed "s/^/cfg=$cfg /"
done
