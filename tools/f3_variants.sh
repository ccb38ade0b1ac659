#!/bin/bash
# Per-layer device times of one 1-hour pass under a list of environment settings (experiment knobs of the kernels):
#   tools/f3_variants.sh "X=0" "BD_F3_NACC=256" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/prof_pass.py fp16x3 2 2>&1 | tail -2
done
