#!/bin/bash
# timing experiments on the fused forward kernel alone (events around the C call), PointNeXt-XL layer shapes
for cfg in "ABL=0" "ABL=1" "ABL=2" "ABL=4" "ABL=7"; do
  unset AMC3D_FUSED_ABLATE
  for kv in $cfg; do k=${kv%%=*}; v=${kv##*=}; case $k in ABL) export AMC3D_FUSED_ABLATE=$v;; esac; done
  echo "== $cfg"
  python tools/diag/fused_prof.py tf32 2>&1 | grep -E "tf32:|fused_sa_forward|transpose"
done
