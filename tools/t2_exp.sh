#!/bin/bash
# One gpurun call: SM cycles per tile pair of the two-tile MLP kernel (tools/t2_exp.py) for the exp build and for any
# build variants given as arguments (csrc/build.py --variant=NAME:MACRO builds gb-nerf_b200/libgbnerf_exp_NAME.so):
#     bash tools/t2_exp.sh [NAME ...]          e.g. after  python gb-nerf_b200/csrc/build.py --variant=foo:GBN_T2_FOO
# Without arguments: the full sweep of the exp build (turn point, ablations), then a wide layer repeated twice, then the
# staggered mode.
mkdir -p gpurun_out
L=gb-nerf_b200/libgbnerf_exp
{
if [ $# -gt 0 ]; then
  for v in "" "$@" ""; do
    echo "== variant '${v:-exp}'"
    GBNERF_LIB=$L${v:+_$v}.so T2_EXP_QUICK=1 timeout 120 python tools/t2_exp.py | sed -n 2,4p
  done
else
  for x in 0 2; do GBNERF_LIB=$L.so GBNERF_T2_DBG_EXTRA=$x timeout 120 python tools/t2_exp.py; done
  for x in 0 2; do GBNERF_LIB=$L.so GBNERF_T2_MODE=2 GBNERF_T2_DBG_EXTRA=$x timeout 120 python tools/t2_exp.py; done
fi
} > gpurun_out/t2_exp.log 2>&1
cut -c1-200 gpurun_out/t2_exp.log
