#!/bin/bash
# one gpurun call: cycle counts of build variants of the two-tile kernel (tools/t2_exp.py), ncu --set full of the stash-writing form
mkdir -p gpurun_out
{
for v in "" _preload _sincos1 _both ""; do echo "== variant '$v'"; GBNERF_LIB=gb-nerf_b200/libgbnerf_exp$v.so T2_EXP_QUICK=1 timeout 120 python tools/t2_exp.py | sed -n 2,4p; done
} > gpurun_out/t2_exp.log 2>&1
cat gpurun_out/t2_exp.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_t2 -s 1 -c 1 -f -o gpurun_out/r2_mlp_t2_stash python tools/train_kernels_once.py > gpurun_out/ncu_stash.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_stash.log
