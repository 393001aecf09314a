#!/bin/bash
mkdir -p gpurun_out
L=gb-nerf_b200/libgbnerf_exp.so
{
for x in 0 2; do GBNERF_LIB=$L GBNERF_T2_MODE=2 GBNERF_T2_DBG_EXTRA=$x timeout 120 python tools/t2_exp.py; done
} > gpurun_out/t2_exp.log 2>&1
tail -60 gpurun_out/t2_exp.log | cut -c1-200
