#!/bin/bash
OUT=${1:-gpurun_out/chaos4}
mkdir -p "$OUT"
export GBNERF_LIB=$PWD/gb-nerf_b200/libgbnerf_diag.so
for seed in 12345 777 31337; do
  HUNT_GATECHECK=1 GBNERF_TS_CHAOS=$seed GBNERF_TS_FIX=16 timeout 200 python tools/dgrad_hunt.py 300 1024 128 warm > "$OUT/order_$seed.log" 2>&1
  grep -E "ORDER|order check|RESULT" "$OUT/order_$seed.log" | head -12
done
HUNT_GATECHECK=1 GBNERF_TS_FIX=16 timeout 200 python tools/dgrad_hunt.py 300 1024 128 warm > "$OUT/order_nochaos.log" 2>&1
grep -E "ORDER|order check|RESULT" "$OUT/order_nochaos.log" | head -12
