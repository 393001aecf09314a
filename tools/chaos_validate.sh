#!/bin/bash
# Validation of the gate-staging fix: diagnostic library, chaos on every role.  The product kernels carry the proxy fence;
# GBNERF_TS_FIX=32 switches it off again (diagnostic library only), which must bring the failures back.
N=${1:-4000}
OUT=${2:-gpurun_out/chaos_validate}
mkdir -p "$OUT"
export GBNERF_LIB=$PWD/gb-nerf_b200/libgbnerf_diag.so
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name" | tee -a "$OUT/summary.txt"
  env "${envs[@]}" timeout 300 python tools/dgrad_hunt.py "$@" > "$OUT/$name.log" 2>&1; echo "rc=$?" >> "$OUT/$name.log"
  grep -E "^RESULT|^rc=|FORWARD" "$OUT/$name.log" | tail -3 | tee -a "$OUT/summary.txt"; }
run fence_chaos_warm GBNERF_TS_CHAOS=12345 -- $N 1024 128 warm
run fence_chaos_cold GBNERF_TS_CHAOS=777 -- $N 1024 128 cold
run fence_chaos_full GBNERF_TS_CHAOS=4242 -- $((N / 2)) 1024 128 full
run fence_chaos_trainshape GBNERF_TS_CHAOS=99 -- $((N / 8)) 4096 192 warm
run nofence_chaos_warm GBNERF_TS_CHAOS=12345 GBNERF_TS_FIX=32 -- $N 1024 128 warm
run nofence_chaos_full GBNERF_TS_CHAOS=4242 GBNERF_TS_FIX=32 -- $((N / 2)) 1024 128 full
