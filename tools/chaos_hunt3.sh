#!/bin/bash
N=${1:-3000}
OUT=${2:-gpurun_out/chaos3}
mkdir -p "$OUT"
export GBNERF_LIB=$PWD/gb-nerf_b200/libgbnerf_diag.so
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name" | tee -a "$OUT/summary.txt"
  env "${envs[@]}" timeout 300 python tools/dgrad_hunt.py "$@" > "$OUT/$name.log" 2>&1; echo "rc=$?" >> "$OUT/$name.log"
  grep -E "^RESULT|^rc=|gate check|   step " "$OUT/$name.log" | tail -12 | tee -a "$OUT/summary.txt"; }
run gatecheck_after HUNT_GATECHECK=1 GBNERF_TS_CHAOS=12345 -- $N 1024 128 warm
run base GBNERF_TS_CHAOS=12345 -- $N 1024 128 warm
run fix1 GBNERF_TS_CHAOS=12345 GBNERF_TS_FIX=1 -- $((N * 2)) 1024 128 warm
run fix1_cold GBNERF_TS_CHAOS=777 GBNERF_TS_FIX=1 -- $N 1024 128 cold
run fix4_sleep GBNERF_TS_CHAOS=12345 GBNERF_TS_FIX=4 -- $N 1024 128 warm
run fix8_producer_fence GBNERF_TS_CHAOS=12345 GBNERF_TS_FIX=8 -- $N 1024 128 warm
run fix1_full GBNERF_TS_CHAOS=4242 GBNERF_TS_FIX=1 -- $((N / 2)) 1024 128 full
