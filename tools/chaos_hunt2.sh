#!/bin/bash
# Second chaos round: which role's delays expose the gate hole, what the wrong gate chunks are, which candidate fix closes it.
N=${1:-1500}
OUT=${2:-gpurun_out/chaos2}
mkdir -p "$OUT"
export GBNERF_LIB=$PWD/gb-nerf_b200/libgbnerf_diag.so
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name" | tee -a "$OUT/summary.txt"
  env "${envs[@]}" timeout 300 python tools/dgrad_hunt.py "$@" > "$OUT/$name.log" 2>&1; echo "rc=$?" >> "$OUT/$name.log"
  grep -E "^RESULT|^rc=|gate check|   step " "$OUT/$name.log" | tail -12 | tee -a "$OUT/summary.txt"; }
run gatecheck HUNT_GATECHECK=1 GBNERF_TS_CHAOS=12345 -- $N 1024 128 warm
for r in 1 2 4 8 16 32; do run roles_$r GBNERF_TS_CHAOS=12345 GBNERF_TS_CHAOS_ROLES=$r -- $N 1024 128 warm; done
run fix1 GBNERF_TS_CHAOS=12345 GBNERF_TS_FIX=1 -- $N 1024 128 warm
run fix2 GBNERF_TS_CHAOS=12345 GBNERF_TS_FIX=2 -- $N 1024 128 warm
run fix3 GBNERF_TS_CHAOS=12345 GBNERF_TS_FIX=3 -- $N 1024 128 warm
