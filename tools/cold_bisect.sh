#!/bin/bash
# First job of the next round: which plan ingredient does the cold-cache nondeterminism of the dgrad
# G stash need?  Each line: variant -> "<k> of <N> repeats differed" (+ the first differing (tile, block) lists).
N=${1:-300}
run() { echo "== $1"; shift; env "$@" timeout 300 python tools/cold_repeat.py $N 2>&1 | tail -4; }
run "default (early acc1 release + split hand-over + ring guard)" X=1
run "no split hand-over" GBNERF_TS_SPLIT=0
run "late acc1 release in the dgrad program" GBNERF_TS_BWD_EARLY=0
run "plain dgrad plan (late release, no split)" GBNERF_TS_BWD_EARLY=0 GBNERF_TS_SPLIT=0
run "forward with late acc1 release too" GBNERF_TS_EARLY=0 GBNERF_TS_BWD_EARLY=0 GBNERF_TS_SPLIT=0
run "shared-memory-operand kernels (GBNERF_MLP=ss) as the control" GBNERF_MLP=ss
