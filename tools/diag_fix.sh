#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
export STEP_SYNC=
echo "== early + split (8 runs)"
STRESS_RANKS=1,2,7,1,2,7,0,3 GBNERF_TS_BWD_EARLY=1 GBNERF_TS_SPLIT=1 timeout 300 python tools/train_stress.py 40
echo "== early only (6 runs)"
STRESS_RANKS=7,2,1,7,2,1 GBNERF_TS_BWD_EARLY=1 timeout 300 python tools/train_stress.py 40
echo "== default (2 runs)"
STRESS_RANKS=7,1 timeout 300 python tools/train_stress.py 40
