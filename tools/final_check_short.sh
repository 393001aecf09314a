#!/bin/bash
# short round-end check on one B200 (a few minutes): GPU tests, smoke, the default bench line
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$? ($(( $(date +%s) - t0 )) s)"
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_final.log; tail -3 gpurun_out/smoke_final.log; tail -c 600 gpurun_out/bench_final.json
