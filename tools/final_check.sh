#!/bin/bash
# round-end check on one B200: GPU tests, smoke, bench line, 8-rank training stress (every rank's data, 40 steps each)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
timeout 300 python tools/train_stress.py 40 > gpurun_out/stress_final.log 2>&1; echo "stress rc=$?"
tail -3 gpurun_out/pytest_final.log; tail -3 gpurun_out/smoke_final.log; tail -12 gpurun_out/stress_final.log
