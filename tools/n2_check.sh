#!/bin/bash
# one gpurun --gpus 2 call: the two-rank GPU parity test, then the N = 2 bench line (ONE frame / ONE batch sharded over 2 ranks)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_mlp_t2.py -m gpu -x -q > gpurun_out/pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-tcnn > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"
tail -c 400 gpurun_out/bench_n2.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
t = d["train_step"]
print("N=2 rays/s", d["value"], "ms/frame", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "| train ms", t["ms_per_step"], "eager", t["eager_dropin"]["ms_per_step"], "| stress", d.get("stress", {}).get("value"))
P
