#!/bin/bash
# one gpurun call: GPU tests, then the bench with the training forward on the two-tile kernel (default) and on the one-tile kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_stash.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_stash.log
timeout 300 python bench.py --no-cpu --no-tcnn --no-stress > gpurun_out/bench_stash_t2.json 2> gpurun_out/bench_stash_t2.err; echo "bench(t2 stash) rc=$?"
GBNERF_T2_STASH=0 timeout 300 python bench.py --no-cpu --no-tcnn --no-stress > gpurun_out/bench_stash_one.json 2> gpurun_out/bench_stash_one.err; echo "bench(one-tile stash) rc=$?"
python - <<'P'
import json
for n in ("t2", "one"):
    try:
        d = json.loads(open(f"gpurun_out/bench_stash_{n}.json").read().strip().splitlines()[-1])
        t = d["train_step"]
        print(n, "train ms", t["ms_per_step"], t["passes_ms_per_step"], "eager", t["eager_dropin"]["ms_per_step"], t["eager_dropin"]["mlp_kernels_ms_per_step"], "wd", t["watchdog_words"], "loss", t["loss_after"], "| frame ms", d["ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
P
