#!/bin/bash
# ncu launch list of the drop-in (eager) training step with the final kernels: shares of fwd+stash / dgrad / wgrad / the rest
mkdir -p gpurun_out
timeout 60 python tools/train_step.py 4096 3 > gpurun_out/train_step_plain.log 2>&1; echo "plain rc=$?"; tail -4 gpurun_out/train_step_plain.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_train_launches.csv python tools/train_step.py 4096 3 > gpurun_out/train_step_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2_train_launches.csv
