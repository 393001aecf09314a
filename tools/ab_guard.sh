#!/bin/bash
# same box: current library vs tools/_ab/libgbnerf_old.so (timing A/B of the forward MLP kernel, TFLOP/s per window).
# The other build is not kept in the tree: `git worktree add /tmp/old <commit>; python /tmp/old/gb-nerf_b200/csrc/build.py;
# cp /tmp/old/gb-nerf_b200/libgbnerf.so tools/_ab/libgbnerf_old.so` (tools/mlp_sustained.py loads it through AB_LIB).
for v in new old new old; do
  if [ $v = new ]; then r=$(python tools/mlp_sustained.py bf16 8 2>&1 | tail -3 | awk '{print $5}' | tr '\n' ' ');
  else r=$(AB_LIB=tools/_ab/libgbnerf_$v.so python tools/mlp_sustained.py bf16 8 2>&1 | tail -3 | awk '{print $5}' | tr '\n' ' '); fi
  echo "$v: $r"
done
