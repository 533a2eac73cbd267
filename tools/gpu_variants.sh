#!/bin/bash
# tuning aid: time every variant library under rrt_b200/variants on the headline (64 spp) and the synthetic (8 spp) workload
mkdir -p gpurun_out
for so in rrt_b200/variants/librrtb200_*.so; do
  v=$(basename $so .so); v=${v#librrtb200_}
  echo "== $v"
  RRTB_LIB=$PWD/$so timeout 300 python tools/gpu_sweep.py ${SPP_FINAL:-64} 2>&1 | grep -E "pool default|rror"
  if [ -z "$NO_SYN" ]; then RRTB_LIB=$PWD/$so timeout 600 python tools/gpu_sweep_wl.py synthetic 8 2>&1 | grep -E "default|rror"; fi
done 2>&1 | tee gpurun_out/variants.log
