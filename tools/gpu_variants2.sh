#!/bin/bash
mkdir -p gpurun_out
for so in rrt_b200/variants/librrtb200_*.so; do
  v=$(basename $so .so); v=${v#librrtb200_}
  echo "== $v"
  RRTB_LIB=$PWD/$so timeout 300 python tools/gpu_sweep.py 64 RRTB_TH_NODE=8,12,16 RRTB_TH_LEAF=4,8,12 2>&1 | grep -E "pool|rror"
  RRTB_LIB=$PWD/$so timeout 600 python tools/gpu_sweep_wl.py synthetic 8 2>&1 | grep -E "default|rror"
done 2>&1 | tee gpurun_out/variants2.log
