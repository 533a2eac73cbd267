#!/bin/bash
mkdir -p gpurun_out
for so in rrt_b200/variants/librrtb200_*.so; do
  v=$(basename $so .so); v=${v#librrtb200_}
  echo "== $v"
  RRTB_LIB=$PWD/$so timeout 300 python tools/gpu_sweep.py 64 RRTB_TH_NODE=8,12,16,20 RRTB_TH_LEAF=4,8,12 RRTB_TH_FETCH=8,16,24 2>&1 | grep -E "pool|rror" | sort -k3 -n | head -6
  RRTB_LIB=$PWD/$so timeout 600 python tools/gpu_sweep_wl.py synthetic 8 RRTB_TH_FETCH=8,16,24 2>&1 | grep -E "default|FETCH|rror"
done 2>&1 | tee gpurun_out/variants3.log
