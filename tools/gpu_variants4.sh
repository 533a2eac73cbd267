#!/bin/bash
for so in rrt_b200/variants/librrtb200_*.so; do
  v=$(basename $so .so); v=${v#librrtb200_}
  echo "== $v"
  RRTB_LIB=$PWD/$so timeout 200 python tools/gpu_sweep.py 64 RRTB_TH_NODE=8,12,16 RRTB_TH_LEAF=4,8 RRTB_TH_FETCH=12,16,24 RRTB_STEP_ITERS=4,8 2>&1 | grep -E "pool" | sort -t' ' -k1,1 | awk '{print}' | sort -k9 -n -r | head -8
done
