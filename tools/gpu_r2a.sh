#!/bin/bash
# round-2 GPU check A: parity tests, then the wide-tree kernel on the headline and the synthetic workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-baselines > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -1 gpurun_out/bench_a.json | cut -c1-400
for v in u1 u2; do
  echo "== variant $v final"
  RRTB_LIB=$PWD/rrt_b200/variants/librrtb200_$v.so timeout 300 python tools/gpu_sweep.py 64 RRTB_TH_NODE=8,12,16,20 RRTB_STEP_ITERS=4,8,16 2>&1 | tail -15
  echo "== variant $v synthetic"
  RRTB_LIB=$PWD/rrt_b200/variants/librrtb200_$v.so timeout 600 python tools/gpu_sweep_wl.py synthetic 8 RRTB_TH_NODE=8,12,16 2>&1 | tail -5
done 2>&1 | tee gpurun_out/sweep_a.log
