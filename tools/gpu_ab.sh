#!/bin/bash
# One GPU call: parity subset on the default library, then every variant library (rrt_b200/variants, `make tune`) timed on
# the headline image and on the 1.1 M-primitive scene with image hashes; threshold sweep on the `cur` tuning build.
mkdir -p gpurun_out
L=gpurun_out/ab.log; : > $L
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_f64.py tests/test_moving_instances.py -m gpu -x -q -k "not full_size and not live" > gpurun_out/ab_pytest.log 2>&1
echo "pytest exit $?" >> $L; tail -3 gpurun_out/ab_pytest.log >> $L
timeout 120 python tools/gpu_ab.py default 64 8 >> $L 2>&1
for so in rrt_b200/variants/librrtb200_*.so; do
  v=$(basename $so .so); v=${v#librrtb200_}
  RRTB_LIB=$PWD/$so timeout 120 python tools/gpu_ab.py $v 64 8 >> $L 2>&1
done
RRTB_LIB=$PWD/rrt_b200/variants/librrtb200_cur.so timeout 200 python tools/gpu_ab.py sweep 64 8 RRTB_TH_FETCH=8,12,16,20 RRTB_TH_NODE=8,12,16 >> $L 2>&1
RRTB_LIB=$PWD/rrt_b200/variants/librrtb200_cur.so timeout 120 python tools/gpu_ab.py sweep2 64 8 RRTB_TH_LEAF=4,8,12 RRTB_STEP_ITERS=4,8,16 >> $L 2>&1
timeout 120 python tools/gpu_ab.py default500 500 0 >> $L 2>&1
RRTB_LIB=$PWD/rrt_b200/variants/librrtb200_legacy.so timeout 120 python tools/gpu_ab.py legacy500 500 0 >> $L 2>&1
cat $L
