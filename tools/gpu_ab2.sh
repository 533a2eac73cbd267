#!/bin/bash
# One GPU call: parity subset on the default library (pass the test files / -k filter in TESTS), then the default library
# against the variant libraries named in VARIANTS (rrt_b200/variants, `make tune`) on the headline image and the
# 1.1 M-primitive scene (tools/gpu_ab.py: time + image hash)
mkdir -p gpurun_out
L=gpurun_out/ab2.log; : > $L
timeout 400 python -m pytest ${TESTS:-tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_f64.py tests/test_moving_instances.py tests/test_synthetic.py} -m gpu -x -q -k "${KEXPR:-not live}" > gpurun_out/ab2_pytest.log 2>&1
echo "pytest exit $?" >> $L; tail -5 gpurun_out/ab2_pytest.log >> $L
timeout 120 python tools/gpu_ab.py default ${SPP:-64 8} >> $L 2>&1
for v in ${VARIANTS:-cur}; do
  RRTB_LIB=$PWD/rrt_b200/variants/librrtb200_$v.so timeout 120 python tools/gpu_ab.py $v ${SPP:-64 8} >> $L 2>&1
done
timeout 120 python tools/gpu_ab.py default ${SPP:-64 8} >> $L 2>&1
cat $L
