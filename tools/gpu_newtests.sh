#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dropin.py tests/test_synthetic.py tests/test_gpu_edge_cases.py "tests/test_gpu_parity.py::test_psnr_vs_rrtd_full_size_live" -m gpu -x -q -s --durations=8 > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log
grep -E "PSNR|passed|failed|Error|error|slowest|s call" gpurun_out/pytest_new.log | tail -25
