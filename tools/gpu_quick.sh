#!/bin/bash
# quick perf + correctness probe of the default library: parity subset, then headline (64 spp and 500 spp) and synthetic (8 spp)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f64.py tests/test_synthetic.py -m gpu -x -q -k "not full_size" 2>&1 | tail -2
python tools/gpu_sweep.py 64 2>&1 | grep -E "pool default|simple"
python tools/gpu_sweep.py 500 2>&1 | grep -E "pool default"
python tools/gpu_sweep_wl.py synthetic 8 2>&1 | grep default
