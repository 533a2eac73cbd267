#!/bin/bash
# cautious check of a new build: every step under its own short timeout (a spinning collapse kernel must not eat the GPU budget)
timeout 120 python -m pytest tests/test_gpu_edge_cases.py tests/test_random_scenes.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python -m pytest tests/test_synthetic.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
timeout 100 python tools/gpu_sweep.py 64 2>&1 | grep -E "pool default|simple"
timeout 100 python tools/gpu_sweep.py 500 2>&1 | grep -E "pool default"
timeout 150 python tools/gpu_sweep_wl.py synthetic 8 2>&1 | grep default
timeout 200 python -m pytest tests/test_synthetic.py tests/test_moving_instances.py tests/test_gpu_parity.py -m gpu -x -q -k "not live" 2>&1 | tail -3
