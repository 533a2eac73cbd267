#!/bin/bash
# round-2 GPU check B: ncu --set full of the render kernel on the synthetic and the headline workload
mkdir -p gpurun_out
python tools/run_one.py synthetic 8 > gpurun_out/plain_syn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 1 -c 1 -f -o gpurun_out/r02_synth python tools/run_one.py synthetic 8 > gpurun_out/ncu_syn.log 2>&1
tail -2 gpurun_out/plain_syn.log
python tools/run_one.py final 64 > gpurun_out/plain_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 1 -c 1 -f -o gpurun_out/r02_final python tools/run_one.py final 64 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/plain_final.log
ls -la gpurun_out/*.ncu-rep
