#!/bin/bash
# round-2 measurement pass on ONE B200 (HEAD): bench lines of every workload, CLI timing, then the ncu launch list and
# one --set full capture of the render kernel on the headline and the 1.1 M-primitive workload (each only after the same
# command has exited 0 without ncu).  Numbers printed under ncu are never used as bench values.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err
for wl in test2 test3; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-baselines 2>/dev/null | tail -1 > gpurun_out/r02_bench_${wl}_n1.json
done
timeout 300 python bench.py --precision f64 --steps 5 --warmup 3 --no-baselines 2>/dev/null | tail -1 > gpurun_out/r02_bench_f64_final.json
timeout 300 python bench.py --workload final_anim --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_final_anim.json
timeout 600 python bench.py --workload synthetic --steps 3 --warmup 3 --no-baselines 2>/dev/null | tail -1 > gpurun_out/r02_bench_synthetic_n1.json
# the shipped executable on the headline workload: "took" vs the bench's ms_per_step
timeout 120 rrt_b200/bin/rrt -i oracle/_ref/scenes/final.txt -w 1200 -h 800 -s 500 -o /tmp/cli.png 2> gpurun_out/r02_cli_final.txt
timeout 120 rrt_b200/bin/rrt -i oracle/_ref/scenes/final.txt -w 1200 -h 800 -s 500 -R -o /tmp/cli.png 2>> gpurun_out/r02_cli_final.txt
grep -E "took|stats" gpurun_out/r02_cli_final.txt
timeout 120 python tools/cold_warm.py > gpurun_out/r02_cold_warm.txt 2>&1
timeout 200 python tools/build_time.py > gpurun_out/r02_build_time.log 2>&1
# ---- ncu
CMD="python bench.py --steps 2 --warmup 3 --no-baselines"
timeout 200 $CMD > gpurun_out/r02_plain_launch.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
timeout 200 $CMD > gpurun_out/r02_plain_full.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 4 -c 1 -f -o gpurun_out/r02_final_head $CMD > gpurun_out/r02_ncu_full.log 2>&1
timeout 200 python tools/run_one.py synthetic 8 > gpurun_out/r02_plain_syn.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 1 -c 1 -f -o gpurun_out/r02_synth_head python tools/run_one.py synthetic 8 > gpurun_out/r02_ncu_syn.log 2>&1
timeout 200 python tools/run_one.py final 64 f64 > gpurun_out/r02_plain_f64.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 1 -c 1 -f -o gpurun_out/r02_f64_head python tools/run_one.py final 64 f64 > gpurun_out/r02_ncu_f64.log 2>&1
ls -la gpurun_out/r02_*
for f in gpurun_out/r02_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "%.0f Mrays/s  %.2f ms  e2e %.0f  frac %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("roofline") or {}).get("frac"), (d.get("clocks") or {}).get("sm_mhz")))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
