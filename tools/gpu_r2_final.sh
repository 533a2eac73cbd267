#!/bin/bash
# Final round-2 measurement pass on ONE B200 (HEAD), ordered by importance so that a cut-off call still leaves the key
# records: headline bench line, ncu launch list + --set full capture of the render kernel (after the same command exited 0
# without ncu), the 1.1 M-primitive bench line, the shipped executable, the other workloads, the second capture.
# Numbers printed under ncu are never used as bench values.  Feeds tools/summarize_profiles.sh.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err
CMD="python bench.py --steps 2 --warmup 3 --no-baselines"
timeout 120 $CMD > gpurun_out/r02_plain_launch.log 2>&1 && {
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 4 -c 1 -f -o gpurun_out/r02_final_head $CMD > gpurun_out/r02_ncu_full.log 2>&1
}
timeout 300 python bench.py --workload synthetic --steps 2 --warmup 3 --no-baselines 2>/dev/null | tail -1 > gpurun_out/r02_bench_synthetic_n1.json
timeout 60 rrt_b200/bin/rrt -i oracle/_ref/scenes/final.txt -w 1200 -h 800 -s 500 -o /tmp/cli.png 2> gpurun_out/r02_cli_final.txt
timeout 60 rrt_b200/bin/rrt -i oracle/_ref/scenes/final.txt -w 1200 -h 800 -s 500 -R -o /tmp/cli.png 2>> gpurun_out/r02_cli_final.txt
timeout 120 python tools/build_time.py > gpurun_out/r02_build_time.log 2>&1
for wl in test2 test3; do
  timeout 120 python bench.py --workload $wl --steps 5 --warmup 3 --no-baselines 2>/dev/null | tail -1 > gpurun_out/r02_bench_${wl}_n1.json
done
timeout 120 python bench.py --precision f64 --steps 5 --warmup 3 --no-baselines 2>/dev/null | tail -1 > gpurun_out/r02_bench_f64_final.json
timeout 120 python bench.py --workload final_anim --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_final_anim.json
timeout 100 python tools/run_one.py synthetic 8 > gpurun_out/r02_plain_syn.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_render_pool -s 1 -c 1 -f -o gpurun_out/r02_synth_head python tools/run_one.py synthetic 8 > gpurun_out/r02_ncu_syn.log 2>&1
ls -la gpurun_out/r02_* | awk '{print $5, $9}'
grep -E "took|stats" gpurun_out/r02_cli_final.txt
cat gpurun_out/r02_build_time.log
for f in gpurun_out/r02_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1].split("/")[-1], "%.0f Mrays/s  %.2f ms  e2e %.0f  frac %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("roofline") or {}).get("frac"), (d.get("clocks") or {}).get("sm_mhz")))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
