#!/bin/bash
# gpurun_out/r02_* (scratch, written by tools/gpu_r2_measure.sh on a B200) -> profiles/r02_* (tracked): bench lines, the ncu
# launch list, raw metric summaries of the --set full captures, the per-source-line aggregation and the DRAM traffic figure
set -e
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
for f in final_n1 test2_n1 test3_n1 f64_final final_anim synthetic_n1; do
  [ -s $G/r02_bench_$f.json ] && tail -1 $G/r02_bench_$f.json > $P/r02_bench_$f.json
done
[ -s $G/r02_launches_bench.csv ] && cp $G/r02_launches_bench.csv $P/r02_launches_bench.csv
[ -s $G/r02_cli_final.txt ] && grep -E "took|stats" $G/r02_cli_final.txt > $P/r02_cli_final.txt
[ -s $G/r02_build_time.log ] && cp $G/r02_build_time.log $P/r02_build_time.txt
[ -s $G/r02_cold_warm.txt ] && cp $G/r02_cold_warm.txt $P/r02_cold_warm.txt
for k in final synth f64; do
  rep=$G/r02_${k}_head.ncu-rep
  [ -s $rep ] || continue
  python tools/ncu_summary.py $rep > $P/r02_ncu_render_pool_${k}_raw.txt
  ncu -i $rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,r=rows[0],rows[1],rows[-1]
for k in ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_read.sum','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','smsp__inst_executed_op_global_ld.sum'):
    if k in h: print('%-86s %-14s %s'%(k,u[h.index(k)],r[h.index(k)]))
" >> $P/r02_ncu_render_pool_${k}_raw.txt
done
# per-source-line view of the headline capture
mkdir -p /tmp/prof_r02 && cd /tmp/prof_r02 && rm -f *.cubin
cuobjdump -xelf all $OLDPWD/rrt_b200/librrtb200.so > /dev/null 2>&1
for f in *.cubin; do if nvdisasm -c $f 2>/dev/null | grep -q "k_render_pool"; then nvdisasm --print-line-info -c $f > render.sass; fi; done
ncu -i $OLDPWD/$G/r02_final_head.ncu-rep --page source --csv > src.csv 2>/dev/null
cd $OLDPWD
python tools/ncu_by_line.py /tmp/prof_r02/src.csv /tmp/prof_r02/render.sass "k_render_poolILb0ELi2ELb0ENS_7PathF32ELb0" 60 > $P/r02_ncu_render_pool_by_line.txt
python - <<'PY'
import csv, json, subprocess
raw = subprocess.run(["ncu", "-i", "gpurun_out/r02_final_head.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h, u, r = rows[0], rows[1], rows[-1]
def val(k):
    v = float(r[h.index(k)]); un = u[h.index(k)]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[un]
t = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
json.dump({"kernel": r[h.index("Kernel Name")], "traffic_bytes_per_launch": t, "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
           "source": "ncu --set full of `python bench.py --steps 2 --warmup 3 --no-baselines`, launch 5 of k_render_pool (gpurun_out/r02_final_head.ncu-rep)"},
          open("profiles/r02_ncu_traffic.json", "w"), indent=1)
print("traffic", t)
PY
ls -la $P/r02_*
