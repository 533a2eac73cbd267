#!/bin/bash
# run under gpurun --gpus 8: headline workload at the N given in NS (default "1 2 4 8"), one bench line each
mkdir -p gpurun_out
out=gpurun_out/scale_final.jsonl; : > $out
port=29710
for n in ${NS:-1 2 4 8}; do
  port=$((port+1))
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps ${STEPS:-10} --warmup 3 --no-baselines 2>gpurun_out/scale_final_n$n.err | tail -1 >> $out
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps ${STEPS:-10} --warmup 3 --no-baselines 2>gpurun_out/scale_final_n$n.err | tail -1 >> $out
  fi
done
python - <<'PY'
import json
base=None
for l in open("gpurun_out/scale_final.jsonl"):
    try: d=json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    if base is None: base=(d["value"], d["e2e"]["value"])
    print("N=%d  %9.0f Mrays/s (x%.2f)  %7.2f ms/step  e2e %9.0f (x%.2f)  clocks: %s MHz, %d samples, %s" % (d["n_gpus"], d["value"], d["value"]/base[0], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["value"]/base[1], d["clocks"]["sm_mhz"], d["clocks"]["samples"], d["clocks"]["reasons"]))
PY
