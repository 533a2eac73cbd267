#!/bin/bash
# Mrays/s of every BASELINE.json config at 1, 2, 4 and 8 GPUs of one box (run under `gpurun --gpus 8`).
# Appends one JSON line per (workload, N) to gpurun_out/scaling.jsonl.
mkdir -p gpurun_out
out=gpurun_out/scaling.jsonl
: > $out
port=29600
for n in 1 2 4 8; do
  for w in final test2 test3 synthetic; do
    port=$((port+1))
    extra="--steps 3 --warmup 3"
    [ $w = synthetic ] && extra="--steps 1 --warmup 1 --spp 256"
    if [ $n = 1 ]; then
      timeout 900 python bench.py --gpus 1 --workload $w $extra --no-baselines 2>/dev/null | tail -1 >> $out
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --workload $w $extra --no-baselines 2>/dev/null | tail -1 >> $out
    fi
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/scaling.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print("%-8s N=%d  %9.0f Mrays/s  %8.2f ms/step  e2e %9.0f" % (d["config"]["workload"].split()[0].split("/")[-1], d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
