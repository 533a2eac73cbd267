#!/bin/bash
# Short scaling check of HEAD (run under `gpurun --gpus 8`): headline workload in both precisions at N = 1, 2, 4, 8.
mkdir -p gpurun_out
out=gpurun_out/scaling_head.jsonl
: > $out
port=29700
for prec in f32 f64; do
  for n in 1 2 4 8; do
    port=$((port+1))
    if [ $n = 1 ]; then
      timeout 600 python bench.py --gpus 1 --precision $prec --steps 5 --warmup 3 --no-baselines 2>/dev/null | tail -1 >> $out
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --precision $prec --steps 5 --warmup 3 --no-baselines 2>/dev/null | tail -1 >> $out
    fi
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/scaling_head.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print("%-6s N=%d  %9.0f Mrays/s  %8.2f ms/step  e2e %9.0f" % ("f64" if "double" in d["metric"] else "f32", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
