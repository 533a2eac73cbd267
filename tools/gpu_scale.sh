#!/bin/bash
# run under gpurun --gpus 8: strong scaling of one image over N = 1, 2, 4, 8 B200 (one process per GPU, torchrun), one
# bench line per N.  usage: gpu_scale.sh <workload> <steps> [extra bench args]
WL=${1:-final}; STEPS=${2:-10}; shift 2
mkdir -p gpurun_out
out=gpurun_out/r02_scale_$WL.jsonl; : > $out
port=29720
for n in ${NS:-1 2 4 8}; do
  port=$((port+1))
  if [ $n = 1 ]; then
    timeout 1200 python bench.py --gpus 1 --workload $WL --steps $STEPS --warmup 3 --no-baselines "$@" 2>gpurun_out/r02_scale_${WL}_n$n.err | tail -1 >> $out
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --workload $WL --steps $STEPS --warmup 3 --no-baselines "$@" 2>gpurun_out/r02_scale_${WL}_n$n.err | tail -1 >> $out
  fi
done
python - "$out" <<'PY'
import json,sys
base=None
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: print("bad line", l[:200]); continue
    if base is None: base=(d["value"], d["e2e"]["value"])
    c=d["clocks"]
    print("%s N=%d  %9.0f Mrays/s (x%.2f)  %9.2f ms/step  e2e %9.0f (x%.2f)  clocks: %s MHz, %d samples, %s" % (d["config"]["workload"][:28], d["n_gpus"], d["value"], d["value"]/base[0], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["value"]/base[1], c["sm_mhz"], c["samples"], c["reasons"]))
PY
