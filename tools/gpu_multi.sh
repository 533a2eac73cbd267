#!/bin/bash
# multi-GPU check (run under gpurun [--gpus N]): the in-library frame path, the drop-in tests, then N=1 and N=NGPU bench lines
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_anim.py tests/test_gpu_dropin.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -15 gpurun_out/pytest_multi.log
N=${NGPU:-2}
port=29650
for n in 1 $N; do
  [ $n = 1 ] && [ "$N" = 1 ] && [ -n "$done1" ] && continue
  done1=1
  port=$((port+1))
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-baselines 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/bench_n1.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 --no-baselines 2>gpurun_out/bench_n$n.err | tail -1 > gpurun_out/bench_n$n.json
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$n.json").read())
    print("N=%d value %.0f ms %.2f e2e %.0f clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/bench_n$n.err").read()[-2000:])
PY
done
