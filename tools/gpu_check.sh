#!/bin/bash
# GPU check: parity tests + smoke + the default bench line (what the driver runs at round end)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py ${BENCH_ARGS:---steps 5 --warmup 3} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], 'refgpu', (d.get('reference_gpu') or {}).get('Mrays/s'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))"
