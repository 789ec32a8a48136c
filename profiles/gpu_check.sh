#!/bin/bash
# usage: profiles/gpu_check.sh <tag>  -- GPU parity tests + short bench on one B200 (logs under gpurun_out/)
TAG=$1
cd /root/repo
/usr/local/graft/bin/gpurun --timeout 600 -- "python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log; python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; tail -c 400 gpurun_out/bench_$TAG.log" 2>&1 | tail -8
python - <<PY
import json
l=[x for x in open('/root/repo/gpurun_out/bench_$TAG.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('VALUE %.3f M/s  ms/step %.3f  e2e %.3f  fwd-only %.3f  elbo steps/s %.1f frac %.4f'%(d['value']/1e6,d['ms_per_step'],d['e2e']['value']/1e6,d['extra']['forward_only_solves_per_s']/1e6,d['extra']['elbo']['steps_per_s'],d['roofline']['frac']))
PY
