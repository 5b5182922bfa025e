#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 400 python -m pytest tests/test_gpu_qnet.py tests/test_gpu_decision.py tests/test_gpu_backward.py -q -m gpu > gpurun_out/t_q.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_q.log | tail -n 5 | cut -c1-250
for u in 4 1; do timeout 120 python bench.py --steps 20 --warmup 3 --units $u --no-cpu-baseline --no-backprop > gpurun_out/bench_u$u.log 2>&1; tail -n 1 gpurun_out/bench_u$u.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('u$u', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"; done
