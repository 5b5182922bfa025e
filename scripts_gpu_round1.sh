#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_qnet.py tests/test_gpu_decision.py -q -s -m gpu > gpurun_out/t_q.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
for u in 1 2 4 8; do timeout 400 python bench.py --steps 10 --warmup 3 --units $u --no-cpu-baseline --no-backprop > gpurun_out/bench_u$u.log 2>&1; echo "bench u$u exit $?" >> gpurun_out/summary.txt; done
cat gpurun_out/summary.txt
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_q.log | tail -n 5 | cut -c1-300
for u in 1 2 4 8; do tail -n 1 gpurun_out/bench_u$u.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('units', d['config']['units_per_step_per_gpu'], round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"; done
