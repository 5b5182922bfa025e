#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_all.log | tail -n 6 | cut -c1-200
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_v8.log 2>&1; tail -n 1 gpurun_out/bench_v8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v8', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('backprop',{}).get('value'), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"
