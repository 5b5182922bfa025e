#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -q -s -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "pytest gpu exit $?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench ref exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_gpu_all.log | tail -n 8 | cut -c1-300
tail -n 2 gpurun_out/smoke.log
tail -n 1 gpurun_out/bench_default.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], d.get('backprop',{}).get('value'), d['roofline']['bound'], round(d['roofline']['frac'],3), d['roofline']['kernel'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()}, d['cpu_baseline']['value'])"
tail -n 1 gpurun_out/bench_reference.log | cut -c1-300
