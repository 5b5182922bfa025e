#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -q -s -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "pytest gpu exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tf32.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python profiles/profile_step.py --mode train --precision tf32 > gpurun_out/ncu_train.log 2>&1; echo "ncu train exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_gpu_all.log | tail -n 8 | cut -c1-300
tail -n 1 gpurun_out/bench_tf32.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], d.get('backprop'), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()}, d['cpu_baseline'])"
