#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -q -s -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "pytest gpu exit $?" >> gpurun_out/summary.txt
for mb in 0 64 96 128; do
  SMG_L2_CHUNK_MB=$mb timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-backprop > gpurun_out/bench_chunk$mb.log 2>&1; echo "bench chunk $mb exit $?" >> gpurun_out/summary.txt
done
SMG_NO_GRAPHS=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-backprop > gpurun_out/bench_nograph.log 2>&1; echo "bench nograph exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tf32.csv python profiles/profile_step.py --precision tf32 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|gradient errors" gpurun_out/t_gpu_all.log | tail -n 8 | cut -c1-300
for f in gpurun_out/bench_chunk*.log gpurun_out/bench_nograph.log; do echo $f; tail -n 1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"; done
