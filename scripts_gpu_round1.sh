#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_qnet.py -q -s -m gpu -x > gpurun_out/t_gpu_all.log 2>&1; echo "pytest gpu exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.log 2>&1; echo "bench tf32 exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tf32.csv python profiles/profile_step.py --precision tf32 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|error" gpurun_out/t_gpu_all.log | tail -n 5
tail -n 1 gpurun_out/bench_tf32.log | cut -c1-300
