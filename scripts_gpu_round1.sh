#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -q -s -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "pytest gpu exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tf32.log 2>&1; echo "bench tf32 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|gradient errors" gpurun_out/t_gpu_all.log | tail -n 12 | cut -c1-400
tail -n 1 gpurun_out/bench_tf32.log
