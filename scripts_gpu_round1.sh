#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?" >> gpurun_out/summary.txt
SMG_NO_GRAPHS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_v10.csv python bench.py --steps 1 --warmup 1 --units 1 --no-cpu-baseline --no-backprop > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_all.log | tail -n 8 | cut -c1-200
tail -n 3 gpurun_out/smoke.log | cut -c1-300
tail -n 1 gpurun_out/bench_default.log | cut -c1-1200
tail -n 1 gpurun_out/bench_ref.log | cut -c1-300
