#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "tf32" > gpurun_out/t_conv.log 2>&1; echo "conv pytest exit $?" >> gpurun_out/summary.txt
for t in 15; do SMG_TMA=$t timeout 400 python bench.py --steps 10 --warmup 3 --units 4 --no-cpu-baseline --no-backprop > gpurun_out/bench_tma$t.log 2>&1; echo "bench tma$t exit $?" >> gpurun_out/summary.txt; done
SMG_TMA=15 timeout 400 python bench.py --steps 10 --warmup 3 --units 1 --no-cpu-baseline --no-backprop > gpurun_out/bench_tma15_u1.log 2>&1
cat gpurun_out/summary.txt
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_conv.log | tail -n 8 | cut -c1-200
for t in tma15 tma15_u1; do tail -n 1 gpurun_out/bench_$t.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"; done
