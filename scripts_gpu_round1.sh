#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for bo in 0 1; do
SMG_TMA=3 SMG_ASYNC=$bo timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "tf32 and test_conv_matches_torch" > gpurun_out/t_conv_bo$bo.log 2>&1; echo "conv pytest bo=$bo exit $?" >> gpurun_out/summary.txt
done
for t in 1 3; do SMG_TMA=$t timeout 400 python bench.py --steps 10 --warmup 3 --units 4 --no-cpu-baseline --no-backprop > gpurun_out/bench_tma$t.log 2>&1; echo "bench tma$t exit $?" >> gpurun_out/summary.txt; done
cat gpurun_out/summary.txt
for bo in 0 1; do grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_conv_bo$bo.log | tail -n 8 | cut -c1-200; done
for t in 1 3; do tail -n 1 gpurun_out/bench_tma$t.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tma $t', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"; done
