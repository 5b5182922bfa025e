#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 200 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "tf32 or tma" > gpurun_out/t_conv.log 2>&1; echo "conv pytest exit $?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_conv.log | tail -n 3 | cut -c1-200
for u in 4 1; do timeout 120 python bench.py --steps 10 --warmup 3 --units $u --no-cpu-baseline --no-backprop > gpurun_out/bench_u$u.log 2>&1; tail -n 1 gpurun_out/bench_u$u.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('u$u', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"; done
