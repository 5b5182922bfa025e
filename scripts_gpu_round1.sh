#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for sl in 32 16; do SMG_CONV3_SLOT=$sl timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "tma" > gpurun_out/t_conv$sl.log 2>&1; echo "conv pytest slot $sl exit $?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_conv$sl.log | tail -n 3 | cut -c1-200; done
for sl in 32 16; do for u in 4 1; do SMG_CONV3_SLOT=$sl timeout 400 python bench.py --steps 10 --warmup 3 --units $u --no-cpu-baseline --no-backprop > gpurun_out/bench_s${sl}_u$u.log 2>&1; tail -n 1 gpurun_out/bench_s${sl}_u$u.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('slot$sl u$u', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"; done; done
