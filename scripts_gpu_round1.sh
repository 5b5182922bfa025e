#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_qnet.py -q -m gpu -s -k "stem" > gpurun_out/t_q.log 2>&1; echo "pytest exit $?"
grep -E "tensor-core stem|passed|failed|^FAILED|^E  " gpurun_out/t_q.log | tail -n 8 | cut -c1-250
for t in 7 23; do SMG_TMA=$t timeout 120 python bench.py --steps 5 --warmup 3 --units 4 --no-cpu-baseline --no-backprop > gpurun_out/bench_t$t.log 2>&1; tail -n 1 gpurun_out/bench_t$t.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tma$t', round(d['value'],1), d['config']['precision_note'])"; done
