#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_conv.py -q -s -m gpu -k "multi_tile" > gpurun_out/t_mt.log 2>&1; echo "pytest mt exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_qnet.py tests/test_gpu_decision.py -q -s -m gpu > gpurun_out/t_q.log 2>&1; echo "pytest q exit $?" >> gpurun_out/summary.txt
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-backprop > gpurun_out/bench_$name.log 2>&1; echo "bench $name exit $?" >> gpurun_out/summary.txt; }
run T1 SMG_TILES_PER_CTA=1
run Tauto SMG_X=1
run T2 SMG_TILES_PER_CTA=2
run T4 SMG_TILES_PER_CTA=4
cat gpurun_out/summary.txt
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_mt.log gpurun_out/t_q.log | tail -n 10 | cut -c1-300
for f in gpurun_out/bench_T*.log; do echo $f; tail -n 1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})"; done
