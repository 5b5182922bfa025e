mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_bn_bwd.py tests/test_gpu_parity_r02.py tests/test_gpu_backward.py tests/test_gpu_conv.py -m gpu -q -s > gpurun_out/r02_pytest2.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error|error|dgrad|fused|kink|benched|hc K" gpurun_out/r02_pytest2.log | head -80
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench2.log 2> gpurun_out/r02_bench2.err; echo "bench exit $?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02_bench2.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    for k in ('value','ms_per_step','e2e','backprop','backprop_fp32','gpu_reference','fp32_mode','precision_err_vs_reference','cpu_baseline'):
        print(k, d.get(k))
PY
tail -n 5 gpurun_out/r02_bench2.err
