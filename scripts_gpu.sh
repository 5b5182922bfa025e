mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_backward.py tests/test_gpu_parity_r02.py tests/test_gpu_qnet.py -m gpu -q -s > gpurun_out/r02_pytest6.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error|error|fused|kink|graph vs" gpurun_out/r02_pytest6.log | head -40
for cfg in "1 1 64" "0 1 64" "1 0 64" "1 1 32" "1 1 148"; do
set -- $cfg
SMG_BN_FUSE=$1 SMG_WGRAD_ASYNC=$2 SMG_WGRAD_CTAS=$3 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-extras > gpurun_out/r02_bench6_$1$2$3.log 2>/dev/null
python - "$cfg" gpurun_out/r02_bench6_$1$2$3.log <<'PY'
import json,sys
l=[x for x in open(sys.argv[2]) if x.startswith('{')]
d=json.loads(l[-1])
print("fuse/async/cap", sys.argv[1], "value %.1f"%d['value'], "backprop", {k:round(v,2) if isinstance(v,float) else v for k,v in d['backprop'].items() if k in('value','ms_per_step','launches_per_step','error')})
PY
done
