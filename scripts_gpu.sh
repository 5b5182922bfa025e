mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest4.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error|error|fused|kink|graph vs" gpurun_out/r02_pytest4.log | head -60
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench4.log 2> gpurun_out/r02_bench4.err; echo "bench exit $?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02_bench4.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    for k in ('value','ms_per_step','e2e','backprop','backprop_fp32'):
        print(k, d.get(k))
PY
tail -n 5 gpurun_out/r02_bench4.err
