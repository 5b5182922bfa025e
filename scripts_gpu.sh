mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest11.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error |error" gpurun_out/r02_pytest11.log | head -20
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench11.log 2>gpurun_out/r02_bench11.err
python - gpurun_out/r02_bench11.log <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
d=json.loads(l[-1])
print("value %.1f e2e %.1f"%(d['value'], d['e2e']['value']))
for k in ('backprop','backprop_fp32','decision','replay','fp32_mode','precision_err_vs_reference'):
    v=d.get(k,{}); print(k, {a:(round(b,6) if isinstance(b,float) else b) for a,b in v.items() if a not in ('what','exchange','err_what')} if isinstance(v,dict) else v)
PY
tail -n 3 gpurun_out/r02_bench11.err
