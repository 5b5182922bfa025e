mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest12.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error |error" gpurun_out/r02_pytest12.log | head -20
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench12.log 2>gpurun_out/r02_bench12.err
python - gpurun_out/r02_bench12.log <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
d=json.loads(l[-1])
print("value %.1f e2e %s launches %d"%(d['value'], d['e2e'], d['gpu_launches']))
print({k:{a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()} for k,v in d['roofline']['classes'].items()})
for k in ('backprop','decision','fp32_mode'):
    v=d.get(k,{}); print(k, {a:(round(b,6) if isinstance(b,float) else b) for a,b in v.items() if a not in ('what','exchange','err_what')} if isinstance(v,dict) else v)
PY
tail -n 3 gpurun_out/r02_bench12.err
