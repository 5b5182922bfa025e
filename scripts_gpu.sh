mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -s -x > gpurun_out/r02_pytest7.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error|error|replay batch|Traceback" gpurun_out/r02_pytest7.log | head -30
for fuse in 0 1; do
SMG_BN_FUSE=$fuse timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench7_f$fuse.log 2>gpurun_out/r02_bench7_f$fuse.err
python - gpurun_out/r02_bench7_f$fuse.log <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
d=json.loads(l[-1])
print("value %.1f e2e %.1f"%(d['value'], d['e2e']['value']))
for k in ('backprop','backprop_fp32','decision','replay','fp32_mode'):
    v=d.get(k,{}); print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('what','exchange')})
PY
tail -n 3 gpurun_out/r02_bench7_f$fuse.err
done
