mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest15.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error |error" gpurun_out/r02_pytest15.log | head -20
for fuse in 1 0 1 0; do
SMG_BN_FUSE=$fuse timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-extras > gpurun_out/r02_bench15_f$fuse.log 2>/dev/null
python - $fuse gpurun_out/r02_bench15_f$fuse.log <<'PY'
import json,sys
l=[x for x in open(sys.argv[2]) if x.startswith('{')]
d=json.loads(l[-1])
print("fuse", sys.argv[1], "value %.1f e2e %.1f"%(d['value'], d['e2e']['value']), "backprop %.1f steps/s %.2f ms launches %d"%(d['backprop']['value'], d['backprop']['ms_per_step'], d['backprop']['launches_per_step']), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['classes'].items()})
PY
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
