mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench8_n2.log 2> gpurun_out/r02_bench8_n2.err; echo "bench exit $?"
python - gpurun_out/r02_bench8_n2.log <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print("value %.1f e2e %.1f"%(d['value'], d['e2e']['value']))
    for k in ('backprop','decision','replay'):
        v=d.get(k,{}); print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('what','exchange')})
PY
tail -n 12 gpurun_out/r02_bench8_n2.err
