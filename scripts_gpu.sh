mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest exit $?"
tail -n 15 gpurun_out/r02_pytest1.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench1.log 2> gpurun_out/r02_bench1.err; echo "bench exit $?"
tail -c 3000 gpurun_out/r02_bench1.log
tail -n 5 gpurun_out/r02_bench1.err
