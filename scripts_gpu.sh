mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multigpu.py tests/test_gpu_train_step.py -m gpu -q -x > gpurun_out/r02_pytest10.log 2>&1; echo "pytest exit $?"
tail -n 25 gpurun_out/r02_pytest10.log
