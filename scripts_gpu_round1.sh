#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/t_k.log 2>&1; echo "kernels pytest exit $?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_k.log | tail -n 5 | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 4 gpurun_out/smoke.log | cut -c1-200
