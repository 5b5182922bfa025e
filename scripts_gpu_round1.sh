#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_backward.py -q -s -m gpu > gpurun_out/t_bwd.log 2>&1; echo "pytest bwd exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|error|worst|Error|assert" gpurun_out/t_bwd.log | tail -n 30
