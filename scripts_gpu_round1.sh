#!/bin/bash
mkdir -p gpurun_out
export SMG_NO_GRAPHS=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3_persist_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r01_persist3 -f python bench.py --steps 1 --warmup 1 --units 4 --no-cpu-baseline --no-backprop > gpurun_out/ncu_p3.log 2>&1
echo "ncu p3 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma_tma_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/r01_tma1 -f python bench.py --steps 1 --warmup 1 --units 4 --no-cpu-baseline --no-backprop > gpurun_out/ncu_t1.log 2>&1
echo "ncu t1 exit $?"
ls -la gpurun_out/*.ncu-rep
