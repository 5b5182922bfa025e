#!/bin/bash
# GPU pass: full gpu test suite, bench, ncu launch list + full capture of the conv kernels
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -q -s -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "pytest gpu exit $?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tf32.log 2>&1; echo "bench tf32 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tf32.csv python profiles/profile_step.py --precision tf32 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma_kernelILi4ELi32ELi9 -c 3 -o gpurun_out/prof_conv3x3 -f python profiles/profile_step.py --precision tf32 > gpurun_out/ncu_full3.log 2>&1; echo "ncu full 3x3 exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma_kernelILi4ELi128ELi1ELi0 -c 3 -o gpurun_out/prof_conv1x1 -f python profiles/profile_step.py --precision tf32 > gpurun_out/ncu_full1.log 2>&1; echo "ncu full 1x1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 5 gpurun_out/t_gpu_all.log
tail -n 2 gpurun_out/bench_tf32.log
