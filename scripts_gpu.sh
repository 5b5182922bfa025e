mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_qnet.py tests/test_gpu_kernels.py tests/test_gpu_parity_r02.py -m gpu -q -x > gpurun_out/r02_pytest13.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/r02_pytest13.log
# sustained line (400 steps, ~5 s of device time) with its clock record
timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-extras --no-backprop > gpurun_out/r02_bench_sustained_400.json 2>/dev/null; tail -c 600 gpurun_out/r02_bench_sustained_400.json; echo
# launch lists: one 1-unit inference step, one training step (graphs replayed; ncu serialises, compare shares)
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_infer.csv python profiles/profile_step.py --precision tf32 > /dev/null 2>&1; echo "ncu infer $?"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_v2.csv python profiles/profile_step.py --mode train --precision tf32 > /dev/null 2>&1; echo "ncu train $?"
# full captures: tensor-core wgrad kernels and the tensor-core dgrad (block-1 instances come late in the backward: skip to them)
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad_umma -s 100 -c 4 -o gpurun_out/r02_prof_wgrad python profiles/profile_step.py --mode train --precision tf32 > /dev/null 2>&1; echo "ncu wgrad $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"prep_rotate|bn_bwd" -s 0 -c 3 -o gpurun_out/r02_prof_misc python profiles/profile_step.py --mode train --precision tf32 > /dev/null 2>&1; echo "ncu misc $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"prep_rotate_kernel|conv1_t_kernel|conv3_wt_kernel" -s 0 -c 12 -o gpurun_out/r02_prof_fwd python profiles/profile_step.py --precision tf32 > /dev/null 2>&1; echo "ncu fwd $?"
ls -la gpurun_out/*.ncu-rep
