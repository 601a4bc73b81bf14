mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
QSIM_DUAL_VERBOSE=1 timeout 300 python tools/config_runs.py c1 c3 c5 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; grep "qsim_b200: pass" gpurun_out/configs.err | cut -c1-220
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-dense --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qsim_jit_pass -s 3 -c 1 -o gpurun_out/jit_c2_final python tools/profile_case.py c2 30 2 > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
QSIM_DUAL=always timeout 400 ncu --set full --clock-control none --import-source on -k regex:qsim_jit_pass -s 17 -c 1 -o gpurun_out/dual_c3_p4 python tools/profile_case.py c3 30 1 > gpurun_out/ncu_dual.log 2>&1; echo "ncu dual rc=$?"
