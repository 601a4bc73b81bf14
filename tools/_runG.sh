mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke rc=$?"
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 500 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qsim_jit_pass -s 18 -c 1 -o gpurun_out/dual_c3_p4 python tools/profile_case.py c3 30 1 > gpurun_out/ncu_dual.log 2>&1; echo "ncu dual rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:qsim_jit_pass -s 24 -c 1 -o gpurun_out/jit_dense_p3 python tools/profile_case.py dense 30 2 > gpurun_out/ncu_dense.log 2>&1; echo "ncu dense rc=$?"
