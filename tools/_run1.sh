mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'chunk_approx|group_stitch|marginal|collapse|trajectory|batched_sample|shot_lookup|sample_kernel' -c 24 -o gpurun_out/readout_full_r02 python tools/profile_readout.py 30 65536 > gpurun_out/ncu_readout.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_readout.log
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json | cut -c1-600; tail -3 gpurun_out/ncu_readout.log
