mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701"
( timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q ) > gpurun_out/pytest_sharded.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_sharded.log
timeout 300 $TR tools/inplace_big.py 30 4 > gpurun_out/inplace_30.json 2> gpurun_out/inplace_30.err; echo "rc=$?" >> gpurun_out/inplace_30.err
timeout 400 $TR tools/inplace_big.py 33 2 > gpurun_out/inplace_33.json 2> gpurun_out/inplace_33.err; echo "rc=$?" >> gpurun_out/inplace_33.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
tail -5 gpurun_out/pytest_sharded.log; tail -2 gpurun_out/inplace_30.err; cat gpurun_out/inplace_30.json | cut -c1-1500; tail -2 gpurun_out/inplace_33.err; cat gpurun_out/inplace_33.json | cut -c1-1500; tail -2 gpurun_out/bench_n2.err
