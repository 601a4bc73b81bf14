mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701"
timeout 700 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n2.err
( timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q ) 2>&1 | tail -3
