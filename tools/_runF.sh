TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721"
timeout 300 $TR tools/stress_sharded.py 70 native-inplace 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
timeout 200 $TR tools/stress_sharded.py 40 native 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
