TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711"
timeout 400 $TR tools/sharded_check.py native-inplace 24 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -8
timeout 300 $TR tools/inplace_big.py 30 3 2>/dev/null | tail -1 | cut -c1-900
