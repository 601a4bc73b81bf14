TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711"
timeout 300 $TR tools/sharded_check.py native-inplace 24 2>&1 | grep "native n=\|sharded check ok\|Error" | tail -6
timeout 200 $TR tools/inplace_big.py 30 2 2>/dev/null | grep "^{" | python -c "
import json,sys;d=json.loads(sys.stdin.readline())
for k in ('split','in_place','separate'): print(k, d[k]['ms_per_step'], d[k].get('split_per_step'), d[k]['in_place_per_step'])
print('agree', d['marginals_agree'])"
