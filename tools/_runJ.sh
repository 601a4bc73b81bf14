mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701"
timeout 300 $TR tools/inplace_big.py 30 4 2>gpurun_out/ib30.err | grep "^{" > gpurun_out/inplace_30.json; echo "ib30 rc=${PIPESTATUS[0]}"; grep -i "error\|Traceback" gpurun_out/ib30.err | head -3
python -c "
import json;d=json.load(open('gpurun_out/inplace_30.json'))
for k in ('split','in_place','separate'): print(k, d[k]['ms_per_step'], d[k].get('split_per_step'), d[k]['in_place_per_step'])
print('agree', d['marginals_agree'])"
timeout 400 $TR tools/inplace_big.py 33 2 2>gpurun_out/ib33.err | grep "^{" > gpurun_out/inplace_33.json; echo "ib33 rc=${PIPESTATUS[0]}"; grep -i "error\|Traceback" gpurun_out/ib33.err | head -3
python -c "
import json;d=json.load(open('gpurun_out/inplace_33.json'))
for k in ('split','in_place','separate'): print(k, d[k]['ms_per_step'], d[k].get('split_per_step'), d[k]['in_place_per_step'])
print('agree', d['marginals_agree'])"
timeout 300 $TR tools/sharded_check.py native-inplace 24 2>&1 | grep "native n=\|sharded check ok\|Error" | tail -6
timeout 200 $TR tools/stress_sharded.py 40 native-inplace 2>&1 | grep "stress\|Error" | tail -2
