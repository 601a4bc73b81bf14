mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke rc=$?"
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['dense_variant']['ms_per_step'])"
