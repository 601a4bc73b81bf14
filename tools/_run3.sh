mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_jit_gpu.py -x -q -k "two_warp" ) > gpurun_out/pytest_dual.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dual.log
tail -15 gpurun_out/pytest_dual.log
for m in off auto always; do
  QSIM_DUAL=$m timeout 200 python tools/pass_times.py dense 30 2>/dev/null | tee gpurun_out/pt_dense_$m.json | cut -c1-80; python -c "import json;d=json.load(open('gpurun_out/pt_dense_$m.json'));print('$m dense', d['pass_ms'], d['total_ms'])"
done
for m in off auto; do
  QSIM_DUAL=$m timeout 200 python tools/pass_times.py c3 30 2>/dev/null > gpurun_out/pt_c3_$m.json; python -c "import json;d=json.load(open('gpurun_out/pt_c3_$m.json'));print('$m c3', d['pass_ms'], d['total_ms'])"
  QSIM_DUAL=$m timeout 200 python tools/pass_times.py c2 30 2>/dev/null > gpurun_out/pt_c2_$m.json; python -c "import json;d=json.load(open('gpurun_out/pt_c2_$m.json'));print('$m c2', d['pass_ms'], d['total_ms'])"
  QSIM_DUAL=$m timeout 200 python tools/config_runs.py c1 > gpurun_out/c1_$m.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/c1_$m.json'));print('$m c1', d['ms_compiled'], d['ms_compiled_specialised'], d['max_abs_err_specialised'])"
done
QSIM_DUAL=always timeout 200 python tools/config_runs.py c1 > gpurun_out/c1_always.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/c1_always.json'));print('always c1', d['ms_compiled'], d['ms_compiled_specialised'], d['max_abs_err_specialised'])"
