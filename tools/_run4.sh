mkdir -p gpurun_out
for m in off auto; do
  QSIM_DUAL=$m timeout 200 python tools/pass_times.py dense 30 2>/dev/null > gpurun_out/pt_dense_$m.json; python -c "import json;d=json.load(open('gpurun_out/pt_dense_$m.json'));print('$m dense', d['pass_ms'], d['total_ms'])"
  QSIM_DUAL=$m timeout 200 python tools/pass_times.py c3 30 2>/dev/null > gpurun_out/pt_c3_$m.json; python -c "import json;d=json.load(open('gpurun_out/pt_c3_$m.json'));print('$m c3', d['pass_ms'], d['total_ms'])"
done
QSIM_DUAL=always timeout 200 python tools/pass_times.py dense 30 > gpurun_out/pt_dense_always.json 2> gpurun_out/pt_dense_always.err; tail -5 gpurun_out/pt_dense_always.err; cut -c1-400 gpurun_out/pt_dense_always.json
# per-kernel times of C1 (graph replay of specialised kernels)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c1_launches.csv python tools/config_runs.py c1 > gpurun_out/c1_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/c1_launches.csv')) if len(r)>5 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, ..., Metric Name, Metric Unit, Metric Value
agg=collections.OrderedDict()
for r in rows[-40:]:
    print(r[4][:60], r[-2], r[-1])
PY
# ncu full capture of a dual kernel: c3 pass 4 at 30 q
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qsim_jit_pass -s 10 -c 1 -o gpurun_out/dual_c3_p4 python tools/profile_case.py c3 30 1 > gpurun_out/ncu_dual.log 2>&1; tail -3 gpurun_out/ncu_dual.log
