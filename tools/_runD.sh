( timeout 900 python -m pytest tests/test_jit_gpu.py tests/test_gates_gpu.py tests/test_noise_gpu.py -x -q ) 2>&1 | tail -3
for m in 0 1; do QSIM_JIT_DEFER_FLIPS=$m QSIM_DUAL_VERBOSE=1 timeout 200 python tools/pass_times.py dense 30 2>gpurun_out/pt_err.txt | python -c "import json,sys;d=json.loads(sys.stdin.readline());print('defer=$m dense', d['pass_ms'], d['total_ms'])"; grep "qsim_b200: pass" gpurun_out/pt_err.txt | cut -c60-200; done
QSIM_JIT_DEFER_FLIPS=1 timeout 200 python tools/pass_times.py c2 30 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.readline());print('c2', d['pass_ms'])"
timeout 100 python tools/dual_survey.py 30 3 2>&1 | grep -v "^    " | tail -1
timeout 100 python tools/dual_survey.py 29 2 2>&1 | grep -v "^    " | tail -1
