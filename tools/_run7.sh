export QSIM_PASS_TIMELINE=1
for s in 1 2 3 4 5 6 7; do timeout 100 python tools/_dbg_dual3.py 30 $s 2>&1 | grep -v "^    " | tail -2; done
timeout 100 python tools/_dbg_dual3.py 30 -1 2>&1 | grep -v "^    " | tail -2
for s in 42 1 2; do timeout 100 python tools/_dbg_dual3.py 29 $s 2>&1 | grep -v "^    " | tail -2; done
for s in 42 1 2; do timeout 100 python tools/_dbg_dual3.py 31 $s 2>&1 | grep -v "^    "| tail -2; done
unset QSIM_PASS_TIMELINE
QSIM_DBG_FLAGS=0 QSIM_DUAL=always timeout 200 python tools/pass_times.py dense 30 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.readline());print('always dense', d['pass_ms'], d['total_ms'])"
