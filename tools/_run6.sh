export QSIM_PASS_TIMELINE=1
echo "== min 210 (p2 dual)"; QSIM_DUAL_MIN_FP64=210 timeout 120 python tools/_dbg_dual2.py 2>&1 | tail -12
echo "== min 230"; QSIM_DUAL_MIN_FP64=230 timeout 120 python tools/_dbg_dual2.py 2>&1 | tail -4
echo "== c1 timeline dual off"; QSIM_DUAL=off timeout 120 python tools/_c1_timeline.py 20 2>&1 | tail -8
echo "== c1 timeline dual auto"; timeout 120 python tools/_c1_timeline.py 20 2>&1 | tail -8
echo "== 14q timeline"; QSIM_DUAL=off timeout 120 python tools/_c1_timeline.py 14 2>&1 | tail -8
