timeout 400 python -m pytest tests/test_jit_gpu.py -x -q -k "conditional_flips or two_warp" 2>&1 | tail -3; echo "rc=${PIPESTATUS[0]}"
