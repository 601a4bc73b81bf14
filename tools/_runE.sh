export QSIM_JIT_CACHE=/tmp/fresh_cache_$$
timeout 300 python -m pytest tests/test_gates_gpu.py -x -q 2>&1 | tail -2; echo "gates rc=${PIPESTATUS[0]}"
export QSIM_JIT_CACHE=/tmp/fresh_cache2_$$
timeout 300 python -m pytest tests/test_gates_gpu.py -x -q -k "c3_30q or 30q_properties" 2>&1 | tail -2; echo "gates-30q rc=${PIPESTATUS[0]}"
