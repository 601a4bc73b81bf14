timeout 400 python -m pytest tests/test_sharded_gpu.py -x -q -k "one_gpu" 2>&1 | tail -15; echo "rc=${PIPESTATUS[0]}"
