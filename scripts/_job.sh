timeout 400 python scripts/cpu_scaling.py 10 8,16,32,64,128 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_tree_parity.py -x -q -m gpu -k "checkpoint" 2>&1 | tail -3
