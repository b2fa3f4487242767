timeout 900 python -m pytest tests/test_gpu_tree_parity.py tests/test_gpu_batch_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 200 python scripts/run_partition.py 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v11',d['device_ms'],d['nwaves'],d['nleaves'],d['nfail'],d['iters_per_solve'],d['nsolves'])"
