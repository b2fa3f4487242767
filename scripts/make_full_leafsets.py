"""Full-size fixture of BASELINE config 2 from the CPU ORACLE alone (test infrastructure): per root
simplex of tests/golden/pendulum_roots.npy the number of leaves and the sha256 of the sorted leaf
vertex sets.  ~2.5 minutes on 16 host threads.  Writes tests/golden/pendulum_full_leafsets.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)


def main():
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from explicit_hybrid_mpc_b200.partition import leafset_digest
    from oracle.oracle_cpu import CpuOracle
    mpc = InvertedPendulumOnCart()
    eps_a, eps_r = 19.21220798973866, 1.0
    roots = np.load(os.path.join(ROOT, 'tests', 'golden', 'pendulum_roots.npy'))
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 8)
    t0 = time.time()
    out = CpuOracle(mpc.problem, eps_a, eps_r, mpc.theta_scale).partition(roots, nthreads=threads)
    assert out['error'] == 0 and out['nfail'] == 0, (out['error'], out['nfail'])
    cnt, dig = leafset_digest(out['left'], out['root_id'], out['vertices'], len(roots))
    fx = dict(source='oracle/ehm_oracle.c (CPU oracle), scripts/make_full_leafsets.py', eps_a=eps_a, eps_r=eps_r,
              nroots=len(roots), nleaves=int(out['nleaves']), nnodes=int(len(out['left'])),
              nsolves=int(out['nsolves']), seconds=round(time.time() - t0, 1), threads=threads,
              leaves_per_root=cnt, sha256_per_root=dig)
    with open(os.path.join(ROOT, 'tests', 'golden', 'pendulum_full_leafsets.json'), 'w') as f:
        json.dump(fx, f, indent=1)
    print('leaves', fx['nleaves'], 'nodes', fx['nnodes'], 'seconds', fx['seconds'])


if __name__ == '__main__':
    main()
