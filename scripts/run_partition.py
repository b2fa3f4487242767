"""Run the device partition loop on the pendulum fixture and print statistics."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--eps-a', type=float, default=19.21220798973866)
    ap.add_argument('--eps-r', type=float, default=1.0)
    ap.add_argument('--roots', type=str, default='all')
    ap.add_argument('--wave', type=int, default=0)
    ap.add_argument('--verbose', type=int, default=0)
    ap.add_argument('--repeat', type=int, default=1)
    ap.add_argument('--max-nodes', type=int, default=0)
    ap.add_argument('--max-frontier', type=int, default=0)
    ap.add_argument('--problem', type=str, default='pendulum', help='pendulum | synthetic')
    ap.add_argument('--nroots', type=int, default=8, help='synthetic: number of Kuhn simplices of [-1,1]^8')
    ap.add_argument('--node-budget', type=int, default=0)
    ap.add_argument('--root-stride', type=int, default=1, help='synthetic: keep every k-th root simplex')
    a = ap.parse_args()
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from explicit_hybrid_mpc_b200.oracle import Oracle
    from explicit_hybrid_mpc_b200.partition import partition
    if a.problem == 'synthetic':
        from explicit_hybrid_mpc_b200.mpc_library import SyntheticMPLP
        from explicit_hybrid_mpc_b200.tools import kuhn_triangulation
        mpc = SyntheticMPLP()
        roots = kuhn_triangulation(-np.ones(8), np.ones(8), limit=a.nroots)[::a.root_stride]
    else:
        mpc = InvertedPendulumOnCart()
        roots = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'pendulum_roots.npy'))
        if a.roots != 'all':
            roots = roots[[int(x) for x in a.roots.split(',')]]
    orc = Oracle(mpc, a.eps_a, a.eps_r)
    for _ in range(a.repeat):
        t0 = time.time()
        res = partition(orc, roots, wave=a.wave, verbose=a.verbose, max_nodes=a.max_nodes,
                        max_frontier=a.max_frontier, node_budget=a.node_budget)
        st = dict(res.stats)
        st['wall_total_s'] = time.time() - t0
        st['leaves_per_s_device'] = st['nleaves'] / (st['device_ms'] * 1e-3)
        st['solves_per_s_device'] = st['nsolves'] / (st['device_ms'] * 1e-3)
        st['iters_per_solve'] = st['niters'] / max(1, st['nsolves'])
        print(json.dumps(st))
        del res


if __name__ == '__main__':
    main()
