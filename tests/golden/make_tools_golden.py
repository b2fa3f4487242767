"""
Generate ``tests/golden/tools_golden.npz`` with the REFERENCE'S OWN geometry helpers
(``lib/tools.py``: ``delaunay`` :152-189, ``join_triangulation`` :191-222, ``simplex_volume``
:134-150, ``split_along_longest_edge`` :224-257, imported from /root/reference): the pendulum's
initial triangulation exactly as ``examples.pendulum_example`` (:120-163) assembles it -- three
velocity slabs, each Delaunay-triangulated, joined -- on itertools.product-ordered box vertices
(cdd, which orders them in the reference, is not installable), plus volumes and splits of random
simplices, and the eps_a that ``examples.create_oracle`` (:18-47) derives from abs_frac.
``tests/golden/pendulum_roots.npy`` is the leaf list of that tree.  Run here only:

    python tests/golden/make_tools_golden.py
"""
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden  # noqa: E402


def main():
    # this repository's packages first (the reference has modules named oracle / tools / tree too)
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from oracle.oracle_cpu import CpuOracle
    mine = InvertedPendulumOnCart()
    for name in [k for k in sys.modules if k == 'oracle' or k.startswith('oracle.')]:
        sys.modules.pop(name)
    ref_p = make_golden.import_reference_pendulum()
    import examples as ref_examples
    import tools as ref_tools
    # examples.create_oracle (:18-47) with its Oracle class answered by the CPU oracle: eps_a
    ref_examples.Oracle = lambda mpc, eps_a, eps_r: CpuOracle(mine.problem, eps_a, eps_r, mine.theta_scale)

    def box(lo_hi):
        return np.array(list(itertools.product(*lo_hi)))
    pe, ae, ve, re_, eps = ref_p.p_err, ref_p.ang_err, ref_p.v_err, ref_p.rate_err, ref_p.v_eps
    slabs = [box([(-pe, pe), (-ae, ae), (eps, ve), (-re_, re_)]),        # examples.py:134-149
             box([(-pe, pe), (-ae, ae), (-ve, -eps), (-re_, re_)]),
             box([(-pe, pe), (-ae, ae), (-eps, eps), (-re_, re_)])]
    trees = [ref_tools.delaunay(s)[0] for s in slabs]
    ref_tools.join_triangulation(trees[1], trees[2])                       # examples.py:159-160
    ref_tools.join_triangulation(trees[0], trees[1])
    leaves, stack = [], [trees[0]]
    while stack:
        t = stack.pop()
        if t.is_leaf():
            leaves.append(np.asarray(t.data.vertices))
        else:
            stack.append(t.right)
            stack.append(t.left)
    leaves = np.array(leaves)
    assert np.array_equal(leaves, np.load(os.path.join(HERE, 'pendulum_roots.npy')))
    full = box([(-pe, pe), (-ae, ae), (-ve, ve), (-re_, re_)])
    orc = ref_examples.create_oracle(ref_p, full, 0.25, None, 1.0)         # BASELINE config 2
    eps_a_025 = orc.eps_a
    eps_a_003 = ref_examples.create_oracle(ref_p, full, 0.03, None, 0.05).eps_a   # config 3
    rng = np.random.default_rng(11)
    simplices, vols, s1, s2, ij = [], [], [], [], []
    for k in range(200):
        R = leaves[k % len(leaves)].copy() if k % 2 else rng.standard_normal((5, 4))
        for _ in range(k % 9):
            R = ref_tools.split_along_longest_edge(R)[k % 2]
        a, b, c = ref_tools.split_along_longest_edge(R)
        simplices.append(R); vols.append(ref_tools.simplex_volume(R)); s1.append(a); s2.append(b); ij.append(c)
    np.savez_compressed(os.path.join(HERE, 'tools_golden.npz'), slabs=np.array(slabs), roots=leaves,
                        simplices=np.array(simplices), volumes=np.array(vols), S1=np.array(s1),
                        S2=np.array(s2), ij=np.array(ij), eps_a=np.array([eps_a_025, eps_a_003]))
    print('saved', len(leaves), 'roots,', len(simplices), 'simplices')


if __name__ == '__main__':
    main()
