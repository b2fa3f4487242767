"""
Generate ``tests/golden/worker_tree_golden_toy.npz``: like make_worker_golden.py -- the REFERENCE'S
OWN ``Worker.ecc`` / ``Worker.lcss`` (``lib/worker.py:241-417``) over the six-method oracle API --
but on the two-mode toy problem of toy_problem.py, whose root triangle admits no common
commutation: the recursion goes through the ECC split branch (``:264-277``: barycentre
feasibility test, ``V_R`` returns None, split, recurse with 'ecc') before L-CSS starts.

    python tests/golden/make_worker_golden_toy.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden  # noqa: E402
from toy_problem import TOY_EPS, TOY_ROOT, toy_problem  # noqa: E402


def main():
    from oracle.oracle_cpu import CpuOracle
    pb = toy_problem()
    cpu = CpuOracle(pb, TOY_EPS[0], TOY_EPS[1], np.ones(2))
    for name in [k for k in sys.modules if k == 'oracle' or k.startswith('oracle.')]:
        sys.modules.pop(name)
    make_golden.import_reference_pendulum()
    import tools as ref_tools
    import tree as ref_tree
    import worker as ref_worker
    ref_tools.info_print = lambda *a, **k: None
    ref_tools.error_print = lambda *a, **k: None
    w = object.__new__(ref_worker.Worker)
    w.oracle = cpu
    w.status_publisher = type('S', (), dict(update=lambda self, **kw: None, data={}))()
    w.alg_call = lambda alg, branch, loc: (w.ecc if alg == 'ecc' else w.lcss)(branch, loc)
    w.offload_child_computation = (
        lambda child, location, which_alg, prioritize_self=False, force='none':
        w.alg_call(which_alg, child, location))
    top = ref_tree.Tree(ref_tree.NodeData(vertices=TOY_ROOT.copy()))
    w.ecc(top, '')
    out = dict(paths=[], vertices=[], cand=[], vcost=[], vinput=[], leaf=[], closed=[])
    stack = [(top, '')]
    while stack:
        t, p = stack.pop()
        d = t.data
        has = hasattr(d, 'commutation')
        out['paths'].append('0:' + p)
        out['vertices'].append(np.asarray(d.vertices, dtype=float))
        out['cand'].append(pb.index_from_commutation(np.asarray(d.commutation)) if has else -1)
        out['vcost'].append(np.asarray(d.vertex_costs, dtype=float) if has else np.zeros(3))
        out['vinput'].append(np.asarray(d.vertex_inputs, dtype=float).reshape(3, 2) if has else np.zeros((3, 2)))
        out['leaf'].append(t.is_leaf())
        out['closed'].append(bool(d.is_epsilon_suboptimal))
        if not t.is_leaf():
            stack.append((t.left, p + '0'))
            stack.append((t.right, p + '1'))
    np.savez_compressed(os.path.join(HERE, 'worker_tree_golden_toy.npz'), roots=TOY_ROOT[None],
                        eps=np.array(TOY_EPS), root_ids=np.array([0]), paths=np.array(out['paths']),
                        vertices=np.array(out['vertices']), cand=np.array(out['cand']),
                        vcost=np.array(out['vcost']), vinput=np.array(out['vinput']),
                        leaf=np.array(out['leaf']), closed=np.array(out['closed']))
    print('saved', len(out['paths']), 'nodes,', int(np.sum(out['leaf'])), 'leaves,',
          int(np.sum(np.array(out['cand']) < 0)), 'data-less (ECC) nodes')


if __name__ == '__main__':
    main()
