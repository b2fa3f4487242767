"""
Generate ``tests/golden/worker_tree_golden.npz``: the partition tree that the REFERENCE'S OWN
recursion code -- ``Worker.ecc`` / ``Worker.lcss`` of ``lib/worker.py:241-417``, with the
reference's own ``tools.split_along_longest_edge`` and ``tree.Tree`` / ``tree.NodeData``,
imported unmodified from /root/reference -- builds on root simplices of the pendulum fixture
when its six oracle calls are answered by the CPU oracle of this repository (``CpuOracle``,
same signatures and return conventions as ``lib/oracle.py:104-414``; MOSEK/CVXPY are not
installable here).  This pins the control flow of the fused partition loop (SURVEY 8a rows
12-14, 16): close / swap in place / split, the inheritance of vertex costs and inputs with the
midpoint entry replaced, the child order -- against the reference's code rather than against a
restatement of it.  Only the MPI task hand-over (``offload_child_computation``) is short-circuited
to "do it yourself", which the reference itself does whenever no worker is idle
(``lib/worker.py:230-239``).  Needs /root/reference; run here only (about two minutes):

    python tests/golden/make_worker_golden.py                 # roots 22, 44 -> worker_tree_golden.npz
    python tests/golden/make_worker_golden.py 0 root0         # root 0 (contains a swap-in-place
                                                              # visit) -> worker_tree_golden_root0.npz
    python tests/golden/make_worker_golden.py 22 deep deep    # tighter tolerances (eps_a 2.3055, eps_r 0.05)
                                                              # on a depth-5 descendant of root 22
                                                              # -> worker_tree_golden_deep.npz
    python tests/golden/make_worker_golden.py 0 config3 config3   # BASELINE config 3: abs_frac 0.03 (eps_a =
                                                              # 0.27937..., examples.create_oracle) / eps_r
                                                              # 0.05 on a depth-9 descendant of root 0
                                                              # -> worker_tree_golden_config3.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden  # noqa: E402

ROOT_IDS = [22, 44]
EPS_A, EPS_R = 19.21220798973866, 1.0


def main():
    global ROOT_IDS, EPS_A, EPS_R
    suffix = ''
    deep = len(sys.argv) > 3 and sys.argv[3] in ('deep', 'config3')
    depth = 5
    if len(sys.argv) > 1:
        ROOT_IDS = [int(x) for x in sys.argv[1].split(',')]
        suffix = '_' + sys.argv[2]
    if deep:
        EPS_A, EPS_R = 19.21220798973866 / 0.25 * 0.03, 0.05
        if sys.argv[3] == 'config3':   # eps_a as examples.create_oracle derives it (tools_golden.npz)
            EPS_A = float(np.load(os.path.join(HERE, 'tools_golden.npz'))['eps_a'][1])
            depth = 9
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from oracle.oracle_cpu import CpuOracle
    mpc = InvertedPendulumOnCart()
    cpu = CpuOracle(mpc.problem, EPS_A, EPS_R, mpc.theta_scale)
    # the reference has modules called oracle / tree / tools as well: hide this repository's
    for name in [k for k in sys.modules if k == 'oracle' or k.startswith('oracle.')]:
        sys.modules.pop(name)
    make_golden.import_reference_pendulum()            # /root/reference/lib first on sys.path
    import tools as ref_tools
    import tree as ref_tree
    import worker as ref_worker
    ref_tools.info_print = lambda *a, **k: None
    ref_tools.error_print = lambda *a, **k: None
    w = object.__new__(ref_worker.Worker)              # no MPI set-up
    w.oracle = cpu
    w.status_publisher = type('S', (), dict(update=lambda self, **kw: None, data={}))()
    w.alg_call = lambda alg, branch, loc: (w.ecc if alg == 'ecc' else w.lcss)(branch, loc)
    w.offload_child_computation = (
        lambda child, location, which_alg, prioritize_self=False, force='none':
        w.alg_call(which_alg, child, location))
    sys.setrecursionlimit(10000)
    roots = np.load(os.path.join(HERE, 'pendulum_roots.npy'))[ROOT_IDS]
    if deep:   # a descendant simplex, cut with the reference's own split
        sub = []
        for R in roots:
            for d in range(depth):
                R = ref_tools.split_along_longest_edge(R)[d % 2]
            sub.append(R)
        roots = np.array(sub)
    out = dict(paths=[], vertices=[], cand=[], vcost=[], vinput=[], leaf=[], closed=[])
    for r, R in enumerate(roots):
        top = ref_tree.Tree(ref_tree.NodeData(vertices=R.copy()))
        w.ecc(top, '')
        stack = [(top, '')]
        while stack:
            t, p = stack.pop()
            d = t.data
            out['paths'].append('%d:%s' % (r, p))
            out['vertices'].append(np.asarray(d.vertices, dtype=float))
            has = hasattr(d, 'commutation')
            out['cand'].append(mpc.problem.index_from_commutation(np.asarray(d.commutation)) if has else -1)
            out['vcost'].append(np.asarray(d.vertex_costs, dtype=float) if has else np.zeros(5))
            out['vinput'].append(np.asarray(d.vertex_inputs, dtype=float).reshape(5, 1) if has else np.zeros((5, 1)))
            out['leaf'].append(t.is_leaf())
            out['closed'].append(bool(d.is_epsilon_suboptimal))
            if not t.is_leaf():
                stack.append((t.left, p + '0'))
                stack.append((t.right, p + '1'))
    np.savez_compressed(os.path.join(HERE, 'worker_tree_golden%s.npz' % suffix), root_ids=np.array(ROOT_IDS),
                        roots=roots,
                        eps=np.array([EPS_A, EPS_R]), paths=np.array(out['paths']),
                        vertices=np.array(out['vertices']), cand=np.array(out['cand']),
                        vcost=np.array(out['vcost']), vinput=np.array(out['vinput']),
                        leaf=np.array(out['leaf']), closed=np.array(out['closed']))
    print('saved', len(out['paths']), 'nodes,', int(np.sum(out['leaf'])), 'leaves')


if __name__ == '__main__':
    main()
