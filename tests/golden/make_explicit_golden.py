"""
Generate ``tests/golden/explicit_mpc_golden.npz``: outputs of the REFERENCE'S OWN online
evaluation class ``ExplicitMPC`` (``lib/mpc_library.py:662-792``, imported from /root/reference
through the stand-ins of make_golden.py) on a small pendulum partition tree.

The tree (root simplices 22 and 44 of the fixture, eps of BASELINE config 2) is built by the CPU
oracle, converted to the reference's own ``tree.Tree`` / ``tree.NodeData`` objects under a
right spine as ``tools.delaunay`` would produce, and handed to the reference class.  Saved:
the flat tree arrays, query points, the reference's control inputs and the vertices of the cell
the reference located.  Needs /root/reference; run here only:

    python tests/golden/make_explicit_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden  # noqa: E402  (installs the stand-in modules on import of the reference)


def main():
    # this repository's packages first: the reference has modules named oracle / tree / tools too
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from oracle.oracle_cpu import CpuOracle
    make_golden.import_reference_pendulum()           # puts /root/reference/lib on sys.path
    import mpc_library as ref_mpc                      # the reference module
    import tree as ref_tree
    mpc = InvertedPendulumOnCart()
    roots = np.load(os.path.join(HERE, 'pendulum_roots.npy'))[[22, 44]]
    cpu = CpuOracle(mpc.problem, 19.21220798973866, 1.0, mpc.theta_scale)
    arr = cpu.partition(roots, nthreads=1)
    assert arr['error'] == 0 and arr['nfail'] == 0
    nn = len(arr['left'])
    rn = [int(np.nonzero((arr['depth'] == 0) & (arr['root_id'] == r))[0][0]) for r in range(2)]

    def data(i):
        return ref_tree.NodeData(vertices=arr['vertices'][i].copy(), vertex_costs=arr['vcost'][i].copy(),
                                 vertex_inputs=arr['vinput'][i].copy())
    # right spine over the two root simplices (lib/tools.py:172-189), then the sub-trees
    top = ref_tree.Tree(ref_tree.NodeData(vertices=np.vstack([roots[0], roots[1]])))
    top.grow(data(rn[0]), data(rn[1]))
    stack = [(top.left, rn[0]), (top.right, rn[1])]
    while stack:
        t, i = stack.pop()
        if arr['left'][i] >= 0:
            t.grow(data(arr['left'][i]), data(arr['right'][i]))
            stack.append((t.left, arr['left'][i]))
            stack.append((t.right, arr['right'][i]))
    stub = type('O', (), dict(mpc=type('M', (), dict(plant=None, T_s=0.1))()))()
    empc = ref_mpc.ExplicitMPC(top, stub)
    rng = np.random.default_rng(20260922)
    pts = []
    for k in range(400):
        R = roots[k % 2]
        w = rng.dirichlet(np.ones(5))
        pts.append(w @ R)
    leaves = np.nonzero(arr['left'] < 0)[0]
    for i in leaves[:: max(1, len(leaves) // 100)]:        # leaf centroids
        pts.append(arr['vertices'][i].mean(axis=0))
    pts = np.array(pts)
    U, cells = [], []
    for x in pts:
        u, _ = empc(x)
        U.append(np.atleast_1d(u))
        cells.append(empc.get_containing_cell(x).vertices)
    # flat full tree: node store + one spine node (id nn) on top
    flat = dict(left=np.append(arr['left'], rn[0]).astype(np.int32),
                right=np.append(arr['right'], rn[1]).astype(np.int32),
                vertices=np.concatenate([arr['vertices'], np.zeros((1, 5, 4))]),
                vinput=np.concatenate([arr['vinput'], np.zeros((1, 5, 1))]), top=nn)
    np.savez_compressed(os.path.join(HERE, 'explicit_mpc_golden.npz'), points=pts, u=np.array(U),
                        cell_vertices=np.array(cells), **flat)
    print('saved', len(pts), 'queries on a tree of', nn + 1, 'nodes')


if __name__ == '__main__':
    main()
