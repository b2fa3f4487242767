"""
The partition driver: replaces the reference's ``scheduler.main()`` / ``worker.main()`` pair
(``lib/scheduler.py:693-708``, ``lib/worker.py:463-466``) by one call into the device-resident
frontier loop ``ehm_partition_run`` and rebuilds the reference's ``Tree`` / ``NodeData``
objects (``lib/tree.py``) from the flat node store, so ``tree.pkl`` consumers keep working.
"""
import ctypes as C
import time

import numpy as np

from . import _capi
from .tree import NodeData, Tree


class PartitionResult:
    """Flat node arrays of one partition run + run statistics."""

    def __init__(self, arrays, stats, problem):
        self.__dict__.update(arrays)
        self.stats = stats
        self.problem = problem

    @property
    def nnodes(self):
        return len(self.left)

    def leaf_ids(self):
        return np.nonzero(self.left < 0)[0]

    def subtree(self, node_id):
        """Reference ``Tree`` of the subtree rooted at node_id (iterative, depth-safe)."""
        pb = self.problem

        def data(i):
            if self.cand[i] >= 0:
                d = NodeData(vertices=self.vertices[i].copy(),
                             commutation=pb.commutation_from_index(int(self.cand[i])),
                             vertex_costs=self.vcost[i].copy(),
                             vertex_inputs=self.vinput[i].copy())
            else:
                d = NodeData(vertices=self.vertices[i].copy())
            d.is_epsilon_suboptimal = bool(self.closed[i])
            return d
        root = Tree(data(node_id), top=False)
        stack = [(root, node_id)]
        while stack:
            t, i = stack.pop()
            if self.left[i] >= 0:
                t.grow(data(self.left[i]), data(self.right[i]))
                stack.append((t.left, self.left[i]))
                stack.append((t.right, self.right[i]))
        return root

    def graft(self, initial_tree):
        """Graft the partitioned subtrees onto the leaves of the initial triangulation tree
        (what ``build_tree`` does with the ``branch_<location>.pkl`` files,
        ``lib/scheduler.py:644-691``).  Leaves are matched in depth-first order, which is the
        order :func:`roots_of` extracted them in."""
        from .tools import leaves
        for r, (_, leaf) in enumerate(leaves(initial_tree)):
            self.subtree(r).copy(leaf)
            leaf.top = False
        initial_tree.top = True
        return initial_tree


def roots_of(initial_tree):
    """Root simplices (nroots,n+1,n) = the open leaves of the initial tree in depth-first
    order (``Scheduler.populate_queue``, ``lib/scheduler.py:620-642``)."""
    from .tools import leaves
    return np.array([leaf.data.vertices for _, leaf in leaves(initial_tree)], dtype=np.float64)


def partition(oracle, roots, wave=0, max_nodes=0, max_frontier=0, node_budget=0, verbose=0):
    """Run ECC + L-CSS to convergence on the device.

    oracle : :class:`explicit_hybrid_mpc_b200.oracle.Oracle`
    roots  : (nroots, n+1, n) array of root simplices (HOST memory; copied to the device,
             the resulting node store is copied back -- both inside this call).
    """
    roots = np.ascontiguousarray(roots, dtype=np.float64)
    n, nu = oracle.n, oracle.nu
    assert roots.ndim == 3 and roots.shape[1:] == (n + 1, n)
    opts = _capi.PartitionOpts(int(wave), int(max_nodes), int(max_frontier), int(node_budget),
                               int(verbose))
    stats = _capi.PartitionStats()
    h = C.c_void_p()
    L = _capi.lib()
    t0 = time.perf_counter()
    rc = L.ehm_partition_run(oracle._p.handle, roots.shape[0],
                             roots.ctypes.data_as(C.POINTER(C.c_double)), C.byref(opts),
                             C.byref(h), C.byref(stats))
    try:
        _capi.check(rc)
        nn = stats.nnodes
        arr = dict(left=np.empty(nn, np.int32), right=np.empty(nn, np.int32),
                   root_id=np.empty(nn, np.int32), depth=np.empty(nn, np.int32),
                   cand=np.empty(nn, np.int32), closed=np.empty(nn, np.uint8),
                   vertices=np.empty((nn, n + 1, n)), vcost=np.empty((nn, n + 1)),
                   vinput=np.empty((nn, n + 1, nu)))
        _capi.check(L.ehm_partition_fetch(h, *[arr[k].ctypes.data for k in
                                               ('left', 'right', 'root_id', 'depth', 'cand',
                                                'closed', 'vertices', 'vcost', 'vinput')]))
    finally:
        if h:
            L.ehm_partition_free(h)
    wall = time.perf_counter() - t0
    arr['vinput'] = arr['vinput'] * oracle.pb.u_scale
    st = {k: (list(getattr(stats, k)) if k.startswith('stage_') else getattr(stats, k))
          for k, _ in _capi.PartitionStats._fields_}
    st['wall_s'] = wall
    return PartitionResult(arr, st, oracle.pb)
