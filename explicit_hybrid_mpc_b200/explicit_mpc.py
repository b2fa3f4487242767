"""
Online evaluation of the explicit control law: mirror of the reference's ``ExplicitMPC``
(``lib/mpc_library.py:662-792``) over the flat node store, evaluated by the sm_100a kernel
``k_locate`` (``csrc/ehm_locate.cuh``) through the C-ABI ``ehm_locator_*``.  SURVEY section 8(f)
row 1.  No CPU path: without the CUDA library / a device the constructor raises.

    empc = ExplicitMPC(tree_or_result)        # reference: ExplicitMPC(tree, oracle)
    u, t = empc(x)                            # reference signature (:770-792)
    U, leaf = empc.evaluate_batch(X)          # B queries in one launch
"""
import ctypes as C
import time

import numpy as np

from . import _capi


def flat_from_result(res):
    """Flat FULL tree of a :class:`partition.PartitionResult`: the node store plus the right
    spine over the root simplices that ``tools.delaunay`` / ``join_triangulation`` build above
    them (``lib/tools.py:152-222``): spine node i has left = root simplex i, right = spine node
    i+1 (the last one has right = the last root simplex)."""
    nn = res.nnodes
    rn = [int(i) for i in res.root_nodes()]
    nr = len(rn)
    ns = max(nr - 1, 0)
    n1, n = res.vertices.shape[1], res.vertices.shape[2]
    nu = res.vinput.shape[2]
    left = np.concatenate([res.left, np.zeros(ns, np.int32)]).astype(np.int32)
    right = np.concatenate([res.right, np.zeros(ns, np.int32)]).astype(np.int32)
    verts = np.concatenate([res.vertices, np.zeros((ns, n1, n))])
    vin = np.concatenate([res.vinput, np.zeros((ns, n1, nu))])
    for i in range(ns):
        left[nn + i] = rn[i]
        right[nn + i] = nn + i + 1 if i < ns - 1 else rn[nr - 1]
    top = nn if ns > 0 else rn[0]
    return dict(left=left, right=right, vertices=verts, vinput=vin, top=top)


def flat_from_tree(tree):
    """Flat arrays of a reference-style ``Tree`` (``lib/tree.py``; e.g. an unpickled
    ``tree.pkl``).  Nodes without data (the triangulation spine) get zero vertices."""
    nodes, stack = [], [tree]
    index = {}
    while stack:
        t = stack.pop()
        index[id(t)] = len(nodes)
        nodes.append(t)
        if not t.is_leaf():
            stack.append(t.right)
            stack.append(t.left)
    some = next(t for t in nodes if t.is_leaf())
    n1, n = np.asarray(some.data.vertices).shape
    nu = np.asarray(some.data.vertex_inputs).reshape(n1, -1).shape[1]
    N = len(nodes)
    left, right = np.full(N, -1, np.int32), np.full(N, -1, np.int32)
    verts, vin = np.zeros((N, n1, n)), np.zeros((N, n1, nu))
    for i, t in enumerate(nodes):
        if not t.is_leaf():
            left[i], right[i] = index[id(t.left)], index[id(t.right)]
        d = t.data
        if d is not None and getattr(d, 'vertices', None) is not None and np.asarray(d.vertices).shape == (n1, n):
            verts[i] = d.vertices
            if getattr(d, 'vertex_inputs', None) is not None:
                vin[i] = np.asarray(d.vertex_inputs).reshape(n1, nu)
    return dict(left=left, right=right, vertices=verts, vinput=vin, top=0)


class ExplicitMPC:
    """``ExplicitMPC(tree, oracle=None)``; ``tree`` may be a reference-style ``Tree``, a
    :class:`partition.PartitionResult` or a dict of flat arrays."""

    def __init__(self, tree, oracle=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('ExplicitMPC needs a CUDA device (no CPU fallback)')
        if isinstance(tree, dict):
            flat = tree
        elif hasattr(tree, 'root_nodes'):
            flat = flat_from_result(tree)
        else:
            flat = flat_from_tree(tree)
        self.flat = flat
        self.oracle = oracle
        self.n = flat['vertices'].shape[2]
        self.nu = flat['vinput'].shape[2]
        self.device = torch.device('cuda', torch.cuda.current_device())
        L = _capi.lib()
        self._h = C.c_void_p()
        f = {k: np.ascontiguousarray(flat[k], dtype=dt) for k, dt in
             (('left', np.int32), ('right', np.int32), ('vertices', np.float64), ('vinput', np.float64))}
        _capi.check(L.ehm_locator_create(self.n, self.nu, len(f['left']), f['left'].ctypes.data,
                                         f['right'].ctypes.data, f['vertices'].ctypes.data,
                                         f['vinput'].ctypes.data, int(flat['top']), C.byref(self._h)))

    def evaluate_batch(self, X):
        """X: (B, n) array or CUDA tensor -> (U (B, nu), leaf (B,)) CUDA tensors."""
        import torch
        Xd = torch.as_tensor(X, dtype=torch.float64, device=self.device).contiguous().reshape(-1, self.n)
        B = Xd.shape[0]
        U = torch.empty(B, self.nu, dtype=torch.float64, device=self.device)
        leaf = torch.empty(B, dtype=torch.int32, device=self.device)
        _capi.check(_capi.lib().ehm_locator_eval(self._h, B, Xd.data_ptr(), U.data_ptr(), leaf.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
        return U, leaf

    def get_containing_cell(self, x):
        """node id of the leaf that the reference's ``get_containing_cell`` would return"""
        _, leaf = self.evaluate_batch(np.asarray(x, dtype=np.float64)[None])
        return int(leaf.cpu()[0])

    def __call__(self, x):
        tic = time.time()
        U, _ = self.evaluate_batch(np.asarray(x, dtype=np.float64)[None])
        u = U.cpu().numpy()[0]
        return u, time.time() - tic

    def close(self):
        if self._h:
            _capi.lib().ehm_locator_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
