"""
Geometry helpers mirroring the reference's ``lib/tools.py:134-257`` (host side, NumPy).
The partition loop does not call these -- the split runs inside the CUDA engine -- they
exist so that code written against the reference keeps working and for the tests.
"""
import itertools
import math

import numpy as np
import numpy.linalg as la

from .tree import NodeData, Tree


def simplex_volume(R):
    """``lib/tools.py:134-150``: |det[v1-v0 ... vn-v0]| / n!"""
    R = np.asarray(R, dtype=float)
    M = np.column_stack([v - R[0] for v in R[1:]])
    return np.abs(1. / math.factorial(R.shape[0] - 1) * la.det(M))   # the reference's expression, bit for bit


def split_along_longest_edge(R):
    """``lib/tools.py:224-257``: bisect the simplex with vertex rows ``R`` at the midpoint of its
    longest edge.  Returns ``(S_1, S_2, (i, j))``: ``S_1`` is R with vertex i replaced by the
    midpoint, ``S_2`` is R with vertex j replaced.  The edge is the first maximum, in
    ``itertools.combinations`` order, of ``la.norm(R[i]-R[j])`` -- the reference's own formula
    (NumPy evaluates it as the square root of a left-to-right fused-multiply-add chain, which is
    what the CUDA kernels and the CPU oracle restate bit for bit)."""
    R = np.asarray(R, dtype=float)
    edges = list(itertools.combinations(range(len(R)), 2))
    lengths = [la.norm(R[i] - R[j]) for i, j in edges]
    i, j = edges[int(np.argmax(lengths))]
    halves = []
    for replaced in (i, j):
        S = R.copy()
        S[replaced] = (R[i] + R[j]) / 2.
        halves.append(S)
    return halves[0], halves[1], (i, j)


def kuhn_triangulation(lower, upper, limit=None):
    """Deterministic triangulation of the box [lower, upper] into n! simplices (Kuhn / Freudenthal):
    one simplex per permutation p of the axes, vertices v_0 = lower, v_k = v_(k-1) + e_p(k) * width.
    Unlike ``delaunay`` (Qhull) the result does not depend on library versions or vertex order
    (SURVEY 8f row 3).  Permutations in ``itertools.permutations`` order; ``limit`` keeps the first
    ``limit`` of them (BASELINE config 5: 1024 of 8! = 40320)."""
    lower, upper = np.asarray(lower, dtype=float), np.asarray(upper, dtype=float)
    n = lower.shape[0]
    out = []
    for q, perm in enumerate(itertools.permutations(range(n))):
        if limit is not None and q >= limit:
            break
        v = lower.copy()
        S = [v.copy()]
        for ax in perm:
            v[ax] = upper[ax]
            S.append(v.copy())
        out.append(np.array(S))
    return np.array(out)


def delaunay(R):
    """``lib/tools.py:152-189``: Delaunay triangulation of the polytope with vertex rows R
    into a right-spine binary tree whose leaves are the simplices."""
    import scipy.spatial as scs
    R = np.asarray(R, dtype=float)
    simplices = scs.Delaunay(R).simplices
    return spine_from_simplices([R[s] for s in simplices], top_vertices=R)


def spine_from_simplices(simplices, top_vertices=None):
    """Right-spine tree over a list of root simplices (what ``delaunay`` +
    ``join_triangulation`` produce, ``lib/tools.py:172-222``)."""
    Nsx = len(simplices)
    vol = sum(simplex_volume(s) for s in simplices)
    root = Tree(NodeData(vertices=top_vertices if top_vertices is not None else simplices[0]))
    if Nsx == 1:
        root.data = NodeData(vertices=simplices[0])
        return root, Nsx, vol
    cursor = root
    for i in range(Nsx - 1):
        left = NodeData(vertices=simplices[i])
        right = None if i < Nsx - 2 else NodeData(vertices=simplices[i + 1])
        cursor.grow(left, right)
        cursor = cursor.right
    return root, Nsx, vol


def join_triangulation(cursor, new_tree):
    """``lib/tools.py:191-222``: hang ``new_tree`` on the rightmost leaf of ``cursor``'s tree.  That
    leaf turns into a data-less inner node whose left child keeps its simplex and whose right
    child is ``new_tree`` (no longer a top)."""
    node = cursor
    while not node.is_leaf():
        node = node.right
    simplex = node.data
    node.grow(simplex, None)
    node.right, node.data = new_tree, None
    new_tree.top = False


def leaves(tree, location=''):
    """Yield ``(location, node)`` for every leaf, depth first, left before right."""
    if tree.is_leaf():
        yield location, tree
    else:
        yield from leaves(tree.left, location + '0')
        yield from leaves(tree.right, location + '1')
