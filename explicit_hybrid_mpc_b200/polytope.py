"""
Front end of the path: the part of the reference's ``Polytope`` (``lib/polytope.py:70-140``) that
the examples use to describe the parameter set Theta -- H-representation ``{x : A x <= b}``,
V-representation (list of vertices) or hyper-rectangle ``R=[(lo, hi), ...]`` -- with the same
constructor and attributes (``A``, ``b``, ``V``, ``n``).  The reference converts with pycddlib
(exact rational double description); here the conversions are SciPy / Qhull
(``HalfspaceIntersection`` about the Chebyshev centre, ``ConvexHull``), which is all a bounded,
full-dimensional Theta needs, and a hyper-rectangle's vertices are enumerated directly in
``itertools.product`` order (the order the committed root fixture was built from).
"""
import itertools

import numpy as np


class PolytopeError(Exception):
    pass


class Polytope:
    def __init__(self, A=None, b=None, V=None, R=None):
        if A is not None:
            self.A, self.b = np.asarray(A, dtype=float), np.asarray(b, dtype=float)
            if V is None:
                self._hrep2vrep()
            elif V is not False:
                self.V = [np.asarray(v, dtype=float) for v in V]
        elif V is not None:
            self.V = [np.asarray(v, dtype=float) for v in V]
            if A is None:
                self._vrep2hrep()
        else:
            n = len(R)
            self.A, self.b = np.zeros((2 * n, n)), np.zeros(2 * n)
            for i, (lo, hi) in enumerate(R):
                if lo > hi:
                    raise PolytopeError('x%d_min > x%d_max' % (i, i))
                self.A[2 * i, i], self.A[2 * i + 1, i] = 1, -1
                self.b[2 * i], self.b[2 * i + 1] = hi, -lo
            self.V = [np.array(v, dtype=float) for v in itertools.product(*[(lo, hi) for lo, hi in R])]
        if V is not False and len(self.V) == 0:
            raise PolytopeError('Empty polytope')
        self.n = self.V[0].size if V is not False else self.A.shape[1]

    def chebyshev_centre(self):
        """centre and radius of the largest inscribed ball (one LP)"""
        from scipy.optimize import linprog
        nrm = np.linalg.norm(self.A, axis=1)
        n = self.A.shape[1]
        res = linprog(np.r_[np.zeros(n), -1.0], A_ub=np.c_[self.A, nrm], b_ub=self.b,
                      bounds=[(None, None)] * n + [(0, None)], method='highs')
        if res.status != 0 or res.x[-1] <= 0:
            raise PolytopeError('empty or lower-dimensional polytope')
        return res.x[:n], float(res.x[-1])

    def _hrep2vrep(self):
        from scipy.spatial import HalfspaceIntersection
        c, _ = self.chebyshev_centre()
        hs = HalfspaceIntersection(np.c_[self.A, -self.b], c)
        V = np.unique(np.round(hs.intersections, 12), axis=0)
        self.V = [v for v in V]

    def _vrep2hrep(self):
        from scipy.spatial import ConvexHull
        hull = ConvexHull(np.vstack(self.V))
        eq = np.unique(np.round(hull.equations, 12), axis=0)
        self.A, self.b = eq[:, :-1], -eq[:, -1]

    def contains(self, x, tol=1e-9):
        return bool(np.all(self.A.dot(np.asarray(x, dtype=float)) <= self.b + tol))

    def volume(self):
        from scipy.spatial import ConvexHull
        return float(ConvexHull(np.vstack(self.V)).volume)
