"""
CPU restatement (NumPy) of the reference's online evaluation ``ExplicitMPC``
(``lib/mpc_library.py:662-792``) over a FLAT tree -- TEST INFRASTRUCTURE ONLY: imported by
tests/, never by the product (``explicit_hybrid_mpc_b200/explicit_mpc.py`` calls the CUDA
kernel through the C-ABI and has no CPU path).

Reference algorithm, restated:
  * ``compute_simplex_basis_inverse`` (:684-712): for every LEFT child and every leaf,
    ``Minv = inv([v1-v0 ... vn-v0])`` (columns), ``c = v0``.
  * ``check_containment`` (:714-736): ``alpha = Minv (x - c)``, ``alpha0 = 1 - sum(alpha)``;
    contained iff every alpha in ``[-eps, 1+eps]``, eps = machine epsilon.
  * ``get_containing_cell`` (:738-768): from the top, at an inner node go LEFT iff x is in the
    left child's simplex, otherwise RIGHT (the right child is never tested).
  * ``__call__`` (:770-792): at the leaf, ``u = alpha0 U[0] + sum_i alpha_i U[i]``.

Pinned against the reference's own class (imported from /root/reference, see
tests/golden/make_explicit_golden.py -> tests/golden/explicit_mpc_golden.npz).
"""
import numpy as np


class ExplicitMPCRef:
    def __init__(self, flat):
        self.left = np.asarray(flat['left'])
        self.right = np.asarray(flat['right'])
        self.vertices = np.asarray(flat['vertices'], dtype=np.float64)
        self.vinput = np.asarray(flat['vinput'], dtype=np.float64)
        self.top = int(flat['top'])
        self.eps = np.finfo(np.float64).eps
        self._minv = {}

    def minv(self, i):
        m = self._minv.get(i)
        if m is None:
            v = self.vertices[i]
            m = np.linalg.inv(np.column_stack([_v - v[0] for _v in v[1:]]))
            self._minv[i] = m
        return m

    def contained(self, x, i):
        alpha = list(self.minv(i).dot(x - self.vertices[i][0]))
        alpha.append(1 - sum(alpha))
        return all(a >= -self.eps and a <= 1 + self.eps for a in alpha)

    def locate(self, x):
        cur = self.top
        while self.left[cur] >= 0:
            L = int(self.left[cur])
            cur = L if self.contained(x, L) else int(self.right[cur])
        return cur

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        i = self.locate(x)
        alpha = self.minv(i).dot(x - self.vertices[i][0])
        alpha0 = 1 - sum(alpha)
        u = alpha0 * self.vinput[i][0] + self.vinput[i][1:].T.dot(alpha)
        return u, i
