"""
ctypes front end of the CPU oracle (``oracle/ehm_oracle.c``).  TEST INFRASTRUCTURE ONLY: used
by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs, never by the product package.

``CpuOracle`` exposes the six public methods of the reference's ``Oracle``
(``lib/oracle.py:104-414``) with the reference's return conventions, so the parity tests
read like calls into the reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile ``libehm_oracle.so`` with the committed Makefile (gcc only)."""
    subprocess.check_call(['make', '-C', _HERE, '-s'])


class _Tree(C.Structure):
    _fields_ = [('nnodes', C.c_long), ('n', C.c_int), ('nu', C.c_int),
                ('left', C.POINTER(C.c_int)), ('right', C.POINTER(C.c_int)),
                ('root_id', C.POINTER(C.c_int)), ('depth', C.POINTER(C.c_int)),
                ('cand', C.POINTER(C.c_int)), ('closed', C.POINTER(C.c_ubyte)),
                ('vertices', C.POINTER(C.c_double)), ('vcost', C.POINTER(C.c_double)),
                ('vinput', C.POINTER(C.c_double)),
                ('nleaves', C.c_long), ('nsolves', C.c_long), ('nfail', C.c_long), ('max_depth', C.c_int),
                ('error', C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libehm_oracle.so')
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ehmo_create.restype = C.c_void_p
        L.ehmo_create.argtypes = [C.c_int] * 5 + [dp] * 11 + [C.c_double, C.c_double]
        L.ehmo_destroy.argtypes = [C.c_void_p]
        L.ehmo_set_eps.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.ehmo_num_candidates.argtypes = [C.c_void_p]
        L.ehmo_solve_count.argtypes = [C.c_void_p]
        L.ehmo_solve_count.restype = C.c_long
        L.ehmo_p_theta_delta.argtypes = [C.c_void_p, dp, C.c_int, dp, dp]
        L.ehmo_p_theta.argtypes = [C.c_void_p, dp, dp, dp]
        L.ehmo_v_r.argtypes = [C.c_void_p, dp, dp, dp]
        L.ehmo_bar_e.argtypes = [C.c_void_p, dp, dp, dp]
        L.ehmo_bar_d.argtypes = [C.c_void_p, dp, dp, C.c_int, dp, dp, dp, ip]
        L.ehmo_min_over_simplex.argtypes = [C.c_void_p, dp, C.c_int, dp]
        L.ehmo_set_presolve.argtypes = [C.c_int]
        L.ehmo_split.argtypes = [C.c_int, dp, dp, dp, ip]
        L.ehmo_simplex_volume.argtypes = [C.c_int, dp]
        L.ehmo_simplex_volume.restype = C.c_double
        L.ehmo_partition.restype = C.POINTER(_Tree)
        L.ehmo_partition.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, C.c_long, C.c_int]
        L.ehmo_tree_free.argtypes = [C.POINTER(_Tree)]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class CpuOracle:
    """CPU oracle over a ``HybridMPCProblem`` description."""

    def __init__(self, problem, eps_a, eps_r, theta_scale=None):
        self.problem = pb = problem
        self.eps_a, self.eps_r = float(eps_a), float(eps_r)
        ts = np.ones(pb.n_x) if theta_scale is None else _f64(theta_scale)
        self._keep = [_f64(x) for x in (pb.A, pb.B, pb.w, pb.C, pb.D, pb.e, pb.Q, pb.P, pb.R,
                                        pb.ru, ts)]
        self._h = lib().ehmo_create(pb.n_x, pb.n_u, pb.N, pb.K, pb.p, *[_dp(a) for a in self._keep],
                                    self.eps_a, self.eps_r)
        if not self._h:
            raise RuntimeError('ehmo_create failed (problem too large for the oracle)')
        self.n, self.nu = pb.n_x, pb.n_u
        self.ncand = lib().ehmo_num_candidates(self._h)

    def __del__(self):
        if getattr(self, '_h', None):
            lib().ehmo_destroy(self._h)
            self._h = None

    def set_eps(self, eps_a, eps_r):
        self.eps_a, self.eps_r = float(eps_a), float(eps_r)
        lib().ehmo_set_eps(self._h, self.eps_a, self.eps_r)

    @property
    def solve_count(self):
        return lib().ehmo_solve_count(self._h)

    # ---- reference-shaped methods -------------------------------------------------------
    def P_theta(self, theta, check_feasibility=False):
        J, u0 = C.c_double(), np.zeros(self.nu)
        c = lib().ehmo_p_theta(self._h, _dp(_f64(theta)), C.byref(J), _dp(u0))
        if check_feasibility:
            return c >= 0
        if c < 0:
            return None, None, np.inf, 0.0
        return self.problem.u_scale * u0, self.problem.commutation_from_index(c), J.value, 0.0

    def P_theta_delta(self, theta, delta, check_feasibility=False):
        c = delta if np.isscalar(delta) else self.problem.index_from_commutation(delta)
        J, u0 = C.c_double(), np.zeros(self.nu)
        ok = lib().ehmo_p_theta_delta(self._h, _dp(_f64(theta)), int(c), C.byref(J), _dp(u0))
        if check_feasibility:
            return ok == 1
        if ok != 1:
            return None, np.inf, 0.0
        return self.problem.u_scale * u0, J.value, 0.0

    def V_R(self, R):
        R = _f64(R)
        Jv, Uv = np.zeros(self.n + 1), np.zeros((self.n + 1, self.nu))
        c = lib().ehmo_v_r(self._h, _dp(R), _dp(Jv), _dp(Uv))
        if c < 0:
            return None, None
        return (self.problem.commutation_from_index(c),
                [(self.problem.u_scale * Uv[i], Jv[i], 0.0) for i in range(self.n + 1)])

    def bar_E_delta_R(self, R, V_delta_R, return_tstar=False):
        R, Vr = _f64(R), _f64(V_delta_R)
        ts = np.zeros(self.ncand)
        r = lib().ehmo_bar_e(self._h, _dp(R), _dp(Vr), _dp(ts))
        return (bool(r), ts) if return_tstar else bool(r)

    def bar_D_delta_R(self, R, V_delta_R, delta_ref):
        R, Vr = _f64(R), _f64(V_delta_R)
        cref = delta_ref if np.isscalar(delta_ref) else self.problem.index_from_commutation(delta_ref)
        th, Jv, Uv = np.zeros(self.n), np.zeros(self.n + 1), np.zeros((self.n + 1, self.nu))
        vs = C.c_int()
        c = lib().ehmo_bar_d(self._h, _dp(R), _dp(Vr), int(cref), _dp(th), _dp(Jv), _dp(Uv),
                             C.byref(vs))
        if c < 0:
            return None, None, None, None
        return (self.problem.commutation_from_index(c), th,
                [(self.problem.u_scale * Uv[i], Jv[i], 0.0) for i in range(self.n + 1)],
                bool(vs.value))

    def min_over_simplex(self, R, cand):
        v = C.c_double()
        st = lib().ehmo_min_over_simplex(self._h, _dp(_f64(R)), int(cand), C.byref(v))
        return v.value if st == 0 else None

    # ---- geometry -----------------------------------------------------------------------
    @staticmethod
    def split_along_longest_edge(R):
        R = _f64(R)
        n = R.shape[1]
        S1, S2, ij = np.zeros_like(R), np.zeros_like(R), (C.c_int * 2)()
        lib().ehmo_split(n, _dp(R), _dp(S1), _dp(S2), ij)
        return S1, S2, (ij[0], ij[1])

    @staticmethod
    def simplex_volume(R):
        R = _f64(R)
        return lib().ehmo_simplex_volume(R.shape[1], _dp(R))

    # ---- whole tree ---------------------------------------------------------------------
    def partition(self, roots, nthreads=1, max_nodes=0, prune=True):
        """Serial-DFS ECC/L-CSS partition of the root simplices ``roots`` (nroots,n+1,n).
        Returns a dict of flat node arrays."""
        roots = _f64(roots)
        tp = lib().ehmo_partition(self._h, roots.shape[0], _dp(roots), int(nthreads),
                                  int(max_nodes), int(bool(prune)))
        t = tp.contents
        nn, n, nu = t.nnodes, t.n, t.nu

        def arr(ptr, shape, dt):
            return np.ctypeslib.as_array(ptr, shape=shape).astype(dt).copy() if nn else np.zeros(shape, dt)
        out = dict(
            left=arr(t.left, (nn,), np.int64), right=arr(t.right, (nn,), np.int64),
            root_id=arr(t.root_id, (nn,), np.int64), depth=arr(t.depth, (nn,), np.int64),
            cand=arr(t.cand, (nn,), np.int64), closed=arr(t.closed, (nn,), np.uint8),
            vertices=arr(t.vertices, (nn, n + 1, n), np.float64),
            vcost=arr(t.vcost, (nn, n + 1), np.float64),
            vinput=arr(t.vinput, (nn, n + 1, nu), np.float64) * self.problem.u_scale,
            nleaves=int(t.nleaves), nsolves=int(t.nsolves), nfail=int(t.nfail),
            max_depth=int(t.max_depth),
            error=int(t.error))
        lib().ehmo_tree_free(tp)
        return out
