"""
Hybrid MPC problem definitions for the partitioning hot path (host side).

Mirror of the part of the reference's ``lib/mpc_library.py`` that the oracle consumes:
the *numbers* of the parametric mixed-integer program, not CVXPY objects.  Every example
is lowered to one flat description, :class:`HybridMPCProblem` -- a switched-affine
prediction model with mode-dependent polyhedral stage constraints and a quadratic (or
linear) cost:

    x_0 = theta,   x_{k+1} = A[i_k] x_k + B[i_k] u_k + w[i_k],          k = 0..N-1
    C[i_k] x_k + D[i_k] u_k <= e[i_k]                                  k = 0..N-1
    V = sum_k u_k' R u_k + ru[i_k]' u_k + sum_{k=1}^{N-1} x_k' Q x_k + x_N' P x_N

where ``i_k in {0..K-1}`` is the mode selected by the one-hot commutation block
``delta[k*K:(k+1)*K]``.  The CUDA engine (``csrc/``) condenses this per commutation
(mode sequence) on the device; nothing here is solved on the host.

Why this form is exact for the reference's pendulum (``lib/mpc_library.py:520-552``):
with ``sum_i z[k][i]==1`` (``:526-527``) the big-M pair of the active mode
(``:530-534`` with ``1-z=0``) is the equality ``x[k+1]==A_i x[k]+B_i u[k]+w_i``; the
pairs of the four inactive modes become the local rows
``|(A_i-A_j) x_k + (B_i-B_j) u_k + (w_i-w_j)| <= bigM``; the eight friction-logic rows
(``:537-548``) and the two input bounds (``:550-551``) are already local in
``(x_k,u_k)``.  Row order inside a stage follows the reference's constraint list.
"""
import json
import os
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


@dataclass
class HybridMPCProblem:
    """Flat switched-affine MPC description handed to the C-ABI (``ehm_problem_create``)."""
    n_x: int                 # state = parameter dimension n
    n_u: int                 # input dimension (engine variable is the *scaled* input uhat)
    N: int                   # prediction horizon
    K: int                   # modes per step (= reference ``delta_size``)
    A: np.ndarray            # (K,n,n)
    B: np.ndarray            # (K,n,n_u)   already multiplied by u_scale
    w: np.ndarray            # (K,n)
    C: np.ndarray            # (K,p,n)     stage rows  C x_k + D uhat_k <= e
    D: np.ndarray            # (K,p,n_u)   already multiplied by u_scale
    e: np.ndarray            # (K,p)
    Q: np.ndarray            # (n,n)       stage weight on x_1..x_{N-1}
    P: np.ndarray            # (n,n)       terminal weight on x_N
    R: np.ndarray            # (n_u,n_u)   weight on uhat_k (already scaled)
    ru: np.ndarray           # (K,n_u)     linear cost on uhat_k per mode (LP examples)
    u_scale: np.ndarray      # (n_u,)      u = u_scale * uhat
    name: str = 'custom'
    meta: dict = field(default_factory=dict)

    @property
    def p(self):
        return self.C.shape[1]

    @property
    def delta_size(self):
        return self.K

    @property
    def num_commutations(self):
        return self.K ** self.N

    # -- commutation <-> candidate index -------------------------------------------------
    def commutation_from_index(self, c):
        """Candidate index (base-K digits, step 0 most significant) -> 0/1 float vector
        laid out like the reference's ``delta`` (``delta[k*K+i]``)."""
        delta = np.zeros(self.K * self.N)
        for k in range(self.N - 1, -1, -1):
            delta[k * self.K + (c % self.K)] = 1.0
            c //= self.K
        return delta

    def index_from_commutation(self, delta):
        d = np.asarray(delta).reshape(self.N, self.K)
        d = np.rint(d).astype(int)
        if not np.all(d.sum(axis=1) == 1):
            return -1                      # not one-hot: violates sum(z[k])==1
        c = 0
        for k in range(self.N):
            c = c * self.K + int(np.argmax(d[k]))
        return c

    def validate(self):
        n, m, N, K, p = self.n_x, self.n_u, self.N, self.K, self.p
        for nm, shape in (('A', (K, n, n)), ('B', (K, n, m)), ('w', (K, n)), ('C', (K, p, n)),
                          ('D', (K, p, m)), ('e', (K, p)), ('Q', (n, n)), ('P', (n, n)),
                          ('R', (m, m)), ('ru', (K, m)), ('u_scale', (m,))):
            a = np.ascontiguousarray(getattr(self, nm), dtype=np.float64)
            if a.shape != shape:
                raise ValueError('%s has shape %s, expected %s' % (nm, a.shape, shape))
            setattr(self, nm, a)
        self.validate_unit_box()
        return self

    def validate_unit_box(self):
        """Every mode must carry the input-bound rows ``+-uhat_j <= 1`` (the reference's ``|u| <= u_max``
        rows, ``lib/mpc_library.py:550-551``, in scaled inputs): the row pre-solve of the oracle kernels
        (DESIGN.md "Row pre-solve") drops rows that cannot bind inside that box, which is exact only
        if the box is part of the problem."""
        for i in range(self.K):
            for j in range(self.n_u):
                for sgn in (1.0, -1.0):
                    hit = [r for r in range(self.p)
                           if not self.C[i, r].any() and sgn * self.D[i, r, j] > 0.0
                           and np.count_nonzero(self.D[i, r]) == 1
                           and self.e[i, r] <= abs(self.D[i, r, j]) * (1.0 + 1e-12)]
                    if not hit:
                        raise ValueError('mode %d has no row %+d*uhat_%d <= 1: scale the inputs so that '
                                         'the input bounds are the unit box (u = u_scale * uhat)' % (i, int(sgn), j))


def _discretize(A_c, B_c, w_c, h):
    """Exact ZOH discretisation of xdot = A_c x + B_c u + w_c (ref ``discrete_plant``,
    ``lib/mpc_library.py:556-586``: three augmented-matrix exponentials)."""
    n = A_c.shape[0]
    A = sla.expm(A_c * h)
    H = np.block([[A_c, B_c.reshape(n, 1)], [np.zeros((1, n + 1))]])
    B = sla.expm(H * h)[:n, n:].flatten()
    H = np.block([[A_c, np.eye(n)], [np.zeros((n, 2 * n))]])
    w = sla.expm(H * h)[:n, n:].dot(w_c)
    return A, B, w


def load_pendulum_linearization():
    """Contents of the reference's absent ``lib/sage/*.pkl`` (``mpc_library.py:416-430``),
    re-derived by ``scripts/derive_pendulum_linearization.py``."""
    with open(os.path.join(_DATA_DIR, 'pendulum_linearization.json')) as f:
        d = json.load(f)
    lin = {k: {kk: np.array(vv, dtype=float) for kk, vv in v.items()}
           for k, v in d['linearization'].items()}
    return d['parameters'], lin


class InvertedPendulumOnCart:
    """Inverted pendulum on a cart with static/kinetic friction
    (reference ``lib/mpc_library.py:410-554``).  ``N`` must be 5: the reference indexes
    ``delta[k*N+i]`` (``:524``) which is a valid one-hot layout only when ``N==delta_size``."""

    def __init__(self, N=5):
        if N != 5:
            raise ValueError('pendulum horizon must be 5 (reference delta[k*N+i] layout, '
                             'mpc_library.py:524)')
        self.N = N
        pars, lin_map = load_pendulum_linearization()
        g, l, m, M = pars['g'], pars['l'], pars['m'], pars['M']
        self.v_eps, self.a_eps, self.T_s = pars['v_eps'], pars['a_eps'], pars['T_s']
        self.num_opts = len(lin_map)
        self.A_c = [lin_map['opt%d' % (i + 1)]['A'] for i in range(self.num_opts)]
        self.B_c = [lin_map['opt%d' % (i + 1)]['B'] for i in range(self.num_opts)]
        self.w_c = [lin_map['opt%d' % (i + 1)]['w'] for i in range(self.num_opts)]
        self.n_x, self.n_u = 4, 1
        self.delta_size = 5
        disc = [_discretize(self.A_c[i], self.B_c[i], self.w_c[i], self.T_s)
                for i in range(self.num_opts)]
        A = [d[0] for d in disc]
        B = [d[1] for d in disc]
        w = [d[2] for d in disc]
        self.A_d, self.B_d, self.w_d = A, B, w

        # DLQR terminal weight (:437-443, :466-471)
        A_lqr = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [0, -m * g / M, 0, 0],
                          [0, g / l * (1 + m / M), 0, 0]])
        B_lqr = np.array([0, 0, 1 / M, -1 / (M * l)])
        A_dlqr = sla.expm(A_lqr * self.T_s)
        H = np.block([[A_lqr, B_lqr.reshape(4, 1)], [np.zeros((1, 5))]])
        B_dlqr = sla.expm(H * self.T_s)[:4, 4:].flatten()

        p_max, v_max, a_max = 1., 1., 30.
        ang_max, rate_max, F_max = np.deg2rad(10), np.deg2rad(50), 20.
        bigM = np.array([p_max, ang_max, v_max, rate_max])
        self.p_err, self.v_err = 1., 1.
        self.ang_err, self.rate_err = np.deg2rad(5), np.deg2rad(20)
        self.D_x = np.diag([self.p_err, self.ang_err, self.v_err, self.rate_err])
        self.D_u = np.diag([F_max])
        self.theta_scale = np.diag(self.D_x).copy()   # natural magnitude of theta
        Qhat = np.diag([0.1, 1, 1, 10])
        Rhat = np.eye(1)
        Dxi, Dui = np.linalg.inv(self.D_x), np.linalg.inv(self.D_u)
        Q = Dxi.T.dot(Qhat).dot(Dxi)
        R = Dui.T.dot(Rhat).dot(Dui)
        P = sla.solve_discrete_are(A_dlqr, B_dlqr.reshape(4, 1), Q, R)
        self.Q, self.R, self.P = Q, R, P
        self.bigM, self.a_max, self.F_max = bigM, a_max, F_max

        # ---- stage rows per active mode, in the reference's constraint order ----------
        K, n = self.num_opts, self.n_x
        rows_C, rows_D, rows_e = [], [], []
        v_eps, a_eps = self.v_eps, self.a_eps
        ev = np.zeros(n)
        ev[2] = 1.0                                    # selects x[k][2] (cart velocity)
        for i in range(K):
            Ci, Di, ei = [], [], []

            def add(cx, du, rhs):
                Ci.append(np.asarray(cx, dtype=float))
                Di.append(float(du))
                ei.append(float(rhs))
            # big-M dynamics rows of the inactive modes (:529-534); the active mode's own
            # pair is the equality that defines x[k+1] and has been eliminated.
            for j in range(K):
                if j == i:
                    continue
                dA, dB, dw = A[i] - A[j], B[i] - B[j], w[i] - w[j]
                for r in range(n):                     # x[k+1] <= A_j x + B_j u + w_j + bigM
                    add(dA[r], dB[r], bigM[r] - dw[r])
            for j in range(K):
                if j == i:
                    continue
                dA, dB, dw = A[i] - A[j], B[i] - B[j], w[i] - w[j]
                for r in range(n):                     # x[k+1] >= A_j x + B_j u + w_j - bigM
                    add(-dA[r], -dB[r], bigM[r] + dw[r])
            z = np.zeros(K)
            z[i] = 1.0
            s234 = z[2] + z[3] + z[4]
            # friction logic (:537-548)
            add(-ev, 0., -(v_eps * z[0] - bigM[2] * (1 - z[0])))          # x2 >= ...
            add(ev, 0., -v_eps * z[1] + bigM[2] * (1 - z[1]))             # x2 <= ...
            add(ev, 0., v_eps * s234 + bigM[2] * (1 - s234))
            add(-ev, 0., v_eps * s234 + bigM[2] * (1 - s234))
            a3 = (self.A_c[2][2], self.B_c[2][2], self.w_c[2][2])
            a4 = (self.A_c[3][2], self.B_c[3][2], self.w_c[3][2])
            a5 = (self.A_c[4][2], self.B_c[4][2], self.w_c[4][2])
            add(-a3[0], -a3[1], a3[2] - (a_eps * z[2] - a_max * (1 - z[2])))   # accel3 >= ...
            add(a4[0], a4[1], -a4[2] + (-a_eps * z[3] + a_max * (1 - z[3])))   # accel4 <= ...
            add(-a5[0], -a5[1], a5[2] + (a_eps * z[4] + a_max * (1 - z[4])))   # accel5 >= -...
            add(a5[0], a5[1], -a5[2] + (a_eps * z[4] + a_max * (1 - z[4])))    # accel5 <= ...
            # input bounds (:550-551)
            add(np.zeros(n), 1., F_max)
            add(np.zeros(n), -1., F_max)
            rows_C.append(np.array(Ci))
            rows_D.append(np.array(Di).reshape(-1, 1))
            rows_e.append(np.array(ei))
        us = np.array([F_max])
        self.problem = HybridMPCProblem(
            n_x=4, n_u=1, N=N, K=K,
            A=np.array(A), B=np.array(B).reshape(K, 4, 1) * us, w=np.array(w),
            C=np.array(rows_C), D=np.array(rows_D) * us, e=np.array(rows_e),
            Q=Q, P=P, R=R * us[0] ** 2, ru=np.zeros((K, 1)), u_scale=us,
            name='pendulum',
            meta=dict(v_eps=v_eps, p_err=self.p_err, ang_err=self.ang_err,
                      v_err=self.v_err, rate_err=self.rate_err)).validate()


class SyntheticMPLP:
    """Synthetic multiparametric mixed-integer LP (no reference counterpart; SURVEY 8d
    config 5): ``min c_k' z  s.t.  A_k z <= b_k + S_k theta,  |z_j| <= zmax``, z in R^nz, K one-hot
    modes, N=1, theta in [-1,1]^n.  Seeded with ``numpy.random.default_rng(seed)``.

    The engine variable is the SCALED input ``uhat = z / zmax`` so that the box rows read
    ``|uhat_j| <= 1`` -- the convention of every :class:`HybridMPCProblem` (the row pre-solve of the
    oracle kernels relies on the input box being the unit box, see ``validate_unit_box``)."""

    def __init__(self, n=8, nz=16, rows=48, K=4, seed=20260922, zmax=10.0):
        rng = np.random.default_rng(seed)
        self.n_x, self.n_u, self.N, self.delta_size = n, nz, 1, K
        self.theta_scale = np.ones(n)
        A = rng.standard_normal((K, rows, nz))
        c = rng.standard_normal((K, nz))
        S = rng.standard_normal((K, rows, n))
        # every row of S is scaled to l1-norm 0.9: with the slack >= 1 below, z0_k stays strictly
        # feasible for EVERY theta in [-1,1]^n, i.e. Theta has no infeasible region (plain N(0,1)
        # entries leave most of the box infeasible for all modes and ECC stops with the reference's
        # 'STOP, Theta contains infeasible regions')
        S *= 0.9 / np.abs(S).sum(axis=2, keepdims=True)
        z0 = rng.standard_normal((K, nz))
        b = np.einsum('krz,kz->kr', A, z0) + 1.0 + np.abs(rng.standard_normal((K, rows)))
        # box rows keep every LP bounded: |z_j| <= zmax, i.e. |uhat_j| <= 1
        Abox = np.concatenate([np.eye(nz), -np.eye(nz)], axis=0)
        D = np.concatenate([A * zmax, np.broadcast_to(Abox, (K,) + Abox.shape)], axis=1)
        S = np.concatenate([S, np.zeros((K, 2 * nz, n))], axis=1)
        b = np.concatenate([b, np.ones((K, 2 * nz))], axis=1)
        self.problem = HybridMPCProblem(
            n_x=n, n_u=nz, N=1, K=K,
            A=np.zeros((K, n, n)), B=np.zeros((K, n, nz)), w=np.zeros((K, n)),
            C=-S, D=D, e=b, Q=np.zeros((n, n)), P=np.zeros((n, n)),
            R=np.zeros((nz, nz)), ru=c * zmax, u_scale=np.full(nz, zmax), name='synthetic_mplp',
            meta=dict(seed=seed, zmax=zmax)).validate()


# name parity with the reference, where ExplicitMPC lives in lib/mpc_library.py:662-792
from .explicit_mpc import ExplicitMPC  # noqa: E402,F401
