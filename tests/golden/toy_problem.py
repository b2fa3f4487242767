"""A two-mode toy problem (theta in R^2, u in R^2, N=1) whose modes share dynamics and cost but
are valid on overlapping half-planes (mode 0: theta_0 <= 0.3, mode 1: theta_0 >= -0.3).  On a root
triangle spanning theta_0 in [-1,1] no single commutation is feasible at all vertices, so the
reference's ECC recursion has to split (lib/worker.py:264-277) -- the branch the pendulum example
never takes (every root simplex of it admits a common commutation)."""
import numpy as np


def toy_problem():
    from explicit_hybrid_mpc_b200.mpc_library import HybridMPCProblem
    n, m, K = 2, 2, 2
    C, D, e = np.zeros((K, 7, n)), np.zeros((K, 7, m)), np.zeros((K, 7))
    for k in range(K):
        D[k, 0, 0], e[k, 0] = 1, 1        # |u_j| <= 1
        D[k, 1, 0], e[k, 1] = -1, 1
        D[k, 2, 1], e[k, 2] = 1, 1
        D[k, 3, 1], e[k, 3] = -1, 1
        C[k, 4, 0], D[k, 4, 0], e[k, 4] = 0.5, -1, 0.2      # u_0 >= 0.5 theta_0 - 0.2
        C[k, 5, 1], D[k, 5, 1], e[k, 5] = -0.4, -1, 0.1     # u_1 >= -0.4 theta_1 - 0.1
    C[0, 6, 0], e[0, 6] = 1.0, 0.3        # mode 0: theta_0 <= 0.3
    C[1, 6, 0], e[1, 6] = -1.0, 0.3       # mode 1: theta_0 >= -0.3
    A = np.stack([np.eye(n) * 0.9] * 2)
    B = np.stack([np.array([[1.0, 0], [0, 0.5]])] * 2)
    return HybridMPCProblem(n_x=n, n_u=m, N=1, K=K, A=A, B=B, w=np.zeros((K, n)), C=C, D=D, e=e,
                            Q=np.eye(n), P=np.eye(n) * 2.0, R=np.eye(m) * 0.5, ru=np.zeros((K, m)),
                            u_scale=np.ones(m), name='toy_two_mode').validate()


TOY_ROOT = np.array([[-1.0, -1.0], [1.0, -1.0], [0.0, 1.0]])
TOY_EPS = (0.005, 0.01)
