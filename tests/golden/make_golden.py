"""
Generate ``tests/golden/pendulum_golden.json``: known answers for the pendulum oracle problems
obtained by (1) running the REFERENCE'S OWN problem-construction code -- ``lib/mpc_library.py``
``InvertedPendulumOnCart.__init__`` / ``make_constraints`` (:410-554) imported from
/root/reference through the CVXPY stand-in ``cvx_shim.py`` (CVXPY/MOSEK/mpi4py/cdd are not
installed here) -- and (2) solving the resulting full-variable big-M problems (29 continuous
variables, 250 inequality rows, reference form, NOT the condensed form the engine uses) with
SciPy: HiGHS (``linprog``/``milp``) for feasibility LPs/MILPs and SLSQP for the fixed-delta QPs.

The CPU oracle (oracle/) is pinned against these vectors by tests/test_oracle_golden.py; the
GPU engine is then checked against the CPU oracle.  Needs /root/reference; run here only:

    python tests/golden/make_golden.py
"""
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import scipy.optimize as so

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import cvx_shim  # noqa: E402


def import_reference_pendulum(N=5):
    """Import lib/mpc_library.py with stand-ins for the uninstallable dependencies."""
    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith('__'):
                raise AttributeError(k)
            m = _Any(self.__name__ + '.' + k)
            setattr(self, k, m)
            return m

        def __call__(self, *a, **kw):
            return _Any('call')
    for name in ('mpi4py', 'mpi4py.MPI', 'cdd', 'matplotlib', 'matplotlib.pyplot', 'mpl_toolkits',
                 'mpl_toolkits.mplot3d', 'progressbar', 'mosek'):
        sys.modules[name] = _Any(name)
    sys.modules['cvxpy'] = cvx_shim
    sys.path.insert(0, os.path.join(REF, 'lib'))
    import global_vars
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, 'lib', 'sage'))
    lin = json.load(open(os.path.join(ROOT, 'explicit_hybrid_mpc_b200', 'data',
                                      'pendulum_linearization.json')))
    with open(os.path.join(tmp, 'lib', 'sage', 'pendulum_parameters.pkl'), 'wb') as f:
        pickle.dump(lin['parameters'], f)
    lm = {k: {kk: np.array(vv) for kk, vv in v.items()} for k, v in lin['linearization'].items()}
    with open(os.path.join(tmp, 'lib', 'sage', 'pendulum_linearization.pkl'), 'wb') as f:
        pickle.dump(lm, f)
    global_vars.PROJECT_DIR = tmp
    global_vars.MPC_HORIZON = N
    import mpc_library
    cvx_shim.reset()
    return mpc_library.InvertedPendulumOnCart()


class RefProblem:
    """Reference-form matrices for fixed numeric theta; delta either fixed or free binaries."""

    def __init__(self, mpc):
        self.mpc = mpc
        self.nz = cvx_shim.nvar()          # uhat(5), xhat(24) and the 25 delta binaries
        d = mpc.delta
        self.idelta = np.arange(d.index0, d.index0 + d.n)
        self.icont = np.setdiff1d(np.arange(self.nz), self.idelta)
        self.H, self.f, self.c = mpc.V.matrices()   # V = z'Hz + f'z + c

    def rows(self, theta, delta=None, delta_sum=True):
        """(A_ub, b_ub, A_eq, b_eq) over the full variable vector z for x0 = theta.
        delta = 0/1 vector fixes the commutation (a Parameter), None keeps it a variable."""
        mpc = self.mpc
        mpc.x0.value = np.asarray(theta, dtype=float)
        dl = mpc.delta if delta is None else cvx_shim.Parameter(len(delta), value=delta)
        cons = mpc.make_constraints(mpc.x0, mpc.x, mpc.u, dl, delta_sum_constraint=delta_sum)
        Aub, bub, Aeq, beq = [], [], [], []
        for kind, e in cons:
            e = cvx_shim.Expr.wrap(e)
            (Aub if kind == 'le' else Aeq).append(e.A)
            (bub if kind == 'le' else beq).append(-e.b)
        return np.vstack(Aub), np.concatenate(bub), np.vstack(Aeq), np.concatenate(beq)

    def solve_fixed(self, theta, delta):
        """P_theta_delta in reference form: feasibility by HiGHS, optimum by SLSQP."""
        Aub, bub, Aeq, beq = self.rows(theta, delta)
        ic = self.icont
        # constant equality rows (sum(z[k])==1 with numeric delta) must hold
        const_eq = np.abs(Aeq[:, ic]).max(axis=1) == 0
        if np.any(np.abs(beq[const_eq]) > 1e-9):
            return dict(feasible=False)
        Aeq, beq = Aeq[~const_eq][:, ic], beq[~const_eq]
        Aub = Aub[:, ic]
        lp = so.linprog(np.zeros(len(ic)), A_ub=Aub, b_ub=bub, A_eq=Aeq, b_eq=beq,
                        bounds=[(None, None)] * len(ic), method='highs')
        if lp.status != 0:
            return dict(feasible=False)
        H, f = self.H[np.ix_(ic, ic)], self.f[ic]
        # eliminate the equalities exactly (null-space), then SLSQP on the inequalities
        z0 = np.linalg.lstsq(Aeq, beq, rcond=None)[0]
        _, sv, Vt = np.linalg.svd(Aeq)
        Z = Vt[np.sum(sv > 1e-12):].T
        Hr, fr = Z.T @ H @ Z, Z.T @ (2 * H @ z0 + f)
        Ar, br = Aub @ Z, bub - Aub @ z0
        # drop rows that vanished after the elimination (the active mode's own big-M pair)
        keep = np.abs(Ar).max(axis=1) > 1e-12
        assert np.all(br[~keep] > -1e-9)
        Ar, br = Ar[keep], br[keep]
        # interior margin of the feasible set in the reduced coordinates: max t with
        # Ar y + t |Ar_row|_inf <= br.  0 <=> the set has empty interior ("thin": feasible only
        # on a measure-zero slice, which the engine's canonical semantics excludes).
        # (the active mode's big-M pair x[k+1] <= .. and x[k+1] >= .. is an equality written as
        # two inequalities: pair rows are detected and excluded from the margin)
        paired = np.zeros(len(br), dtype=bool)
        for i in range(len(br)):
            if paired[i]:
                continue
            j = np.nonzero((np.abs(Ar + Ar[i]).max(axis=1) < 1e-12) & (np.abs(br + br[i]) < 1e-12))[0]
            if len(j):
                paired[i] = True
                paired[j] = True
        Aq, bq = Ar[paired], br[paired]
        Am, bm = Ar[~paired], br[~paired]
        nr = np.abs(Am).max(axis=1)
        mg = so.linprog(np.r_[np.zeros(Z.shape[1]), -1.0],
                        A_ub=np.r_[np.c_[Am, nr], np.c_[Aq, np.zeros(len(bq))]], b_ub=np.r_[bm, bq],
                        bounds=[(None, None)] * Z.shape[1] + [(None, 1.0)], method='highs')
        margin = float(mg.x[-1]) if mg.status == 0 else 0.0
        y0 = np.linalg.lstsq(Z, lp.x - z0, rcond=None)[0]
        best = None
        for start in (y0, np.zeros(Z.shape[1])):
            r = so.minimize(lambda y: y @ Hr @ y + fr @ y, start, jac=lambda y: 2 * Hr @ y + fr,
                            constraints=[dict(type='ineq', fun=lambda y: br - Ar @ y, jac=lambda y: -Ar)],
                            method='SLSQP', options=dict(ftol=1e-14, maxiter=500))
            if r.success and (best is None or r.fun < best.fun):
                best = r
        if best is None:
            return dict(feasible=True, J=None, margin=margin)
        z = z0 + Z @ best.x
        J = float(z @ H @ z + f @ z + self.c)
        viol = float(np.max(Aub @ z - bub))
        u0 = float((self.mpc.u[0]._pad().A[:, ic] @ z + self.mpc.u[0].b)[0])
        return dict(feasible=True, J=J, u0=u0, max_violation=viol, margin=margin)

    def milp_feasible(self, thetas):
        """V_R-type feasibility MILP: one shared binary delta, one (x,u) copy per theta
        (lib/oracle.py:57-66).  Returns a feasible delta or None (HiGHS branch & bound)."""
        ic, idl = self.icont, self.idelta
        nc, nd, V = len(ic), len(idl), len(thetas)
        rowsA, lo, hi = [], [], []
        for v, th in enumerate(thetas):
            Aub, bub, Aeq, beq = self.rows(th, None, delta_sum=(v == 0))
            for A, b, eq in ((Aub, bub, False), (Aeq, beq, True)):
                blk = np.zeros((A.shape[0], nd + V * nc))
                blk[:, :nd] = A[:, idl]
                blk[:, nd + v * nc: nd + (v + 1) * nc] = A[:, ic]
                rowsA.append(blk)
                hi.append(b)
                lo.append(b if eq else np.full(len(b), -np.inf))
        A = np.vstack(rowsA)
        integrality = np.concatenate([np.ones(nd), np.zeros(V * nc)])
        bounds = so.Bounds(np.concatenate([np.zeros(nd), np.full(V * nc, -np.inf)]),
                           np.concatenate([np.ones(nd), np.full(V * nc, np.inf)]))
        r = so.milp(np.zeros(nd + V * nc), constraints=so.LinearConstraint(A, np.concatenate(lo), np.concatenate(hi)),
                    integrality=integrality, bounds=bounds)
        if r.status != 0 or r.x is None:
            return None
        return np.rint(r.x[:nd])


def main():
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from oracle.oracle_cpu import CpuOracle
    ref = RefProblem(import_reference_pendulum())
    mine = InvertedPendulumOnCart()
    orc = CpuOracle(mine.problem, 27.96235132870629, 1.0, mine.theta_scale)
    pb = mine.problem
    rng = np.random.default_rng(20260922)
    ts = mine.theta_scale
    out = dict(note='reference-form solutions (SciPy HiGHS/SLSQP) of problems built by the '
                    "reference's own lib/mpc_library.py code; see make_golden.py",
               fixed=[], milp_point=[], milp_simplex=[])
    # ---- fixed-delta problems: feasible ones (argmin commutations) and random ones ---------
    thetas = [np.array([0.3, 0.02, 0.5, 0.1]), 0.25 * ts * np.array([1, -1, 1, -1]),
              np.array([-0.5, 0.05, -0.4, 0.2]), ts * np.array([1, 1, 1e-3, 1]),
              np.array([0.1, -0.03, 5e-4, -0.1]), np.array([0.7, 0.01, -0.8, 0.05])]
    thetas += [rng.uniform(-1, 1, 4) * ts for _ in range(6)]
    for th in thetas:
        d_opt = orc.P_theta(th)[1]
        cands = [pb.index_from_commutation(d_opt)] + [int(c) for c in rng.integers(0, pb.K ** pb.N, 3)]
        # a few near-optimal alternatives: flip the last step's mode
        cands += [cands[0] - cands[0] % 5 + k for k in range(5)]
        for c in sorted(set(cands)):
            g = ref.solve_fixed(th, pb.commutation_from_index(c))
            g.update(theta=th.tolist(), cand=int(c))
            out['fixed'].append(g)
    # ---- P_theta feasibility as the reference poses it: MILP over binary delta --------------
    for th in thetas[:8] + [ts * np.array([1.2, 0, 0, 0]), ts * np.array([0, 0, 1.05, 0])]:
        d = ref.milp_feasible([th])
        out['milp_point'].append(dict(theta=th.tolist(), feasible=d is not None,
                                      delta=None if d is None else d.tolist()))
    # ---- V_R feasibility MILP on root simplices ---------------------------------------------
    roots = np.load(os.path.join(HERE, 'pendulum_roots.npy'))
    for r in (0, 22, 44, 50):
        d = ref.milp_feasible(list(roots[r]))
        out['milp_simplex'].append(dict(root=int(r), feasible=d is not None,
                                        delta=None if d is None else d.tolist()))
    with open(os.path.join(HERE, 'pendulum_golden.json'), 'w') as f:
        json.dump(out, f, indent=0)
    nf = sum(1 for g in out['fixed'] if g['feasible'])
    print('fixed problems: %d (%d feasible); milp points: %d; milp simplices: %d'
          % (len(out['fixed']), nf, len(out['milp_point']), len(out['milp_simplex'])))
    # quick self-check against the CPU oracle
    worst = 0.0
    for g in out['fixed']:
        ok = orc.P_theta_delta(np.array(g['theta']), g['cand'], check_feasibility=True)
        fat = g['feasible'] and g['margin'] > 1e-5
        thin = (not g['feasible']) or g['margin'] < 1e-9
        if (fat and not ok) or (thin and ok):
            print('FEASIBILITY MISMATCH', g)
        if g['feasible'] and g.get('J') is not None:
            J = orc.P_theta_delta(np.array(g['theta']), g['cand'])[1]
            worst = max(worst, abs(J - g['J']) / max(1, abs(J)))
    print('worst relative cost difference oracle vs reference-form SLSQP: %.2e' % worst)


if __name__ == '__main__':
    main()
