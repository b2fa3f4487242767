"""
Generate ``tests/golden/epigraph_golden.json``: known answers for the eps-suboptimality problems
of ``Oracle.bar_E_delta_R`` / ``bar_D_delta_R`` (``lib/oracle.py:88-102,285-414``) in REFERENCE
FORM: the full-variable big-M rows produced by the reference's own
``InvertedPendulumOnCart.make_constraints`` (29 continuous variables per problem, imported from
/root/reference through the stand-ins of make_golden.py), the simplex multipliers alpha with
``theta = sum_i alpha_i v_i`` (``:70-74``) and the two inequalities ``bar_V - V >= eps_a``,
``bar_V - V >= eps_r V`` (``:94-95``) written with the epigraph variable tau -- solved by SciPy
(HiGHS for feasibility, SLSQP after eliminating the equalities) for a FIXED commutation:

    t*_c = min V(z) + tau   s.t.  C(theta(alpha), z, delta_c),  alpha in the unit simplex,
                                  tau >= eps_a - bar_V(alpha),  tau >= -bar_V(alpha)/(1+eps_r)

``t*_c <= 0`` somewhere <=> the reference's MIQCP is feasible for that commutation.  The CPU
oracle's condensed 10-variable problems are pinned against these values by
tests/test_oracle_golden.py.  Needs /root/reference; run here only (about a minute):

    python tests/golden/make_epigraph_golden.py
"""
import json
import os
import sys

import numpy as np
import scipy.optimize as so

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden  # noqa: E402

EPS_A, EPS_R = 19.21220798973866, 1.0


def reference_form_tstar(ref, R, Vr, delta, with_tau=True):
    """(t*, theta*, V*, max violation) of the epigraph problem in reference-form variables, or None.
    with_tau=False: min over the simplex of V alone (lib/oracle.py:75-79, in_variability_ball)."""
    n = R.shape[1]
    ic = ref.icont
    Aub, b0, Aeq, e0 = ref.rows(np.zeros(n), delta)
    dBu, dBe = np.zeros((len(b0), n)), np.zeros((len(e0), n))
    for j in range(n):
        _, bj, _, ej = ref.rows(np.eye(n)[j], delta)
        dBu[:, j], dBe[:, j] = bj - b0, ej - e0
    const_eq = np.abs(Aeq[:, ic]).max(axis=1) == 0
    assert np.all(np.abs(e0[const_eq]) < 1e-9) and np.all(np.abs(dBe[const_eq]) < 1e-12)
    Aeq, e0, dBe = Aeq[~const_eq][:, ic], e0[~const_eq], dBe[~const_eq]
    Aub = Aub[:, ic]
    nz, na = len(ic), n + 1
    k2 = 1.0 / (1.0 + EPS_R)
    # variables y = [z (nz); alpha (n+1); tau]
    N = nz + na + 1
    G = np.zeros((len(b0) + na + 2, N)); h = np.zeros(len(b0) + na + 2)
    G[:len(b0), :nz] = Aub; G[:len(b0), nz:nz + na] = -dBu @ R.T; h[:len(b0)] = b0
    G[len(b0):len(b0) + na, nz:nz + na] = -np.eye(na)                      # alpha >= 0
    r1, r2 = len(b0) + na, len(b0) + na + 1
    if with_tau:
        G[r1, nz:nz + na] = -Vr; G[r1, -1] = -1.0; h[r1] = -EPS_A           # tau >= eps_a - bar_V
        G[r2, nz:nz + na] = -k2 * Vr; G[r2, -1] = -1.0; h[r2] = 0.0         # tau >= -bar_V/(1+eps_r)
    else:
        G[r1, -1] = -1.0; G[r2, -1] = 1.0                                   # tau = 0
    E = np.zeros((len(e0) + 1, N)); d = np.zeros(len(e0) + 1)
    E[:len(e0), :nz] = Aeq; E[:len(e0), nz:nz + na] = -dBe @ R.T; d[:len(e0)] = e0
    E[-1, nz:nz + na] = 1.0; d[-1] = 1.0                                    # sum alpha = 1
    lp = so.linprog(np.zeros(N), A_ub=G, b_ub=h, A_eq=E, b_eq=d, bounds=[(None, None)] * N, method='highs')
    if lp.status != 0:
        return None
    H = np.zeros((N, N)); f = np.zeros(N)
    H[:nz, :nz] = ref.H[np.ix_(ic, ic)]; f[:nz] = ref.f[ic]; f[-1] = 1.0
    y0 = np.linalg.lstsq(E, d, rcond=None)[0]
    _, sv, Vt = np.linalg.svd(E)
    Z = Vt[np.sum(sv > 1e-12):].T
    Hr, fr = Z.T @ H @ Z, Z.T @ (2 * H @ y0 + f)
    Gr, hr = G @ Z, h - G @ y0
    keep = np.abs(Gr).max(axis=1) > 1e-12
    Gr, hr = Gr[keep], hr[keep]
    # scale the objective to O(1) (costs reach 1e3) and try SLSQP from the LP point; fall back to
    # trust-constr, which is slower but does not give up on degenerate active sets
    sc = max(1.0, float(np.abs(Vr).max()))
    best = None
    start = np.linalg.lstsq(Z, lp.x - y0, rcond=None)[0]
    fun = lambda y: (y @ Hr @ y + fr @ y) / sc
    jac = lambda y: (2 * Hr @ y + fr) / sc
    r = so.minimize(fun, start, jac=jac,
                    constraints=[dict(type='ineq', fun=lambda y: hr - Gr @ y, jac=lambda y: -Gr)],
                    method='SLSQP', options=dict(ftol=1e-13, maxiter=2000))
    if r.success:
        best = r
    r2 = so.minimize(fun, best.x if best is not None else start, jac=jac, hess=lambda y: 2 * Hr / sc,
                     constraints=[so.LinearConstraint(Gr, -np.inf, hr)], method='trust-constr',
                     options=dict(gtol=1e-10, xtol=1e-12, maxiter=3000))
    if np.max(Gr @ r2.x - hr) <= 1e-8 and (best is None or r2.fun < best.fun - 1e-12):
        best = r2
    if best is None:
        return None
    y = y0 + Z @ best.x
    t = float(y @ H @ y + f @ y + ref.c)
    alpha, tau = y[nz:nz + na], y[-1]
    return t, (R.T @ alpha).tolist(), t - float(tau), float(np.max(G @ y - h))


def main():
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from explicit_hybrid_mpc_b200.tools import split_along_longest_edge
    from oracle.oracle_cpu import CpuOracle
    mine = InvertedPendulumOnCart()
    cpu = CpuOracle(mine.problem, EPS_A, EPS_R, mine.theta_scale)
    roots = np.load(os.path.join(HERE, 'pendulum_roots.npy'))
    for name in [k for k in sys.modules if k == 'oracle' or k.startswith('oracle.')]:
        sys.modules.pop(name)
    ref = make_golden.RefProblem(make_golden.import_reference_pendulum())
    pb = mine.problem
    rng = np.random.default_rng(7)
    out = dict(eps_a=EPS_A, eps_r=EPS_R, cases=[])
    for base, depth in ((22, 0), (22, 3), (0, 2), (0, 5), (44, 4), (7, 6)):
        R = roots[base]
        for d in range(depth):
            R = split_along_longest_edge(R)[d % 2]
        dref, vx = cpu.V_R(R)
        if dref is None:
            continue
        Vr = np.array([v[1] for v in vx])
        conv, ts = cpu.bar_E_delta_R(R, Vr, return_tstar=True)
        alive = np.nonzero(np.isfinite(ts))[0]
        cref = pb.index_from_commutation(dref)
        pick = [cref] + [int(c) for c in alive[np.argsort(ts[alive])[:3]]] + \
               [int(c) for c in rng.choice(alive, size=min(3, len(alive)), replace=False)]
        for c in sorted(set(pick)):
            g = reference_form_tstar(ref, R, Vr, pb.commutation_from_index(c))
            if g is None:
                continue
            out['cases'].append(dict(R=R.tolist(), Vr=Vr.tolist(), cand=int(c), is_ref=bool(c == cref),
                                     tstar=g[0], theta_star=g[1], Vstar=g[2], max_violation=g[3],
                                     oracle_tstar=float(ts[c])))
            print(base, depth, c, 'ref-form t* %.8f  oracle %.8f  viol %.1e' % (g[0], ts[c], g[3]))
    # min over the simplex of V*_delta (first problem of in_variability_ball, lib/oracle.py:220-283)
    out['minv'] = []
    for base, depth in ((22, 1), (0, 3), (44, 2), (7, 5), (30, 4)):
        R = roots[base]
        for d in range(depth):
            R = split_along_longest_edge(R)[(d + 1) % 2]
        dref, vx = cpu.V_R(R)
        if dref is None:
            continue
        c = pb.index_from_commutation(dref)
        g = reference_form_tstar(ref, R, np.zeros(5), dref, with_tau=False)
        if g is None:
            continue
        mine_v = cpu.min_over_simplex(R, c)
        out['minv'].append(dict(R=R.tolist(), cand=int(c), minV=g[0], theta_star=g[1], max_violation=g[3]))
        print('minV', base, depth, c, 'ref-form %.8f  oracle %.8f' % (g[0], mine_v))
    with open(os.path.join(HERE, 'epigraph_golden.json'), 'w') as f:
        json.dump(out, f)
    print('saved', len(out['cases']), 'cases,', len(out['minv']), 'minV cases')


if __name__ == '__main__':
    main()
