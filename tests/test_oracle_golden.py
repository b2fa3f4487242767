"""Pin the CPU oracle (and the pendulum problem lowering) against tests/golden/
pendulum_golden.json: solutions of the REFERENCE-FORM problems -- built by the reference's own
lib/mpc_library.py code, solved by SciPy HiGHS / SLSQP (see tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

G = json.load(open(os.path.join(GOLDEN, 'pendulum_golden.json')))


def test_fixed_delta_problems(pendulum, cpu_oracle):
    nfat = nthin = 0
    for g in G['fixed']:
        th, c = np.array(g['theta']), g['cand']
        ok = cpu_oracle.P_theta_delta(th, c, check_feasibility=True)
        if not g['feasible'] or g['margin'] < 1e-9:
            # infeasible, or feasible only on a measure-zero slice (excluded by the canonical
            # interior-margin rule TAU_INT)
            assert not ok, g
            nthin += 1
        elif g['margin'] > 1e-5:
            assert ok, g
            nfat += 1
            if g.get('J') is not None:
                u0, J, _ = cpu_oracle.P_theta_delta(th, c)
                assert abs(J - g['J']) <= 2e-6 * max(1.0, abs(g['J'])), (g, J)
                assert abs(u0[0] - g['u0']) <= 2e-3 * max(1.0, abs(g['u0'])), (g, u0)
    assert nfat >= 30 and nthin >= 20


def test_micp_feasibility_points(pendulum, cpu_oracle):
    """P_theta(check_feasibility=True): HiGHS branch & bound over binary delta on the
    reference-form MILP vs enumeration in the oracle."""
    for g in G['milp_point']:
        th = np.array(g['theta'])
        ok = cpu_oracle.P_theta(th, check_feasibility=True)
        if not g['feasible']:
            assert not ok, g
        else:
            # the commutation HiGHS found must be accepted by the oracle unless it is thin
            c = pendulum.problem.index_from_commutation(np.array(g['delta']))
            assert c >= 0
            if cpu_oracle.P_theta_delta(th, c, check_feasibility=True):
                assert ok


def test_v_r_feasibility_simplices(pendulum, pendulum_roots, cpu_oracle):
    for g in G['milp_simplex']:
        R = pendulum_roots[g['root']]
        d, vx = cpu_oracle.V_R(R)
        assert (d is not None) == g['feasible']
        # the oracle's canonical choice is feasible at every vertex by construction; it must
        # be one-hot like every solution of the reference MILP (sum z[k] == 1)
        assert np.allclose(np.asarray(d).reshape(5, 5).sum(axis=1), 1.0)
        assert np.allclose(np.array(g['delta']).reshape(5, 5).sum(axis=1), 1.0)


def test_problem_lowering_matches_appendix_a(pendulum):
    """Spot values of the re-derived linearisation / discretisation."""
    A, B, w = pendulum.A_d, pendulum.B_d, pendulum.w_d
    assert abs(A[0][1, 1] - 1.0213246230) < 1e-9 and abs(A[0][2, 1] + 0.29604221141) < 1e-9
    assert np.allclose(B[0], [0.0050040885, -0.0016725776, 0.1001636561, -0.0335699384], atol=1e-9)
    assert np.allclose(w[2], [-0.0102104742, 0.0034127716, -0.2043711966, 0.068495202], atol=1e-9)
    assert abs(pendulum.P[1, 1] - 2699.666188493) < 1e-5
    assert np.array_equal(B[4], np.zeros(4)) and np.array_equal(w[4], np.zeros(4))


def test_split_matches_reference_rule(cpu_oracle):
    """tools.split_along_longest_edge (lib/tools.py:245-250): first maximum, in
    itertools.combinations order, of la.norm(R[i]-R[j]) -- evaluated here with NumPy exactly as
    the reference writes it, against the oracle's restatement (sqrt of an FMA chain)."""
    import itertools
    import numpy.linalg as la
    from explicit_hybrid_mpc_b200 import tools
    rng = np.random.default_rng(3)
    for trial in range(600):
        n = (2, 4, 6, 8)[trial % 4]
        if trial % 3 == 0:
            R = np.round(rng.standard_normal((n + 1, n)) * 4) / 4     # many exact ties
        elif trial % 3 == 1:
            # repeated bisection of a regular-ish simplex: edges equal up to rounding noise
            R = rng.standard_normal((n + 1, n)) * np.array([1.0, 0.087, 1.0, 0.35, 1, 1, 1, 1][:n])
            for _ in range(trial % 7):
                R = cpu_oracle.split_along_longest_edge(R)[trial % 2]
        else:
            R = rng.standard_normal((n + 1, n))
        S1, S2, ij = cpu_oracle.split_along_longest_edge(R)
        combos = list(itertools.combinations(range(n + 1), 2))
        k = int(np.argmax([la.norm(R[a] - R[b]) for a, b in combos]))     # the reference's line
        assert ij == combos[k]
        mid = (R[ij[0]] + R[ij[1]]) / 2.
        assert np.array_equal(S1[ij[0]], mid) and np.array_equal(S2[ij[1]], mid)
        m1, m2, mij = tools.split_along_longest_edge(R)                   # host mirror
        assert mij == ij and np.array_equal(m1, S1) and np.array_equal(m2, S2)


WORKER_FIXTURES = ['worker_tree_golden.npz', 'worker_tree_golden_root0.npz', 'worker_tree_golden_deep.npz',
                   'worker_tree_golden_config3.npz']


def worker_fixture(name, pendulum_roots):
    """(golden npz, root simplices, eps_a, eps_r) of a make_worker_golden.py fixture"""
    z = np.load(os.path.join(GOLDEN, name))
    roots = z['roots'] if 'roots' in z.files else pendulum_roots[z['root_ids']]
    return z, roots, float(z['eps'][0]), float(z['eps'][1])


@pytest.mark.parametrize('fixture', WORKER_FIXTURES)
def test_worker_recursion_golden(fixture, pendulum, pendulum_roots):
    """The fused ECC/L-CSS loop of the oracle reproduces, node for node, the tree that the
    REFERENCE'S OWN Worker.ecc / Worker.lcss / tools.split_along_longest_edge code
    (lib/worker.py:241-417, imported from /root/reference) built over the six-method oracle API
    (tests/golden/make_worker_golden.py).  The root-0 fixture (2035 nodes) contains a swap-in-place
    visit (lib/worker.py:396-401); the deep fixture uses tighter tolerances (eps_a 2.3055,
    eps_r 0.05) on a depth-5 descendant of root 22; the config3 fixture is BASELINE config 3 proper
    (abs_frac 0.03 -> eps_a 0.27937 by examples.create_oracle, eps_r 0.05) on a depth-9 descendant
    of root 0."""
    from oracle.oracle_cpu import CpuOracle
    z, roots, eps_a, eps_r = worker_fixture(fixture, pendulum_roots)
    cpu = CpuOracle(pendulum.problem, eps_a, eps_r, pendulum.theta_scale)
    r = cpu.partition(roots, nthreads=4)
    assert r['error'] == 0 and r['nfail'] == 0
    check_against_worker_golden(r, z, len(roots))


def check_against_worker_golden(r, z, nroots, cost_rtol=0.0, input_atol=0.0):
    """r: flat node arrays of a partition of the golden's root simplices; z: the golden npz."""
    from test_gpu_tree_parity import _by_path
    pc = _by_path(r, nroots)
    gp = {str(p): i for i, p in enumerate(z['paths'])}
    assert set(pc) == set(gp), 'topology differs: %d vs %d nodes' % (len(pc), len(gp))
    for p, i in gp.items():
        j = pc[p]
        assert np.array_equal(z['vertices'][i], r['vertices'][j]), p          # bit exact
        assert bool(z['leaf'][i]) == (r['left'][j] < 0), p
        assert int(z['cand'][i]) == int(r['cand'][j]), p
        if z['leaf'][i]:
            assert bool(z['closed'][i]) and int(r['closed'][j]) == 1, p
        if z['cand'][i] >= 0:
            tol = cost_rtol * np.maximum(1.0, np.abs(z['vcost'][i]))
            assert (np.abs(z['vcost'][i] - r['vcost'][j]) <= tol).all(), p
            assert np.abs(z['vinput'][i] - r['vinput'][j]).max() <= input_atol, p


def test_row_presolve_is_exact(pendulum, pendulum_roots, cpu_oracle):
    """Dropping rows that cannot bind inside the input box |uhat| <= 1 (DESIGN.md "Row
    pre-solve") must not change any decision: identical tree with and without it."""
    from oracle.oracle_cpu import lib
    from test_gpu_tree_parity import compare_trees
    roots = pendulum_roots[[22]]
    try:
        lib().ehmo_set_presolve(0)
        full = cpu_oracle.partition(roots, nthreads=4)
    finally:
        lib().ehmo_set_presolve(1)
    pre = cpu_oracle.partition(roots, nthreads=4)
    assert full['nfail'] == 0 and pre['nfail'] == 0
    assert compare_trees(pre, full, 1) <= 1e-6


def test_phase1_early_exit_is_exact(pendulum_roots, cpu_oracle):
    """Stopping a phase-1 LP at the first iterate that certifies an interior point (DESIGN.md
    "Phase-1 early exit") must not change any decision: identical tree, fewer iterations."""
    import ctypes as C
    from oracle.oracle_cpu import lib
    from test_gpu_tree_parity import compare_trees
    L = lib()
    L.ehmo_total_iters.restype = C.c_long
    roots = pendulum_roots[[22]]
    try:
        L.ehmo_set_early_exit(0)
        i0 = L.ehmo_total_iters()
        full = cpu_oracle.partition(roots, nthreads=4)
        i1 = L.ehmo_total_iters()
    finally:
        L.ehmo_set_early_exit(1)
    early = cpu_oracle.partition(roots, nthreads=4)
    i2 = L.ehmo_total_iters()
    assert full['nfail'] == 0 and early['nfail'] == 0 and full['nsolves'] == early['nsolves']
    assert compare_trees(early, full, 1) <= 1e-9
    assert (i2 - i1) < 0.85 * (i1 - i0)


def test_tangent_plane_bounds_are_result_invariant(pendulum_roots, cpu_oracle):
    """Tangent-plane lower bounds of t*_c (DESIGN.md "Tangent-plane bounds") only skip work:
    identical tree without them (0), with planes inherited down the lineage (1) and with the
    per-visit planes the CUDA engine uses (2, the default); fewer sub-problems with them."""
    from oracle.oracle_cpu import lib
    from test_gpu_tree_parity import compare_trees
    roots = pendulum_roots[[22, 8]]
    out = {}
    try:
        for mode in (0, 1, 2):
            lib().ehmo_set_planes(mode)
            out[mode] = cpu_oracle.partition(roots, nthreads=4)
            assert out[mode]['nfail'] == 0
    finally:
        lib().ehmo_set_planes(2)
    assert compare_trees(out[1], out[0], 2) <= 1e-9 and compare_trees(out[2], out[0], 2) <= 1e-9
    assert out[2]['nsolves'] < 0.95 * out[0]['nsolves'] and out[1]['nsolves'] <= out[2]['nsolves']


def test_geometry_helpers_match_reference_outputs(pendulum_roots, cpu_oracle):
    """tools.delaunay / join_triangulation / simplex_volume / split_along_longest_edge: the host
    mirror, the CPU oracle and the roots fixture against outputs of the reference's own
    lib/tools.py functions (tests/golden/make_tools_golden.py)."""
    from explicit_hybrid_mpc_b200 import tools
    from explicit_hybrid_mpc_b200.partition import roots_of
    z = np.load(os.path.join(GOLDEN, 'tools_golden.npz'))
    assert np.array_equal(z['roots'], pendulum_roots)
    trees = [tools.delaunay(s)[0] for s in z['slabs']]                  # examples.py:155-160
    tools.join_triangulation(trees[1], trees[2])
    tools.join_triangulation(trees[0], trees[1])
    assert np.array_equal(np.array(roots_of(trees[0])), z['roots'])
    for R, vol, S1, S2, ij in zip(z['simplices'], z['volumes'], z['S1'], z['S2'], z['ij']):
        assert tools.simplex_volume(R) == vol                               # bit exact
        assert abs(cpu_oracle.simplex_volume(R) - vol) <= 1e-12 * abs(vol)
        for impl in (tools.split_along_longest_edge, cpu_oracle.split_along_longest_edge):
            a, b, c = impl(R)
            assert tuple(c) == tuple(ij) and np.array_equal(a, S1) and np.array_equal(b, S2)


def test_eps_a_as_create_oracle_derives_it(pendulum):
    """examples.create_oracle (lib/examples.py:18-47): eps_a = max over the box vertices v of the
    MICP optimal cost at abs_frac * v.  Fixture values produced by the reference's own function
    (tests/golden/make_tools_golden.py); the bench and the tests use the 0.25 value."""
    import itertools
    from oracle.oracle_cpu import CpuOracle
    z = np.load(os.path.join(GOLDEN, 'tools_golden.npz'))
    assert abs(z['eps_a'][0] - 19.21220798973866) <= 1e-9 * 19.2
    cpu = CpuOracle(pendulum.problem, 1.0, 1.0, pendulum.theta_scale)
    box = np.array(list(itertools.product(*[(-s, s) for s in pendulum.theta_scale])))
    for frac, want in zip((0.25, 0.03), z['eps_a']):
        got = max(cpu.P_theta(frac * v)[2] for v in box)
        assert abs(got - want) <= 1e-9 * max(1.0, want)


def test_synthetic_mplp_partition_cpu():
    """BASELINE config 5's problem family (synthetic n=8 mp-LP, 4 modes) through the CPU oracle:
    the data generator is well posed (no infeasible region in [-1,1]^8) and the partition of a few
    Kuhn simplices closes with leaves that tile them.  (The CUDA engine's run of the same problem is
    tests/test_gpu_tree_parity.py::test_synthetic_mplp_tree_equals_cpu_oracle.)"""
    from explicit_hybrid_mpc_b200.mpc_library import SyntheticMPLP
    from explicit_hybrid_mpc_b200.tools import kuhn_triangulation, simplex_volume
    from oracle.oracle_cpu import CpuOracle
    m = SyntheticMPLP()
    roots = kuhn_triangulation(-np.ones(8), np.ones(8), limit=3)
    cpu = CpuOracle(m.problem, 0.03, 0.01, m.theta_scale)
    r = cpu.partition(roots, nthreads=4)
    assert r['error'] == 0 and r['nfail'] == 0 and r['nleaves'] > 3
    leaves = np.nonzero(r['left'] < 0)[0]
    assert (r['closed'][leaves] == 1).all()
    vol = sum(simplex_volume(r['vertices'][i]) for i in leaves)
    want = sum(simplex_volume(R) for R in roots)
    assert abs(vol - want) <= 1e-9 * want


def test_epigraph_problems_reference_form(cpu_oracle):
    """bar_E / bar_D's per-commutation epigraph optimum t*_c (lib/oracle.py:88-102,285-414):
    the oracle's condensed 10-variable problem against the REFERENCE-FORM problem -- the
    reference's own big-M rows over 29 continuous variables + simplex multipliers + the two
    eps-suboptimality inequalities, solved by SciPy (tests/golden/make_epigraph_golden.py)."""
    E = json.load(open(os.path.join(GOLDEN, 'epigraph_golden.json')))
    assert len(E['cases']) >= 30
    seen = {}
    for g in E['cases']:
        R, Vr = np.array(g['R']), np.array(g['Vr'])
        key = (R.tobytes(), Vr.tobytes())
        if key not in seen:
            seen[key] = cpu_oracle.bar_E_delta_R(R, Vr, return_tstar=True)
        conv, ts = seen[key]
        t = ts[g['cand']]
        assert np.isfinite(t), g['cand']
        assert abs(t - g['tstar']) <= 2e-6 * max(1.0, abs(g['tstar'])), (g['cand'], t, g['tstar'])
    # the convergence flag is what the t* say: converged <=> no candidate at or below zero
    for conv, ts in seen.values():
        assert conv == (not np.any(ts[np.isfinite(ts)] <= 0.0))
    # min over the simplex of V*_delta (first problem of in_variability_ball, :220-283)
    assert len(E['minv']) >= 5
    for g in E['minv']:
        v = cpu_oracle.min_over_simplex(np.array(g['R']), g['cand'])
        assert v is not None and abs(v - g['minV']) <= 2e-6 * max(1.0, abs(g['minV']))


def test_worker_recursion_golden_ecc_branch():
    """The ECC split branch (lib/worker.py:264-277: V_R finds no common commutation -> split and
    recurse with 'ecc'), which the pendulum never takes: the reference's own Worker code on the
    two-mode toy problem of tests/golden/toy_problem.py against the fused CPU loop."""
    sys_path_golden = os.path.join(GOLDEN)
    import sys
    if sys_path_golden not in sys.path:
        sys.path.insert(0, sys_path_golden)
    from toy_problem import TOY_EPS, TOY_ROOT, toy_problem
    from oracle.oracle_cpu import CpuOracle
    z = np.load(os.path.join(GOLDEN, 'worker_tree_golden_toy.npz'))
    assert int((z['cand'] < 0).sum()) >= 5                       # data-less ECC nodes exist
    cpu = CpuOracle(toy_problem(), TOY_EPS[0], TOY_EPS[1], np.ones(2))
    r = cpu.partition(TOY_ROOT[None], nthreads=2)
    assert r['error'] == 0 and r['nfail'] == 0 and r['nleaves'] == int(z['leaf'].sum())
    check_against_worker_golden(r, z, 1)
