"""SURVEY 8f rows 3-4 on the CPU: the Polytope / examples front end reproduces the committed root
fixture; status / statistics records have the reference's fields."""
import os
import pickle

import numpy as np

from conftest import GOLDEN


def test_polytope_representations():
    from explicit_hybrid_mpc_b200.polytope import Polytope, PolytopeError
    P = Polytope(R=[(-1, 1), (-2, 2), (0, 3)])
    assert P.n == 3 and len(P.V) == 8 and abs(P.volume() - 24.0) < 1e-12
    Q = Polytope(A=P.A, b=P.b)                      # H-rep -> V-rep (lib/polytope.py:131-140)
    assert sorted(map(tuple, np.round(Q.V, 9))) == sorted(map(tuple, P.V))
    H = Polytope(V=P.V)                             # V-rep -> H-rep
    assert H.A.shape == (6, 3) and all(H.contains(v) for v in P.V) and not H.contains([2, 0, 1])
    tri = Polytope(V=[[0, 0], [1, 0], [0, 1]])
    assert tri.A.shape == (3, 2) and tri.contains([0.2, 0.2]) and not tri.contains([0.6, 0.6])
    try:
        Polytope(R=[(1, 0)])
        assert False
    except PolytopeError:
        pass


def test_pendulum_example_reproduces_root_fixture(pendulum_roots):
    """examples.pendulum_example (lib/examples.py:120-163): Polytope slabs -> Delaunay -> joined spine;
    its open leaves are the 66 root simplices of the bench, in order, bit for bit."""
    from explicit_hybrid_mpc_b200 import examples
    from explicit_hybrid_mpc_b200.partition import roots_of

    class FakeOracle:        # examples.create_oracle without a GPU: eps_a = max over abs_frac * vertices
        def __init__(self, mpc, a, r):
            self.eps = (a, r)

        def P_theta(self, theta):
            return None, None, float(np.abs(theta).sum()), 0.0

        def set_eps(self, a, r):
            self.eps = (a, r)
    full, tree, orc = examples.pendulum_example(abs_frac=0.25, rel_err=1.0, oracle_cls=FakeOracle)
    assert np.array_equal(roots_of(tree), pendulum_roots)
    assert len(full.V) == 16 and orc.eps[1] == 1.0
    want = max(np.abs(0.25 * np.asarray(v)).sum() for v in full.V)
    assert abs(orc.eps[0] - want) < 1e-15


def test_statistics_record_fields(tmp_path):
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    from explicit_hybrid_mpc_b200.partition import PartitionResult
    z = np.load(os.path.join(GOLDEN, 'worker_tree_golden.npz'))
    n = len(z['paths'])
    arr = dict(left=np.where(z['leaf'], -1, 1).astype(np.int32), right=np.where(z['leaf'], -1, 1).astype(np.int32),
               root_id=np.zeros(n, np.int32), depth=np.array([len(p.split(':')[1]) for p in z['paths']], np.int32),
               cand=z['cand'].astype(np.int32), closed=z['closed'].astype(np.uint8), vertices=z['vertices'],
               vcost=z['vcost'], vinput=z['vinput'])
    res = PartitionResult(arr, dict(nleaves=int(z['leaf'].sum()), nsolves=123, nfail=0, device_ms=10.0, open_left=0,
                                    stage_ms=[1.0] * 8), InvertedPendulumOnCart().problem)
    path = str(tmp_path / 'statistics.pkl')
    res.append_statistics(path)
    rec = res.append_statistics(path)
    with open(path, 'rb') as f:
        a, b = pickle.load(f), pickle.load(f)
    assert a == b == rec
    for key in ('num_proc_active', 'num_proc_failed', 'num_tasks_in_queue', 'volume_filled_total',
                'volume_filled_frac', 'simplex_count_total', 'time_elapsed', 'algorithms_running',
                'time_active_total', 'time_idle_total', 'eta', 'scheduler_looprate'):       # lib/scheduler.py:293-304
        assert key in rec['overall']
    assert rec['overall']['simplex_count_total'] == int((z['leaf'] & z['closed']).sum())
    assert rec['process'][0]['simplex_count_total'] == int(z['leaf'].sum())


def test_unit_box_convention_is_enforced():
    """HybridMPCProblem.validate: every mode must carry +-uhat_j <= 1 (the row pre-solve relies on it)."""
    import pytest
    from explicit_hybrid_mpc_b200.mpc_library import HybridMPCProblem, SyntheticMPLP
    m = SyntheticMPLP(n=3, nz=4, rows=6, K=2, seed=1).problem       # valid by construction
    assert m.p == 6 + 8 and np.allclose(m.u_scale, 10.0)
    bad = HybridMPCProblem(n_x=m.n_x, n_u=m.n_u, N=1, K=m.K, A=m.A, B=m.B, w=m.w, C=m.C, D=m.D.copy(), e=m.e.copy(),
                           Q=m.Q, P=m.P, R=m.R, ru=m.ru, u_scale=m.u_scale)
    bad.e[1, -1] = 3.0                                              # mode 1: -uhat_3 <= 3 instead of <= 1
    with pytest.raises(ValueError, match='unit box'):
        bad.validate()
