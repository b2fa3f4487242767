"""GPU parity of the batched oracle sub-problems (through the C-ABI) against the CPU oracle.

Tolerance: both sides converge the same convex programs to 1e-9 relative (gap + primal
residual); values must agree to 1e-7 * max(1,|v|) and feasibility decisions exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-7


def _close(a, b):
    return abs(a - b) <= RTOL * max(1.0, abs(a), abs(b))


def test_point_batch_vs_cpu(pendulum, cpu_oracle, gpu_oracle):
    rng = np.random.default_rng(1)
    ts = pendulum.theta_scale
    B = 600
    theta = rng.uniform(-1, 1, size=(B, 4)) * ts
    cand = rng.integers(0, gpu_oracle.ncand, size=B).astype(np.int32)
    # make a good share feasible: reuse the argmin commutation of P_theta at some points
    for b in range(0, B, 3):
        d = cpu_oracle.P_theta(theta[b])[1]
        if d is not None:
            cand[b] = pendulum.problem.index_from_commutation(d)
    J, u0, st = gpu_oracle.point_batch(theta, cand)
    J, u0, st = J.cpu().numpy(), u0.cpu().numpy(), st.cpu().numpy()
    nfeas = 0
    for b in range(B):
        ok = cpu_oracle.P_theta_delta(theta[b], int(cand[b]), check_feasibility=True)
        assert (st[b] == 1) == ok, (b, st[b], ok)
        if ok:
            nfeas += 1
            u_c, J_c, _ = cpu_oracle.P_theta_delta(theta[b], int(cand[b]))
            assert _close(J[b], J_c), (b, J[b], J_c)
            assert abs(u0[b, 0] * 20.0 - u_c[0]) <= 1e-5 * max(1.0, abs(u_c[0]))
    assert nfeas > 100


def test_p_theta_matches_cpu(pendulum, cpu_oracle, gpu_oracle):
    ts = pendulum.theta_scale
    for th in (np.array([0.3, 0.02, 0.5, 0.1]), 0.25 * ts * np.array([1, -1, 1, -1]),
               np.array([-0.5, 0.05, -0.4, 0.2])):
        u_g, d_g, J_g, _ = gpu_oracle.P_theta(th)
        u_c, d_c, J_c, _ = cpu_oracle.P_theta(th)
        assert np.array_equal(d_g, d_c)
        assert _close(J_g, J_c)
        assert abs(u_g[0] - u_c[0]) <= 1e-5 * max(1.0, abs(u_c[0]))


def test_tstar_all_candidates(pendulum, pendulum_roots, cpu_oracle, gpu_oracle):
    """bar_E's per-candidate epigraph optimum t*_c on two root simplices."""
    for ridx in (0, 50):
        R = pendulum_roots[ridx]
        d, vx = cpu_oracle.V_R(R)
        Vr = np.array([v[1] for v in vx])
        sub_c, t_c = cpu_oracle.bar_E_delta_R(R, Vr, return_tstar=True)
        t_g, _ = gpu_oracle._tstar_all(R, Vr)
        t_g = t_g.cpu().numpy()
        assert np.array_equal(np.isnan(t_c), np.isnan(t_g))
        live = ~np.isnan(t_c)
        assert live.sum() > 50
        err = np.abs(t_c[live] - t_g[live]) / np.maximum(1.0, np.abs(t_c[live]))
        assert err.max() <= RTOL, err.max()
        assert gpu_oracle.bar_E_delta_R(R, Vr) == sub_c


def test_v_r_and_bar_d(pendulum, pendulum_roots, cpu_oracle, gpu_oracle):
    for ridx in (0, 30, 50):
        R = pendulum_roots[ridx]
        d_c, vx_c = cpu_oracle.V_R(R)
        d_g, vx_g = gpu_oracle.V_R(R)
        assert np.array_equal(d_c, d_g)
        for (u1, J1, _), (u2, J2, _) in zip(vx_c, vx_g):
            assert _close(J1, J2)
        Vr = np.array([v[1] for v in vx_c])
        r_c = cpu_oracle.bar_D_delta_R(R, Vr, d_c)
        r_g = gpu_oracle.bar_D_delta_R(R, Vr, d_c)
        assert (r_c[0] is None) == (r_g[0] is None)
        if r_c[0] is not None:
            assert np.array_equal(r_c[0], r_g[0])
            assert np.allclose(r_c[1], r_g[1], rtol=0, atol=1e-5 * np.abs(R).max())
            assert r_c[3] == r_g[3]
            for (u1, J1, _), (u2, J2, _) in zip(r_c[2], r_g[2]):
                assert _close(J1, J2)


def test_min_over_simplex(pendulum, pendulum_roots, cpu_oracle, gpu_oracle):
    from explicit_hybrid_mpc_b200._capi import SX_MINV
    R = pendulum_roots[3]
    d, _ = cpu_oracle.V_R(R)
    c = pendulum.problem.index_from_commutation(d)
    v_c = cpu_oracle.min_over_simplex(R, c)
    v_g, _, st = gpu_oracle.simplex_batch(R[None], None, np.array([c], np.int32), SX_MINV, False)
    assert int(st[0]) == 1 and _close(float(v_g[0]), v_c)


@pytest.mark.parametrize('n', [2, 4, 6, 8])     # cwh_z, pendulum / cwh_xy, cwh_xyz, synthetic
def test_split_bit_exact(n, gpu_oracle, cpu_oracle):
    import torch
    from explicit_hybrid_mpc_b200 import _capi
    rng = np.random.default_rng(7)
    B = 512
    R = rng.standard_normal((B, n + 1, n))
    # force ties: regular-ish simplices with repeated edge lengths
    R[:64] = np.round(R[:64] * 2) / 2
    Rt = torch.as_tensor(R).cuda()
    S1t, S2t = torch.empty_like(Rt), torch.empty_like(Rt)
    ijt = torch.empty(B, 2, dtype=torch.int32, device='cuda')
    _capi.check(_capi.lib().ehm_split_batch(n, B, Rt.data_ptr(), S1t.data_ptr(), S2t.data_ptr(),
                                            ijt.data_ptr(), None))
    S1, S2, ij = S1t.cpu().numpy(), S2t.cpu().numpy(), ijt.cpu().numpy()
    for b in range(B):
        a1, a2, aij = cpu_oracle.split_along_longest_edge(R[b])
        assert tuple(ij[b]) == tuple(aij)
        assert np.array_equal(S1[b], a1) and np.array_equal(S2[b], a2)   # bit exact
    # empty batch: a no-op, not an error
    _capi.check(_capi.lib().ehm_split_batch(n, 0, Rt.data_ptr(), S1t.data_ptr(), S2t.data_ptr(),
                                            ijt.data_ptr(), None))


def test_tstar_against_reference_form_golden(gpu_oracle):
    """The device epigraph problems against the reference-form known answers of
    tests/golden/epigraph_golden.json (see tests/test_oracle_golden.py)."""
    import json
    import os
    from conftest import GOLDEN
    E = json.load(open(os.path.join(GOLDEN, 'epigraph_golden.json')))
    seen = {}
    for g in E['cases']:
        R, Vr = np.array(g['R']), np.array(g['Vr'])
        key = (R.tobytes(), Vr.tobytes())
        if key not in seen:
            seen[key] = gpu_oracle._tstar_all(R, Vr)[0].cpu().numpy()
        t = seen[key][g['cand']]
        assert abs(t - g['tstar']) <= 2e-6 * max(1.0, abs(g['tstar'])), (g['cand'], t, g['tstar'])


def test_edge_cases_of_the_c_abi(gpu_oracle):
    """Empty batches are no-ops, bad arguments are error codes (never exceptions inside the library),
    unsupported problem shapes are refused with a message."""
    import torch
    from explicit_hybrid_mpc_b200 import _capi
    from explicit_hybrid_mpc_b200.mpc_library import HybridMPCProblem, SyntheticMPLP
    from explicit_hybrid_mpc_b200.oracle import Oracle
    L, h = _capi.lib(), gpu_oracle._p.handle
    z = torch.zeros(8, dtype=torch.float64, device='cuda')
    zi = torch.zeros(8, dtype=torch.int32, device='cuda')
    assert L.ehm_p_theta_batch(h, 0, z.data_ptr(), zi.data_ptr(), None, None, None, None) == 0
    assert L.ehm_v_r_batch(h, 0, z.data_ptr(), zi.data_ptr(), None, None, None, None) == 0
    assert L.ehm_bar_e_batch(h, 0, z.data_ptr(), z.data_ptr(), zi.data_ptr(), None, None) == 0
    assert L.ehm_qp_theta_delta_batch(h, 0, z.data_ptr(), zi.data_ptr(), z.data_ptr(), None, zi.data_ptr(), None) == 0
    assert L.ehm_split_batch(4, 0, z.data_ptr(), z.data_ptr(), z.data_ptr(), zi.data_ptr(), None) == 0
    assert L.ehm_p_theta_batch(h, -1, z.data_ptr(), zi.data_ptr(), None, None, None, None) == _capi.EHM_EINVAL
    assert L.ehm_bar_d_batch(h, 1, z.data_ptr(), z.data_ptr(), None, zi.data_ptr(), z.data_ptr(), z.data_ptr(),
                             z.data_ptr(), zi.data_ptr(), None, None) == _capi.EHM_EINVAL
    assert L.ehm_split_batch(17, 1, z.data_ptr(), z.data_ptr(), z.data_ptr(), zi.data_ptr(), None) == _capi.EHM_EINVAL
    # a shape without a compiled kernel
    odd = SyntheticMPLP(n=3, nz=4, rows=6, K=2, seed=1)
    with pytest.raises(_capi.EhmError, match='no compiled kernel'):
        Oracle(odd, 0.1, 0.1)
    # the synthetic n=8 shape is compiled for LP costs only
    m = SyntheticMPLP()
    pb = m.problem
    qp = HybridMPCProblem(n_x=pb.n_x, n_u=pb.n_u, N=1, K=pb.K, A=pb.A, B=pb.B, w=pb.w, C=pb.C, D=pb.D, e=pb.e,
                          Q=pb.Q, P=pb.P, R=np.eye(pb.n_u), ru=pb.ru, u_scale=pb.u_scale).validate()

    class _M:
        problem, theta_scale = qp, m.theta_scale
    with pytest.raises(_capi.EhmError, match='LP costs only'):
        Oracle(_M(), 0.1, 0.1)
