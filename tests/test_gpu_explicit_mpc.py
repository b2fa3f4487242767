"""Online evaluation kernel (ehm_locator_*) against the golden outputs of the
reference's ExplicitMPC class, the NumPy restatement, and size-independent properties on the
full-size partition."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_kernel_reproduces_reference_class():
    from explicit_hybrid_mpc_b200.mpc_library import ExplicitMPC
    z = np.load(os.path.join(GOLDEN, 'explicit_mpc_golden.npz'))
    flat = {k: z[k] for k in ('left', 'right', 'vertices', 'vinput')}
    flat['top'] = int(z['top'])
    empc = ExplicitMPC(flat)
    U, leaf = empc.evaluate_batch(z['points'])
    U, leaf = U.cpu().numpy(), leaf.cpu().numpy()
    assert np.array_equal(flat['vertices'][leaf], z['cell_vertices'])      # located cells: bit-exact
    assert np.abs(U - z['u']).max() <= 1e-9 * max(1.0, np.abs(z['u']).max())
    u, t = empc(z['points'][3])                                             # reference signature
    assert u.shape == (1,) and abs(u[0] - z['u'][3, 0]) <= 1e-9 * max(1.0, abs(z['u'][3, 0])) and t >= 0
    assert np.array_equal(flat['vertices'][empc.get_containing_cell(z['points'][5])], z['cell_vertices'][5])


def test_full_size_properties(pendulum, pendulum_roots, gpu_oracle):
    from explicit_hybrid_mpc_b200.explicit_mpc import flat_from_result
    from explicit_hybrid_mpc_b200.mpc_library import ExplicitMPC
    from explicit_hybrid_mpc_b200.partition import partition
    from oracle.explicit_mpc_ref import ExplicitMPCRef
    res = partition(gpu_oracle, pendulum_roots)
    empc = ExplicitMPC(res)
    flat = flat_from_result(res)
    rng = np.random.default_rng(5)
    ts = pendulum.theta_scale
    B = 200000
    X = (rng.random((B, 4)) * 2 - 1) * ts
    U, leaf = empc.evaluate_batch(X)
    U, leaf = U.cpu().numpy(), leaf.cpu().numpy()
    assert (flat['left'][leaf] < 0).all()                                   # every answer is a leaf
    # the returned leaf contains its query point and U is the barycentric interpolation
    V = flat['vertices'][leaf]                                              # (B,5,4)
    M = np.transpose(V[:, 1:] - V[:, :1], (0, 2, 1))
    al = np.linalg.solve(M, (X - V[:, 0])[..., None])[..., 0]
    al0 = 1 - al.sum(axis=1)
    assert al.min() >= -1e-9 and al0.min() >= -1e-9
    Ui = al0[:, None] * flat['vinput'][leaf][:, 0] + np.einsum('bi,bik->bk', al, flat['vinput'][leaf][:, 1:])
    assert np.abs(U - Ui).max() <= 1e-9 * 20.0
    # same leaves as the restatement of the reference's walk, on a sample
    ref = ExplicitMPCRef(flat)
    for q in range(0, B, B // 300):
        assert ref.locate(X[q]) == leaf[q]
