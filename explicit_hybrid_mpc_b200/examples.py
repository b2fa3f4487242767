"""
Mirror of the reference's ``lib/examples.py`` for the examples this repository lowers: builds the
parameter set, the oracle with ``eps_a`` derived from ``abs_frac`` and the initial triangulation,
and returns the reference's triple ``(full_set, partition, oracle)`` -- so that the replacement
``main`` of INTEGRATION.md reads like the reference's.
"""
import numpy as np

from . import mpc_library, tools
from .polytope import Polytope


def create_oracle(mpc, set_vrep, abs_frac=None, abs_err=None, rel_err=None, oracle_cls=None):
    """``lib/examples.py:14-47``: eps_a = max over the vertices of ``abs_frac * Theta`` of the MICP
    optimal cost, unless ``abs_err`` is given -- ONE batched ``ehm_p_theta_batch`` call on the device
    instead of 2^n MICP solves."""
    if oracle_cls is None:
        from .oracle import Oracle as oracle_cls
    if abs_err is None:
        oracle = oracle_cls(mpc, 1.0, 1.0)
        if hasattr(oracle, 'P_theta_batch'):
            _, J, _ = oracle.P_theta_batch(abs_frac * np.asarray(set_vrep, dtype=float))
            abs_err = float(J.max().item())
        else:
            abs_err = max(oracle.P_theta(theta=abs_frac * np.asarray(v))[2] for v in set_vrep)
        oracle.set_eps(abs_err, rel_err)
        return oracle
    return oracle_cls(mpc, abs_err, rel_err)


def pendulum_example(abs_frac=0.5, abs_err=None, rel_err=2.0, oracle_cls=None):
    """``lib/examples.py:120-163``: Theta = [-p,p] x [-ang,ang] x [-v,v] x [-rate,rate] cut into the
    three velocity slabs v >= v_eps, v <= -v_eps, |v| <= v_eps (one commutation cannot straddle the
    friction switch), each Delaunay-triangulated, the three spines joined."""
    pend = mpc_library.InvertedPendulumOnCart()
    box = lambda lo, hi: Polytope(R=[(-pend.p_err, pend.p_err), (-pend.ang_err, pend.ang_err), (lo, hi),
                                     (-pend.rate_err, pend.rate_err)])
    full_set = box(-pend.v_err, pend.v_err)
    sets = [box(pend.v_eps, pend.v_err), box(-pend.v_err, -pend.v_eps), box(-pend.v_eps, pend.v_eps)]
    oracle = create_oracle(pend, np.vstack(full_set.V), abs_frac, abs_err, rel_err, oracle_cls)
    parts = [tools.delaunay(np.vstack(s.V))[0] for s in sets]
    tools.join_triangulation(parts[1], parts[2])
    tools.join_triangulation(parts[0], parts[1])
    return full_set, parts[0], oracle


def example(*args, **kwargs):
    return pendulum_example(*args, **kwargs)
