import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def pendulum():
    from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart
    return InvertedPendulumOnCart()


@pytest.fixture(scope='session')
def pendulum_roots():
    return np.load(os.path.join(GOLDEN, 'pendulum_roots.npy'))


@pytest.fixture(scope='session')
def cpu_oracle(pendulum):
    """The CPU oracle (test infrastructure) at the eps of BASELINE config 2."""
    from oracle.oracle_cpu import CpuOracle
    return CpuOracle(pendulum.problem, 19.21220798973866, 1.0, pendulum.theta_scale)


@pytest.fixture(scope='session')
def gpu_oracle(pendulum):
    from explicit_hybrid_mpc_b200.oracle import Oracle
    return Oracle(pendulum, 19.21220798973866, 1.0)
