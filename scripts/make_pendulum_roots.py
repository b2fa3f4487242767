"""
Generate the root-simplex fixture ``tests/golden/pendulum_roots.npy``.

The reference builds its initial partition with ``scipy.spatial.Delaunay`` on the vertices
of the three velocity slabs of the pendulum's parameter box (``lib/examples.py:134-163``,
``lib/tools.py:152-189``).  Qhull's output depends on vertex order and library version, so
the roots are treated as *input* shared by the CUDA engine and the CPU oracle and are
committed here (66 simplices with SciPy 1.18 on ``itertools.product`` vertex order).
"""
import itertools
import os
import sys

import numpy as np
import scipy.spatial as scs

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from explicit_hybrid_mpc_b200.mpc_library import InvertedPendulumOnCart  # noqa: E402


def main():
    mpc = InvertedPendulumOnCart()
    p, a, v, r, ve = mpc.p_err, mpc.ang_err, mpc.v_err, mpc.rate_err, mpc.v_eps
    roots = []
    for lo, hi in ((ve, v), (-v, -ve), (-ve, ve)):      # set_1, set_2, set_3 (examples.py:138-149)
        V = np.array(list(itertools.product((-p, p), (-a, a), (lo, hi), (-r, r))))
        roots += [V[s] for s in scs.Delaunay(V).simplices]
    roots = np.array(roots)
    out = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'pendulum_roots.npy')
    np.save(out, roots)
    print(roots.shape, '->', os.path.abspath(out))


if __name__ == '__main__':
    main()
