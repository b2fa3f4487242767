"""
Re-derive the two artefacts the reference loads from ``lib/sage/*.pkl`` and which are
absent from its checkout (git-ignored, produced offline by SageMath):

* ``pendulum_parameters.pkl``     -- reference ``lib/pendulum_linearize.py:29-38``
* ``pendulum_linearization.pkl``  -- reference ``lib/pendulum_linearize.py:112-143``

SymPy replaces Sage.  Output: ``explicit_hybrid_mpc_b200/data/pendulum_linearization.json``
(committed; the package loads it the way the reference loads its pickles,
``lib/mpc_library.py:416-430``).  Run once:  ``python scripts/derive_pendulum_linearization.py``
"""
import json
import os
import sympy as sp

x, theta, dxdt, dthetadt, F = sp.symbols('x theta dxdt dthetadt F')
STATE = [x, theta, dxdt, dthetadt]

pars = dict(mu_s=0.16, mu_d=0.03, g=9.81, l=3., m=0.3, M=1., v_eps=1e-3, a_eps=1e-3,
            T_s=1 / 10., T_s_plant=1 / 100.)


def linearize(d2x, d2th):
    """Jacobian linearisation about state=0, F=0 (ref make_linearized_system :44-107)."""
    at0 = {x: 0, theta: 0, dxdt: 0, dthetadt: 0, F: 0}
    rows = [dxdt, dthetadt, d2x, d2th]
    A = [[float(sp.diff(r, s).subs(at0)) for s in STATE] for r in rows]
    B = [float(sp.diff(r, F).subs(at0)) for r in rows]
    # w = f(0) - A*0 - B*0 = f(0)
    w = [float(sp.sympify(r).subs(at0)) for r in rows]
    return A, B, w


def main():
    mu_s, mu_d, g, l, m, M = (sp.Float(pars[k]) if False else pars[k]
                              for k in ('mu_s', 'mu_d', 'g', 'l', 'm', 'M'))
    sin, cos = sp.sin, sp.cos
    d2x = dict(
        opt1=(F - m * g / 2 * sin(2 * theta) + m * l * dthetadt**2 * (sin(theta) + mu_d * cos(theta)) -
              mu_d * (M + m * cos(theta)**2) * g) / (M + m * (sin(theta)**2 + mu_d * sin(2 * theta) / 2)),
        opt2=(F - m * g / 2 * sin(2 * theta) + m * l * dthetadt**2 * (sin(theta) - mu_d * cos(theta)) +
              mu_d * (M + m * cos(theta)**2) * g) / (M + m * (sin(theta)**2 - mu_d * sin(2 * theta) / 2)),
        opt3=(F - m * g / 2 * sin(2 * theta) + m * l * dthetadt**2 * (sin(theta) + mu_s * cos(theta)) -
              mu_s * (M + m * cos(theta)**2) * g) / (M + m * (sin(theta)**2 + mu_s / 2 * sin(2 * theta))),
        opt4=(F - m * g / 2 * sin(2 * theta) + m * l * dthetadt**2 * (sin(theta) - mu_s * cos(theta)) +
              mu_s * (M + m * cos(theta)**2) * g) / (M + m * (sin(theta)**2 - mu_s / 2 * sin(2 * theta))),
        opt5=0 * F)
    out = dict(parameters=pars, linearization={})
    for opt, expr in d2x.items():
        d2th = g / l * sin(theta) - expr / l * cos(theta)
        A, B, w = linearize(expr, d2th)
        out['linearization'][opt] = dict(A=A, B=B, w=w)
    path = os.path.join(os.path.dirname(__file__), '..', 'explicit_hybrid_mpc_b200', 'data',
                        'pendulum_linearization.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote', os.path.abspath(path))
    for k, v in out['linearization'].items():
        print(k, v['A'][2], v['A'][3], v['B'], v['w'])


if __name__ == '__main__':
    main()
