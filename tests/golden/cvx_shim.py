"""
A minimal stand-in for the parts of CVXPY 1.0 that the reference's ``lib/mpc_library.py``
uses to *build* the pendulum MPC problem (Variable, Parameter, quad_form, Minimize, affine
arithmetic, ==/<=/>= constraints).  It lets ``make_golden.py`` import and run the reference's
own problem-construction code (``InvertedPendulumOnCart.__init__`` and ``make_constraints``,
lib/mpc_library.py:410-554) in this container, where CVXPY/MOSEK are not installed, and read
off the exact constraint matrices -- nothing in here solves anything.

Semantics reproduced from CVXPY 1.0.x: ``*`` is matrix multiplication unless one operand is
scalar-shaped (all dims 1), in which case it is scalar multiplication; 1-D * 1-D is the dot
product.
"""
import numpy as np

_REG = {'nvar': 0, 'names': []}


def reset():
    _REG['nvar'] = 0
    _REG['names'] = []


def nvar():
    return _REG['nvar']


class Expr:
    """Affine expression  A z + P p + b  over the global variable vector z; parameters are
    bound to numeric values before use (see Parameter), so only (A, b) are tracked."""
    __array_ufunc__ = None   # make ndarray op Expr defer to Expr.__r*__

    def __init__(self, A, b):
        self.A = np.atleast_2d(np.asarray(A, dtype=float))
        self.b = np.atleast_1d(np.asarray(b, dtype=float))
        assert self.A.shape[0] == self.b.shape[0]

    # -- helpers
    @property
    def size(self):
        return self.b.shape[0]

    def _pad(self):
        n = nvar()
        if self.A.shape[1] < n:
            self.A = np.hstack([self.A, np.zeros((self.A.shape[0], n - self.A.shape[1]))])
        return self

    @staticmethod
    def wrap(x):
        if isinstance(x, Expr):
            return x._pad()
        x = np.atleast_1d(np.asarray(x, dtype=float))
        return Expr(np.zeros((x.shape[0], nvar())), x)

    def is_scalar(self):
        return self.size == 1

    # -- arithmetic
    def __add__(self, o):
        a, o = Expr.wrap(self), Expr.wrap(o)
        if a.size == 1 and o.size > 1:
            a = Expr(np.repeat(a.A, o.size, 0), np.repeat(a.b, o.size))
        if o.size == 1 and a.size > 1:
            o = Expr(np.repeat(o.A, a.size, 0), np.repeat(o.b, a.size))
        return Expr(a.A + o.A, a.b + o.b)

    __radd__ = __add__

    def __neg__(self):
        return Expr(-self._pad().A, -self.b)

    def __sub__(self, o):
        return self + (-Expr.wrap(o))

    def __rsub__(self, o):
        return Expr.wrap(o) + (-self)

    def __rmul__(self, c):       # constant * Expr
        a = self._pad()
        c = np.asarray(c, dtype=float)
        if c.size == 1:              # scalar-shaped constant: scalar multiplication
            k = float(c.reshape(-1)[0])
            return Expr(k * a.A, k * a.b)
        if a.size == 1:              # array * scalar-shaped expression
            cf = c.reshape(-1)
            return Expr(np.outer(cf, a.A[0]), cf * a.b[0])
        if c.ndim == 1:              # 1-D * 1-D = dot product
            return Expr(c @ a.A, [c @ a.b])
        return Expr(c @ a.A, c @ a.b)

    def __mul__(self, c):        # Expr * constant
        c = np.asarray(c, dtype=float)
        if c.size == 1:
            return self.__rmul__(c)
        if self.size == 1:
            return self.__rmul__(c)
        raise NotImplementedError

    def __getitem__(self, i):
        a = self._pad()
        return Expr(a.A[i:i + 1] if isinstance(i, int) else a.A[i],
                    a.b[i:i + 1] if isinstance(i, int) else a.b[i])

    # -- constraints:  stored as (kind, Expr) meaning Expr <= 0 / Expr == 0
    def __le__(self, o):
        return ('le', self - Expr.wrap(o))

    def __ge__(self, o):
        return ('le', Expr.wrap(o) - self)

    def __eq__(self, o):
        return ('eq', self - Expr.wrap(o))

    __hash__ = None

    @property
    def value(self):
        return None


class Variable(Expr):
    def __init__(self, n=1, boolean=False, nonneg=False, name=None):
        n = int(n) if not isinstance(n, tuple) else int(np.prod(n))
        i0 = _REG['nvar']
        _REG['nvar'] += n
        _REG['names'].append((name, i0, n, boolean))
        A = np.zeros((n, _REG['nvar']))
        A[np.arange(n), i0 + np.arange(n)] = 1.0
        Expr.__init__(self, A, np.zeros(n))
        self.index0, self.n, self.boolean = i0, n, boolean


class Parameter(Expr):
    """A parameter with a numeric ``value`` set BEFORE the expressions using it are built."""

    def __init__(self, n=1, value=None):
        n = int(n)
        v = np.zeros(n) if value is None else np.asarray(value, dtype=float)
        Expr.__init__(self, np.zeros((n, nvar())), v.copy())
        self._v = v

    @property
    def value(self):
        return self._v

    @value.setter
    def value(self, v):
        self._v = np.asarray(v, dtype=float)
        self.b = self._v.copy()


class QuadForm:
    """sum of  e' P e  terms."""

    def __init__(self, terms):
        self.terms = terms

    def __add__(self, o):
        if isinstance(o, (int, float)) and o == 0:
            return self
        return QuadForm(self.terms + o.terms)

    __radd__ = __add__

    def matrices(self):
        n = nvar()
        H = np.zeros((n, n))
        f = np.zeros(n)
        c = 0.0
        for e, P in self.terms:
            e = Expr.wrap(e)
            H += e.A.T @ P @ e.A
            f += 2.0 * e.A.T @ P @ e.b
            c += e.b @ P @ e.b
        return H, f, c   # V(z) = z'Hz + f'z + c


def quad_form(e, P):
    return QuadForm([(e, np.atleast_2d(np.asarray(P, dtype=float)))])


class Minimize:
    def __init__(self, expr):
        self.expr = expr


class Problem:
    def __init__(self, cost, constraints):
        self.cost, self.constraints = cost, constraints


OPTIMAL = 'optimal'
OPTIMAL_INACCURATE = 'optimal_inaccurate'
MOSEK = 'MOSEK'


class SolverError(Exception):
    pass
