"""Minimal eager re-implementation of the slice of the JAX API that CMAD's
material-point / element hot path uses, on NumPy fp64 with nested forward-mode
dual numbers.

TEST INFRASTRUCTURE ONLY.  JAX is not installable in the build container, so the
reference (`/root/reference/cmad`, pure Python on JAX) cannot be imported as is.
This package is put on `sys.path` under the name `jax` by
`tests/golden/make_reference_golden.py` so that the reference's OWN, UNMODIFIED
source files execute and produce golden vectors (`tests/golden/ref_*.npz`).
What is shimmed is the array library and the AD transforms, never the
constitutive code:

* arrays: `Array` wraps a NumPy array; every differentiable primitive carries
  its JVP rule, nested perturbation tags give `jacfwd(jacfwd(...))`;
* `jacfwd`, `jacrev`, `grad`, `hessian`, `jvp`, `value_and_grad`: all computed
  in forward mode, one basis direction at a time (same values as reverse mode
  up to rounding), same pytree structure of the result as JAX;
* `custom_jvp`: the user rule is applied at the outermost live perturbation;
* `lax.while_loop` / `cond` / `scan` / `fori_loop` / `switch`, `vmap`: eager
  Python loops (no tracing: `jit` is the identity);
* pytrees: dict keys in sorted order, `None` is an empty subtree, lists,
  tuples, namedtuples, classes registered with `register_pytree_node_class`.

Nothing here is imported by `cmad_b200/`, by the `-m gpu` tests or at bench
time; only the fixture generator uses it.
"""
from __future__ import annotations

import itertools

import numpy as np

_TAG = [0]              # live perturbation depth
WHILE_LOG: list = []    # (trip_count, final_carry) of every completed lax.while_loop


# --------------------------------------------------------------------------- #
#  Array                                                                      #
# --------------------------------------------------------------------------- #
class Array:
    """Constant (tag 0: `p` is an ndarray) or dual (tag>0: `p`, `t` are Arrays
    of lower tag and equal shape)."""
    __slots__ = ("p", "t", "tag")

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        """NumPy ufuncs reaching a shim array (``ndarray * Array``, ``np.log(Array)``):
        arithmetic is routed to the differentiable primitives, anything else acts on
        the primal values like NumPy does on a real jax.Array through ``__array__``."""
        if method == "__call__" and not kw and ufunc in _UFUNC_MAP:
            return _UFUNC_MAP[ufunc](*inputs)
        args = [_raw(x) if isinstance(x, Array) else x for x in inputs]
        return getattr(ufunc, method)(*args, **kw)

    def __init__(self, p, t=None, tag=0):
        self.p, self.t, self.tag = p, t, tag

    # ---- introspection ----
    @property
    def shape(self):
        return _raw(self).shape

    @property
    def ndim(self):
        return _raw(self).ndim

    @property
    def size(self):
        return _raw(self).size

    @property
    def dtype(self):
        return _raw(self).dtype

    @property
    def T(self):
        return transpose(self)

    @property
    def at(self):
        return _At(self)

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        for i in range(self.shape[0]):
            yield self[i]

    def __array__(self, dtype=None, copy=None):
        r = _raw(self)
        return r.astype(dtype) if dtype is not None else r

    def __float__(self):
        return float(_raw(self))

    def __int__(self):
        return int(_raw(self))

    def __index__(self):
        return int(_raw(self))

    def __bool__(self):
        return bool(_raw(self))

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return f"ShimArray({_raw(self)!r}, tag={self.tag})"

    def __format__(self, spec):
        return format(float(_raw(self)), spec) if self.ndim == 0 else repr(self)

    def item(self):
        return _raw(self).item()

    def tolist(self):
        return _raw(self).tolist()

    def copy(self):
        return self

    def block_until_ready(self):
        return self

    def astype(self, dt):
        if self.tag == 0:
            return Array(self.p.astype(dt))
        return self

    # ---- structural ----
    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return reshape(self, shape)

    def flatten(self):
        return reshape(self, (-1,))

    ravel = flatten

    def squeeze(self, axis=None):
        return _lin1(lambda a: np.squeeze(a, axis=axis))(self)

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        return transpose(self, axes or None)

    def sum(self, axis=None):
        return sum_(self, axis)

    def dot(self, o):
        return matmul(self, o)

    def __getitem__(self, idx):
        idx = _norm_index(idx)
        return _lin1(lambda a: a[idx])(self)

    # ---- arithmetic ----
    def __neg__(self):
        return _lin1(np.negative)(self)

    def __pos__(self):
        return self

    def __abs__(self):
        return abs_(self)

    def __add__(self, o):
        return add(self, o)

    def __radd__(self, o):
        return add(o, self)

    def __sub__(self, o):
        return sub(self, o)

    def __rsub__(self, o):
        return sub(o, self)

    def __mul__(self, o):
        return mul(self, o)

    def __rmul__(self, o):
        return mul(o, self)

    def __truediv__(self, o):
        return div(self, o)

    def __rtruediv__(self, o):
        return div(o, self)

    def __pow__(self, o):
        return power(self, o)

    def __rpow__(self, o):
        return power(o, self)

    def __matmul__(self, o):
        return matmul(self, o)

    def __rmatmul__(self, o):
        return matmul(o, self)

    def __floordiv__(self, o):
        return Array(_raw(self) // _raw(o))

    def __mod__(self, o):
        return Array(_raw(self) % _raw(o))

    # ---- comparisons (not differentiable: on primal values) ----
    def __lt__(self, o):
        return Array(_raw(self) < _raw(o))

    def __le__(self, o):
        return Array(_raw(self) <= _raw(o))

    def __gt__(self, o):
        return Array(_raw(self) > _raw(o))

    def __ge__(self, o):
        return Array(_raw(self) >= _raw(o))

    def __eq__(self, o):
        return Array(_raw(self) == _raw(o))

    def __ne__(self, o):
        return Array(_raw(self) != _raw(o))

    def __and__(self, o):
        return Array(_raw(self) & _raw(o))

    def __or__(self, o):
        return Array(_raw(self) | _raw(o))

    def __rand__(self, o):
        return Array(_raw(o) & _raw(self))

    def __ror__(self, o):
        return Array(_raw(o) | _raw(self))

    def __invert__(self):
        return Array(~_raw(self))


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, _norm_index(idx))


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def _upd(self, v, how):
        idx = self.idx

        def f(xs):
            out = np.array(xs[0], dtype=np.result_type(xs[0], xs[1]), copy=True)
            if how == "set":
                out[idx] = xs[1]
            else:
                np.add.at(out, idx, xs[1])
            return out
        return _linN(f, [self.a, v])

    def set(self, v):
        return self._upd(v, "set")

    def add(self, v):
        return self._upd(v, "add")

    def get(self):
        return self.a[self.idx]


def _norm_index(idx):
    if isinstance(idx, tuple):
        return tuple(_norm_index(i) for i in idx)
    if isinstance(idx, Array):
        r = _raw(idx)
        return int(r) if r.ndim == 0 and r.dtype != bool else r
    if isinstance(idx, list):
        return [_norm_index(i) for i in idx]
    return idx


# --------------------------------------------------------------------------- #
#  helpers                                                                    #
# --------------------------------------------------------------------------- #
def _raw(x):
    while isinstance(x, Array):
        x = x.p
    return x if isinstance(x, np.ndarray) else np.asarray(x)


def _has_array(x):
    if isinstance(x, Array):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_array(i) for i in x)
    return False


def asarray(x, dtype=None):
    if isinstance(x, Array):
        return x if dtype is None else x.astype(dtype)
    if isinstance(x, (list, tuple)) and _has_array(x):
        return stack([asarray(i) for i in x])
    a = np.asarray(x, dtype=dtype)
    if a.dtype == np.float32 or a.dtype == object:
        a = a.astype(np.float64)
    return Array(a)


array = asarray
_A = asarray


def _tag(x):
    return x.tag if isinstance(x, Array) else 0


def _p(x, tag):
    return x.p if x.tag == tag else x


def _t(x, tag):
    return x.t if x.tag == tag else None


def _mk(p, t, tag):
    return p if t is None else Array(p, t, tag)


def _zeros(x):
    return Array(np.zeros(_raw(x).shape))


def _lin1(npf):
    def op(x):
        x = _A(x)
        if x.tag == 0:
            return Array(np.asarray(npf(x.p)))
        return Array(op(x.p), op(x.t), x.tag)
    return op


def _linN(npf, arrs):
    arrs = [_A(a) for a in arrs]
    tag = max(a.tag for a in arrs)
    if tag == 0:
        return Array(np.asarray(npf([a.p for a in arrs])))
    ps = [_p(a, tag) for a in arrs]
    ts = [_t(a, tag) for a in arrs]
    ts = [_zeros(p) if t is None else t for p, t in zip(ps, ts)]
    return Array(_linN(npf, ps), _linN(npf, ts), tag)


def _bc(t, shape):
    if t is None or t.shape == tuple(shape):
        return t
    return _lin1(lambda a: np.broadcast_to(a, shape))(t)


def _addt(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return add(a, b)


def _unbc(x, shape):      # not needed in forward mode; kept for clarity
    return x


# --------------------------------------------------------------------------- #
#  arithmetic primitives                                                      #
# --------------------------------------------------------------------------- #
def _binary(a, b, npf, rule):
    a, b = _A(a), _A(b)
    tag = max(a.tag, b.tag)
    if tag == 0:
        return Array(np.asarray(npf(a.p, b.p)))
    pa, pb, ta, tb = _p(a, tag), _p(b, tag), _t(a, tag), _t(b, tag)
    out = _binary(pa, pb, npf, rule)
    t = rule(pa, pb, ta, tb, out)
    return _mk(out, _bc(t, out.shape), tag)


def add(a, b):
    return _binary(a, b, np.add, lambda pa, pb, ta, tb, o: _addt(ta, tb))


def sub(a, b):
    return _binary(a, b, np.subtract,
                   lambda pa, pb, ta, tb, o: _addt(ta, None if tb is None else -tb))


def mul(a, b):
    return _binary(a, b, np.multiply,
                   lambda pa, pb, ta, tb, o: _addt(None if ta is None else ta * pb,
                                                   None if tb is None else pa * tb))


def div(a, b):
    def rule(pa, pb, ta, tb, o):
        return _addt(None if ta is None else ta / pb,
                     None if tb is None else -(o / pb) * tb)
    with np.errstate(divide="ignore", invalid="ignore"):
        return _binary(a, b, np.true_divide, rule)


def _np_pow(x, y):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return np.power(x, y)


def power(a, b):
    def rule(pa, pb, ta, tb, o):
        # integer exponents follow lax.integer_pow (what `x ** 2` lowers to): d/dx x**n =
        # n x**(n-1), and x**0 is a constant - its derivative is an exact zero, never
        # 0 * x**(-1) (which is NaN at x = 0 and would poison third derivatives of x**2)
        b_ = _A(pb)
        if ta is not None and b_.tag == 0 and np.issubdtype(b_.p.dtype, np.integer) and np.all(b_.p == 0):
            da = ta * 0.0
        else:
            da = None if ta is None else ta * (pb * power(pa, pb - 1))
        db = None if tb is None else tb * (o * log(pa))
        return _addt(da, db)
    a, b = _A(a), _A(b)
    if b.tag == 0 and np.issubdtype(b.p.dtype, np.integer) and np.all(b.p < 0):
        b = Array(b.p.astype(np.float64))
    if a.tag == 0 and np.issubdtype(a.p.dtype, np.integer) and b.tag == 0 and \
            not np.issubdtype(b.p.dtype, np.integer):
        a = Array(a.p.astype(np.float64))
    return _binary(a, b, _np_pow, rule)


def _unary(npf, dfn):
    def op(x):
        x = _A(x)
        if x.tag == 0:
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                return Array(np.asarray(npf(x.p)))
        out = op(x.p)
        return Array(out, dfn(x.p, out) * x.t, x.tag)
    return op


sqrt = _unary(np.sqrt, lambda x, o: 0.5 / o)
exp = _unary(np.exp, lambda x, o: o)
log = _unary(np.log, lambda x, o: 1.0 / x)
cbrt = _unary(np.cbrt, lambda x, o: o / (3.0 * x))
abs_ = _unary(np.abs, lambda x, o: Array(np.sign(_raw(x))))
sin = _unary(np.sin, lambda x, o: cos(x))
cos = _unary(np.cos, lambda x, o: -sin(x))
tanh = _unary(np.tanh, lambda x, o: 1.0 - o * o)
arccos = _unary(np.arccos, lambda x, o: -1.0 / sqrt(1.0 - x * x))
arctan = _unary(np.arctan, lambda x, o: 1.0 / (1.0 + x * x))
sign = lambda x: Array(np.sign(_raw(x)))  # noqa: E731


def where(c, x, y):
    cr = _raw(c)
    return _linN(lambda xs: np.where(cr, xs[0], xs[1]), [x, y])


def maximum(x, y):
    return where(Array(_raw(x) >= _raw(y)), x, y)


def minimum(x, y):
    return where(Array(_raw(x) <= _raw(y)), x, y)


def clip(x, lo=None, hi=None):
    if lo is not None:
        x = maximum(x, lo)
    if hi is not None:
        x = minimum(x, hi)
    return x


# ---- structural / linear ----
def reshape(x, shape):
    return _lin1(lambda a: np.reshape(a, shape))(x)


def transpose(x, axes=None):
    return _lin1(lambda a: np.transpose(a, axes))(x)


def sum_(x, axis=None, keepdims=False):
    return _lin1(lambda a: np.sum(a, axis=axis, keepdims=keepdims))(x)


def prod(x, axis=None):
    x = _A(x)
    if axis is None:
        x, axis = ravel(x), 0
    n = x.shape[axis]
    out = take(x, 0, axis=axis)
    for i in range(1, n):
        out = out * take(x, i, axis=axis)
    return out


def checkpoint(fun, **kw):
    return fun


def mean(x, axis=None):
    return _lin1(lambda a: np.mean(a, axis=axis))(x)


def trace(x):
    return _lin1(np.trace)(x)


def diag(x, k=0):
    return _lin1(lambda a: np.diag(a, k))(x)


def stack(xs, axis=0):
    return _linN(lambda a: np.stack(a, axis=axis), list(xs))


def concatenate(xs, axis=0):
    return _linN(lambda a: np.concatenate(a, axis=axis), list(xs))


def hstack(xs):
    return _linN(lambda a: np.hstack(a), list(xs))


def vstack(xs):
    return _linN(lambda a: np.vstack(a), list(xs))


def broadcast_to(x, shape):
    return _lin1(lambda a: np.broadcast_to(a, shape))(x)


def atleast_1d(x):
    return _lin1(np.atleast_1d)(x)


def atleast_2d(x):
    return _lin1(np.atleast_2d)(x)


def expand_dims(x, axis):
    return _lin1(lambda a: np.expand_dims(a, axis))(x)


def squeeze(x, axis=None):
    return _A(x).squeeze(axis)


def ravel(x):
    return reshape(x, (-1,))


def take(x, idx, axis=None):
    i = _raw(idx)
    return _lin1(lambda a: np.take(a, i, axis=axis))(x)


def tril(x, k=0):
    return _lin1(lambda a: np.tril(a, k))(x)


def triu(x, k=0):
    return _lin1(lambda a: np.triu(a, k))(x)


def cumsum(x, axis=None):
    return _lin1(lambda a: np.cumsum(a, axis=axis))(x)


def roll(x, shift, axis=None):
    return _lin1(lambda a: np.roll(a, shift, axis=axis))(x)


def flip(x, axis=None):
    return _lin1(lambda a: np.flip(a, axis=axis))(x)


def swapaxes(x, a1, a2):
    return _lin1(lambda a: np.swapaxes(a, a1, a2))(x)


def moveaxis(x, s, d):
    return _lin1(lambda a: np.moveaxis(a, s, d))(x)


def repeat(x, n, axis=None):
    return _lin1(lambda a: np.repeat(a, n, axis=axis))(x)


def tile(x, reps):
    return _lin1(lambda a: np.tile(a, reps))(x)


# ---- multilinear ----
def einsum(spec, *ops):
    ops = [_A(o) for o in ops]
    tag = max(o.tag for o in ops)
    if tag == 0:
        return Array(np.asarray(np.einsum(spec, *[o.p for o in ops])))
    ps = [_p(o, tag) for o in ops]
    out = einsum(spec, *ps)
    t = None
    for i, o in enumerate(ops):
        ti = _t(o, tag)
        if ti is not None:
            t = _addt(t, einsum(spec, *(ps[:i] + [ti] + ps[i + 1:])))
    return _mk(out, t, tag)


def matmul(a, b):
    a, b = _A(a), _A(b)
    tag = max(a.tag, b.tag)
    if tag == 0:
        return Array(np.asarray(np.matmul(a.p, b.p)))
    pa, pb, ta, tb = _p(a, tag), _p(b, tag), _t(a, tag), _t(b, tag)
    out = matmul(pa, pb)
    t = _addt(None if ta is None else matmul(ta, pb), None if tb is None else matmul(pa, tb))
    return _mk(out, t, tag)


dot = matmul


def outer(a, b):
    return einsum("i,j->ij", ravel(a), ravel(b))


def tensordot(a, b, axes=2):
    a, b = _A(a), _A(b)
    tag = max(a.tag, b.tag)
    if tag == 0:
        return Array(np.asarray(np.tensordot(a.p, b.p, axes)))
    pa, pb, ta, tb = _p(a, tag), _p(b, tag), _t(a, tag), _t(b, tag)
    out = tensordot(pa, pb, axes)
    t = _addt(None if ta is None else tensordot(ta, pb, axes),
              None if tb is None else tensordot(pa, tb, axes))
    return _mk(out, t, tag)


def cross(a, b):
    a, b = _A(a), _A(b)
    return stack([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


# ---- linalg ----
def solve(A, b):
    A, b = _A(A), _A(b)
    tag = max(A.tag, b.tag)
    if tag == 0:
        return Array(np.linalg.solve(A.p, b.p))
    pA, pb, tA, tb = _p(A, tag), _p(b, tag), _t(A, tag), _t(b, tag)
    x = solve(pA, pb)
    rhs = _addt(tb, None if tA is None else -matmul(tA, x))
    return _mk(x, solve(pA, rhs), tag)


def inv(A):
    A = _A(A)
    if A.tag == 0:
        return Array(np.linalg.inv(A.p))
    Ai = inv(A.p)
    return Array(Ai, -matmul(matmul(Ai, A.t), Ai), A.tag)


def det(A):
    A = _A(A)
    if A.tag == 0:
        return Array(np.asarray(np.linalg.det(A.p)))
    d = det(A.p)
    return Array(d, d * trace(solve(A.p, A.t)), A.tag)


def eigh(A, UPLO=None, symmetrize_input=True):
    """`jnp.linalg.eigh` with JAX's JVP rule (jax/_src/lax/linalg.py, `_eigh_jvp_rule`):
    dw = diag(V^T dA V), dV = V (F o V^T dA V), F_ij = 1 / (w_j - w_i) off the diagonal
    (inf / NaN at repeated eigenvalues, like JAX).  The input is symmetrised first, as
    JAX does, so the derivative with respect to a single off-diagonal entry is half the
    symmetric one.  Written on shim primitives: nests for second derivatives."""
    A = _A(A)
    if A.tag == 0:
        w, v = np.linalg.eigh(0.5 * (A.p + A.p.T))
        return Array(w), Array(v)
    w, v = eigh(A.p)
    At = 0.5 * (A.t + transpose(A.t))
    vAv = matmul(matmul(transpose(v), At), v)
    n = A.shape[0]
    I = Array(np.eye(n))
    with np.errstate(divide="ignore", invalid="ignore"):
        F = 1.0 / (I + reshape(w, (1, n)) - reshape(w, (n, 1))) - I
        dv = matmul(v, F * vAv)
    return Array(w, diag(vAv), A.tag), Array(v, dv, A.tag)


def norm(x, ord=None, axis=None):
    x = _A(x)
    if ord not in (None, 2, "fro"):
        raise NotImplementedError("shim norm: only 2/fro")
    return sqrt(sum_(x * x, axis=axis))


# ---- non-differentiable ----
def _nd(npf):
    def f(*a, **k):
        return Array(np.asarray(npf(*[_raw(x) if isinstance(x, Array) else x for x in a], **k)))
    return f


logical_and = _nd(np.logical_and)
logical_or = _nd(np.logical_or)
logical_not = _nd(np.logical_not)
isfinite = _nd(np.isfinite)
isnan = _nd(np.isnan)
all_ = _nd(np.all)
any_ = _nd(np.any)
argmax = _nd(np.argmax)
argmin = _nd(np.argmin)
argsort = _nd(np.argsort)
floor = _nd(np.floor)


def isclose(a, b, rtol=1e-5, atol=1e-8):
    return Array(np.asarray(np.isclose(_raw(a), _raw(b), rtol=rtol, atol=atol)))


def allclose(a, b, rtol=1e-5, atol=1e-8):
    return bool(np.allclose(_raw(a), _raw(b), rtol=rtol, atol=atol))


def sort(x, axis=-1):
    x = _A(x)
    if x.tag:
        raise NotImplementedError("shim sort of a perturbed array")
    return Array(np.sort(x.p, axis=axis))


def setdiff1d(a, b, size=None, **_):
    r = np.setdiff1d(_raw(a), _raw(b))
    return Array(r if size is None else r[:size])


def unique(a, **k):
    return Array(np.unique(_raw(a)))


# ---- constructors ----
def zeros(shape, dtype=float):
    return Array(np.zeros(shape, dtype=dtype))


def ones(shape, dtype=float):
    return Array(np.ones(shape, dtype=dtype))


def full(shape, v, dtype=None):
    return broadcast_to(_A(v), shape if isinstance(shape, tuple) else (shape,))


def eye(n, m=None, dtype=float):
    return Array(np.eye(n, m, dtype=dtype))


def arange(*a, **k):
    return Array(np.arange(*a, **k))


def linspace(*a, **k):
    return Array(np.linspace(*a, **k))


def zeros_like(x, dtype=None):
    return Array(np.zeros(_raw(x).shape, dtype=dtype or _raw(x).dtype))


def ones_like(x, dtype=None):
    return Array(np.ones(_raw(x).shape, dtype=dtype or _raw(x).dtype))


class _RClass:
    def __getitem__(self, items):
        if not isinstance(items, tuple):
            items = (items,)
        return concatenate([atleast_1d(i) for i in items], axis=0)


class _CClass:
    def __getitem__(self, items):
        if not isinstance(items, tuple):
            items = (items,)
        cols = []
        for i in items:
            i = _A(i)
            cols.append(reshape(i, (-1, 1)) if i.ndim < 2 else i)
        return concatenate(cols, axis=-1)


r_, c_ = _RClass(), _CClass()


# --------------------------------------------------------------------------- #
#  pytrees                                                                    #
# --------------------------------------------------------------------------- #
_REGISTRY: dict = {}


def register_pytree_node_class(cls):
    _REGISTRY[cls] = (lambda o: o.tree_flatten(), lambda aux, ch: cls.tree_unflatten(aux, ch))
    return cls


def register_pytree_node(cls, flat, unflat):
    _REGISTRY[cls] = (flat, unflat)


class DictKey:
    def __init__(self, key):
        self.key = key

    def __str__(self):
        return f"[{self.key!r}]"

    __repr__ = __str__


class SequenceKey:
    def __init__(self, idx):
        self.idx = idx

    def __str__(self):
        return f"[{self.idx}]"

    __repr__ = __str__


class GetAttrKey:
    def __init__(self, name):
        self.name = name

    def __str__(self):
        return f".{self.name}"

    __repr__ = __str__


class TreeDef:
    """kind in {leaf, none, dict, list, tuple, namedtuple, custom}."""

    def __init__(self, kind, meta=None, children=()):
        self.kind, self.meta, self.children = kind, meta, tuple(children)
        self.num_leaves = 1 if kind == "leaf" else sum(c.num_leaves for c in self.children)

    def unflatten(self, leaves):
        it = iter(leaves)
        out = self._build(it)
        return out

    def _build(self, it):
        k = self.kind
        if k == "leaf":
            return next(it)
        if k == "none":
            return None
        ch = [c._build(it) for c in self.children]
        if k == "dict":
            return dict(zip(self.meta, ch))
        if k == "list":
            return ch
        if k == "tuple":
            return tuple(ch)
        if k == "namedtuple":
            return self.meta(*ch)
        flat_unflat, aux = self.meta
        return flat_unflat[1](aux, tuple(ch))

    def __eq__(self, o):
        return isinstance(o, TreeDef) and self.kind == o.kind and self.children == o.children and \
            (self.meta == o.meta if self.kind in ("dict",) else True)

    def __hash__(self):
        return hash((self.kind, self.children))

    def __repr__(self):
        return f"TreeDef({self.kind}, {self.num_leaves} leaves)"


def _flatten(tree, is_leaf, path, out):
    if is_leaf is not None and is_leaf(tree):
        out.append((path, tree))
        return TreeDef("leaf")
    if tree is None:
        return TreeDef("none")
    if isinstance(tree, dict):
        keys = sorted(tree.keys())
        return TreeDef("dict", keys, [_flatten(tree[k], is_leaf, path + (DictKey(k),), out) for k in keys])
    if type(tree) in _REGISTRY:
        fu = _REGISTRY[type(tree)]
        children, aux = fu[0](tree)
        return TreeDef("custom", (fu, aux),
                       [_flatten(c, is_leaf, path + (SequenceKey(i),), out) for i, c in enumerate(children)])
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return TreeDef("namedtuple", type(tree),
                       [_flatten(c, is_leaf, path + (GetAttrKey(n),), out) for n, c in zip(tree._fields, tree)])
    if isinstance(tree, list):
        return TreeDef("list", None, [_flatten(c, is_leaf, path + (SequenceKey(i),), out) for i, c in enumerate(tree)])
    if isinstance(tree, tuple):
        return TreeDef("tuple", None, [_flatten(c, is_leaf, path + (SequenceKey(i),), out) for i, c in enumerate(tree)])
    out.append((path, tree))
    return TreeDef("leaf")


def tree_flatten(tree, is_leaf=None):
    out: list = []
    td = _flatten(tree, is_leaf, (), out)
    return [l for _, l in out], td


def tree_flatten_with_path(tree, is_leaf=None):
    out: list = []
    td = _flatten(tree, is_leaf, (), out)
    return out, td


def tree_unflatten(treedef, leaves):
    return treedef.unflatten(leaves)


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_structure(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[1]


def _flatten_up_to(td, tree):
    """Flatten `tree` only as deep as `td` goes (JAX's prefix rule for tree_map)."""
    if td.kind == "leaf":
        return [tree]
    if td.kind == "none":
        return []
    if td.kind == "dict":
        ch = [tree[k] for k in td.meta]
    elif td.kind == "custom":
        ch = list(td.meta[0][0](tree)[0])
    else:
        ch = list(tree)
    out = []
    for c, t in zip(td.children, ch):
        out += _flatten_up_to(c, t)
    return out


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, td = tree_flatten(tree, is_leaf)
    others = [_flatten_up_to(td, r) for r in rest]
    return td.unflatten([f(*xs) for xs in zip(leaves, *others)])


def ravel_pytree(tree):
    leaves, td = tree_flatten(tree)
    shapes = [np.shape(_raw(l)) for l in leaves]
    sizes = [int(np.prod(s)) for s in shapes]
    flat = concatenate([ravel(_A(l)) for l in leaves]) if leaves else Array(np.zeros(0))

    def unravel(v):
        v = _A(v)
        out, o = [], 0
        for s, n in zip(shapes, sizes):
            out.append(reshape(v[o:o + n], s))
            o += n
        return td.unflatten(out)
    return flat, unravel


# --------------------------------------------------------------------------- #
#  transforms                                                                 #
# --------------------------------------------------------------------------- #
def _max_tag(tree):
    return max([_tag(l) for l in tree_leaves(tree)] + [0])


def jvp(fun, primals, tangents, has_aux=False):
    _TAG[0] += 1
    tag = _TAG[0]
    try:
        pl, td = tree_flatten(tuple(primals))
        tl = _flatten_up_to(td, tuple(tangents))
        duals = []
        for p, t in zip(pl, tl):
            p = _A(p)
            if t is None or (not isinstance(t, Array) and np.asarray(t).dtype.kind not in "fc"):
                duals.append(p)
            else:
                duals.append(Array(p, _bc(_A(t), p.shape), tag))
        out = fun(*td.unflatten(duals))
        ol, otd = tree_flatten(out)
        ol = [_A(o) for o in ol]
        po = [_p(o, tag) for o in ol]
        to = [_t(o, tag) for o in ol]
        to = [_zeros(p) if t is None else t for p, t in zip(po, to)]
        return otd.unflatten(po), otd.unflatten(to)
    finally:
        _TAG[0] -= 1


def _jac(fun, argnums, has_aux=False):
    single = isinstance(argnums, int)
    nums = (argnums,) if single else tuple(argnums)

    def jf(*args, **kw):
        args = list(args)
        per_arg = []
        out_td = out_shapes = None
        for an in nums:
            leaves, td = tree_flatten(args[an])
            leaves = [_A(l) for l in leaves]
            cols = [[] for _ in leaves]
            for li, leaf in enumerate(leaves):
                for k in range(max(leaf.size, 1) if leaf.ndim else 1):
                    e = np.zeros(leaf.shape)
                    e.reshape(-1)[k] = 1.0
                    tans = [Array(e) if j == li else None for j in range(len(leaves))]

                    def g(*lv, _an=an, _td=td):
                        a2 = list(args)
                        a2[_an] = _td.unflatten(list(lv))
                        return fun(*a2, **kw)
                    _, tout = jvp(g, tuple(leaves), tuple(tans))
                    ol, out_td = tree_flatten(tout)
                    out_shapes = [o.shape for o in ol]
                    cols[li].append(ol)
            # assemble: for every output leaf, a tree like arg with blocks out_shape + leaf_shape
            blocks_per_out = []
            for oi, osh in enumerate(out_shapes):
                blk = []
                for li, leaf in enumerate(leaves):
                    st = stack([c[oi] for c in cols[li]], axis=-1)
                    blk.append(reshape(st, tuple(osh) + tuple(leaf.shape)))
                blocks_per_out.append(td.unflatten(blk))
            per_arg.append(blocks_per_out)
        outs = []
        for oi in range(len(out_shapes)):
            outs.append(per_arg[0][oi] if single else tuple(pa[oi] for pa in per_arg))
        return out_td.unflatten(outs)
    return jf


def jacfwd(fun, argnums=0, has_aux=False, holomorphic=False):
    return _jac(fun, argnums)


def jacrev(fun, argnums=0, has_aux=False, holomorphic=False, allow_int=False):
    return _jac(fun, argnums)


jacobian = jacrev


def grad(fun, argnums=0, has_aux=False, holomorphic=False, allow_int=False):
    if has_aux:
        def f0(*a, **k):
            return fun(*a, **k)[0]

        def g(*a, **k):
            return _jac(f0, argnums)(*a, **k), fun(*a, **k)[1]
        return g
    return _jac(fun, argnums)


def value_and_grad(fun, argnums=0, has_aux=False, holomorphic=False):
    def vg(*a, **k):
        if has_aux:
            v = fun(*a, **k)
            return v, _jac(lambda *aa, **kk: fun(*aa, **kk)[0], argnums)(*a, **k)
        return fun(*a, **k), _jac(fun, argnums)(*a, **k)
    return vg


def hessian(fun, argnums=0, has_aux=False, holomorphic=False):
    return jacfwd(jacrev(fun, argnums), argnums)


class custom_jvp:
    def __init__(self, fun, nondiff_argnums=()):
        self.fun, self.rule = fun, None
        self.__name__ = getattr(fun, "__name__", "custom_jvp")

    def defjvp(self, rule, symbolic_zeros=False):
        self.rule = rule
        return rule

    def __call__(self, *args):
        tag = _max_tag(args)
        if tag == 0 or self.rule is None:
            return self.fun(*args)
        leaves, td = tree_flatten(tuple(args))
        leaves = [_A(l) for l in leaves]
        ps = [_p(l, tag) for l in leaves]
        ts = [_t(l, tag) for l in leaves]
        ts = [_zeros(p) if t is None else t for p, t in zip(ps, ts)]
        po, to = self.rule(td.unflatten(ps), td.unflatten(ts))
        pl, otd = tree_flatten(po)
        tl = _flatten_up_to(otd, to)
        return otd.unflatten([Array(_A(p), _bc(_A(t), _A(p).shape), tag) for p, t in zip(pl, tl)])


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


class ShapeDtypeStruct:
    def __init__(self, shape, dtype, **kw):
        self.shape, self.dtype = tuple(shape), dtype


def pure_callback(callback, result_shape_dtypes, *args, **kw):
    """Eager: the host callback sees plain NumPy operands (forward evaluation only)."""
    kw.pop("vmap_method", None)
    kw.pop("vectorized", None)
    res = callback(*[tree_map(lambda l: np.asarray(_raw(l)) if isinstance(l, Array) else l, a) for a in args], **kw)
    return tree_map(lambda l: asarray(l), res)


def custom_linear_solve(matvec, b, solve, transpose_solve=None, symmetric=False, has_aux=False):
    """Forward evaluation of lax.custom_linear_solve: x = solve(matvec, b)."""
    return solve(matvec, b)


def stop_gradient(x):
    return tree_map(lambda l: Array(_raw(l)) if isinstance(l, Array) else l, x)


# ---- control flow (eager) ----
def while_loop(cond_fun, body_fun, init):
    c, n = init, 0
    while bool(cond_fun(c)):
        c = body_fun(c)
        n += 1
    WHILE_LOG.append((n, c))
    return c


def cond(pred, true_fun, false_fun, *operands, **kw):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def switch(index, branches, *operands):
    i = int(np.clip(int(index), 0, len(branches) - 1))
    return branches[i](*operands)


def fori_loop(lo, hi, body, init):
    v = init
    for i in range(int(lo), int(hi)):
        v = body(i, v)
    return v


def _stack_trees(trees):
    if not trees:
        return None
    ls = [tree_flatten(t) for t in trees]
    td = ls[0][1]
    n = len(ls[0][0])
    return td.unflatten([stack([_A(l[0][i]) for l in ls]) for i in range(n)])


def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
    if xs is None:
        n = int(length)
        get = lambda i: None  # noqa: E731
    else:
        xl, xtd = tree_flatten(xs)
        n = int(_A(xl[0]).shape[0]) if xl else int(length)
        get = lambda i: xtd.unflatten([_A(x)[i] for x in xl])  # noqa: E731
    carry, ys = init, []
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        carry, y = f(carry, get(i))
        ys.append(y)
    if reverse:
        ys = ys[::-1]
    return carry, _stack_trees(ys)


def vmap(fun, in_axes=0, out_axes=0, **kw):
    def vf(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            axl = _flatten_up_to(tree_structure(ax, is_leaf=lambda x: x is None), a) \
                if not isinstance(ax, int) else None
            lv = tree_leaves(a)
            if isinstance(ax, int):
                if lv:
                    n = int(_A(lv[0]).shape[ax])
                    break
            else:
                for sub_ax, sub in zip(tree_leaves(ax, is_leaf=lambda x: x is None), axl):
                    if sub_ax is not None and tree_leaves(sub):
                        n = int(_A(tree_leaves(sub)[0]).shape[sub_ax])
                        break
                if n is not None:
                    break
        outs = []
        for i in range(n):
            sl = []
            for a, ax in zip(args, axes):
                sl.append(_slice_tree(a, ax, i))
            outs.append(fun(*sl))
        res = _stack_trees(outs)
        if out_axes != 0 and isinstance(out_axes, int):
            res = tree_map(lambda l: moveaxis(l, 0, out_axes), res)
        return res
    return vf


def _slice_tree(a, ax, i):
    if ax is None:
        return a
    if isinstance(ax, int):
        return tree_map(lambda l: _lin1(lambda x: np.take(x, i, axis=ax))(_A(l)), a)
    td = tree_structure(ax, is_leaf=lambda x: x is None)
    subs = _flatten_up_to(td, a)
    axl = tree_leaves(ax, is_leaf=lambda x: x is None)
    return td.unflatten([_slice_tree(s, sa, i) for s, sa in zip(subs, axl)])


_UFUNC_MAP = {np.add: add, np.subtract: sub, np.multiply: mul, np.true_divide: div, np.power: power,
              np.matmul: matmul, np.negative: lambda x: -_A(x), np.sqrt: sqrt, np.exp: exp,
              np.absolute: abs_,
              np.less: lambda a, b: _A(a) < b, np.less_equal: lambda a, b: _A(a) <= b,
              np.greater: lambda a, b: _A(a) > b, np.greater_equal: lambda a, b: _A(a) >= b,
              np.equal: lambda a, b: _A(a) == b, np.not_equal: lambda a, b: _A(a) != b}


class _Debug:
    @staticmethod
    def print(fmt, *a, **k):
        pass

    @staticmethod
    def callback(f, *a, **k):
        return f(*a, **k)


debug = _Debug()


class _Config:
    def update(self, *a, **k):
        pass

    jax_enable_x64 = True


config = _Config()
_ = itertools
