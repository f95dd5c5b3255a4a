"""refshim -- execute the UNMODIFIED reference model bodies without jax / numpyro / funsor.

TEST INFRASTRUCTURE ONLY (same rule as the rest of ``oracle/``: only ``tests/`` and the fixture
generators under ``tests/golden/`` import this; nothing under ``biolith_b200/`` may).

The reference (timmh/biolith) only *declares* its probabilistic programs; numpyro traces them, funsor
sums out the enumerated discrete latents and jax differentiates.  None of the three is installable here
(SURVEY.md section 8c).  This module provides *functional* numpy stand-ins for exactly the primitives the
reference's model bodies touch

    jax.numpy (-> numpy), jax.nn.sigmoid, jax.scipy.special.logsumexp,
    numpyro.sample / plate / deterministic / factor, numpyro.handlers.mask,
    numpyro.distributions.{Normal, HalfNormal, Beta, Exponential, Gamma, Bernoulli, Poisson, Binomial,
                           Categorical, TruncatedDistribution}  (+ .expand / .to_event / .support),
    numpyro's enumerate="parallel" dim allocation and a plate-aware sum-product (funsor's job),
    biject_to(support) transforms with their log-Jacobians (numpyro's potential_fn),

so that ``biolith.models.occu`` / ``occu_rn`` / ``occu_cop`` / ``nmixture`` / ``occu_cs`` (and
``biolith.regression.LinearRegression``, ``biolith.utils.modeling``, ``biolith.utils.distributions``) are
imported from /root/reference and RUN AS THEY ARE: every mask resolution, transpose, flatten, reshape,
linear predictor, false-positive formula, enumeration shape and plate nesting in the result comes from
reference lines, not from a restatement.  What is still restated (from numpyro >= 0.18's published
source, recalled, not vendored) is the arithmetic *inside* the distribution objects:

    clamp_probs(p)          = clip(p, finfo.tiny, 1 - finfo.eps)                 (distributions/util.py)
    BernoulliProbs.log_prob = xlogy(v, p~) + xlog1py(1 - v, -p~)                 (distributions/discrete.py)
    BinomialProbs.log_prob  = gammaln(n+1) - gammaln(v+1) - gammaln(n-v+1) + xlogy(v, p~) + xlog1py(n-v, -p~)
    Poisson.log_prob        = log(rate) * v - gammaln(v + 1) - rate
    CategoricalLogits       = logits - logsumexp(logits), gathered at the value
    Normal / HalfNormal / Beta (via Dirichlet) / Exponential / Gamma / LeftTruncatedDistribution log_prob
    MaskedDistribution      = where(m, base.log_prob(where(m, v, feasible)), 0)  (handlers.mask)
    plate                   : broadcasts a site's log_prob to the plate size (expand)
    enum                    : support placed at dim -(max_plate_nesting + 1 + i) for the i-th enumerated site
    potential_fn            : unconstrained substitution through SigmoidTransform (clipped expit) /
                              ExpTransform / exp + low, log|J| added

Derivatives: the bodies are evaluated on complex128 inputs theta + i*h*e_k (complex-step differentiation,
h = 1e-30): every operation on the path is analytic away from the clip/where decisions, which are taken on
the real part exactly like jnp.clip / jnp.where propagate zero / selected cotangents.  The result is the
directional derivative of the *executed reference body* to machine precision, with no subtraction error.

``float_dtype`` selects the working precision of the stand-in (float64, or float32 to mimic the reference's
default x64-disabled arithmetic op by op); ``clamp_dtype`` selects whose ``finfo`` clamp_probs uses (the repo's
fp32 kernels are checked against float64 arithmetic with float32 clamp constants).
"""

from __future__ import annotations

import contextlib
import importlib.abc
import importlib.machinery
import math
import sys
import types
from collections import OrderedDict
from unittest import mock

import numpy as np
from scipy import special as _sp

REFERENCE_ROOT = "/root/reference"
_STUB_ROOTS = ("jax", "numpyro", "funsor", "rpy2", "optax", "flax")


# ------------------------------------------------------------------------------------------------
# configuration
# ------------------------------------------------------------------------------------------------
class _Config:
    float_dtype = np.float64
    clamp_dtype = None  # None -> float_dtype

    @classmethod
    def finfo(cls):
        return np.finfo(cls.clamp_dtype or cls.float_dtype)


@contextlib.contextmanager
def precision(float_dtype=np.float64, clamp_dtype=None):
    old = (_Config.float_dtype, _Config.clamp_dtype)
    _Config.float_dtype, _Config.clamp_dtype = float_dtype, clamp_dtype
    try:
        yield
    finally:
        _Config.float_dtype, _Config.clamp_dtype = old


def _is_complex(x):
    return np.iscomplexobj(x)


def _re(x):
    return np.real(x)


def _default_float(a):
    """jnp.asarray semantics: python/np float64 data becomes the configured default float type."""
    a = np.asarray(a)
    if a.dtype == np.float64 and _Config.float_dtype != np.float64:
        return a.astype(_Config.float_dtype)
    return a


# ------------------------------------------------------------------------------------------------
# jax.numpy / jax.nn / jax.scipy.special stand-ins (numpy, complex-step safe)
# ------------------------------------------------------------------------------------------------
class ConcretizationTypeError(TypeError):
    pass


def _clip(x, a_min=None, a_max=None, *, min=None, max=None):  # noqa: A002 - jnp.clip's keyword names
    lo = a_min if a_min is not None else min
    hi = a_max if a_max is not None else max
    x = np.asarray(x)
    if not _is_complex(x):
        return np.clip(x, lo, hi)
    out = x
    if lo is not None:
        out = np.where(_re(out) < lo, lo, out)  # constant (zero tangent) outside, like jnp.clip's vjp
    if hi is not None:
        out = np.where(_re(out) > hi, hi, out)
    return out


def _ceil(x):
    if _is_complex(x):  # a traced value under jit: the reference catches exactly this (distributions.py:19-29)
        raise ConcretizationTypeError("abstract value")
    return np.ceil(x)


def _sigmoid(x):
    x = np.asarray(x)
    pos = _re(x) >= 0
    with np.errstate(over="ignore", invalid="ignore"):
        e = np.exp(np.where(pos, -x, x))
    return np.where(pos, 1.0 / (1.0 + e), e / (1.0 + e))


def _logsumexp(a, axis=None, b=None, keepdims=False):
    assert b is None
    a = np.asarray(a)
    m = np.max(_re(a), axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore"):
        out = np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True)) + m
    if not keepdims:
        out = np.squeeze(out, axis=axis) if axis is not None else out.reshape(())
    return out


def _xlogy(x, y):
    x = np.asarray(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(x == 0, 0.0, x * np.log(np.where(x == 0, 1.0, y)))


def _xlog1py(x, y):
    x = np.asarray(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(x == 0, 0.0, x * np.log1p(np.where(x == 0, 0.0, y)))


def _gammaln(x):
    x = np.asarray(x)
    if _is_complex(x):
        return _sp.loggamma(x)
    return _sp.gammaln(x)


def _softplus(x):
    x = np.asarray(x)
    return np.where(_re(x) > 0, x, 0.0) + np.log1p(np.exp(-np.where(_re(x) > 0, x, -x)))


class _JnpModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(np, name)


def _make_jnp():
    jnp = _JnpModule("jax.numpy")
    jnp.__path__ = []
    jnp.ndarray = np.ndarray
    jnp.array = lambda x, dtype=None: _default_float(np.array(x, dtype=dtype))
    jnp.asarray = lambda x, dtype=None: _default_float(np.asarray(x, dtype=dtype))
    jnp.zeros = lambda shape, dtype=None: np.zeros(shape, dtype=dtype or _Config.float_dtype)
    jnp.ones = lambda shape, dtype=None: np.ones(shape, dtype=dtype or _Config.float_dtype)
    jnp.clip = _clip
    jnp.ceil = _ceil
    jnp.nan = np.nan
    jnp.inf = np.inf
    return jnp


# ------------------------------------------------------------------------------------------------
# numpyro.distributions stand-ins
# ------------------------------------------------------------------------------------------------
class _Constraint:
    def __init__(self, kind, low=None):
        self.kind, self.low = kind, low

    def __repr__(self):
        return f"constraint({self.kind})"


real = _Constraint("real")
positive = _Constraint("positive")
unit_interval = _Constraint("unit_interval")
boolean = _Constraint("boolean")
nonnegative_integer = _Constraint("nonnegative_integer")


def _clamp_probs(p):
    fi = _Config.finfo()
    return _clip(p, fi.tiny, 1.0 - fi.eps)


class Distribution:
    support = real
    event_dim = 0
    has_enumerate_support = False

    @property
    def batch_shape(self):
        return ()

    def expand(self, batch_shape):
        return _Expanded(self, tuple(batch_shape))

    def to_event(self, n=None):
        return _Independent(self, int(n)) if n else self

    def log_prob(self, value):  # pragma: no cover - abstract
        raise NotImplementedError


class _Expanded(Distribution):
    def __init__(self, base, shape):
        self.base, self._shape = base, shape
        self.support, self.event_dim = base.support, base.event_dim

    @property
    def batch_shape(self):
        return self._shape

    def log_prob(self, value):
        lp = self.base.log_prob(value)
        return np.broadcast_to(lp, np.broadcast_shapes(np.shape(lp), self._shape))


class _Independent(Distribution):
    def __init__(self, base, n):
        self.base, self.n = base, n
        self.support, self.event_dim = base.support, base.event_dim + n

    @property
    def batch_shape(self):
        return self.base.batch_shape[: len(self.base.batch_shape) - self.n]

    def log_prob(self, value):
        return self.base.log_prob(value).sum(axis=tuple(range(-self.n, 0)))


class Normal(Distribution):
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = loc, scale

    @property
    def batch_shape(self):
        return np.broadcast_shapes(np.shape(self.loc), np.shape(self.scale))

    def log_prob(self, value):
        normalize_term = np.log(math.sqrt(2 * math.pi) * self.scale)
        value_scaled = (value - self.loc) / self.scale
        return -0.5 * value_scaled**2 - normalize_term

    def log_sf(self, x):
        return _sp.log_ndtr(-(x - self.loc) / self.scale)


class HalfNormal(Distribution):
    support = positive

    def __init__(self, scale=1.0):
        self.scale = scale

    def log_prob(self, value):
        return Normal(0.0, self.scale).log_prob(value) + math.log(2.0)


class Beta(Distribution):
    support = unit_interval

    def __init__(self, concentration1, concentration0):
        self.c1, self.c0 = concentration1, concentration0

    def log_prob(self, value):  # Dirichlet([c1, c0]).log_prob([v, 1 - v])
        norm = _sp.gammaln(self.c1) + _sp.gammaln(self.c0) - _sp.gammaln(self.c1 + self.c0)
        return _xlogy(self.c1 - 1.0, value) + _xlogy(self.c0 - 1.0, 1.0 - value) - norm


class Exponential(Distribution):
    support = positive

    def __init__(self, rate=1.0):
        self.rate = rate

    def log_prob(self, value):
        return np.log(self.rate) - self.rate * value


class Gamma(Distribution):
    support = positive

    def __init__(self, concentration, rate=1.0):
        self.concentration, self.rate = concentration, rate

    def log_prob(self, value):
        c, r = self.concentration, self.rate
        return -_sp.gammaln(c) + c * np.log(r) + (c - 1.0) * np.log(value) - r * value


class TruncatedDistribution(Distribution):
    """Only the left-truncated Normal the reference uses (occu_cs.py:148): support greater_than(low)."""

    def __init__(self, base_dist, low=None, high=None):
        assert high is None and low is not None and isinstance(base_dist, Normal)
        self.base_dist, self.low = base_dist, low
        self.support = _Constraint("greater_than", low)

    def log_prob(self, value):
        return self.base_dist.log_prob(value) - self.base_dist.log_sf(self.low)


class Bernoulli(Distribution):
    support = boolean
    has_enumerate_support = True

    def __init__(self, probs=None, logits=None):
        assert logits is None, "the reference only uses Bernoulli(probs)"
        self.probs = probs

    @property
    def batch_shape(self):
        return np.shape(self.probs)

    def log_prob(self, value):
        ps = _clamp_probs(self.probs)
        return _xlogy(value, ps) + _xlog1py(1 - value, -ps)

    def enumerate_support(self):
        return np.arange(2)


class Poisson(Distribution):
    support = nonnegative_integer

    def __init__(self, rate):
        self.rate = rate

    @property
    def batch_shape(self):
        return np.shape(self.rate)

    def log_prob(self, value):
        with np.errstate(divide="ignore", invalid="ignore"):
            return (np.log(self.rate) * value) - _sp.gammaln(value + 1) - self.rate


class Binomial(Distribution):
    support = nonnegative_integer  # integer_interval(0, n): feasible_like -> 0

    def __init__(self, total_count=1, probs=None):
        self.total_count, self.probs = total_count, probs

    def log_prob(self, value):
        n = self.total_count
        ps = _clamp_probs(self.probs)
        with np.errstate(divide="ignore", invalid="ignore"):
            return (_sp.gammaln(n + 1) - _sp.gammaln(value + 1) - _sp.gammaln(n - value + 1)
                    + _xlogy(value, ps) + _xlog1py(n - value, -ps))


class Categorical(Distribution):
    support = nonnegative_integer
    has_enumerate_support = True

    def __init__(self, probs=None, logits=None):
        assert probs is None, "the reference only uses Categorical(logits)"
        self.logits = np.asarray(logits)

    @property
    def batch_shape(self):
        return self.logits.shape[:-1]

    def log_prob(self, value):
        value = np.asarray(value)
        batch_shape = np.broadcast_shapes(value.shape, self.batch_shape)
        v = np.broadcast_to(value[..., None], batch_shape + (1,))
        log_pmf = self.logits - _logsumexp(self.logits, axis=-1, keepdims=True)
        log_pmf = np.broadcast_to(log_pmf, batch_shape + log_pmf.shape[-1:])
        return np.take_along_axis(log_pmf, v, axis=-1)[..., 0]

    def enumerate_support(self):
        return np.arange(self.logits.shape[-1])


def _biject(support, x):
    """numpyro.distributions.transforms.biject_to(support)(x) and log|det J| (elementwise)."""
    k = support.kind
    if k == "real":
        return x, 0.0
    if k == "positive":
        return np.exp(x), x
    if k == "greater_than":
        return np.exp(x) + support.low, x
    if k == "unit_interval":
        fi = np.finfo(_Config.float_dtype)
        y = _clip(_sigmoid(x), fi.tiny, 1.0 - fi.eps)  # SigmoidTransform._clipped_expit
        return y, -_softplus(x) - _softplus(-x)
    raise NotImplementedError(k)


# ------------------------------------------------------------------------------------------------
# numpyro primitives: a single tracing runtime
# ------------------------------------------------------------------------------------------------
class _Trace:
    def __init__(self, params, max_plate_nesting):
        self.params = params
        self.max_plate_nesting = max_plate_nesting
        self.plates = []   # active (name, size, dim)
        self.masks = []    # active mask arrays
        self.sites = OrderedDict()
        self.n_enum = 0
        self.enum_ordinal = {}  # enum dim -> frozenset of plate dims of the site that introduced it
        self.deterministic = OrderedDict()


_ACTIVE: list[_Trace] = []


def _rt() -> _Trace:
    if not _ACTIVE:
        raise RuntimeError("numpyro primitive called outside refshim.trace()")
    return _ACTIVE[-1]


class plate:  # noqa: N801 - numpyro's name
    def __init__(self, name, size, subsample_size=None, dim=None):
        assert dim is not None and dim < 0, "the reference always passes dim"
        self.name, self.size, self.dim = name, int(size), int(dim)

    def __enter__(self):
        _rt().plates.append((self.name, self.size, self.dim))
        return np.arange(self.size)

    def __exit__(self, *exc):
        _rt().plates.pop()
        return False


@contextlib.contextmanager
def _mask_handler(mask=None):
    _rt().masks.append(np.asarray(mask))
    try:
        yield
    finally:
        _rt().masks.pop()


def _expand_to_plates(lp, plates):
    lp = np.asarray(lp)
    nd = max([lp.ndim] + [-d for _, _, d in plates])
    shape = [1] * (nd - lp.ndim) + list(lp.shape)
    for _, size, d in plates:
        if shape[d] == 1:
            shape[d] = size
        else:
            assert shape[d] == size, f"plate size mismatch at dim {d}: {shape[d]} != {size}"
    return np.broadcast_to(lp, shape)


def sample(name, fn, obs=None, rng_key=None, sample_shape=(), infer=None, obs_mask=None):
    t = _rt()
    assert name not in t.sites, f"duplicate site {name}"
    plates = list(t.plates)
    ordinal = frozenset(d for _, _, d in plates)
    site = dict(type="sample", name=name, fn=fn, plates=plates, ordinal=ordinal, observed=obs is not None,
                enumerated=False)
    ladj = 0.0
    if obs is not None:
        value = obs
    elif infer and infer.get("enumerate") == "parallel":
        assert fn.has_enumerate_support
        dim = -(t.max_plate_nesting + 1 + t.n_enum)
        t.n_enum += 1
        sup = fn.enumerate_support()
        value = sup.reshape((-1,) + (1,) * (-dim - 1))
        t.enum_ordinal[dim] = ordinal
        site.update(enumerated=True, enum_dim=dim)
    else:
        if name not in t.params:
            raise KeyError(f"no value supplied for latent site '{name}'")
        x = np.asarray(t.params[name])
        value, ladj = _biject(fn.support, x)
        if fn.event_dim and np.ndim(ladj):
            ladj = np.sum(ladj, axis=tuple(range(-fn.event_dim, 0)))
    if t.masks:
        m = t.masks[0]
        for mm in t.masks[1:]:
            m = m & mm
        feasible = np.where(m, value, 0)  # feasible_like() is 0 for every support on this path
        lp = np.where(m, fn.log_prob(feasible), 0.0)
    else:
        lp = fn.log_prob(value)
    lp = _expand_to_plates(lp + ladj, plates)
    site.update(value=value, log_prob=lp)
    t.sites[name] = site
    return value


def deterministic(name, value):
    _rt().deterministic[name] = value
    return value


def factor(name, log_factor):
    t = _rt()
    plates = list(t.plates)
    lp = _expand_to_plates(log_factor, plates)
    t.sites[name] = dict(type="factor", name=name, plates=plates, ordinal=frozenset(d for _, _, d in plates),
                         observed=True, enumerated=False, log_prob=lp, value=None)


def _enum_dims(lp, mpn):
    return [d for d in range(-lp.ndim, -mpn) if lp.shape[d] > 1]


def _sum_product(factors, enum_ordinal, mpn):
    """Plate-aware sum-product over tree-structured plates (what funsor.sum_product does for numpyro's
    enumerated log-density): combine the deepest ordinal, logsumexp the enumerated dims that live there,
    sum out the plates the parent ordinal does not have, repeat."""
    total = 0.0
    factors = list(factors)
    while True:
        rest = []
        for lp, pl in factors:
            if _enum_dims(lp, mpn):
                rest.append((lp, pl))
            else:
                total = total + lp.sum()
        if not rest:
            return total
        leaf = max((pl for _, pl in rest), key=len)
        group = [lp for lp, pl in rest if pl == leaf]
        others = [(lp, pl) for lp, pl in rest if pl != leaf]
        for _, pl in others:
            assert pl < leaf or not (pl & leaf) or True
        s = group[0]
        for g in group[1:]:
            s = s + g
        elim = tuple(d for d in _enum_dims(s, mpn) if enum_ordinal[d] == leaf)
        if elim:
            s = _logsumexp(s, axis=elim, keepdims=True)
        remaining = _enum_dims(s, mpn)
        parent = frozenset().union(*[enum_ordinal[d] for d in remaining]) if remaining else frozenset()
        assert parent <= leaf and (parent < leaf or elim), "sum-product made no progress"
        drop = tuple(sorted(leaf - parent))
        if drop:
            s = s.sum(axis=drop, keepdims=True)
        factors = others + [(s, parent)]


class TraceResult:
    def __init__(self, trace: _Trace):
        self.sites = trace.sites
        self.deterministic = trace.deterministic
        self._t = trace

    def log_density(self, include=None):
        """Enumerated log joint (= -potential_energy in unconstrained space); `include` filters site names."""
        t = self._t
        facs = [(s["log_prob"], s["ordinal"]) for n, s in t.sites.items() if include is None or include(n, s)]
        return _sum_product(facs, t.enum_ordinal, t.max_plate_nesting)

    @staticmethod
    def is_prior(name, site):
        return site["type"] == "sample" and not site["observed"] and not site["enumerated"]

    def log_prior(self):
        return self.log_density(include=self.is_prior)

    def log_likelihood_marginal(self):
        return self.log_density(include=lambda n, s: not self.is_prior(n, s))


def trace(model, params, *args, max_plate_nesting=4, **kwargs) -> TraceResult:
    """Run `model(*args, **kwargs)` with latent (non-enumerated) sites substituted from `params`
    (UNCONSTRAINED values, numpyro shapes) and every site's log_prob recorded."""
    t = _Trace(params, max_plate_nesting)
    _ACTIVE.append(t)
    try:
        model(*args, **kwargs)
    finally:
        _ACTIVE.pop()
    return TraceResult(t)


# ------------------------------------------------------------------------------------------------
# module installation
# ------------------------------------------------------------------------------------------------
class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _pkg(name, **attrs):
    m = _StubModule(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _build_modules():
    jnp = _make_jnp()
    jax = _pkg("jax", numpy=jnp, local_device_count=lambda: 1)
    jax.nn = _pkg("jax.nn", sigmoid=_sigmoid, softplus=_softplus)
    jsp_special = _pkg("jax.scipy.special", logsumexp=_logsumexp, gammaln=_gammaln, xlogy=_xlogy,
                       xlog1py=_xlog1py, expit=_sigmoid)
    jax.scipy = _pkg("jax.scipy", special=jsp_special)
    jax.errors = _pkg("jax.errors", ConcretizationTypeError=ConcretizationTypeError)
    jax.random = _pkg("jax.random", PRNGKey=lambda seed: int(seed))
    dist = _pkg(
        "numpyro.distributions", Distribution=Distribution, Normal=Normal, HalfNormal=HalfNormal, Beta=Beta,
        Exponential=Exponential, Gamma=Gamma, TruncatedDistribution=TruncatedDistribution, Bernoulli=Bernoulli,
        Poisson=Poisson, Binomial=Binomial, Categorical=Categorical,
    )
    handlers = _pkg("numpyro.handlers", mask=_mask_handler)
    numpyro = _pkg("numpyro", sample=sample, plate=plate, deterministic=deterministic, factor=factor,
                   distributions=dist, handlers=handlers)
    return {
        "jax": jax, "jax.numpy": jnp, "jax.nn": jax.nn, "jax.scipy": jax.scipy,
        "jax.scipy.special": jsp_special, "jax.errors": jax.errors, "jax.random": jax.random,
        "numpyro": numpyro, "numpyro.distributions": dist, "numpyro.handlers": handlers,
    }


_INSTALLED = {}


def install():
    """Register the stand-ins in sys.modules (anything else under jax/numpyro/funsor/... is a MagicMock)."""
    if _INSTALLED:
        return _INSTALLED
    for root in _STUB_ROOTS:
        if root in sys.modules and not isinstance(sys.modules[root], _StubModule):
            raise RuntimeError(f"a real '{root}' is already imported; refshim is for boxes without it")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    mods = _build_modules()
    sys.modules.update(mods)
    _INSTALLED.update(mods)
    return _INSTALLED


def uninstall():
    """Drop every stand-in and every reference module again (scipy probes sys.modules['jax'])."""
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _StubFinder)]
    for name in list(sys.modules):
        root = name.split(".")[0]
        if root in _STUB_ROOTS and isinstance(sys.modules[name], _StubModule):
            del sys.modules[name]
        elif root == "biolith" and getattr(sys.modules[name], "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[name]
    _INSTALLED.clear()


def import_reference(reference_root=REFERENCE_ROOT):
    """`import biolith.models` from the read-only reference checkout with the stand-ins installed."""
    install()
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import biolith.models as models  # noqa: E402

    return models


# ------------------------------------------------------------------------------------------------
# theta (this repo's flat layout, oracle/occupancy.py docstring) -> numpyro's unconstrained site dict
# ------------------------------------------------------------------------------------------------
def theta_to_params(model: str, theta, Ks: int, Ko: int, *, fp_constant=False, fp_unoccupied=False,
                    site_random_effects=False, obs_random_effects=False, dims=None):
    th = np.asarray(theta)
    i = Ks + Ko + 2
    p = {"beta": th[None, : Ks + 1], "alpha": th[None, Ks + 1 : i]}
    if model == "occu_cs":
        for n in ("mu0", "mu1", "sigma0", "sigma1"):
            p[n] = th[i]; i += 1
    elif model == "occu_cop":
        if fp_constant:
            p["rate_fp_constant"] = th[i]; i += 1
        if fp_unoccupied:
            p["rate_fp_unoccupied"] = th[i]; i += 1
    else:
        if fp_constant:
            p["prob_fp_constant"] = th[i]; i += 1
        if fp_unoccupied:
            p["prob_fp_unoccupied"] = th[i]; i += 1
    if site_random_effects or obs_random_effects:
        S, P, J = dims
        occ_name = "site_re_abu" if model in ("occu_rn", "nmixture") else "site_re_occ"
        if site_random_effects:
            p["site_re_sd"] = th[i]; i += 1
        if obs_random_effects:
            p["obs_re_sd"] = th[i]; i += 1
        if site_random_effects:
            p[occ_name] = th[i : i + S].reshape(S, 1); i += S
            p["site_re_det"] = th[i : i + S].reshape(S, 1); i += S
        if obs_random_effects:  # oracle layout (S, P, J) -> numpyro plate layout (J, P, S, Sp)
            p["obs_re"] = th[i : i + S * P * J].reshape(S, P, J).transpose(2, 1, 0)[..., None]; i += S * P * J
    assert i == th.size, f"theta has {th.size} entries, consumed {i}"
    return p


def model_kwargs_for(model: str, *, fp_constant=False, fp_unoccupied=False, max_abundance=None,
                     site_random_effects=False, obs_random_effects=False):
    kw = {}
    if fp_constant:
        kw["false_positives_constant"] = True
    if fp_unoccupied:
        kw["false_positives_unoccupied"] = True
    if max_abundance is not None:
        kw["max_abundance"] = int(max_abundance)
    if site_random_effects:
        kw["site_random_effects"] = True
    if obs_random_effects:
        kw["obs_random_effects"] = True
    return kw


def value_and_grad(model_fn, model: str, theta, data: dict, *, h=1e-30, **flags):
    """(log joint, log likelihood, grad joint, grad likelihood) of the executed reference body at theta.

    log joint = -potential_energy of numpyro (priors + Jacobians + enumerated likelihood); "likelihood" drops
    the latent sample sites (what a numpyro.factor drop-in has to supply)."""
    theta = np.asarray(theta, np.float64)
    X, W, y = data["site_covs"], data["obs_covs"], data["obs"]
    Ks, Ko = X.shape[1], W.shape[3]
    dims = (X.shape[0], W.shape[1], W.shape[2])
    pflags = {k: flags.get(k, False) for k in ("fp_constant", "fp_unoccupied", "site_random_effects",
                                               "obs_random_effects")}
    kw = model_kwargs_for(model, max_abundance=flags.get("max_abundance"), **pflags)
    fd = _Config.float_dtype
    args = dict(site_covs=np.asarray(X, fd), obs_covs=np.asarray(W, fd), obs=np.asarray(y, fd))
    if data.get("session_duration") is not None:
        args["session_duration"] = np.asarray(data["session_duration"], fd)

    def run(th):
        tr = trace(model_fn, theta_to_params(model, th, Ks, Ko, dims=dims, **pflags), **args, **kw)
        return tr.log_density(), tr.log_likelihood_marginal()

    lj, ll = run(theta.astype(fd))
    gj, gl = np.zeros(theta.size), np.zeros(theta.size)
    if h:
        for k in range(theta.size):
            th = theta.astype(np.complex128)
            th[k] += 1j * h
            a, b = run(th)
            gj[k], gl[k] = np.imag(a) / h, np.imag(b) / h
    return float(np.real(lj)), float(np.real(ll)), gj, gl
