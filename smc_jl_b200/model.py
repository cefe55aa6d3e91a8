"""Host-side model description: parameters, priors and likelihood descriptors.

Mirrors what `smc(loglikelihood, parameters, data; ...)` receives in the reference
(src/smc_main.jl:118-161): a `ParameterVector` built with `ModelConstructors.parameter(...)`
(examples/regression_model/estimate_regression.jl:9-10, test/modelsetup.jl:14-30) and a
log-likelihood.  An arbitrary host closure cannot run inside a CUDA kernel, so the
log-likelihood is given as a *descriptor* naming one of the device likelihood families and
carrying its data; anything else raises (there is no CPU fallback by design).
"""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# prior kinds (values are part of the C ABI, include/smcb200.h)
PRIOR_NORMAL, PRIOR_UNIFORM, PRIOR_GAMMA, PRIOR_ROOT_INV_GAMMA, PRIOR_BETA, PRIOR_INV_GAMMA = range(6)
LIK_NONE, LIK_GAUSSREG, LIK_AS_DSGE = 0, 1, 2


@dataclass(frozen=True)
class Normal:
    mu: float
    sigma: float
    kind = PRIOR_NORMAL

    def params(self):
        return float(self.mu), float(self.sigma)


@dataclass(frozen=True)
class Uniform:
    a: float
    b: float
    kind = PRIOR_UNIFORM

    def params(self):
        return float(self.a), float(self.b)


@dataclass(frozen=True)
class Gamma:
    """Gamma(shape alpha, scale theta) as in Distributions.jl."""
    alpha: float
    theta: float
    kind = PRIOR_GAMMA

    def params(self):
        return float(self.alpha), float(self.theta)


@dataclass(frozen=True)
class RootInverseGamma:
    """ModelConstructors.RootInverseGamma(nu, tau)."""
    nu: float
    tau: float
    kind = PRIOR_ROOT_INV_GAMMA

    def params(self):
        return float(self.nu), float(self.tau)


@dataclass(frozen=True)
class Beta:
    a: float
    b: float
    kind = PRIOR_BETA

    def params(self):
        return float(self.a), float(self.b)


@dataclass(frozen=True)
class InverseGamma:
    """InverseGamma(shape, scale) as in Distributions.jl."""
    a: float
    b: float
    kind = PRIOR_INV_GAMMA

    def params(self):
        return float(self.a), float(self.b)


@dataclass
class Parameter:
    """One entry of the ParameterVector (ModelConstructors.parameter)."""
    key: str
    value: float
    valuebounds: Tuple[float, float] = (-1e5, 1e5)
    prior: Optional[object] = None
    fixed: bool = False


def parameter(key, value, valuebounds=(-1e5, 1e5), transform_parameterization=None, transform=None,
              prior=None, fixed=False):
    """Same argument order as ModelConstructors.parameter(key, value, valuebounds,
    transform_parameterization, transform, prior; fixed).  Transforms are irrelevant to SMC
    (it works in model space) and are ignored."""
    if fixed:
        valuebounds = (value, value)
    if prior is None and not fixed:
        raise ValueError("free parameter %s needs a prior" % key)
    return Parameter(str(key), float(value), (float(valuebounds[0]), float(valuebounds[1])), prior, bool(fixed))


# ------------------------------------------------------------------------------------------------
# likelihood descriptors
# ------------------------------------------------------------------------------------------------
@dataclass
class GaussRegLogLik:
    """Gaussian regression family in centred sufficient-statistic form.

    Equation e has coefficients theta[coef_off + e*stride + (0..k-1)] and, when sig_off >= 0,
    standard deviation sigma_e = theta[sig_off + e*stride]:

        ll = sum_e  -T_e/2 log(2 pi) - T_e log(sigma_e)
                    - qscale_e/2 * (rss_e + |U_e (b_e - bhat_e)|^2) / sigma_e^2

    eqdata holds, per equation, [T, qscale, rss, sigma_fixed, bhat(k), U(k*k row-major, upper)].
    """
    neq: int
    k: int
    stride: int
    coef_off: int
    sig_off: int
    eqdata: np.ndarray
    n_para: int
    kind: int = LIK_GAUSSREG

    def iparams(self):
        return np.array([self.neq, self.k, self.stride, self.coef_off, self.sig_off], dtype=np.int32)


def _suffstats(y, Z):
    """Centred sufficient statistics of y ~ Z b: (bhat, U upper with U'U = Z'Z, rss)."""
    y = np.asarray(y, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    Q, R = np.linalg.qr(Z)
    sgn = np.sign(np.diag(R))
    sgn[sgn == 0] = 1.0
    R = R * sgn[:, None]
    Q = Q * sgn[None, :]
    bhat = np.linalg.solve(R, Q.T @ y)
    res = y - Z @ bhat
    return bhat, np.triu(R), float(res @ res)


def _pack_eq(T, qscale, rss, sigma_fixed, bhat, U):
    return np.concatenate([[float(T), float(qscale), float(rss), float(sigma_fixed)], bhat, U.ravel()])


def LinearGaussianLogLik(y, X, sigma2=1.0):
    """y = X beta + eps, eps ~ N(0, sigma2) with sigma2 known; parameters = beta (k = X.shape[1]).
    Generalises examples/regression_model/estimate_regression.jl:46-53 (there X = [1, x])."""
    X = np.asarray(X, dtype=np.float64)
    T, k = X.shape
    bhat, U, rss = _suffstats(y, X)
    eq = _pack_eq(T, 1.0, rss, np.sqrt(sigma2), bhat, U)
    return GaussRegLogLik(1, k, k, 0, -1, eq, k)


def LinearEquationsLogLik(data, X):
    """test/modelsetup.jl:119-138: n_eq equations y_it = alpha_i + beta_i x_it + e_it, e ~ N(0, sigma_i^2),
    parameters ordered (alpha_i, beta_i, sigma_i).  data, X are n_eq x T (X may have more columns)."""
    data = np.asarray(data, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    neq, T = data.shape
    eqs = []
    for i in range(neq):
        Z = np.column_stack([np.ones(T), X[i, :T]])
        bhat, U, rss = _suffstats(data[i], Z)
        eqs.append(_pack_eq(T, 1.0, rss, 0.0, bhat, U))
    return GaussRegLogLik(neq, 2, 3, 0, 2, np.concatenate(eqs), 3 * neq)


def CAPMLogLik(lik_data, market_data, as_written=False):
    """examples/capm_model/estimate_capm.jl:48-69.  as_written=False is the per-period likelihood the
    example intends (identical in form to test/modelsetup.jl:119-138 with the market return as the
    regressor of every asset).  as_written=True reproduces the example's code literally: beta_i aliases
    alpha_i (`p[i*3-2]`, :57) and the whole-sample quadratic form is added once per period (:65-67)."""
    lik_data = np.asarray(lik_data, dtype=np.float64)
    market = np.asarray(market_data, dtype=np.float64).reshape(-1)
    neq, T = lik_data.shape
    if not as_written:
        return LinearEquationsLogLik(lik_data, np.tile(market[:T], (neq, 1)))
    eqs = []
    for i in range(neq):
        Z = (1.0 + market[:T]).reshape(T, 1)
        bhat, U, rss = _suffstats(lik_data[i], Z)
        eqs.append(_pack_eq(T, float(T), rss, 0.0, bhat, U))
    return GaussRegLogLik(neq, 1, 3, 0, 2, np.concatenate(eqs), 3 * neq)


@dataclass
class AnSchorfheideLogLik:
    """Log-likelihood of the three-equation An-Schorfheide DSGE model (BASELINE config C4):
    examples/dsge_models/small_dsge_model.jl:35-50 -> `DSGE.likelihood(m, data; sampler=false,
    catch_errors=true, use_chand_recursion=true)`.  `data` is 3 x T (gdp growth, inflation, nominal rate),
    the first `n_presample` periods are filtered but not scored.  The ParameterVector must be the model's
    16 parameters (tau, kappa, psi_1, psi_2, rA, pi_star, gamma_Q, rho_R, rho_g, rho_z, sigma_R, sigma_g,
    sigma_z, e_y, e_pi, e_R).  NaN entries are missing observations (dropped from that period's update, as DSGE.jl's filter
    does).  Solved and filtered on the device (csrc/aslik.cuh)."""
    data: np.ndarray
    n_presample: int = 2
    n_para: int = 16
    kind: int = LIK_AS_DSGE

    def __post_init__(self):
        self.data = np.ascontiguousarray(np.asarray(self.data, dtype=np.float64))
        if self.data.ndim != 2 or self.data.shape[0] != 3 or self.data.shape[1] < 1:
            raise ValueError("An-Schorfheide data must be 3 x T")
        if np.any(np.isinf(self.data)):
            raise ValueError("infinite observations")        # NaN = missing: the filter drops that series for the period

    def iparams(self):
        return np.array([self.data.shape[1], self.n_presample], dtype=np.int32)

    @property
    def eqdata(self):
        """column-major 3 x T (the Julia matrix bytes)"""
        return np.ascontiguousarray(self.data.T).ravel()


# ------------------------------------------------------------------------------------------------
# flat model spec handed to the C ABI (and, in tests, to the oracle)
# ------------------------------------------------------------------------------------------------
@dataclass
class ModelSpec:
    d: int
    fixed: np.ndarray
    lo: np.ndarray
    hi: np.ndarray
    kind: np.ndarray
    p1: np.ndarray
    p2: np.ndarray
    values: np.ndarray
    keys: List[str]
    liks: List[Optional[object]] = field(default_factory=lambda: [None, None])

    @property
    def free_inds(self):
        return np.flatnonzero(self.fixed == 0)

    @property
    def n_free(self):
        return int((self.fixed == 0).sum())


def make_spec(parameters: Sequence[Parameter], loglikelihood=None, old_loglikelihood=None) -> ModelSpec:
    d = len(parameters)
    fixed = np.zeros(d, np.int32)
    lo = np.zeros(d); hi = np.zeros(d); p1 = np.zeros(d); p2 = np.ones(d)
    kind = np.zeros(d, np.int32)
    values = np.zeros(d)
    for i, p in enumerate(parameters):
        fixed[i] = 1 if p.fixed else 0
        lo[i], hi[i] = p.valuebounds
        values[i] = p.value
        if p.prior is not None:
            kind[i] = p.prior.kind
            p1[i], p2[i] = p.prior.params()
    for lk in (loglikelihood, old_loglikelihood):
        if lk is not None:
            if not isinstance(lk, (GaussRegLogLik, AnSchorfheideLogLik)):
                raise TypeError("loglikelihood must be a device likelihood descriptor (e.g. LinearGaussianLogLik); "
                                "arbitrary host callables cannot run on the GPU and there is no CPU fallback")
            if lk.n_para != d:
                raise ValueError("likelihood expects %d parameters, ParameterVector has %d" % (lk.n_para, d))
    return ModelSpec(d, fixed, lo, hi, kind, p1, p2, values, [p.key for p in parameters],
                     [loglikelihood, old_loglikelihood])
