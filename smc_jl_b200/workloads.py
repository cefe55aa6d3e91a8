"""Synthetic workloads of BASELINE.json's configs + host-side prior sampling.

Host-side sampling here is stage-0 set-up (src/initialization.jl:23-63 draws from the prior once);
the per-stage hot path never touches it.
"""
import numpy as np

from . import model as M


def linear_gaussian(d=20, T=256, seed_x=20260001, seed_e=20260002, prior_sd=10.0):
    """Config C2: d free betas ~ N(0, prior_sd), sigma^2 = 1 known, X (T x d) with a constant first
    column and iid N(0,1) others, beta_true_k = k/10, y = X beta_true + eps."""
    rx = np.random.Generator(np.random.Philox(seed_x))
    re = np.random.Generator(np.random.Philox(seed_e))
    X = rx.standard_normal((T, d))
    X[:, 0] = 1.0
    beta_true = np.arange(1, d + 1) / 10.0
    y = X @ beta_true + re.standard_normal(T)
    params = [M.parameter("β%d" % (k + 1), 0.0, (-1e5, 1e5), (-1e5, 1e5), None, M.Normal(0.0, prior_sd))
              for k in range(d)]
    return params, M.LinearGaussianLogLik(y, X, 1.0), (y, X, beta_true)


def regression_example(T=100, seed=1793):
    """Config C1: examples/regression_model/estimate_regression.jl:9-53 (y = 1 + x exactly, sigma^2 = 1)."""
    rng = np.random.Generator(np.random.Philox(seed))
    x = rng.uniform(size=T)
    y = 1.0 + 1.0 * x
    X = np.column_stack([np.ones(T), x])
    params = [M.parameter("α1", 0.0, (-1e5, 1e5), (-1e5, 1e5), None, M.Normal(0, 10)),
              M.parameter("β1", 0.0, (-1e5, 1e5), (-1e5, 1e5), None, M.Normal(0, 10))]
    return params, M.LinearGaussianLogLik(y, X, 1.0), (y, X)


def three_equation_parameters(prior_para=1e3):
    """test/modelsetup.jl:9-30 and examples/capm_model/estimate_capm.jl:16-33 (same ParameterVector)."""
    ps = []
    for i in (1, 2, 3):
        ps.append(M.parameter("α%d" % i, 0.0, (-1e5, 1e5), (-1e5, 1e5), None, M.Normal(0, prior_para)))
        ps.append(M.parameter("β%d" % i, 0.0, (-1e5, 1e5), (-1e5, 1e5), None, M.Normal(0, prior_para)))
        ps.append(M.parameter("σ%d" % i, 1.0, (1e-5, 1e5), (1e-5, 1e5), None, M.Uniform(0, prior_para)))
    return ps


def synthetic_three_equation(T=100, seed=1793):
    """Same generating process as test/modelsetup.jl:69-80 (data = beta .* X .+ alpha .+ err)."""
    rng = np.random.Generator(np.random.Philox(seed))
    X = rng.standard_normal((3, T))
    err = rng.standard_normal((3, T))
    a = np.arange(1.0, 4.0)[:, None]
    return a * X + a + err, X


def an_schorfheide_parameters():
    """Config C4: the An-Schorfheide ParameterVector stored in the reference's fixtures
    (test/reference/one_draw_in.jld2): 13 free parameters, 3 fixed measurement errors."""
    big, rho_hi = (1e-20, 1e5), (1e-20, 0.9999999)
    spec = [("τ", 1.6735264339099827, big, M.Gamma(16.0, 0.125)), ("κ", 0.017686826714964354, (1e-20, 10.0), M.Uniform(0.0, 1.0)),
            ("ψ_1", 1.405751678057145, big, M.Gamma(36.0, 1.0 / 24.0)), ("ψ_2", 0.21072315533472247, big, M.Gamma(4.0, 0.125)),
            ("rA", 0.09536887782738207, big, M.Gamma(1.0, 0.5)), ("π_star", 2.536422981325555, big, M.Gamma(12.25, 0.5714285714285714)),
            ("γ_Q", 0.6164762411216859, big, M.Normal(0.4, 0.2)), ("ρ_R", 0.1600060971420918, rho_hi, M.Uniform(0.0, 1.0)),
            ("ρ_g", 0.4229562655081418, rho_hi, M.Uniform(0.0, 1.0)), ("ρ_z", 0.602297580266383, rho_hi, M.Uniform(0.0, 1.0)),
            ("σ_R", 0.40411072598543146, big, M.RootInverseGamma(4.0, 0.4)), ("σ_g", 1.3307835750296042, big, M.RootInverseGamma(4.0, 1.0)),
            ("σ_z", 0.22539373793338302, big, M.RootInverseGamma(4.0, 0.5))]
    ps = [M.parameter(k, v, b, b, None, pr) for k, v, b, pr in spec]
    for k, v in (("e_y", 0.1159846), ("e_π", 0.2941664), ("e_R", 0.4475874)):
        ps.append(M.parameter(k, v, (v, v), (v, v), None, None, fixed=True))
    return ps


def prior_draw(parameters, n, rng):
    """rand(parameters, n) (src/initialization.jl:27): every free parameter from its prior, redrawn until
    inside its valuebounds; fixed parameters keep their value.  Returns n x n_para."""
    out = np.zeros((n, len(parameters)))
    for k, p in enumerate(parameters):
        if p.fixed:
            out[:, k] = p.value
            continue
        lo, hi = p.valuebounds
        todo = np.arange(n)
        while todo.size:
            m = todo.size
            pr = p.prior
            if isinstance(pr, M.Normal):
                x = rng.normal(pr.mu, pr.sigma, m)
            elif isinstance(pr, M.Uniform):
                x = rng.uniform(pr.a, pr.b, m)
            elif isinstance(pr, M.Gamma):
                x = rng.gamma(pr.alpha, pr.theta, m)
            elif isinstance(pr, M.Beta):
                x = rng.beta(pr.a, pr.b, m)
            elif isinstance(pr, M.InverseGamma):
                x = pr.b / rng.gamma(pr.a, 1.0, m)
            elif isinstance(pr, M.RootInverseGamma):
                # x^2 ~ InverseGamma(nu/2, nu tau^2 / 2)
                x = np.sqrt((pr.nu * pr.tau ** 2 / 2.0) / rng.gamma(pr.nu / 2.0, 1.0, m))
            else:
                raise NotImplementedError(type(pr))
            out[todo, k] = x
            ok = (x > lo) & (x < hi)
            todo = todo[~ok]
    return out


def initial_cloud(parameters, n, rng):
    """n x (n_para+5) Fortran-ordered particle matrix: prior draws, old_loglh = 0, weight = 1
    (src/initialization.jl:111-118); loglh / logprior are filled by Engine.evaluate()."""
    d = len(parameters)
    P = np.zeros((n, d + 5), order="F")
    P[:, :d] = prior_draw(parameters, n, rng)
    P[:, d + 4] = 1.0
    return P
