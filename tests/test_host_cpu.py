"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host-side model
description is consistent with the reference's example likelihoods, and the oracle's stage loop is
self-consistent (so that the GPU parity tests compare against something meaningful)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from smc_jl_b200 import cloud as CL
from smc_jl_b200 import model as M
from smc_jl_b200 import workloads as W


def test_library_exports_every_declared_symbol():
    from smc_jl_b200 import _lib
    declared = _lib.declared_symbols()
    assert len(declared) >= 25
    assert set(declared) == set(_lib.BOUND_SYMBOLS)          # binding and header agree
    for name in declared:
        assert hasattr(_lib.lib, name)                        # dlsym succeeds
    assert _lib.lib.smcb200_abi_version() == 2
    assert _lib.lib.smcb200_status_string(4).decode().startswith("proposal covariance")


def test_no_cpu_fallback_without_gpu():
    import torch
    from smc_jl_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert _lib.lib.smcb200_create(C.byref(h), 0) == _lib.ERR_CUDA and not h
    from smc_jl_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)


def test_arbitrary_callable_is_rejected():
    params, lk, _ = W.regression_example()
    with pytest.raises(TypeError, match="no CPU fallback"):
        M.make_spec(params, lambda p, d: 0.0)


def test_sufficient_statistics_match_direct_likelihoods():
    L = O.lib()
    rng = np.random.default_rng(0)
    # linear-Gaussian, d = 20 (C2)
    params, lk, (y, X, _) = W.linear_gaussian(d=20, T=256)
    mod = O.Model(M.make_spec(params, lk))
    for _ in range(200):
        th = rng.normal(0, 3, 20)
        direct = L.orc_loglik_linreg_direct(th, np.ascontiguousarray(y), np.ascontiguousarray(X), 256, 20, 1.0)
        assert mod.loglik(th) == pytest.approx(direct, rel=1e-12)
    # CAPM as written (examples/capm_model/estimate_capm.jl:48-69), restated literally in numpy
    g = np.load(O.ROOT + "/tests/golden/capm_data.npz")
    lik_data, market = g["lik_data"], g["market_data"]
    lkw = M.CAPMLogLik(lik_data, market, as_written=True)
    mod = O.Model(M.make_spec(W.three_equation_parameters(), lkw))
    for _ in range(100):
        p = np.abs(rng.normal(1, 0.5, 9))
        alpha = p[[0, 3, 6]]; beta = p[[0, 3, 6]]; sig2 = p[[2, 5, 8]] ** 2
        term1 = -3 / 2 * np.log(2 * np.pi) - 0.5 * np.log(np.prod(sig2))
        errors = lik_data - alpha[:, None] - beta[:, None] * market
        want = sum(term1 - 0.5 * np.sum(errors * (errors / sig2[:, None])) for _ in range(lik_data.shape[1]))
        assert mod.loglik(p) == pytest.approx(want, rel=1e-11)
    # per-period form == test/modelsetup.jl form
    lkc = M.CAPMLogLik(lik_data, market, as_written=False)
    mod = O.Model(M.make_spec(W.three_equation_parameters(), lkc))
    Xm = np.tile(market.reshape(1, -1), (3, 1))
    for _ in range(100):
        p = np.abs(rng.normal(1, 0.5, 9))
        want = L.orc_loglik_lineq_direct(p, np.ascontiguousarray(lik_data.T).ravel(), np.ascontiguousarray(Xm.T).ravel(), 3,
                                         lik_data.shape[1])
        assert mod.loglik(p) == pytest.approx(want, rel=1e-11)


def test_cloud_container_layout():
    c = CL.Cloud.empty(3, 10)
    assert c.particles.shape == (10, 8) and c.particles.flags.f_contiguous
    assert (c.stage_index, c.n_Φ, c.resamples, c.c, c.accept) == (1, 0, 0, 0.0, 0.25)      # particle.jl:50-53
    c.particles[:] = np.arange(80).reshape(10, 8)
    assert np.array_equal(CL.get_loglh(c), c.particles[:, 3]) and np.array_equal(CL.get_weights(c), c.particles[:, 7])
    assert CL.get_vals(c).shape == (3, 10) and CL.get_vals(c, transpose=False).shape == (10, 3)
    CL.update_draws(c, np.ones((3, 10)))
    assert np.all(c.particles[:, :3] == 1.0)
    with pytest.raises(ValueError):
        CL.update_draws(c, np.ones((4, 4)))
    CL.reset_weights(c)
    assert np.all(CL.get_weights(c) == 1.0) and len(c) == 10


def test_canonical_orders_are_shard_invariant():
    """The canonical sums are binary trees over power-of-two aligned shards: combining per-shard results
    in rank order reproduces the single-shard result bit-for-bit (basis of multi-GPU invariance)."""
    L = O.lib()
    rng = np.random.default_rng(2)
    N = 1 << 15
    x = np.exp(rng.normal(0, 3, N))
    full = L.orc_canon_sum(x, N)
    for G in (2, 4, 8):
        parts = [L.orc_canon_sum(np.ascontiguousarray(x[g * N // G:(g + 1) * N // G]), N // G) for g in range(G)]
        while len(parts) > 1:
            parts = [parts[i] + parts[i + 1] for i in range(0, len(parts), 2)]
        assert parts[0] == full
    assert abs(full - np.sum(x)) <= 1e-12 * np.sum(x)
    c = np.zeros(N)
    L.orc_cumsum(x / full, N, c)
    np.testing.assert_allclose(c, np.cumsum(x / full), rtol=1e-13)
    assert abs(c[-1] - 1.0) < 1e-13


def test_oracle_stage_loop_recovers_posterior():
    """Oracle-only end-to-end check of config C1 (regression example, N = 1000): the weighted posterior
    mean reaches the OLS solution (alpha, beta) = (1, 1) of examples/regression_model."""
    params, lk, _ = W.regression_example()
    spec = M.make_spec(params, lk)
    mod = O.Model(spec)
    N = 1000
    P = W.initial_cloud(params, N, np.random.default_rng(0))
    buf = O.cloud_f(P)
    O.lib().orc_evaluate(mod.h, buf, N)
    scratch = np.zeros_like(buf)
    sched = (np.arange(100) / 99.0) ** 2.1
    io = O.StageIO(threshold_ratio=0.5, target=0.25, alpha=1.0, tempering_target=0.95, n_mh_steps=1, n_blocks=1,
                   resample_method=0, adaptive=0, has_old=0, nthreads=0, seed=1, c=0.5, accept=0.25, ess_prev=float(N), j=2)
    for s in range(99):
        io.phi_n1, io.phi_n, io.stage = float(sched[s]), float(sched[s + 1]), s + 2
        assert O.lib().orc_stage(mod.h, buf, scratch, N, sched, 100, C.byref(io), None, None, None, None) == 0
    got = O.cloud_m(buf, N, 2)
    mean = np.average(got[:, :2], axis=0, weights=got[:, -1])
    assert np.allclose(mean, [1.0, 1.0], atol=0.1)
    assert 0.05 < io.accept < 0.8


def test_oracle_initial_draw_distributions():
    """initial_draw! (src/initialization.jl:88-119) in the oracle: every prior family's sampler has the right first two
    moments, draws respect valuebounds, fixed parameters keep their value, weights = 1 and old_loglh = 0."""
    rng = np.random.default_rng(0)
    X = rng.standard_normal((40, 7)); y = X @ np.full(7, 0.1) + rng.standard_normal(40)
    pri = [M.Normal(1.0, 2.0), M.Uniform(-1.0, 3.0), M.Gamma(2.5, 0.4), M.Gamma(0.6, 2.0), M.RootInverseGamma(6.0, 0.5),
           M.Beta(2.0, 5.0), M.InverseGamma(5.0, 2.0)]
    ps = [M.parameter("p%d" % k, 0.5, (-1e5, 1e5), (-1e5, 1e5), None, pr) for k, pr in enumerate(pri)]
    ps[1] = M.parameter("p1", 0.5, (0.0, 3.0), (0.0, 3.0), None, pri[1])       # truncates Uniform(-1, 3) to (0, 3)
    ps.append(M.parameter("fx", 0.25, (0.25, 0.25), (0.25, 0.25), None, None, fixed=True))
    X8 = np.column_stack([X, np.ones(40)])
    spec = M.make_spec(ps, M.LinearGaussianLogLik(y, X8, 1.0))
    N = 200_000
    buf = np.zeros(N * 13)
    mod = O.Model(spec)
    assert O.lib().orc_initial_draw(mod.h, buf, N, 0, np.ascontiguousarray(spec.values), 11, 10) == 0
    P = O.cloud_m(buf, N, 8)
    assert np.all(P[:, 7] == 0.25) and np.all(P[:, 12] == 1.0) and np.all(P[:, 10] == 0.0)
    assert np.all(np.isfinite(P[:, 8])) and np.all(np.isfinite(P[:, 9]))
    assert P[:, 1].min() > 0.0 and P[:, 1].max() < 3.0
    se = 5.0 / np.sqrt(N)
    def chk(col, mean, var):
        assert abs(P[:, col].mean() - mean) < se * np.sqrt(var) + 1e-12, (col, P[:, col].mean(), mean)
        assert abs(P[:, col].var() - var) < 0.05 * var, (col, P[:, col].var(), var)
    chk(0, 1.0, 4.0)
    chk(1, 1.5, 9.0 / 12.0)
    chk(2, 2.5 * 0.4, 2.5 * 0.16)
    chk(3, 0.6 * 2.0, 0.6 * 4.0)
    x2 = P[:, 4] ** 2                                    # x^2 ~ InverseGamma(nu/2 = 3, nu tau^2/2 = 0.75)
    assert abs(x2.mean() - 0.75 / 2.0) < 5 * np.sqrt((0.75 ** 2 / (4.0 * 1.0)) / N)
    chk(5, 2.0 / 7.0, 2.0 * 5.0 / (49.0 * 8.0))
    chk(6, 2.0 / 4.0, 4.0 / (16.0 * 3.0))
    # full distributions (Kolmogorov-Smirnov against scipy's reference CDFs; 200 000 draws => D_crit(1e-3) ~ 0.0044)
    from scipy import stats
    ks = lambda x, dist: stats.kstest(x, dist.cdf).statistic
    assert ks(P[:, 0], stats.norm(1.0, 2.0)) < 0.0044
    assert ks(P[:, 1], stats.uniform(0.0, 3.0)) < 0.0044                      # Uniform(-1, 3) truncated to its valuebounds (0, 3)
    assert ks(P[:, 2], stats.gamma(2.5, scale=0.4)) < 0.0044
    assert ks(P[:, 3], stats.gamma(0.6, scale=2.0)) < 0.0044                  # shape < 1: boosted Marsaglia-Tsang
    assert ks(P[:, 4] ** 2, stats.invgamma(3.0, scale=0.75)) < 0.0044         # RootInverseGamma(6, 0.5)
    assert ks(P[:, 5], stats.beta(2.0, 5.0)) < 0.0044
    assert ks(P[:, 6], stats.invgamma(5.0, scale=2.0)) < 0.0044
    # the row's loglh / logprior are the model's
    for r in (0, 17, N - 1):
        th = np.ascontiguousarray(P[r, :8])
        assert P[r, 8] == mod.loglik(th) and P[r, 9] == mod.logprior(th)
    # shard invariance: particles 1000.. drawn as their own shard are identical
    buf2 = np.zeros(500 * 13)
    O.lib().orc_initial_draw(mod.h, buf2, 500, 1000, np.ascontiguousarray(spec.values), 11, 10)
    assert np.array_equal(O.cloud_m(buf2, 500, 8), P[1000:1500])


def test_as_prototype_agrees_with_oracle(golden):
    """tools/as_reduced_prototype.py (numpy, generic matrix algebra) and oracle/as_model.c (fixed operation order,
    L D L' innovations) are two independent restatements of the same reduced An-Schorfheide solution + Kalman filter."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("as_proto", os.path.join(os.path.dirname(__file__), "..", "tools", "as_reduced_prototype.py"))
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    g = golden("as_clouds.npz")
    data = g["data"]
    flat = np.ascontiguousarray(data.T).ravel()
    for r in range(0, 600, 60):
        th = np.ascontiguousarray(g["cloud600"][r, :16])
        a = proto.loglik(th, data)
        b = O.lib().orc_as_loglik(th, flat, 230, 2)
        assert a == pytest.approx(b, rel=1e-11)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs next to ours): exactly one JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "particle_mh_steps_per_sec_per_stage" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_log_prior_families_against_scipy():
    """All six prior families' log-densities (ModelConstructors.prior) against scipy.stats: Normal / Uniform /
    Gamma / RootInverseGamma are also pinned by the reference's fixtures; Beta and InverseGamma only here."""
    from scipy import stats
    fams = [(M.Normal(0.4, 0.2), lambda x: stats.norm(0.4, 0.2).logpdf(x), (-1.0, 2.0)),
            (M.Uniform(0.0, 3.0), lambda x: stats.uniform(0.0, 3.0).logpdf(x), (0.1, 2.9)),
            (M.Gamma(2.5, 0.4), lambda x: stats.gamma(2.5, scale=0.4).logpdf(x), (0.01, 5.0)),
            (M.Beta(2.0, 5.0), lambda x: stats.beta(2.0, 5.0).logpdf(x), (0.01, 0.99)),
            (M.InverseGamma(5.0, 2.0), lambda x: stats.invgamma(5.0, scale=2.0).logpdf(x), (0.05, 4.0)),
            # RootInverseGamma(nu, tau): x^2 ~ InverseGamma(nu/2, nu tau^2/2)  =>  p(x) = 2 x p_IG(x^2)
            (M.RootInverseGamma(4.0, 0.5), lambda x: np.log(2 * x) + stats.invgamma(2.0, scale=0.5).logpdf(x * x), (0.05, 3.0))]
    rng = np.random.default_rng(1)
    for prior, ref, (lo, hi) in fams:
        ps = [M.parameter("p", 0.5, (-1e5, 1e5), (-1e5, 1e5), None, prior)]
        mod = O.Model(M.make_spec(ps))
        for x in rng.uniform(lo, hi, 200):
            assert mod.logprior(np.array([x])) == pytest.approx(float(ref(x)), rel=1e-12, abs=1e-12), (prior, x)
    # outside the support
    for prior, x in ((M.Uniform(0.0, 3.0), 3.5), (M.Gamma(2.5, 0.4), -0.1), (M.Beta(2.0, 5.0), 1.2), (M.InverseGamma(5.0, 2.0), -1.0),
                     (M.RootInverseGamma(4.0, 0.5), -0.3)):
        mod = O.Model(M.make_spec([M.parameter("p", 0.5, (-1e5, 1e5), (-1e5, 1e5), None, prior)]))
        assert mod.logprior(np.array([x])) == -np.inf


@pytest.mark.parametrize("alpha,n_blocks", [(1.0, 1), (0.5, 1), (0.7, 2)])
def test_mutation_leaves_the_posterior_invariant(alpha, n_blocks):
    """Metropolis-Hastings with the (mixture) proposal and its density correction q0 - q1 (mutation.jl:123,
    helpers.jl:128-164) must leave the target invariant: particles drawn from the exact Gaussian posterior of a linear
    model keep its mean and covariance after several mutation sweeps at phi = 1.  A wrong proposal-density ratio (the
    mixture is NOT symmetric for alpha < 1) shows up here as a shifted / shrunk cloud."""
    rng = np.random.default_rng(3)
    d, T, N = 3, 40, 40000
    X = rng.standard_normal((T, d)); X[:, 0] = 1.0
    y = X @ np.array([0.5, -0.3, 0.8]) + rng.standard_normal(T)
    prior_sd = 2.0
    ps = [M.parameter("b%d" % k, 0.0, (-1e5, 1e5), (-1e5, 1e5), None, M.Normal(0.0, prior_sd)) for k in range(d)]
    spec = M.make_spec(ps, M.LinearGaussianLogLik(y, X, 1.0))
    prec = X.T @ X + np.eye(d) / prior_sd ** 2
    cov_post = np.linalg.inv(prec)
    mean_post = cov_post @ (X.T @ y)
    P = np.zeros((N, d + 5), order="F")
    P[:, :d] = rng.multivariate_normal(mean_post, cov_post, N)
    P[:, d + 4] = 1.0
    mod = O.Model(spec)
    buf = O.cloud_f(P)
    L = O.lib()
    L.orc_evaluate(mod.h, buf, N)
    perm = np.arange(d, dtype=np.int32)
    sizes = np.array([d], np.int32) if n_blocks == 1 else np.array([2, 1], np.int32)
    st = C.c_int()
    # a deliberately mis-centred / mis-scaled proposal: the correction, not the proposal, must carry the target
    mu_prop = mean_post + np.array([0.3, -0.2, 0.1])
    cov_prop = np.ascontiguousarray(1.5 * cov_post)
    # alpha < 1 runs at c = 1: the reference evaluates the diagonal component's density with variance Sigma_ii where it
    # draws with c^2 Sigma_ii (helpers.jl:146 vs :93), so for c != 1 its q0 - q1 is not exactly the Hastings correction of
    # the mixture (a bias of ~3.5 standard errors at this N with c = 0.8, reproduced on purpose -- DESIGN.md); at c = 1 the
    # two coincide and the kernel must be exactly invariant
    cfac = 0.8 if alpha == 1.0 else 1.0
    pr = L.orc_proposal_create(d, d, np.ascontiguousarray(mu_prop), cov_prop, n_blocks, sizes, perm, perm, cfac, C.byref(st))
    assert pr and st.value == 0
    for sweep in range(6):
        L.orc_mutate(mod.h, pr, buf, N, 0, 1.0, 1.0, alpha, 2, d, 0, 100 + sweep, 2 + sweep, 0)
    L.orc_proposal_free(pr)
    Q = O.cloud_m(buf, N, d)
    se = np.sqrt(np.diag(cov_post) / N)
    assert np.all(np.abs(Q[:, :d].mean(axis=0) - mean_post) < 5 * se), (Q[:, :d].mean(axis=0), mean_post)
    np.testing.assert_allclose(np.cov(Q[:, :d].T), cov_post, rtol=0.06, atol=0.06 * np.sqrt(np.outer(np.diag(cov_post), np.diag(cov_post))).max())
    assert 0.05 < Q[:, d + 3].mean() / 2 < 0.9                     # accept column: sum over the 2 MH steps / n_free (per sweep)


def test_log_marginal_likelihood_from_weight_history():
    """End-to-end statistical check of the stage loop (correction -> selection -> mutation) in the oracle: the tempering
    estimate of the log marginal data density, sum_n log((1/N) sum_i W_{n-1,i} w_{n,i}) -- what the reference's w / W
    matrices exist for (smc_main.jl:363-366,419-420) -- reproduces the analytic marginal likelihood of a Bayesian linear
    regression."""
    rng = np.random.default_rng(5)
    d, T, N = 3, 30, 20000
    X = rng.standard_normal((T, d)); X[:, 0] = 1.0
    y = X @ np.array([0.4, -0.6, 0.9]) + rng.standard_normal(T)
    s0 = 1.5
    ps = [M.parameter("b%d" % k, 0.0, (-1e5, 1e5), (-1e5, 1e5), None, M.Normal(0.0, s0)) for k in range(d)]
    spec = M.make_spec(ps, M.LinearGaussianLogLik(y, X, 1.0))
    S = np.eye(T) + s0 ** 2 * X @ X.T
    exact = -0.5 * (T * np.log(2 * np.pi) + np.linalg.slogdet(S)[1] + y @ np.linalg.solve(S, y))
    mod = O.Model(spec)
    L = O.lib()
    buf = np.zeros(N * (d + 5))
    assert L.orc_initial_draw(mod.h, buf, N, 0, np.ascontiguousarray(spec.values), 17, 10) == 0
    scratch = np.zeros_like(buf)
    n_phi = 60
    sched = (np.arange(n_phi) / (n_phi - 1.0)) ** 2.1
    # alpha = 1 (symmetric random walk: exact Hastings ratio).  With alpha = 0.9 the same run is biased by about +0.04
    # (12 seeds: mean +0.042, std 0.010, independent of N; +0.003 at n_Phi = 300): the reference's diagonal-component
    # proposal density omits c^2 (helpers.jl:146), so its q0 - q1 is not exactly the Hastings correction of the mixture it
    # draws from -- a property of the reference that oracle and device reproduce on purpose.
    io = O.StageIO(threshold_ratio=0.5, target=0.25, alpha=1.0, tempering_target=0.95, n_mh_steps=2, n_blocks=1, resample_method=0,
                   nthreads=0, seed=17, c=0.5, accept=0.25, ess_prev=float(N), j=2)
    W_prev = np.ones(N)
    log_mdd = 0.0
    for s in range(n_phi - 1):
        io.phi_n1, io.phi_n, io.stage = float(sched[s]), float(sched[s + 1]), s + 2
        inc, nw = np.zeros(N), np.zeros(N)
        assert L.orc_stage(mod.h, buf, scratch, N, sched, n_phi, C.byref(io), inc.ctypes.data_as(C.c_void_p),
                           nw.ctypes.data_as(C.c_void_p), None, None) == 0
        log_mdd += np.log(np.mean(W_prev * inc))
        W_prev = nw                                     # reset to 1 by the stage when it resampled (smc_main.jl:445)
    assert log_mdd == pytest.approx(exact, abs=0.05), (log_mdd, exact)


def test_incremental_weights_all_prior_weight_branches():
    """The three incremental-weight formulas of src/smc_main.jl:401-410 (prior weight 0, 1 and in between) against a
    direct numpy restatement, followed by update_weights! / normalize_weights! / ESS (particle.jl:250-259,362-369,
    smc_main.jl:427)."""
    rng = np.random.default_rng(8)
    N, d = 3000, 2
    for pw, lpod in ((0.0, 0.0), (1.0, 0.0), (0.35, -3.2)):
        P = np.zeros((N, d + 5), order="F")
        P[:, d] = rng.normal(-50, 3, N)            # loglh
        P[:, d + 2] = rng.normal(-30, 2, N)        # old_loglh
        P[:, d + 4] = rng.uniform(0.2, 1.8, N)     # weights
        phi_n1, phi_n = 0.21, 0.34
        ll, old, w0 = P[:, d].copy(), P[:, d + 2].copy(), P[:, d + 4].copy()
        if pw == 0.0:
            inc = np.exp((phi_n1 - phi_n) * old + (phi_n - phi_n1) * ll)
        elif pw == 1.0:
            inc = np.exp((phi_n - phi_n1) * ll)
        else:
            inc = np.exp((phi_n1 - phi_n) * np.log(np.exp(old - lpod + np.log(1 - pw)) + pw) + (phi_n - phi_n1) * ll)
        wn = w0 * inc
        wn = wn * N / wn.sum()
        buf = O.cloud_f(P)
        oinc, onw, out = np.zeros(N), np.zeros(N), np.zeros(3)
        assert O.lib().orc_correct(buf, N, d, phi_n1, phi_n, pw, lpod, oinc.ctypes.data_as(C.c_void_p), onw.ctypes.data_as(C.c_void_p), out) == 0
        np.testing.assert_allclose(oinc, inc, rtol=1e-12)
        np.testing.assert_allclose(onw, wn, rtol=1e-12)
        assert out[1] == pytest.approx(N * N / np.sum(wn ** 2), rel=1e-12) and out[2] == pytest.approx(N, rel=1e-12)


def test_one_pass_moments_oracle_against_numpy_and_its_own_definition():
    """orc_moments_shifted (the fused stage's weighted mean / covariance, src/particle.jl:481-532 in one pass): (i) equal to numpy's
    weighted moments and to the oracle's two-pass form to rounding, also for a cloud 1e4 standard deviations from the origin;
    (ii) bit-identical to a plain Python restatement of its canonical order -- per quantity a sequential fma chain over each
    sub-chunk of 256 consecutive particles, then the adjacent-pair tree over sub-chunks -- which is the order the tensor-core
    SYRK on the device accumulates in."""
    from fractions import Fraction
    L = O.lib()

    def fma(a, b, c):                                  # exact: one rounding of a * b + c (float(Fraction) rounds to nearest even)
        return float(Fraction(a) * Fraction(b) + Fraction(c))

    for N, d, offset in ((1000, 3, 0.0), (777, 5, 1e4), (2048 + 13, 2, -3e3)):
        rng = np.random.default_rng(N)
        P = np.zeros((N, d + 5), order="F")
        P[:, :d] = rng.normal(size=(N, d)) * rng.uniform(0.5, 2.0, d) + offset
        P[:, d + 4] = rng.uniform(0.2, 2.0, N)
        shift = np.ascontiguousarray(P[0, :d])
        mean, cov = np.zeros(d), np.zeros((d, d))
        L.orc_moments_shifted(O.cloud_f(P), N, d, shift, mean, cov)
        w = P[:, d + 4]
        m_np = np.average(P[:, :d], axis=0, weights=w)
        c_np = np.cov(P[:, :d].T, aweights=w, bias=True).reshape(d, d)
        sd = np.sqrt(np.diag(c_np))
        assert np.max(np.abs(mean - m_np) / sd) < 1e-11 and np.max(np.abs(cov - c_np) / np.outer(sd, sd)) < 1e-10
        m2, c2 = np.zeros(d), np.zeros((d, d))
        L.orc_moments(O.cloud_f(P), N, d, m2, c2)
        assert np.max(np.abs(mean - m2) / sd) < 1e-11 and np.max(np.abs(cov - c2) / np.outer(sd, sd)) < 1e-10

        def tree(v):                                   # adjacent-pair tree, zero padded to a power of two
            v = list(v) + [0.0] * ((1 << max(0, (len(v) - 1).bit_length())) - len(v))
            while len(v) > 1:
                v = [v[i] + v[i + 1] for i in range(0, len(v), 2)]
            return v[0]

        def chain(term):                               # sequential fma chain per sub-chunk of 256 particles
            out = []
            for c0 in range(0, N, 256):
                acc = 0.0
                for i in range(c0, min(c0 + 256, N)):
                    a, b = term(i)
                    acc = fma(a, b, acc)
                out.append(acc)
            return tree(out)

        sw = chain(lambda i: (w[i], 1.0))
        a_, b_ = d - 1, 0
        mk = chain(lambda i: (w[i], P[i, a_] - shift[a_]))
        cab = chain(lambda i: (w[i] * (P[i, a_] - shift[a_]), P[i, b_] - shift[b_]))
        ea, eb = mk / sw, chain(lambda i: (w[i], P[i, b_] - shift[b_])) / sw
        assert mean[a_] == shift[a_] + ea
        assert cov[a_, b_] == fma(-ea, eb, cab / sw)
