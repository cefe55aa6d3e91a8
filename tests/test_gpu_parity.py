"""GPU parity tests: every C-ABI entry point of the CUDA engine against the CPU oracle on the same
seeded inputs.  Bar: BIT-EXACT (np.array_equal on float64) for weights, ESS, cumsum, ancestor
indices, moments and whole mutated clouds -- the engine and the oracle share a numerical contract
(DESIGN.md), so no tolerance is needed or used unless stated.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from smc_jl_b200 import model as M
from smc_jl_b200 import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from smc_jl_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def rand_cloud(rng, N, d, ll_scale=50.0, with_old=False):
    P = np.zeros((N, d + 5), order="F")
    P[:, :d] = rng.normal(size=(N, d))
    P[:, d] = -np.abs(rng.normal(size=N)) * ll_scale
    P[:, d + 1] = rng.normal(size=N)
    P[:, d + 2] = -np.abs(rng.normal(size=N)) * ll_scale if with_old else 0.0
    P[:, d + 3] = rng.uniform(size=N)
    P[:, d + 4] = rng.uniform(0.2, 2.0, size=N)
    return P


# --------------------------------------------------------------------------------------------------
def test_device_math_bitexact(eng):
    L = O.lib()
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-745, 709, 50000), rng.normal(0, 2, 50000), [0.0, -0.0, 800.0, -800.0, np.inf, -np.inf]])
    got = eng.debug_math(0, x)
    want = np.array([L.orc_exp(v) for v in x])
    assert np.array_equal(got, want)
    x = np.concatenate([rng.uniform(0, 2, 50000), 10.0 ** rng.uniform(-300, 300, 50000), [0.0, 1.0, np.inf, 5e-324]])
    got = eng.debug_math(1, x)
    want = np.array([L.orc_log(v) for v in x])
    assert np.array_equal(got, want)
    u = rng.uniform(0, 1, 50000)
    out = np.zeros(2)
    s, c = eng.debug_math(2, u), eng.debug_math(3, u)
    for i in range(0, len(u), 7):
        L.orc_sincos2pi_v(u[i], out)
        assert s[i] == out[0] and c[i] == out[1]
    slots = np.arange(4096, dtype=np.float64) % 11
    z0, z1 = eng.debug_math(4, slots, seed=99), eng.debug_math(5, slots, seed=99)
    for i in range(4096):
        L.orc_normal_pair(99, i, 0, int(slots[i]), out)
        assert z0[i] == out[0] and z1[i] == out[1]
    zq = [eng.debug_math(6 + j, slots, seed=99) for j in range(4)]          # proposal normals (binary32 Box-Muller)
    out4 = np.zeros(4)
    for i in range(4096):
        L.orc_normal_quad(99, i, 0, int(slots[i]), out4)
        assert all(zq[j][i] == out4[j] for j in range(4))


@pytest.mark.parametrize("N", [400, 5000, 65536 + 17])
@pytest.mark.parametrize("pw", [0.0, 1.0, 0.3])
def test_correct_bitexact(eng, N, pw):
    rng = np.random.default_rng(N + int(pw * 10))
    d = 3
    P = rand_cloud(rng, N, d, with_old=True)
    eng.cloud_create(N, d)
    eng.upload(P)
    out, inc, nw = eng.correct(0.2, 0.35, pw, -1.5, want_inc=True, want_normw=True)
    buf = O.cloud_f(P)
    oout, oinc, onw = np.zeros(3), np.zeros(N), np.zeros(N)
    st = O.lib().orc_correct(buf, N, d, 0.2, 0.35, pw, -1.5, oinc.ctypes.data_as(C.c_void_p), onw.ctypes.data_as(C.c_void_p), oout)
    assert st == 0
    assert np.array_equal(out, oout)
    assert np.array_equal(inc, oinc) and np.array_equal(nw, onw)
    assert np.array_equal(eng.download(), O.cloud_m(buf, N, d))


def test_correct_nan_ess_is_an_error(eng):
    P = rand_cloud(np.random.default_rng(0), 1000, 2)
    P[:, 2] = -np.inf       # every incremental weight is 0 => ESS NaN => check_nan_ess assertion
    eng.cloud_create(1000, 2)
    eng.upload(P)
    with pytest.raises(AssertionError):
        eng.correct(0.0, 0.5)


def test_compute_ess_golden(eng, golden):
    g = golden("compute_ess.npz")
    n = len(g["loglh"])
    P = np.zeros((n, 6), order="F")
    P[:, 1], P[:, 3], P[:, 5] = g["loglh"], g["old_loglh"], g["current_weights"]
    eng.cloud_create(n, 1)
    eng.upload(P)
    ess = eng.ess_at([float(g["phi_n"])], float(g["phi_n1"]))[0]
    assert ess == pytest.approx(float(g["ess"]), rel=1e-13)          # reference golden (test/helpers.jl:133-175)
    want = O.lib().orc_compute_ess(np.ascontiguousarray(g["loglh"]), np.ascontiguousarray(g["current_weights"]),
                                   np.ascontiguousarray(g["old_loglh"]), n, float(g["phi_n"]), float(g["phi_n1"]), np.zeros(n))
    assert ess == want                                                  # oracle, bit-exact


def test_solve_adaptive_phi_golden(eng, golden):
    g = golden("solve_adaptive_phi.npz")
    P = np.asfortranarray(g["particles"])
    N, d = P.shape[0], P.shape[1] - 5
    eng.cloud_create(N, d)
    eng.upload(P)
    i = int(g["i"])
    phi_n, rl, j, phi_prop = eng.solve_adaptive_phi(g["proposed_fixed_schedule"], int(g["j"]), float(g["phi_prop"]),
                                                    float(g["phi_n1"]), float(g["tempering_target"]),
                                                    float(g["cloud_ESS"][i - 2]), int(g["resampled_last_period"]))
    assert phi_n == pytest.approx(float(g["out_phi_n"]), rel=1e-12)   # reference golden (test/helpers.jl:15-53)
    assert j == int(g["out_j"]) and phi_prop == float(g["out_phi_prop"])
    jj, pp, pn = C.c_int64(int(g["j"])), C.c_double(float(g["phi_prop"])), C.c_double()
    O.lib().orc_solve_adaptive_phi(O.cloud_f(P), N, d, np.ascontiguousarray(g["proposed_fixed_schedule"]), 100, C.byref(jj),
                                   C.byref(pp), float(g["phi_n1"]), float(g["tempering_target"]), float(g["cloud_ESS"][i - 2]),
                                   int(g["resampled_last_period"]), C.byref(pn), None)
    assert (phi_n, j, phi_prop) == (pn.value, jj.value, pp.value)       # oracle, bit-exact


@pytest.mark.parametrize("N", [3000, 40000])
def test_solve_adaptive_phi_random(eng, N):
    rng = np.random.default_rng(N)
    d = 2
    P = rand_cloud(rng, N, d, ll_scale=300.0, with_old=True)
    P[:, d + 4] = 1.0
    eng.cloud_create(N, d)
    eng.upload(P)
    sched = (np.arange(200) / 199.0) ** 2.1
    for phi_n1, j0, prop0, resampled in [(0.0, 2, 0.0, True), (float(sched[40]), 42, float(sched[40]), False)]:
        got = eng.solve_adaptive_phi(sched, j0, prop0, phi_n1, 0.95, 0.9 * N, resampled)
        jj, pp, pn = C.c_int64(j0), C.c_double(prop0), C.c_double()
        O.lib().orc_solve_adaptive_phi(O.cloud_f(P), N, d, sched, 200, C.byref(jj), C.byref(pp), phi_n1, 0.95, 0.9 * N,
                                       int(resampled), C.byref(pn), None)
        assert (got[0], got[2], got[3]) == (pn.value, jj.value, pp.value)
        assert phi_n1 < got[0] <= 1.0


@pytest.mark.parametrize("N", [64, 400, 5000, 8192, 131072 + 3])
@pytest.mark.parametrize("method", ["systematic", "multinomial"])
def test_resample_weights_bitexact(eng, N, method):
    rng = np.random.default_rng(N)
    for case in range(3):
        if case == 0:
            w = rng.uniform(size=N)
        elif case == 1:                                  # degenerate: few heavy particles, many zeros
            w = np.where(rng.uniform(size=N) < 0.01, rng.uniform(size=N), 0.0)
            w[rng.integers(N)] = 5.0
        else:                                            # widely varying magnitudes
            w = np.exp(rng.normal(0, 8, size=N))
        w = w / w.sum()
        idx, cum = eng.resample_weights(w, method, seed=11, stage=7 + case, u=-1.0, want_cum=True)
        oidx, ocum = np.zeros(N, np.int64), np.zeros(N)
        O.lib().orc_resample(w, N, 0 if method == "systematic" else 1, 11, 7 + case, -1.0, oidx, ocum.ctypes.data_as(C.c_void_p))
        assert np.array_equal(cum, ocum)
        assert np.array_equal(idx, oidx)
        assert idx.min() >= 1 and idx.max() <= N
        if method == "systematic":
            assert np.all(np.diff(idx) >= 0)
            assert np.all(np.abs(np.bincount(idx - 1, minlength=N) - N * w) < 1 + 1e-6)


def test_resample_explicit_offset_edges(eng):
    """u = 0 and u -> 1: first/last thresholds; 'not found' clamps to N (reference would return 0)."""
    N = 1000
    w = np.random.default_rng(5).uniform(size=N)
    w /= w.sum()
    for u in (0.0, 0.5, 1.0 - 2 ** -53):
        idx = eng.resample_weights(w, "systematic", u=u)
        oidx = np.zeros(N, np.int64)
        O.lib().orc_resample(w, N, 0, 0, 0, u, oidx, None)
        assert np.array_equal(idx, oidx)
    with pytest.raises(ValueError):
        eng.resample_weights(w, "stratified")


def test_polyalgo_is_served_by_the_multinomial_kernel(eng):
    """:polyalgo = StatsBase.sample(1:n, Weights(w), n): i.i.d. categorical draws (src/resample.jl:73-75)."""
    rng = np.random.default_rng(5)
    w = rng.gamma(0.5, 1.0, 3000)
    a = eng.resample_weights(w, "polyalgo", seed=9, stage=4)
    b = eng.resample_weights(w, "multinomial", seed=9, stage=4)
    assert np.array_equal(a, b)
    # i.i.d. categorical draws: counts over 20 weight-sorted groups of particles follow the multinomial law
    order = np.argsort(w)
    groups = np.array_split(order, 20)
    p = np.array([w[g].sum() for g in groups]) / w.sum()
    counts = np.array([np.isin(a - 1, g).sum() for g in groups])
    big = 3000 * p > 5
    z = (counts[big] - 3000 * p[big]) / np.sqrt(3000 * p[big] * (1 - p[big]))
    assert big.sum() >= 5 and np.abs(z).max() < 4.5


@pytest.mark.parametrize("method", ["systematic", "multinomial"])
def test_resample_with_n_parts_bitexact(eng, method):
    """`resample(weights; n_parts = n_out)` with n_out != length(weights) (bridge, smc_main.jl:262-268)."""
    rng = np.random.default_rng(77)
    w = rng.gamma(0.7, 1.0, 5000)
    for n_out in (1, 2500, 5000, 12000):
        idx = eng.resample_weights(w, method, seed=5, stage=3, n_parts=n_out)
        oidx = np.zeros(n_out, np.int64)
        assert O.lib().orc_resample_n(w, 5000, n_out, 0 if method == "systematic" else 1, 5, 3, -1.0, oidx, None) == 0
        assert idx.shape == (n_out,) and np.array_equal(idx, oidx)
        assert idx.min() >= 1 and idx.max() <= 5000
        if method == "systematic":
            assert np.all(np.diff(idx) >= 0)
            if n_out >= 2500:                      # offspring counts track n_out * w within 1
                counts = np.bincount(idx - 1, minlength=5000)
                assert np.all(np.abs(counts - n_out * w / w.sum()) < 1 + 1e-9)


@pytest.mark.parametrize("N,d", [(5000, 9), (4096, 20), (70000, 2), (3000, 5)])
def test_selection_and_moments_bitexact(eng, N, d):
    rng = np.random.default_rng(N + d)
    P = rand_cloud(rng, N, d)
    P[:, d + 4] *= N / P[:, d + 4].sum()
    eng.cloud_create(N, d)
    eng.upload(P)
    mean, cov = eng.moments()
    omean, ocov = np.zeros(d), np.zeros((d, d))
    O.lib().orc_moments(O.cloud_f(P), N, d, omean, ocov)
    assert np.array_equal(mean, omean) and np.array_equal(cov, ocov)
    np.testing.assert_allclose(mean, np.average(P[:, :d], axis=0, weights=P[:, -1]), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(cov, np.cov(P[:, :d].T, aweights=P[:, -1], bias=True), rtol=1e-10, atol=1e-14)
    # selection: indices, gather of all columns, weights reset to 1
    idx = eng.resample("systematic", seed=3, stage=9, want_indices=True)
    oidx = np.zeros(N, np.int64)
    wn = np.ascontiguousarray(P[:, -1] / N)
    O.lib().orc_resample(wn, N, 0, 3, 9, -1.0, oidx, None)
    assert np.array_equal(idx, oidx)
    dst = np.zeros(N * (d + 5))
    O.lib().orc_gather(O.cloud_f(P), dst, N, d, oidx)
    got = eng.download()
    assert np.array_equal(got, O.cloud_m(dst, N, d))
    assert np.all(got[:, -1] == 1.0)


def _mutation_case(eng, spec, P, blocks, phi_n, phi_n1, c, n_mh, has_old, seed, stage, alpha=1.0):
    N, d = P.shape[0], spec.d
    free = spec.free_inds
    mean = np.average(P[:, :d], axis=0, weights=P[:, -1])
    cov = np.cov(P[:, :d].T, aweights=P[:, -1], bias=True).reshape(d, d)
    mean_fr, cov_fr = mean[free], np.ascontiguousarray(cov[np.ix_(free, free)])
    blocks_free = [np.asarray(b, np.int32) for b in blocks]
    blocks_all = [free[b].astype(np.int32) for b in blocks_free]
    eng.cloud_create(N, d)
    eng.set_model(spec)
    eng.upload(P)
    acc = eng.mutate(mean_fr, cov_fr, blocks_free, blocks_all, phi_n, phi_n1, c=c, alpha=alpha, n_mh_steps=n_mh,
                     has_old_data=has_old, seed=seed, stage=stage)
    got = eng.download()
    L = O.lib()
    mod = O.Model(spec)
    st = C.c_int()
    pr = L.orc_proposal_create(d, len(free), np.ascontiguousarray(mean_fr), cov_fr, len(blocks),
                               np.array([len(b) for b in blocks], np.int32), np.concatenate(blocks_free).astype(np.int32),
                               np.concatenate(blocks_all).astype(np.int32), c, C.byref(st))
    assert pr and st.value == 0
    buf = O.cloud_f(P)
    L.orc_mutate(mod.h, pr, buf, N, 0, phi_n, phi_n1, alpha, n_mh, len(free), int(has_old), seed, stage, 0)
    L.orc_proposal_free(pr)
    want = O.cloud_m(buf, N, d)
    oacc = L.orc_mean_accept(buf, N, d, len(free))
    return got, want, acc, oacc


def _evaluated_cloud(eng, spec, params, N, rng):
    P = W.initial_cloud(params, N, rng)
    eng.cloud_create(N, spec.d)
    eng.set_model(spec)
    eng.upload(P)
    eng.evaluate(0)
    return eng.download()


def test_evaluate_matches_oracle_and_direct(eng):
    params, lk, (y, X, _) = W.linear_gaussian(d=20, T=256)
    spec = M.make_spec(params, lk)
    rng = np.random.default_rng(1)
    N = 5000
    P = _evaluated_cloud(eng, spec, params, N, rng)
    mod = O.Model(spec)
    buf = O.cloud_f(W_reset(P))
    O.lib().orc_evaluate(mod.h, buf, N)
    want = O.cloud_m(buf, N, 20)
    assert np.array_equal(P, want)
    # against the per-observation form of the reference's example likelihood (relative 1e-12)
    Xf = np.ascontiguousarray(X)
    for r in range(0, N, 97):
        direct = O.lib().orc_loglik_linreg_direct(np.ascontiguousarray(P[r, :20]), np.ascontiguousarray(y), Xf, 256, 20, 1.0)
        assert P[r, 20] == pytest.approx(direct, rel=1e-12)


def W_reset(P):
    Q = P.copy(order="F")
    d = P.shape[1] - 5
    Q[:, d] = 0.0
    Q[:, d + 1] = 0.0
    return Q


@pytest.mark.parametrize("d,n_mh", [(20, 3), (2, 1), (8, 2), (7, 2), (11, 1), (31, 1)])      # 7, 11, 31: the general-kernel-only sizes
def test_mutation_linear_gaussian_bitexact(eng, d, n_mh):
    params, lk, _ = W.linear_gaussian(d=d, T=64)
    spec = M.make_spec(params, lk)
    rng = np.random.default_rng(d)
    N = 4096 + 37
    P = _evaluated_cloud(eng, spec, params, N, rng)
    P[:, -1] = rng.uniform(0.5, 1.5, N)
    got, want, acc, oacc = _mutation_case(eng, spec, P, [np.arange(d)], 0.05, 0.02, 0.4, n_mh, False, 1793, 5)
    assert np.array_equal(got, want)
    assert acc == oacc
    assert 0.0 < acc < n_mh + 1e-9
    assert not np.array_equal(got[:, :d], P[:, :d])            # something moved


def test_mutation_three_equation_blocks_and_old_data_bitexact(eng):
    data, X = W.synthetic_three_equation(T=100)
    params = W.three_equation_parameters()
    spec = M.make_spec(params, M.LinearEquationsLogLik(data, X), M.LinearEquationsLogLik(data[:, :50], X))
    rng = np.random.default_rng(4)
    N = 3000
    P = W.initial_cloud(params, N, rng)
    P[:, :9] = np.abs(rng.normal(1.5, 0.5, (N, 9)))             # near the posterior so that moves are accepted
    eng.cloud_create(N, 9)
    eng.set_model(spec)
    eng.upload(P)
    eng.evaluate(0)
    P = eng.download()
    # old_loglh column as initialize_likelihoods! would leave it: likelihood of the old data
    mod = O.Model(spec)
    P[:, 11] = [mod.loglik(np.ascontiguousarray(P[r, :9]), 1) for r in range(N)]
    blocks = [np.array([6, 1, 4]), np.array([0, 8, 3]), np.array([2, 7, 5])]
    got, want, acc, oacc = _mutation_case(eng, spec, P, blocks, 0.3, 0.2, 0.3, 2, True, 7, 11)
    assert np.array_equal(got, want)
    assert acc == oacc and acc > 0.0


@pytest.mark.parametrize("d,n_mh,alpha", [(20, 2, 0.9), (8, 3, 0.5), (2, 1, 0.0)])
def test_mutation_mixture_proposal_bitexact(eng, d, n_mh, alpha):
    """alpha < 1: three-component mixture draw (helpers.jl:87-100) and proposal densities (:128-164)."""
    params, lk, _ = W.linear_gaussian(d=d, T=64)
    spec = M.make_spec(params, lk)
    rng = np.random.default_rng(100 + d)
    N = 4096 + 11
    P = _evaluated_cloud(eng, spec, params, N, rng)
    P[:, -1] = rng.uniform(0.5, 1.5, N)
    got, want, acc, oacc = _mutation_case(eng, spec, P, [np.arange(d)], 0.05, 0.02, 0.4, n_mh, False, 1793, 5, alpha=alpha)
    assert np.array_equal(got, want)
    assert acc == oacc and acc > 0.0
    # the mixture moved particles differently from the pure random walk
    got1, _, _, _ = _mutation_case(eng, spec, P, [np.arange(d)], 0.05, 0.02, 0.4, n_mh, False, 1793, 5, alpha=1.0)
    assert not np.array_equal(got, got1)


def test_mutation_mixture_blocks_old_data_bitexact(eng):
    """The reference's test-suite settings (alpha = 0.9, test/smc.jl:29) with 3 random blocks and old data."""
    data, X = W.synthetic_three_equation(T=100)
    params = W.three_equation_parameters()
    spec = M.make_spec(params, M.LinearEquationsLogLik(data, X), M.LinearEquationsLogLik(data[:, :50], X))
    rng = np.random.default_rng(5)
    N = 3000
    P = W.initial_cloud(params, N, rng)
    P[:, :9] = np.abs(rng.normal(1.5, 0.5, (N, 9)))
    eng.cloud_create(N, 9)
    eng.set_model(spec)
    eng.upload(P)
    eng.evaluate(0)
    P = eng.download()
    mod = O.Model(spec)
    P[:, 11] = [mod.loglik(np.ascontiguousarray(P[r, :9]), 1) for r in range(N)]
    blocks = [np.array([6, 1, 4]), np.array([0, 8, 3]), np.array([2, 7, 5])]
    got, want, acc, oacc = _mutation_case(eng, spec, P, blocks, 0.3, 0.2, 0.3, 2, True, 7, 11, alpha=0.9)
    assert np.array_equal(got, want)
    assert acc == oacc and acc > 0.0


def test_mutation_golden(eng, golden):
    """Reference golden test/mutation.jl:1-59 through the CUDA path (reject-all pass-through, accept = 0)."""
    g = golden("mutation.npz")
    lm = golden("linear_model_rows.npz")
    params = W.three_equation_parameters()
    spec = M.make_spec(params, M.LinearEquationsLogLik(lm["data"], lm["X"]), M.LinearEquationsLogLik(g["old_data"], lm["X"]))
    P = np.asfortranarray(g["particles_in"])
    eng.cloud_create(P.shape[0], 9)
    eng.set_model(spec)
    eng.upload(P)
    bf = [(g["blocks_free"] - 1).astype(np.int32)]
    ba = [(g["blocks_all"] - 1).astype(np.int32)]
    acc = eng.mutate(g["mu"], g["Sigma"], bf, ba, float(g["phi_n"]), float(g["phi_n1"]), c=float(g["c"]), alpha=1.0,
                     n_mh_steps=1, has_old_data=True, seed=42, stage=2)
    got = eng.download()
    assert np.array_equal(got, g["particles_out"])
    assert acc == 0.0


def test_not_posdef_is_an_error(eng):
    params, lk, _ = W.linear_gaussian(d=4, T=32)
    spec = M.make_spec(params, lk)
    eng.cloud_create(256, 4)
    eng.set_model(spec)
    eng.upload(W.initial_cloud(params, 256, np.random.default_rng(0)))
    bad = -np.eye(4)
    with pytest.raises(np.linalg.LinAlgError):
        eng.mutate(np.zeros(4), bad, [np.arange(4)], [np.arange(4)], 0.1, 0.0)


def _run_stages(eng, spec, P0, sched, n_stage, cfg_kw, adaptive=False, alpha=1.0):
    """Run stages on the GPU (fused smcb200_stage) and in the oracle (orc_stage); compare everything."""
    from smc_jl_b200._lib import StageConfig, StageState
    N, d = P0.shape[0], spec.d
    eng.cloud_create(N, d)
    eng.set_model(spec)
    eng.upload(P0)
    state = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2, resampled_last_period=0)
    io = O.StageIO(threshold_ratio=0.5, target=0.25, alpha=alpha, tempering_target=0.95, pw=0.0, log_prob_old_data=0.0,
                   n_mh_steps=cfg_kw["n_mh_steps"], n_blocks=cfg_kw["n_blocks"], resample_method=0, adaptive=int(adaptive),
                   has_old=cfg_kw.get("has_old", 0), nthreads=0, seed=1793, c=0.5, accept=0.25, ess_prev=float(N),
                   resampled_last=0, j=2, phi_prop=0.0)
    mod = O.Model(spec)
    buf = O.cloud_f(P0)
    scratch = np.zeros_like(buf)
    phi_prev = 0.0
    n_resampled = 0
    for s in range(n_stage):
        stage = s + 2
        phi_n = float(sched[s + 1])
        cfg = StageConfig(phi_n1=phi_prev, phi_n=phi_n, threshold_ratio=0.5, target=0.25, alpha=alpha, tempering_target=0.95,
                          prior_weight=0.0, log_prob_old_data=0.0, n_mh_steps=cfg_kw["n_mh_steps"], n_blocks=cfg_kw["n_blocks"],
                          resample_method=0, adaptive=int(adaptive), has_old_data=cfg_kw.get("has_old", 0), seed=1793, stage=stage)
        res, inc, nw = eng.stage(cfg, state, schedule=sched, want_inc=True, want_normw=True)
        io.phi_n1, io.phi_n, io.stage = phi_prev, phi_n, stage
        oinc, onw = np.zeros(N), np.zeros(N)
        st = O.lib().orc_stage(mod.h, buf, scratch, N, np.ascontiguousarray(sched), len(sched), C.byref(io),
                               oinc.ctypes.data_as(C.c_void_p), onw.ctypes.data_as(C.c_void_p), None, None)
        assert st == 0
        assert res.phi_n == io.phi_out
        assert res.ess == io.ess and res.sum_weights == io.sum_w
        assert res.resampled == io.resampled
        assert res.c == io.c and res.accept == io.accept
        assert np.array_equal(inc, oinc) and np.array_equal(nw, onw)
        assert np.array_equal(eng.download(), O.cloud_m(buf, N, d)), "cloud differs at stage %d" % stage
        assert (state.j, state.phi_prop, state.resampled_last_period) == (io.j, io.phi_prop, io.resampled_last)
        n_resampled += res.resampled
        phi_prev = res.phi_n
        if phi_prev >= 1.0:
            break
    return n_resampled, phi_prev


def test_stage_trajectory_linear_gaussian_bitexact(eng):
    """C2-shaped run (d = 20, n_mh = 3, fixed schedule) at a size the oracle finishes in seconds: identical
    phi/ESS/c/accept trajectory, identical resample decisions, identical clouds after every stage."""
    params, lk, _ = W.linear_gaussian(d=20, T=256, prior_sd=1.0)   # 40-point schedule: a N(0,10) prior would collapse the ESS
    spec = M.make_spec(params, lk)
    N = 8192
    P0 = _evaluated_cloud(eng, spec, params, N, np.random.default_rng(7))
    sched = (np.arange(40) / 39.0) ** 2.1
    n_res, phi = _run_stages(eng, spec, P0, sched, 39, dict(n_mh_steps=3, n_blocks=1))
    assert phi == 1.0 and n_res >= 3
    # posterior mean of the final cloud is near the OLS estimate
    mean, cov = eng.moments()
    bhat = lk.eqdata[4:24]
    assert np.max(np.abs(mean - bhat) / np.sqrt(np.diag(cov))) < 0.5


def test_stage_trajectory_blocks_old_data_bitexact(eng):
    data, X = W.synthetic_three_equation(T=100)
    params = W.three_equation_parameters(prior_para=10.0)
    spec = M.make_spec(params, M.LinearEquationsLogLik(data, X), M.LinearEquationsLogLik(data[:, :50], X))
    N = 5000
    P0 = _evaluated_cloud(eng, spec, params, N, np.random.default_rng(8))
    mod = O.Model(spec)
    P0[:, 11] = [mod.loglik(np.ascontiguousarray(P0[r, :9]), 1) for r in range(N)]
    sched = (np.arange(25) / 24.0) ** 2.0
    n_res, phi = _run_stages(eng, spec, P0, sched, 24, dict(n_mh_steps=2, n_blocks=3, has_old=1))
    assert phi == 1.0 and n_res >= 2


def test_stage_trajectory_mixture_blocks_bitexact(eng):
    """test/smc.jl:13-87 settings (3-equation model, alpha = 0.9) with 3 blocks: whole trajectory bit-exact."""
    data, X = W.synthetic_three_equation(T=100)
    params = W.three_equation_parameters(prior_para=10.0)
    spec = M.make_spec(params, M.LinearEquationsLogLik(data, X))
    N = 5000
    P0 = _evaluated_cloud(eng, spec, params, N, np.random.default_rng(18))
    sched = (np.arange(25) / 24.0) ** 2.1
    n_res, phi = _run_stages(eng, spec, P0, sched, 24, dict(n_mh_steps=1, n_blocks=3), alpha=0.9)
    assert phi == 1.0 and n_res >= 2


def test_stage_trajectory_adaptive_bitexact(eng):
    params, lk, _ = W.regression_example()
    spec = M.make_spec(params, lk)
    N = 1000                                                   # config C1 size
    P0 = _evaluated_cloud(eng, spec, params, N, np.random.default_rng(9))
    sched = (np.arange(60) / 59.0) ** 2.1
    n_res, phi = _run_stages(eng, spec, P0, sched, 59, dict(n_mh_steps=1, n_blocks=1), adaptive=True)
    assert 0.0 < phi <= 1.0


def test_moments_onepass_bitexact(eng):
    """The fused stage's one-pass moments (shift = particle 0): bit-identical to the oracle's restatement of the same sums,
    equal to the reference's two-pass weighted_mean / weighted_cov (src/particle.jl:481-532) to rounding -- also for a cloud
    whose mean is 1e4 standard deviations away from the origin (where an unshifted one-pass formula would lose 8 digits)."""
    for N, d, offset in ((5000, 9, 0.0), (4096, 20, 0.0), (70001, 2, 0.0), (3000, 5, 0.0), (12289, 7, 1e4), (6000, 20, -3e3),
                         (2500, 25, 0.0), (3001, 32, 0.0)):     # 25, 32: four and five variable tiles of the tensor-core SYRK
        rng = np.random.default_rng(N + d)
        P = rand_cloud(rng, N, d)
        P[:, :d] += offset
        P[:, d + 4] *= N / P[:, d + 4].sum()
        params, lk, _ = W.linear_gaussian(d=d, T=32)
        eng.cloud_create(N, d)
        eng.set_model(M.make_spec(params, lk))
        eng.upload(P)
        mean, cov = eng.moments_onepass()
        omean, ocov = np.zeros(d), np.zeros((d, d))
        O.lib().orc_moments_shifted(O.cloud_f(P), N, d, np.ascontiguousarray(P[0, :d]), omean, ocov)
        assert np.array_equal(mean, omean) and np.array_equal(cov, ocov), (N, d)
        m2, c2 = eng.moments()                                          # two-pass form
        scale = np.sqrt(np.outer(np.diag(c2), np.diag(c2)))
        assert np.max(np.abs(mean - m2) / np.sqrt(np.diag(c2))) < 1e-10
        assert np.max(np.abs(cov - c2) / scale) < 1e-10
        assert np.array_equal(eng.download(), P)                       # the cloud itself is untouched


@pytest.mark.parametrize("adaptive", [False, True])
def test_run_stages_matches_the_oracle_trajectory(eng, adaptive):
    """smcb200_run_stages: the whole recursion in one call (no host in the loop on a fixed schedule) gives the per-stage
    results, the w / W history columns and the final cloud of the oracle's stage-by-stage loop, bit for bit."""
    from smc_jl_b200._lib import StageConfig, StageState
    data, X = W.synthetic_three_equation(T=100)
    params = W.three_equation_parameters(prior_para=10.0)
    spec = M.make_spec(params, M.LinearEquationsLogLik(data, X))
    N, d = 6000, 9
    P0 = _evaluated_cloud(eng, spec, params, N, np.random.default_rng(21))
    n_phi = 30
    sched = (np.arange(n_phi) / (n_phi - 1.0)) ** 2.1
    eng.upload(P0)
    state = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2, resampled_last_period=0)
    cfg = StageConfig(phi_n1=0.0, phi_n=0.0, threshold_ratio=0.5, target=0.25, alpha=0.9, tempering_target=0.9,
                      n_mh_steps=2, n_blocks=2, resample_method=0, adaptive=int(adaptive), seed=99, stage=0)
    n_max = n_phi - 1 if not adaptive else 200
    inc_h, nw_h = np.zeros((n_max, N)), np.zeros((n_max, N))
    results = eng.run_stages(cfg, state, sched, 2, n_max, inc_hist=inc_h, normw_hist=nw_h)
    assert results[-1].phi_n == 1.0
    if not adaptive:
        assert len(results) == n_phi - 1
    got = eng.download()
    # oracle, one stage at a time
    io = O.StageIO(threshold_ratio=0.5, target=0.25, alpha=0.9, tempering_target=0.9, pw=0.0, log_prob_old_data=0.0,
                   n_mh_steps=2, n_blocks=2, resample_method=0, adaptive=int(adaptive), has_old=0, nthreads=0, seed=99, c=0.5,
                   accept=0.25, ess_prev=float(N), resampled_last=0, j=2, phi_prop=0.0)
    mod = O.Model(spec)
    buf = O.cloud_f(P0)
    scratch = np.zeros_like(buf)
    phi_prev, n_res = 0.0, 0
    for k, res in enumerate(results):
        io.phi_n1, io.phi_n, io.stage = phi_prev, float(sched[min(k + 1, n_phi - 1)]), k + 2
        oinc, onw = np.zeros(N), np.zeros(N)
        assert O.lib().orc_stage(mod.h, buf, scratch, N, np.ascontiguousarray(sched), n_phi, C.byref(io),
                                 oinc.ctypes.data_as(C.c_void_p), onw.ctypes.data_as(C.c_void_p), None, None) == 0
        assert (res.phi_n, res.ess, res.sum_weights, res.c, res.accept, res.resampled) == \
               (io.phi_out, io.ess, io.sum_w, io.c, io.accept, io.resampled), k
        assert np.array_equal(inc_h[k], oinc) and np.array_equal(nw_h[k], onw), k
        phi_prev = io.phi_out
        n_res += io.resampled
    assert n_res >= 2 and phi_prev == 1.0
    assert np.array_equal(got, O.cloud_m(buf, N, d))
    assert (state.c, state.accept, state.ess_prev) == (io.c, io.accept, io.ess)


def test_full_size_c2_bitexact_against_the_oracle(eng):
    """BASELINE config C2 at its FULL size (d = 20, N = 2^20, n_mh_steps = 3): the first stages of the real schedule,
    including one that resamples, against the CPU oracle -- identical ESS / c / accept, identical clouds
    (np.array_equal on all 2^20 x 25 doubles), hence posterior mean / std far inside the 1e-6 relative target."""
    from smc_jl_b200._lib import StageConfig, StageState
    params, lk, _ = W.linear_gaussian(d=20, T=256)
    spec = M.make_spec(params, lk)
    N, d = 1 << 20, 20
    P0 = _evaluated_cloud(eng, spec, params, N, np.random.default_rng(11))
    mod = O.Model(spec)
    buf = O.cloud_f(P0)
    scratch = np.zeros_like(buf)
    sched = ((np.arange(300)) / 299.0) ** 2.1
    thr = 0.9                                                      # resample early (the reference default 0.5 needs more stages)
    state = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2, resampled_last_period=0)
    io = O.StageIO(threshold_ratio=thr, target=0.25, alpha=1.0, tempering_target=0.95, pw=0.0, log_prob_old_data=0.0,
                   n_mh_steps=3, n_blocks=1, resample_method=0, adaptive=0, has_old=0, nthreads=0, seed=1793, c=0.5, accept=0.25,
                   ess_prev=float(N), resampled_last=0, j=2, phi_prop=0.0)
    eng.upload(P0)
    resamples = 0
    for s in range(8):
        cfg = StageConfig(phi_n1=float(sched[s]), phi_n=float(sched[s + 1]), threshold_ratio=thr, target=0.25, alpha=1.0,
                          tempering_target=0.95, n_mh_steps=3, n_blocks=1, resample_method=0, seed=1793, stage=s + 2)
        res, _, _ = eng.stage(cfg, state)
        io.phi_n1, io.phi_n, io.stage = float(sched[s]), float(sched[s + 1]), s + 2
        assert O.lib().orc_stage(mod.h, buf, scratch, N, sched, 300, C.byref(io), None, None, None, None) == 0
        assert (res.ess, res.c, res.accept, res.resampled) == (io.ess, io.c, io.accept, io.resampled), s
        resamples += res.resampled
        if resamples >= 1 and s >= 2:
            break
    assert resamples >= 1
    got, want = eng.download(), O.cloud_m(buf, N, d)
    assert np.array_equal(got, want)
    w = got[:, -1]
    gm, wm = np.average(got[:, :d], axis=0, weights=w), np.average(want[:, :d], axis=0, weights=want[:, -1])
    gs = np.sqrt(np.average((got[:, :d] - gm) ** 2, axis=0, weights=w))
    ws = np.sqrt(np.average((want[:, :d] - wm) ** 2, axis=0, weights=want[:, -1]))
    assert np.max(np.abs(gm - wm) / np.abs(wm)) <= 1e-6 and np.max(np.abs(gs - ws) / ws) <= 1e-6    # north_star tolerance


# ---- full-size (BASELINE config C2) size-independent properties ------------------
def test_full_size_properties(eng):
    from smc_jl_b200._lib import StageConfig, StageState
    params, lk, _ = W.linear_gaussian(d=20, T=256)
    spec = M.make_spec(params, lk)
    N = 1 << 20
    P0 = _evaluated_cloud(eng, spec, params, N, np.random.default_rng(10))
    assert np.all(np.isfinite(P0[:, 20])) and np.all(P0[:, -1] == 1.0)
    sched = ((np.arange(300)) / 299.0) ** 2.1
    state = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2, resampled_last_period=0)
    eng.upload(P0)
    ess_seq, resamples = [], 0
    for s in range(6):
        cfg = StageConfig(phi_n1=float(sched[s]), phi_n=float(sched[s + 1]), threshold_ratio=0.5, target=0.25, alpha=1.0,
                          tempering_target=0.95, n_mh_steps=3, n_blocks=1, resample_method=0, seed=1793, stage=s + 2)
        res, _, nw = eng.stage(cfg, state, want_normw=True)
        assert abs(nw.sum() - N) < 1e-6 * N                         # weights normalised to N (particle.jl:362-369)
        assert res.ess == pytest.approx(N * N / np.sum(nw.astype(np.longdouble) ** 2) if not res.resampled else res.ess,
                                        rel=1e-12)
        assert 0 < res.ess <= N * (1 + 1e-12)
        ess_seq.append(res.ess)
        resamples += res.resampled
    # selection at full size: sortedness, range, offspring counts within 1 of N*w, gather idempotence
    w = eng.read_column(-1)
    idx = eng.resample("systematic", seed=5, stage=99, want_indices=True)
    assert np.all(np.diff(idx) >= 0) and idx[0] >= 1 and idx[-1] <= N
    counts = np.bincount(idx - 1, minlength=N)
    assert np.all(np.abs(counts - w / w.sum() * N) < 1 + 1e-6)
    before = eng.download()
    assert np.all(before[:, -1] == 1.0)
    idx2 = eng.resample("systematic", seed=5, stage=100, u=0.5, want_indices=True)   # uniform weights: identity
    assert np.array_equal(idx2, np.arange(1, N + 1))
    assert np.array_equal(eng.download(), before)


# ---- config C4: An-Schorfheide DSGE likelihood (device decision rule + Kalman filter) -------------------
def _as_spec(g, with_old=False):
    data = g["data"]
    ps = W.an_schorfheide_parameters()
    if with_old:
        return M.make_spec(ps, M.AnSchorfheideLogLik(data), M.AnSchorfheideLogLik(data[:, :115]))
    return M.make_spec(ps, M.AnSchorfheideLogLik(data))


def test_as_evaluate_golden_and_oracle(eng, golden):
    """The reference-produced (theta -> loglh, logprior) rows of the saved An-Schorfheide clouds through the CUDA
    likelihood: relative 2e-11 against the reference's numbers, bit-exact against the oracle."""
    g = golden("as_clouds.npz")
    for name, T in (("cloud600", 230), ("cloud1000", 115), ("prior_draws", 230)):
        P = np.asfortranarray(g[name])
        N = P.shape[0]
        ps = W.an_schorfheide_parameters()
        spec = M.make_spec(ps, M.AnSchorfheideLogLik(g["data"][:, :T]))
        eng.cloud_create(N, 16)
        eng.set_model(spec)
        eng.upload(W_reset(P))
        eng.evaluate(0)
        got = eng.download()
        rel = np.abs(got[:, 16] - P[:, 16]) / np.maximum(1.0, np.abs(P[:, 16]))
        assert np.median(rel) < 1e-14
        assert np.sort(rel)[-2] < 2e-11 and rel.max() < 1e-6     # one near-unit-root prior draw
        np.testing.assert_allclose(got[:, 17], P[:, 17], rtol=1e-13, atol=2e-13)
        buf = O.cloud_f(W_reset(P))
        mod = O.Model(spec)
        O.lib().orc_evaluate(mod.h, buf, N)
        assert np.array_equal(got, O.cloud_m(buf, N, 16))


def test_as_missing_observations_bitexact(eng, golden):
    """NaN observations: the device filter's inert-pivot masking equals the oracle's reduced system bit for bit (evaluation and
    a mutation with old data whose vintage also has gaps); the values differ from the complete-data ones."""
    g = golden("as_clouds.npz")
    data = g["data"].copy()
    rng = np.random.default_rng(8)
    data[0, 7] = np.nan; data[1, 13] = np.nan; data[2, 19] = np.nan
    data[[0, 2], 40] = np.nan; data[[1, 2], 41] = np.nan; data[[0, 1], 42] = np.nan; data[:, 60] = np.nan; data[1, 1] = np.nan
    data[rng.integers(0, 3, 25), rng.integers(70, 230, 25)] = np.nan
    ps = W.an_schorfheide_parameters()
    P = np.asfortranarray(g["cloud600"]).copy(order="F")
    N = P.shape[0]
    spec = M.make_spec(ps, M.AnSchorfheideLogLik(data))
    eng.cloud_create(N, 16)
    eng.set_model(spec)
    eng.upload(W_reset(P))
    eng.evaluate(0)
    got = eng.download()
    buf = O.cloud_f(W_reset(P))
    mod = O.Model(spec)                                        # (keep the handle alive across the call)
    O.lib().orc_evaluate(mod.h, buf, N)
    assert np.array_equal(got, O.cloud_m(buf, N, 16))
    assert np.all(np.isfinite(got[:, 16])) and not np.any(got[:, 16] == P[:, 16])
    spec2 = M.make_spec(ps, M.AnSchorfheideLogLik(data), M.AnSchorfheideLogLik(data[:, :115]))
    Q = got.copy(order="F")
    Q[:, -1] = 1.0
    got2, want2, acc, oacc = _mutation_case(eng, spec2, Q, [np.arange(13)], 0.6, 0.5, 0.3, 1, True, 99, 7)
    assert np.array_equal(got2, want2) and acc == oacc and acc > 0.0


def test_as_initialize_likelihoods_golden(eng, golden):
    """Online update: initialize_likelihoods! moves loglh (first vintage, 115 periods) to old_loglh and
    re-evaluates on the full sample; the saved cloud of the reference's second-vintage run holds both columns."""
    g = golden("as_clouds.npz")
    P = np.asfortranarray(g["cloud600"])
    N = P.shape[0]
    eng.cloud_create(N, 16)
    eng.set_model(M.make_spec(W.an_schorfheide_parameters(), M.AnSchorfheideLogLik(g["data"][:, :115])))
    eng.upload(W_reset(P))
    eng.evaluate(0)
    first = eng.download()
    np.testing.assert_allclose(first[:, 16], P[:, 18], rtol=2e-11)           # loglh on the old data == stored old_loglh
    eng.set_model(_as_spec(g))
    eng.evaluate(1)
    got = eng.download()
    assert np.array_equal(got[:, 18], first[:, 16])
    np.testing.assert_allclose(got[:, 16], P[:, 16], rtol=2e-11)


@pytest.mark.parametrize("blocks,alpha,n_mh", [(1, 1.0, 2), (3, 0.9, 1)])
def test_as_mutation_bitexact(eng, golden, blocks, alpha, n_mh):
    g = golden("as_clouds.npz")
    spec = _as_spec(g, with_old=True)
    P = np.asfortranarray(g["cloud600"]).copy(order="F")
    P[:, -1] = 1.0
    rng = np.random.default_rng(blocks)
    perm = rng.permutation(13)
    blk = [np.sort(b) for b in np.array_split(perm, blocks)]
    got, want, acc, oacc = _mutation_case(eng, spec, P, blk, 0.6, 0.5, 0.3, n_mh, True, 99, 7, alpha=alpha)
    assert np.array_equal(got, want)
    assert acc == oacc and acc > 0.0
    assert np.array_equal(got[:, 13:16], P[:, 13:16])                       # fixed measurement errors never move


def test_as_stage_trajectory_bitexact(eng, golden):
    """C4-shaped run (13 free parameters, n_mh_steps = 5) from the reference's own prior draws: identical
    trajectories and clouds, GPU vs oracle."""
    g = golden("as_clouds.npz")
    spec = _as_spec(g)
    P0 = np.asfortranarray(g["prior_draws"]).copy(order="F")
    P0[:, 18:20] = 0.0
    P0[:, 20] = 1.0
    sched = (np.arange(12) / 11.0) ** 3.0
    n_res, phi = _run_stages(eng, spec, P0, sched, 6, dict(n_mh_steps=5, n_blocks=1))
    assert n_res >= 1


# ---- stage 0 on the device: initial_draw! (src/initialization.jl:88-119) ---------------------------------
def _initial_draw_case(eng, spec, N, seed, max_tries=50):
    eng.cloud_create(N, spec.d)
    eng.set_model(spec)
    eng.initial_draw(spec.values, seed, max_tries)
    got = eng.download()
    buf = np.zeros(N * (spec.d + 5))
    mod = O.Model(spec)
    assert O.lib().orc_initial_draw(mod.h, buf, N, 0, np.ascontiguousarray(spec.values), seed, max_tries) == 0
    return got, O.cloud_m(buf, N, spec.d)


def test_initial_draw_bitexact_linear_and_three_equation(eng):
    params, lk, _ = W.linear_gaussian(d=20, T=256)
    got, want = _initial_draw_case(eng, M.make_spec(params, lk), 5000, 1793)
    assert np.array_equal(got, want)
    assert np.all(got[:, -1] == 1.0) and np.all(got[:, 22] == 0.0) and np.all(np.isfinite(got[:, 20]))
    assert abs(got[:, :20].std() - 10.0) < 0.2                                  # N(0, 10) priors
    data, X = W.synthetic_three_equation(T=100)
    spec = M.make_spec(W.three_equation_parameters(), M.LinearEquationsLogLik(data, X))
    got, want = _initial_draw_case(eng, spec, 4000, 5)
    assert np.array_equal(got, want)
    assert np.all((got[:, [2, 5, 8]] > 1e-5) & (got[:, [2, 5, 8]] < 1e3))       # sigma ~ U(0, 1e3) inside valuebounds


def test_initial_draw_bitexact_an_schorfheide(eng, golden):
    """Gamma / RootInverseGamma / Uniform / Normal prior samplers, fixed parameters, and the redraw of draws without a
    unique stable solution (loglh = -Inf, initialization.jl:56-60)."""
    g = golden("as_clouds.npz")
    spec = _as_spec(g)
    got, want = _initial_draw_case(eng, spec, 3000, 42)
    assert np.array_equal(got, want)
    assert np.all(np.isfinite(got[:, 16])) and np.all(got[:, 13:16] == spec.values[13:16])
    # the prior puts ~2 % of its mass on psi_1 < 1 (indeterminacy): those draws were replaced
    assert got[:, 2].min() > 0.9
    assert abs(got[:, 0].mean() - 2.0) < 0.05 and abs(got[:, 2].mean() - 1.5) < 0.03


def test_initial_draw_failure_is_reported(eng):
    """A likelihood that is never finite (sigma fixed at a non-positive value) exhausts max_tries."""
    data, X = W.synthetic_three_equation(T=20)
    ps = W.three_equation_parameters()
    ps[2] = M.parameter("σ1", -1.0, (-1.0, -1.0), (-1.0, -1.0), None, None, fixed=True)
    spec = M.make_spec(ps, M.LinearEquationsLogLik(data, X))
    eng.cloud_create(256, 9)
    eng.set_model(spec)
    with pytest.raises(ValueError):
        eng.initial_draw(spec.values, 1, 3)
