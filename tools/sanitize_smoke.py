#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family, sized for `compute-sanitizer` (memcheck / racecheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Covers: device initial_draw!, evaluate, correction, the multi-trial adaptive-phi solve, systematic and multinomial
selection (+ resample(weights; n_parts)), both moment passes, proposal preparation, the mutation kernel in its
single-block / multi-block / mixture / old-data variants (Gaussian-regression and An-Schorfheide functors).
Ragged sizes on purpose (tails of tiles, chunks and scan blocks)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smc_jl_b200 import model as M  # noqa: E402
from smc_jl_b200 import workloads as W  # noqa: E402
from smc_jl_b200._lib import StageConfig, StageState  # noqa: E402
from smc_jl_b200.engine import Engine  # noqa: E402


def run(eng, spec, N, n_stage, **kw):
    eng.cloud_create(N, spec.d)
    eng.set_model(spec)
    eng.initial_draw(spec.values, 7, 200)
    if kw.get("has_old"):
        eng.evaluate(1)
    sched = (np.arange(30) / 29.0) ** 2.1
    state = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2)
    phi = 0.0
    for s in range(n_stage):
        cfg = StageConfig(phi_n1=phi, phi_n=float(sched[s + 1]), threshold_ratio=kw.get("thr", 0.9), target=0.25, alpha=kw.get("alpha", 1.0),
                          tempering_target=0.9, n_mh_steps=kw.get("n_mh", 1), n_blocks=kw.get("n_blocks", 1), resample_method=s % 2,
                          adaptive=kw.get("adaptive", 0), has_old_data=kw.get("has_old", 0), seed=3, stage=s + 2)
        res, inc, nw = eng.stage(cfg, state, schedule=sched, want_inc=True, want_normw=True)
        phi = res.phi_n
        if phi >= 1.0:
            break
    mean, cov = eng.moments()
    assert np.all(np.isfinite(mean)) and np.all(np.isfinite(cov))
    return res


eng = Engine(0)
params, lk, _ = W.linear_gaussian(d=20, T=64, prior_sd=1.0)
run(eng, M.make_spec(params, lk), 5000 + 37, 4, n_mh=2)                                   # single full block (BLK = 2)
params8, lk8, _ = W.linear_gaussian(d=8, T=32, prior_sd=1.0)
run(eng, M.make_spec(params8, lk8), 3000 + 5, 4, n_blocks=3, alpha=0.9, adaptive=1)        # blocks + mixture + adaptive phi
data, X = W.synthetic_three_equation(T=60)
spec3 = M.make_spec(W.three_equation_parameters(prior_para=10.0), M.LinearEquationsLogLik(data, X), M.LinearEquationsLogLik(data[:, :30], X))
run(eng, spec3, 2000 + 11, 3, n_blocks=2, alpha=0.9, has_old=1)                            # old data, generic priors
g = np.load(os.path.join(ROOT, "tests", "golden", "as_clouds.npz"))
run(eng, M.make_spec(W.an_schorfheide_parameters(), M.AnSchorfheideLogLik(g["data"][:, :40])), 1000 + 3, 2, n_mh=1, alpha=0.9)
w = np.random.default_rng(0).gamma(0.5, 1.0, 4097)
for method in ("systematic", "multinomial"):
    for n_out in (100, 4097, 9000):
        idx = eng.resample_weights(w, method, seed=1, stage=2, n_parts=n_out)
        assert idx.min() >= 1 and idx.max() <= 4097
eng.close()
print("sanitize_smoke ok")
