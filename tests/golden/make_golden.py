#!/usr/bin/env python
"""Extract the reference's own known-answer vectors into small committed fixtures.

Run ONLY inside the build container (it reads /root/reference, which does not exist on
the GPU box):  python tests/golden/make_golden.py

Each output is an .npz under tests/golden/.  Source fixtures (all under
/root/reference/test/reference unless noted) and the reference test that owns them:

  proposal_densities.npz  proposal_densities_in.jld2 -> proposal_densities_output_version=150.jld2   (test/helpers.jl:101-127)
  compute_ess.npz         ess_inputs_version=150.jld2 -> ess_output_version=150.jld2                  (test/helpers.jl:133-175)
  solve_adaptive_phi.npz  solve_adaptive_phi.jld2 -> helpers_output_version=150.jld2                  (test/helpers.jl:15-53)
  linear_model_rows.npz   initial_draw_out_version=150.jld2 + test_data.h5                            (test/initialization.jl:26-60, test/modelsetup.jl:119-138)
  correction_history.npz  smc_cloud_fix=true_version=150.jld2 (w, W, ESS, schedule, resamples)        (test/smc.jl:13-87)
  mutation.npz            mutation_inputs.jld2 -> mutation_outputs_version=150.jld2                   (test/mutation.jl:1-59)
  init_likelihoods.npz    initialize_likelihood_out_version=150.jld2 / initial_draw_out               (test/initialization.jl)
  as_clouds.npz           AS-model clouds: (theta -> loglh, old_loglh, logprior) rows (test/save/..., solve_adaptive_phi.jld2) + 3x230 data
  capm_data.npz           examples/data/capm.jld2 (lik_data, market_data)                             (examples/capm_model/estimate_capm.jl:38-41)
  resample_structural.npz resample_version=150.jld2 (sys/multi index vectors; inputs are dSFMT draws => structural use only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
from fixture_reader import JLD2, read_h5_v0  # noqa: E402

REF = "/root/reference"
R = REF + "/test/reference/"


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print("%-28s %8.1f KB  %s" % (name, os.path.getsize(path) / 1024, sorted(kw)))


def upper_to_lower(chol):
    f = np.array(chol["factors"])
    assert chol["uplo"][-1:] == b"U"
    return np.triu(f).T.copy()


def main():
    # --- proposal densities ------------------------------------------------------------
    j = JLD2(R + "proposal_densities_in.jld2")
    o = JLD2(R + "proposal_densities_output_version=150.jld2")
    d = j["d_subset"]
    save("proposal_densities.npz", para_draw=j["para_draw"], para_subset=j["para_subset"],
         mu=d["μ"], Sigma=d["Σ"]["mat"], chol_lower=upper_to_lower(d["Σ"]["chol"]),
         c=j["c"], alpha=j["α"], q0=o["q0"], q1=o["q1"])

    # --- compute_ESS -------------------------------------------------------------------
    j = JLD2(R + "ess_inputs_version=150.jld2")
    o = JLD2(R + "ess_output_version=150.jld2")
    save("compute_ess.npz", loglh=j["loglh"], current_weights=j["current_weights"],
         phi_n=j["ϕ_n"], phi_n1=j["ϕ_n1"], old_loglh=j["old_loglh"], ess=o["ess"])

    # --- solve_adaptive_phi ------------------------------------------------------------
    j = JLD2(R + "solve_adaptive_phi.jld2")
    o = JLD2(R + "helpers_output_version=150.jld2")
    cl = j["cloud"]
    save("solve_adaptive_phi.npz", particles=cl["particles"], cloud_ESS=cl["ESS"],
         i=j["i"], j=j["j"], phi_prop=j["phi_prop"], phi_n1=j["phi_n1"],
         proposed_fixed_schedule=j["proposed_fixed_schedule"],
         tempering_target=j["tempering_target"],
         resampled_last_period=int(j["resampled_last_period"][0]),
         out_phi_n=o["phi_n"], out_j=o["j"], out_phi_prop=o["phi_prop"],
         out_resampled_last_period=int(o["resampled_last_period"][0]))

    # --- linear 3-equation test model: (theta -> loglh, logprior) rows -----------------
    h = read_h5_v0(R + "test_data.h5")
    c = JLD2(R + "initial_draw_out_version=150.jld2")["cloud"]
    save("linear_model_rows.npz", data=h["data"], X=h["X"], particles=c["particles"])

    # initialize_likelihoods! golden: old_loglh <- loglh, loglh re-evaluated
    try:
        c2 = JLD2(R + "initialize_likelihood_out_version=150.jld2")
        key = c2.keys()[0]
        save("init_likelihoods.npz", particles=c2[key]["particles"])
    except Exception as e:  # pragma: no cover
        print("init_likelihoods skipped:", e)

    # --- correction history of the full linear-model run -------------------------------
    j = JLD2(R + "smc_cloud_fix=true_version=150.jld2")
    cl = j["cloud"]
    w, W = j["w"], j["W"]
    n_stage = w.shape[1]
    # every one of the 119 correction stages: W[:, n-1], w[:, n] -> W[:, n] (float64, 5000 x 120 each; savez_compressed)
    keep = list(range(1, n_stage))
    save("correction_history.npz",
         ESS=cl["ESS"], tempering_schedule=cl["tempering_schedule"], resamples=cl["resamples"],
         n_parts=w.shape[0], n_stage=n_stage, c=cl["c"], accept=cl["accept"],
         stages=np.array(keep), w=w, W=W,
         final_particles=cl["particles"][::10].copy(),       # thinned: statistical anchor only
         final_mean=np.average(cl["particles"][:, :9], axis=0, weights=cl["particles"][:, -1]),
         sumsq_W=np.sum(W.astype(np.float64) ** 2, axis=0))

    # --- file-format golden: every metadata byte of the reference's JLD2 output (everything except the three big
    #     Float64 payloads), so that smc_jl_b200/jld2.py can be checked byte for byte without the reference tree -------------
    raw = open(R + "smc_cloud_fix=true_version=150.jld2", "rb").read()
    n, ncol, nst = cl["particles"].shape[0], cl["particles"].shape[1], w.shape[1]
    p0 = 512 + 4742 + 87                        # first byte of the particle payload (cloud object at 4623, array header 87 bytes)
    p1 = p0 + n * ncol * 8
    links = j._obj(j.root)["links"]
    w0 = 512 + links["w"] + 87
    w1 = w0 + n * nst * 8
    W0 = 512 + links["W"] + 87
    W1 = W0 + n * nst * 8
    save("jld2_structure.npz", head=np.frombuffer(raw[:p0], np.uint8), mid=np.frombuffer(raw[p1:w0], np.uint8),
         whdr=np.frombuffer(raw[w1:W0], np.uint8), tail=np.frombuffer(raw[W1:], np.uint8), shape=np.array([n, ncol, nst]),
         tempering_schedule=cl["tempering_schedule"], ESS=cl["ESS"],
         scalars=np.array([cl["stage_index"], cl["n_Φ"], cl["resamples"]]), fscalars=np.array([cl["c"], cl["accept"], cl["total_sampling_time"]]),
         file_size=len(raw))

    # --- mutation golden ---------------------------------------------------------------
    j = JLD2(R + "mutation_inputs.jld2")
    o = JLD2(R + "mutation_outputs_version=150.jld2")
    d = j["d"]
    save("mutation.npz", particles_in=j["particles"]["particles"], mu=d["μ"], Sigma=d["Σ"]["mat"],
         blocks_free=np.array(j["blocks_free"][0]), blocks_all=np.array(j["blocks_all"][0]),
         c=j["c"], alpha=j["α"], phi_n=j["ϕ_n"], phi_n1=j["ϕ_n1"], old_data=j["old_data"],
         particles_out=o["particles"]["particles"])

    # --- An-Schorfheide clouds: priors + Kalman likelihood known answers ---------------
    S = REF + "/test/save/output_data/an_schorfheide/ss0/estimate/raw/"
    data = JLD2(R + "one_draw_in.jld2")["data"]
    prior_draws = JLD2(R + "solve_adaptive_phi.jld2")["cloud"]["particles"]
    c600 = JLD2(S + "smc_cloud_vint=200218.jld2")["cloud"]["particles"]
    c1000 = JLD2(S + "smc_cloud_npart=1000_vint=000000.jld2")["cloud"]["particles"]
    # the An-Schorfheide ParameterVector (priors, bounds, fixed flags) serialised in one_draw_in.jld2
    plist = JLD2(R + "one_draw_in.jld2")["loglikelihood"]["m"]["parameters"]
    kind, p1, p2, lo, hi, fixed, value, keys = [], [], [], [], [], [], [], []
    for prm in plist:
        pv = (prm.get("prior") or {}).get("value") or {}
        if "α" in pv: k, a, b = 2, pv["α"], pv["θ"]          # Gamma(shape, scale)
        elif "a" in pv: k, a, b = 1, pv["a"], pv["b"]          # Uniform
        elif "μ" in pv: k, a, b = 0, pv["μ"], pv["σ"]          # Normal
        elif "ν" in pv: k, a, b = 3, pv["ν"], pv["τ"]          # RootInverseGamma
        else: k, a, b = 0, 0.0, 1.0
        kind.append(k); p1.append(a); p2.append(b)
        lo.append(prm["valuebounds"]["1"]); hi.append(prm["valuebounds"]["2"])
        fixed.append(int(prm["fixed"][0])); value.append(prm["value"]); keys.append(prm["key"])
    save("as_clouds.npz", data=data, prior_draws=prior_draws, cloud600=c600, cloud1000=c1000,
         prior_kind=np.array(kind, np.int32), prior_p1=np.array(p1), prior_p2=np.array(p2),
         lo=np.array(lo), hi=np.array(hi), fixed=np.array(fixed, np.int32), value=np.array(value),
         keys=np.array(keys))

    # --- CAPM data ---------------------------------------------------------------------
    j = JLD2(REF + "/examples/data/capm.jld2")
    save("capm_data.npz", lik_data=j["lik_data"], market_data=j["market_data"])

    # --- resample (structural only) ----------------------------------------------------
    j = JLD2(R + "resample_version=150.jld2")
    save("resample_structural.npz", sys=j["sys"], multi=j["multi"])


if __name__ == "__main__":
    main()
