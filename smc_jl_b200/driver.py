"""`smc(loglikelihood, parameters, data; ...)` -- host driver with the reference's signature
(src/smc_main.jl:118-161), keyword names and defaults.  It only sequences C-ABI calls: every
per-particle operation of the stage loop (src/smc_main.jl:377-497) runs in the CUDA engine.

Differences from the reference, all at the edges of the hot path:
  * `loglikelihood` / `old_loglikelihood` are device likelihood descriptors (smc_jl_b200.model);
    `data` / `old_data` are accepted for signature parity (the descriptors carry the data);
  * the reference returns `nothing` and writes JLD2/HDF5 files; this driver returns the final `Cloud`
    (plus `w`, `W` when `testing=False` it also writes them to `savepath` as .npz -- JLD2 writers are
    SURVEY 8(f)2 "next");
  * `resampling_method=:polyalgo`, regime switching and the prior-mixing bridge are not available
    (NotImplementedError);
  * randomness is the engine's Philox stream keyed by `seed` (the reference uses the global dSFMT).
"""
import time

import numpy as np

from ._lib import StageConfig, StageState
from .cloud import Cloud, cloud_isempty
from .engine import RESAMPLERS, Engine
from .model import make_spec


def _println(verbose, level, msg):
    order = {"none": 0, "low": 1, "high": 2}
    if order.get(verbose, 1) >= order[level]:
        print(msg)


def initial_draw(engine, spec, seed, max_tries=1000):
    """initial_draw! (src/initialization.jl:88-119) on the device: every particle draws its free parameters from
    the prior (inside valuebounds) until the log-likelihood is finite (one_draw, :43-60); old_loglh = 0, weight = 1."""
    engine.initial_draw(spec.values, seed, max_tries)


def smc(loglikelihood, parameters, data=None, *, verbose="low", testing=False, data_vintage="", parallel=False,
        n_parts=5_000, n_blocks=1, n_mh_steps=1, λ=2.1, n_Φ=300, resampling_method="systematic",
        threshold_ratio=0.5, c=0.5, α=1.0, target=0.25, use_fixed_schedule=True, tempering_target=0.97,
        old_data=None, old_cloud=None, old_loglikelihood=None, old_vintage="", smc_iteration=1, run_test=False,
        filestring_addl=(), loadpath="", savepath="smc_cloud.npz", particle_store_path="smcsave.npz",
        save_intermediate=False, intermediate_stage_increment=10, continue_intermediate=False,
        intermediate_stage_start=0, tempered_update_prior_weight=0.0, regime_switching=False, toggle=True,
        debug_assertion=False, log_prob_old_data=0.0, seed=1793, device=0, weight_history=True, engine=None):
    if regime_switching:
        raise NotImplementedError("regime switching is out of scope of the device engine")
    resampling_method = str(resampling_method).lstrip(":")
    if resampling_method not in RESAMPLERS:
        if resampling_method == "polyalgo":
            raise NotImplementedError(":polyalgo resampling (StatsBase.sample) has no device kernel")
        raise ValueError("Invalid resampler in SMC. Options are :systematic, :multinomial, or :polyalgo")
    if not (0.0 <= tempered_update_prior_weight <= 1.0):
        raise ValueError("The keyword tempered_update_prior_weight must be within the interval [0, 1] but is currently "
                         "set to %r" % (tempered_update_prior_weight,))

    tempered_update = old_loglikelihood is not None or (old_data is not None and np.size(old_data) > 0)
    if tempered_update and old_loglikelihood is None:
        raise TypeError("tempered update: pass old_loglikelihood as a device descriptor built on old_data")
    spec = make_spec(parameters, loglikelihood, old_loglikelihood if tempered_update else None)
    n_para = spec.d
    if spec.n_free == 0:
        raise AssertionError("All model parameters are fixed!")            # smc_main.jl:237

    own = engine is None
    eng = engine or Engine(device)
    try:
        eng.cloud_create(n_parts, n_para)
        eng.set_model(spec)
        _println(verbose, "low", "\n\n SMC " + ("testing " if testing else "") + "starts ....\n\n")

        # ---- initialisation (smc_main.jl:244-345) --------------------------------------------------------
        if tempered_update:
            if old_cloud is None or cloud_isempty(old_cloud):
                raise ValueError("tempered update needs old_cloud")
            if tempered_update_prior_weight != 0.0 or len(old_cloud) != n_parts:
                raise NotImplementedError("bridge initialisation with prior mixing / a different n_parts (smc_main.jl:260-329)")
            cloud = Cloud(np.array(old_cloud.particles, order="F", copy=True), ESS=np.array([old_cloud.ESS[-1]]))
            eng.upload(cloud.particles)
            eng.evaluate(1)                                               # initialize_likelihoods!
            ess0 = float(old_cloud.ESS[-1])                               # initialize_cloud_settings!(tempered_update=true)
            w0 = cloud.particles[:, -1]
            W_hist = [w0 * n_parts if w0.sum() <= 1.0 else w0.copy()]
        else:
            cloud = Cloud.empty(n_para, n_parts)
            initial_draw(eng, spec, seed)
            ess0 = float(n_parts)
            W_hist = [np.ones(n_parts)]
        w_hist = [np.zeros(n_parts)]
        cloud.ESS = np.array([ess0])
        cloud.stage_index, cloud.n_Φ, cloud.resamples, cloud.c, cloud.accept = 1, n_Φ, 0, c, target
        cloud.total_sampling_time = 0.0

        schedule = ((np.arange(1, n_Φ + 1) - 1.0) / (n_Φ - 1.0)) ** λ          # smc_main.jl:348-352
        cloud.tempering_schedule = schedule.copy() if use_fixed_schedule else np.zeros(1)
        state = StageState(c=c, accept=target, ess_prev=ess0, phi_prop=0.0, j=2, resampled_last_period=0)
        ess_list, sched_list = [ess0], [0.0]
        _println(verbose, "low", "\n\n SMC recursion starts... \n\n")

        # ---- recursion (smc_main.jl:377-508) ---------------------------------------------------------------
        i, phi_n = 1, 0.0
        while phi_n < 1.0:
            t0 = time.perf_counter()
            i += 1
            phi_n1 = sched_list[-1]
            cfg = StageConfig(phi_n1=phi_n1, phi_n=float(schedule[i - 1]) if use_fixed_schedule else 0.0,
                              threshold_ratio=threshold_ratio, target=target, alpha=α, tempering_target=tempering_target,
                              prior_weight=tempered_update_prior_weight, log_prob_old_data=log_prob_old_data,
                              n_mh_steps=n_mh_steps, n_blocks=n_blocks, resample_method=RESAMPLERS[resampling_method],
                              adaptive=0 if use_fixed_schedule else 1, has_old_data=1 if tempered_update else 0,
                              seed=seed, stage=i)
            res, inc, nw = eng.stage(cfg, state, schedule=schedule, want_inc=weight_history, want_normw=weight_history)
            phi_n = res.phi_n
            sched_list.append(phi_n)
            ess_list.append(res.ess)
            cloud.resamples += res.resampled
            cloud.c, cloud.accept, cloud.stage_index = res.c, res.accept, i
            if weight_history:
                w_hist.append(inc)
                W_hist.append(nw)
            cloud.total_sampling_time += time.perf_counter() - t0
            _println(verbose, "low", " stage %4d  phi %.6g  c %.4f  accept %.4f  ESS %.1f  (%d resamples)"
                     % (i, phi_n, res.c, res.accept, res.ess, cloud.resamples))
            if run_test and i == 3:
                break
        cloud.particles = eng.download()
        cloud.ESS = np.array(ess_list)
        if not use_fixed_schedule:
            cloud.tempering_schedule = np.array(sched_list)
        w = np.column_stack(w_hist) if weight_history else None
        W = np.column_stack(W_hist) if weight_history else None
        if not testing:
            np.savez(savepath, particles=cloud.particles, tempering_schedule=cloud.tempering_schedule, ESS=cloud.ESS,
                     stage_index=cloud.stage_index, n_Phi=cloud.n_Φ, resamples=cloud.resamples, c=cloud.c, accept=cloud.accept,
                     total_sampling_time=cloud.total_sampling_time, **({"w": w, "W": W} if weight_history else {}))
            np.savez(particle_store_path, smcparams=cloud.particles[:, :n_para])
        return cloud, w, W
    finally:
        if own:
            eng.close()
