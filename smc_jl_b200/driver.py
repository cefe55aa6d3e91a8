"""`smc(loglikelihood, parameters, data; ...)` -- host driver with the reference's signature
(src/smc_main.jl:118-161), keyword names and defaults.  It only sequences C-ABI calls: every
per-particle operation of the stage loop (src/smc_main.jl:377-497) runs in the CUDA engine.

Differences from the reference, all at the edges of the hot path:
  * `loglikelihood` / `old_loglikelihood` are device likelihood descriptors (smc_jl_b200.model);
    `data` / `old_data` are accepted for signature parity (the descriptors carry the data);
  * the reference returns `nothing` and writes JLD2/HDF5 files; this driver returns the final `Cloud`
    (plus `w`, `W` when `testing=False` it also writes them to `savepath` as .npz -- JLD2 writers are
    SURVEY 8(f)2 "next");
  * `resampling_method=:polyalgo` (StatsBase's alias-table sampler: i.i.d. categorical draws) is served by the
    multinomial kernel -- same distribution, different (unpinned) random stream; regime switching is not available;
  * checkpoints (`save_intermediate`, `continue_intermediate`, smc_main.jl:334-361,499-507) are `.npz` files with the
    reference's keys (`cloud` fields, `w`, `W`, `j`);
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


def bridge_cloud(eng, spec, old_spec, old_cloud, n_parts, prior_weight, resampling_method, seed):
    """Bridge initialisation of a tempered update (src/smc_main.jl:260-329): resample round((1 - pw) n_parts)
    particles from the old cloud, draw the rest from the CURRENT prior scored with the OLD likelihood on the old data
    (device initial_draw!), initialize_likelihoods! on the new data, zero the weights of -Inf particles, normalise,
    resample, reset.  Leaves the bridged cloud on the device (weights 1, ESS = n_parts)."""
    if eng.world > 1:
        raise NotImplementedError("bridge initialisation runs on one GPU (stage-0 set-up); shard the run after it")
    d = spec.d
    n_res = int(round((1.0 - prior_weight) * n_parts))
    n_prior = n_parts - n_res
    parts = []
    if n_res > 0:
        inds = eng.resample_weights(np.ascontiguousarray(old_cloud.particles[:, -1]), resampling_method, seed=seed, stage=0,
                                    n_parts=n_res)
        parts.append(np.asarray(old_cloud.particles)[inds - 1, :])
    if n_prior > 0:
        eng.cloud_create(n_prior, d)
        eng.set_model(old_spec)                       # old_loglikelihood on old_data, current prior (:286-299)
        eng.initial_draw(old_spec.values, seed ^ 0x9E3779B9, 1000)
        parts.append(eng.download())
    P = np.asfortranarray(np.vstack(parts))
    eng.cloud_create(n_parts, d)
    eng.set_model(spec)
    eng.upload(P)
    eng.evaluate(1)                                   # initialize_likelihoods! (:307)
    ll = eng.read_column(d)
    w = eng.read_column(d + 4)
    w[~(ll > -np.inf)] = 0.0                          # zero_bad_loglh_weights! (particle.jl:392-399)
    eng.write_column(d + 4, w)
    eng.resample(resampling_method, seed=seed, stage=1)   # normalize_weights! + resample + reset_weights! (:315-322)


def _save_checkpoint(path, cloud, w, W, j):
    np.savez(path, particles=cloud.particles, tempering_schedule=cloud.tempering_schedule, ESS=cloud.ESS,
             stage_index=cloud.stage_index, n_Phi=cloud.n_Φ, resamples=cloud.resamples, c=cloud.c, accept=cloud.accept,
             total_sampling_time=cloud.total_sampling_time, j=j, **({"w": w, "W": W} if w is not None else {}))


def load_cloud(path):
    """`load(path, "cloud")` for the .npz files this driver writes; returns (cloud, w, W, j)."""
    z = np.load(path)
    cloud = Cloud(np.asfortranarray(z["particles"]), z["tempering_schedule"], z["ESS"], int(z["stage_index"]), int(z["n_Phi"]),
                  int(z["resamples"]), float(z["c"]), float(z["accept"]), float(z["total_sampling_time"]))
    return cloud, (z["w"] if "w" in z.files else None), (z["W"] if "W" in z.files else None), (int(z["j"]) if "j" in z.files else 2)


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
        resumed = None
        if tempered_update:
            if old_cloud is None or cloud_isempty(old_cloud):
                if not loadpath:
                    raise ValueError("tempered update needs old_cloud (or loadpath)")
                old_cloud = load_cloud(loadpath)[0]                       # smc_main.jl:246
            if tempered_update_prior_weight == 0.0 and len(old_cloud) == n_parts:
                cloud = Cloud(np.array(old_cloud.particles, order="F", copy=True), ESS=np.array([old_cloud.ESS[-1]]))
                eng.upload(cloud.particles)
                eng.evaluate(1)                                           # initialize_likelihoods!
                ess0 = float(old_cloud.ESS[-1])                           # initialize_cloud_settings!(tempered_update=true)
                w0 = cloud.particles[:, -1]
                W_hist = [w0 * n_parts if w0.sum() <= 1.0 else w0.copy()]
            else:                                                         # bridge, smc_main.jl:260-329
                old_spec = make_spec(parameters, old_loglikelihood)
                bridge_cloud(eng, spec, old_spec, old_cloud, n_parts, tempered_update_prior_weight, resampling_method, seed)
                cloud = Cloud.empty(n_para, n_parts)
                ess0 = float(n_parts)                                     # push!(cloud.ESS, n_parts), :325
                W_hist = [np.ones(n_parts)]
        elif continue_intermediate:                                       # smc_main.jl:334-335,355-361
            cloud, w_old, W_old, j_old = load_cloud(loadpath)
            if len(cloud) != n_parts or cloud.n_para != n_para:
                raise ValueError("checkpoint does not match n_parts / the ParameterVector")
            eng.upload(cloud.particles)
            resumed = (w_old, W_old, j_old)
            ess0 = float(cloud.ESS[-1])
            W_hist = []
        else:
            cloud = Cloud.empty(n_para, n_parts)
            initial_draw(eng, spec, seed)
            ess0 = float(n_parts)
            W_hist = [np.ones(n_parts)]
        schedule = ((np.arange(1, n_Φ + 1) - 1.0) / (n_Φ - 1.0)) ** λ          # smc_main.jl:348-352
        if resumed is None:
            w_hist = [np.zeros(n_parts)]
            cloud.ESS = np.array([ess0])
            cloud.stage_index, cloud.n_Φ, cloud.resamples, cloud.c, cloud.accept = 1, n_Φ, 0, c, target
            cloud.total_sampling_time = 0.0
            cloud.tempering_schedule = schedule.copy() if use_fixed_schedule else np.zeros(1)
            state = StageState(c=c, accept=target, ess_prev=ess0, phi_prop=0.0, j=2, resampled_last_period=0)
            ess_list, sched_list = [ess0], [0.0]
            i, phi_n = 1, 0.0
        else:                                                                  # resume: i, c, j, phi_prop from the checkpoint
            w_old, W_old, j_old = resumed
            w_hist = [w_old[:, k] for k in range(w_old.shape[1])] if w_old is not None else []
            W_hist = [W_old[:, k] for k in range(W_old.shape[1])] if W_old is not None else []
            i = cloud.stage_index
            ess_list = [float(v) for v in cloud.ESS]
            sched_list = [float(v) for v in (schedule[:i] if use_fixed_schedule else cloud.tempering_schedule)]
            phi_n = sched_list[-1]
            # resampled_last_period restarts as false, as in the reference (smc_main.jl:202 is not part of the checkpoint)
            state = StageState(c=cloud.c, accept=cloud.accept, ess_prev=ess_list[-1], phi_prop=float(schedule[j_old - 1]), j=j_old,
                               resampled_last_period=0)
        _println(verbose, "low", "\n\n SMC recursion starts... \n\n")

        # ---- recursion (smc_main.jl:377-508) ---------------------------------------------------------------
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
            if save_intermediate and i % intermediate_stage_increment == 0 and not testing:   # smc_main.jl:499-507
                cloud.particles = eng.download()
                cloud.ESS = np.array(ess_list)
                if not use_fixed_schedule:
                    cloud.tempering_schedule = np.array(sched_list)
                base = savepath[:-4] if savepath.endswith(".npz") else savepath
                _save_checkpoint("%s_stage=%d.npz" % (base, i), cloud,
                                 np.column_stack(w_hist) if weight_history else None,
                                 np.column_stack(W_hist) if weight_history else None, int(state.j))
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
