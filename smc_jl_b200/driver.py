"""`smc(loglikelihood, parameters, data; ...)` -- host driver with the reference's signature
(src/smc_main.jl:118-161), keyword names and defaults.  It only sequences C-ABI calls: every
per-particle operation of the stage loop (src/smc_main.jl:377-497) runs in the CUDA engine.

Differences from the reference, all at the edges of the hot path:
  * `loglikelihood` / `old_loglikelihood` are device likelihood descriptors (smc_jl_b200.model);
    `data` / `old_data` are accepted for signature parity (the descriptors carry the data);
  * the reference returns `nothing` and writes files; this driver writes the same files -- `savepath` as JLD2 with the keys
    `cloud` (type tag `SMC.Cloud`), `w`, `W` (intermediate checkpoints add `j`) and `particle_store_path` as HDF5 with the
    dataset `smcparams` (src/smc_main.jl:499-526; smc_jl_b200/jld2.py; a path ending in `.npz` selects numpy containers
    with the same keys) -- and also returns `(cloud, w, W)`;
  * `parallel=true` / Distributed.jl workers become `n_gpus=G`: one process per GPU, the cloud sharded in contiguous ranges,
    every per-stage reduction and the post-resample row exchange done by the engine over NVLink.  Under `torchrun` the ranks
    of the job are used as they are; otherwise G worker processes are spawned.  Results do not depend on G;
  * `resampling_method=:polyalgo` (StatsBase's alias-table sampler: i.i.d. categorical draws) is served by the
    multinomial kernel -- same distribution, different (unpinned) random stream; regime switching is not available;
  * randomness is the engine's Philox stream keyed by `seed` (the reference uses the global dSFMT).
Resuming (`continue_intermediate`) restores exactly what the reference's checkpoint holds (cloud, w, W, j); like the reference
it restarts `resampled_last_period = false`, so on an ADAPTIVE schedule a run resumed right after a resampling stage can choose
its next phi differently from the uninterrupted run (fixed schedules resume bit for bit).
"""
import os
import tempfile
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
    resample, reset.  Leaves the bridged cloud on the device (weights 1, ESS = n_parts).
    On a sharded engine every rank builds the same global matrix (its own GPU does the old-cloud resampling and the prior
    draws redundantly: both are keyed on global indices) and uploads its rows."""
    d = spec.d
    n_res = int(round((1.0 - prior_weight) * n_parts))
    n_prior = n_parts - n_res
    parts = []
    side = eng if eng.world == 1 else Engine(eng.device)       # stage-0 work on whole (small) clouds: a plain one-GPU context
    try:
        if n_res > 0:
            inds = side.resample_weights(np.ascontiguousarray(old_cloud.particles[:, -1]), resampling_method, seed=seed, stage=0,
                                         n_parts=n_res)
            parts.append(np.asarray(old_cloud.particles)[inds - 1, :])
        if n_prior > 0:
            side.cloud_create(n_prior, d)
            side.set_model(old_spec)                       # old_loglikelihood on old_data, current prior (:286-299)
            side.initial_draw(old_spec.values, seed ^ 0x9E3779B9, 1000)
            parts.append(side.download())
    finally:
        if side is not eng:
            side.close()
    P = np.asfortranarray(np.vstack(parts))
    eng.cloud_create(n_parts, d)
    eng.set_model(spec)
    eng.upload(P)
    eng.evaluate(1)                                   # initialize_likelihoods! (:307)
    ll = eng.read_column(d)
    w = eng.read_column(d + 4)
    w[ll == -np.inf] = 0.0                            # zero_bad_loglh_weights! (particle.jl:392-399)
    eng.write_column(d + 4, w)
    eng.resample(resampling_method, seed=seed, stage=1)   # normalize_weights! + resample + reset_weights! (:315-322)


def _save(path, cloud, w, W, j=None):
    """`jldopen(path, ...) do file; write(file, "cloud", cloud); write(file, "w", w); write(file, "W", W)[; write(file, "j", j)]`
    (src/smc_main.jl:499-507,521-525)."""
    if path.endswith(".npz"):
        np.savez(path, particles=cloud.particles, tempering_schedule=cloud.tempering_schedule, ESS=cloud.ESS,
                 stage_index=cloud.stage_index, n_Phi=cloud.n_Φ, resamples=cloud.resamples, c=cloud.c, accept=cloud.accept,
                 total_sampling_time=cloud.total_sampling_time, **({} if j is None else {"j": j}),
                 **({"w": w, "W": W} if w is not None else {}))
    else:
        from .jld2 import write_jld2
        write_jld2(path, cloud, w, W, j)


def load_cloud(path):
    """`load(path, "cloud")` (+ "w", "W", "j" when present) for the files this driver -- or the reference -- writes;
    returns (cloud, w, W, j)."""
    if path.endswith(".npz"):
        z = np.load(path)
        cloud = Cloud(np.asfortranarray(z["particles"]), z["tempering_schedule"], z["ESS"], int(z["stage_index"]), int(z["n_Phi"]),
                      int(z["resamples"]), float(z["c"]), float(z["accept"]), float(z["total_sampling_time"]))
        return cloud, (z["w"] if "w" in z.files else None), (z["W"] if "W" in z.files else None), (int(z["j"]) if "j" in z.files else 2)
    from .jld2 import read_jld2
    r = read_jld2(path)
    return r["cloud"], r.get("w"), r.get("W"), int(r.get("j", 2))


def _stage_path(savepath, i):
    base, ext = os.path.splitext(savepath)
    return "%s_stage=%d%s" % (base, i, ext)


# ---- process group of a sharded run (one process per GPU) ----------------------------------------------------------------------
class _Group:
    """rank / world of this process plus the little host-side communication a sharded run needs (communicator id, gathering the
    shards of the results on rank 0) over torch.distributed (gloo or the job's own backend)."""

    def __init__(self, rank=0, world=1, dist=None, tmpdir=None):
        self.rank, self.world, self.dist, self.tmpdir = rank, world, dist, tmpdir

    def comm_id(self):
        obj = [Engine.unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def gather_rows(self, a, tag):
        """row blocks of every rank (rank order) -> the concatenated array on rank 0, None elsewhere (through files in a
        directory shared by the ranks of one node: the shards can be hundreds of MB)."""
        if self.world == 1:
            return a
        d = [self.tmpdir if self.rank == 0 else None]
        self.dist.broadcast_object_list(d, src=0)
        path = os.path.join(d[0], "%s_%d.npy" % (tag, self.rank))
        np.save(path, np.asarray(a))
        self.dist.barrier()
        out = None
        if self.rank == 0:
            out = np.concatenate([np.load(os.path.join(d[0], "%s_%d.npy" % (tag, r))) for r in range(self.world)], axis=0)
        self.dist.barrier()
        os.remove(path)
        return out


def _worker(rank, world, port, tmpdir, args, kwargs, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kwargs = dict(kwargs)
        kwargs["device"] = rank
        out = _smc(_Group(rank, world, dist, tmpdir), *args, **kwargs)
        if rank == 0:
            cloud, w, W = out
            np.save(os.path.join(tmpdir, "ret_particles.npy"), cloud.particles)
            if w is not None:
                np.save(os.path.join(tmpdir, "ret_w.npy"), w)
                np.save(os.path.join(tmpdir, "ret_W.npy"), W)
            cloud.particles = None
            ret["cloud"] = cloud
        dist.barrier()
    finally:
        dist.destroy_process_group()


def nan_ess_message(incremental_weights, normalized_weights):
    """The assertion text of check_nan_ess (src/helpers.jl:270-305)."""
    inc, nw = np.asarray(incremental_weights, float), np.asarray(normalized_weights, float)
    text = "No particles have non-zero weight."
    if np.any(np.isinf(inc)):
        text += " Some particles have approximately infinite log-likelihoods."
    if np.any(np.isnan(inc)):
        text += " Some particles have approximately NaN log-likelihoods."
    with np.errstate(all="ignore"):
        q = float(np.sum(nw ** 2))
    if q <= np.finfo(float).eps:
        text += " The squared sum of the normalized weights is at machine-error."
    if np.isnan(q):
        text += " The squared sum of the normalized weights is returning a NaN."
        if np.any(np.isnan(nw)):
            text += " Part of the reason is that one of the normalized weights is a NaN"
    return text


def check_nan_ess(cloud, incremental_weights, normalized_weights, savepath, debug_assertion):
    """check_nan_ess (src/helpers.jl:270-305) once the engine has reported a NaN ESS: with `debug_assertion` the incremental
    weights, the normalised weights and the cloud go to `<savepath>_debug_assertion.jld2` (`.npz` paths: numpy container);
    always raises the reference's AssertionError."""
    text = nan_ess_message(incremental_weights, normalized_weights)
    if debug_assertion:
        if savepath.endswith(".npz"):
            np.savez(savepath.replace(".npz", "_debug_assertion.npz"), incremental_weights=incremental_weights,
                     normalized_weights=normalized_weights, particles=cloud.particles)
        else:
            from .jld2 import write_jld2
            write_jld2(savepath.replace(".jld2", "_debug_assertion.jld2"), cloud, np.asarray(incremental_weights, float),
                       np.asarray(normalized_weights, float), array_keys=("incremental_weights", "normalized_weights"))
    raise AssertionError(text)


def smc(loglikelihood, parameters, data=None, *, n_gpus=1, **kwargs):
    """See the module docstring; keyword arguments as in src/smc_main.jl:118-161 plus `seed`, `device`, `weight_history`,
    `engine` (an existing single-GPU Engine) and `n_gpus`."""
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if world_env > 1 and kwargs.get("engine") is None and (n_gpus in (1, world_env) or kwargs.get("parallel")):
        # launched under torchrun: the job's ranks are the GPUs
        import torch.distributed as dist
        own = not dist.is_initialized()
        if own:
            dist.init_process_group("gloo")
        try:
            tmp = tempfile.mkdtemp(prefix="smcb200_") if dist.get_rank() == 0 else None
            kwargs.setdefault("device", int(os.environ.get("LOCAL_RANK", "0")))
            return _smc(_Group(dist.get_rank(), dist.get_world_size(), dist, tmp), loglikelihood, parameters, data, **kwargs)
        finally:
            if own:
                dist.destroy_process_group()
    if n_gpus <= 1:
        return _smc(_Group(), loglikelihood, parameters, data, **kwargs)
    if kwargs.get("engine") is not None:
        raise ValueError("n_gpus > 1 creates its own engines (one process per GPU)")
    import multiprocessing as mp
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory(prefix="smcb200_") as tmp, ctx.Manager() as man:
        ret = man.dict()
        procs = [ctx.Process(target=_worker, args=(r, n_gpus, port, tmp, (loglikelihood, parameters, data), kwargs, ret)) for r in range(n_gpus)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        if any(p.exitcode != 0 for p in procs) or "cloud" not in ret:
            raise RuntimeError("a GPU worker of the sharded smc() run failed (exit codes %r)" % [p.exitcode for p in procs])
        cloud = ret["cloud"]
        cloud.particles = np.asfortranarray(np.load(os.path.join(tmp, "ret_particles.npy")))
        w = W = None
        if os.path.exists(os.path.join(tmp, "ret_w.npy")):
            w, W = np.load(os.path.join(tmp, "ret_w.npy")), np.load(os.path.join(tmp, "ret_W.npy"))
        return cloud, w, W


def _smc(grp, loglikelihood, parameters, data=None, *, verbose="low", testing=False, data_vintage="", parallel=False,
         n_parts=5_000, n_blocks=1, n_mh_steps=1, λ=2.1, n_Φ=300, resampling_method="systematic",
         threshold_ratio=0.5, c=0.5, α=1.0, target=0.25, use_fixed_schedule=True, tempering_target=0.97,
         old_data=None, old_cloud=None, old_loglikelihood=None, old_vintage="", smc_iteration=1, run_test=False,
         filestring_addl=(), loadpath="", savepath="smc_cloud.jld2", particle_store_path="smcsave.h5",
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
    lead = grp.rank == 0
    if not lead:
        verbose = "none"

    tempered_update = old_loglikelihood is not None or (old_data is not None and np.size(old_data) > 0)
    if tempered_update and old_loglikelihood is None:
        raise TypeError("tempered update: pass old_loglikelihood as a device descriptor built on old_data")
    spec = make_spec(parameters, loglikelihood, old_loglikelihood if tempered_update else None)
    n_para = spec.d
    if spec.n_free == 0:
        raise AssertionError("All model parameters are fixed!")            # smc_main.jl:237

    own = engine is None
    if not own and (grp.world > 1 or engine.world > 1):
        raise ValueError("smc(engine=...) takes a single-GPU Engine; use n_gpus=G (or torchrun) for a sharded run")
    eng = engine or Engine(device)
    try:
        if grp.world > 1:
            eng.comm_init(grp.rank, grp.world, grp.comm_id())
        eng.cloud_create(n_parts, n_para)
        eng.set_model(spec)
        lo, hi = eng.first, eng.first + eng.count          # this rank's rows of the global cloud
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
                W_hist = [(w0 * n_parts if w0.sum() <= 1.0 else w0.copy())[lo:hi]]
            else:                                                         # bridge, smc_main.jl:260-329
                old_spec = make_spec(parameters, old_loglikelihood)
                bridge_cloud(eng, spec, old_spec, old_cloud, n_parts, tempered_update_prior_weight, resampling_method, seed)
                cloud = Cloud.empty(n_para, n_parts)
                ess0 = float(n_parts)                                     # push!(cloud.ESS, n_parts), :325
                W_hist = [np.ones(hi - lo)]
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
            W_hist = [np.ones(hi - lo)]
        schedule = ((np.arange(1, n_Φ + 1) - 1.0) / (n_Φ - 1.0)) ** λ          # smc_main.jl:348-352
        if resumed is None:
            w_hist = [np.zeros(hi - lo)]
            cloud.ESS = np.array([ess0])
            cloud.stage_index, cloud.n_Φ, cloud.resamples, cloud.c, cloud.accept = 1, n_Φ, 0, c, target
            cloud.total_sampling_time = 0.0
            cloud.tempering_schedule = schedule.copy() if use_fixed_schedule else np.zeros(1)
            state = StageState(c=c, accept=target, ess_prev=ess0, phi_prop=0.0, j=2, resampled_last_period=0)
            ess_list, sched_list = [ess0], [0.0]
            i, phi_n = 1, 0.0
        else:                                                                  # resume: i, c, j, phi_prop from the checkpoint
            w_old, W_old, j_old = resumed
            w_hist = [w_old[lo:hi, k] for k in range(w_old.shape[1])] if w_old is not None else []
            W_hist = [W_old[lo:hi, k] for k in range(W_old.shape[1])] if W_old is not None else []
            i = cloud.stage_index
            ess_list = [float(v) for v in cloud.ESS]
            sched_list = [float(v) for v in (schedule[:i] if use_fixed_schedule else cloud.tempering_schedule)]
            phi_n = sched_list[-1]
            # resampled_last_period restarts as false, as in the reference (smc_main.jl:202 is not part of the checkpoint)
            state = StageState(c=cloud.c, accept=cloud.accept, ess_prev=ess_list[-1], phi_prop=float(schedule[j_old - 1]), j=j_old,
                               resampled_last_period=0)
        _println(verbose, "low", "\n\n SMC recursion starts... \n\n")

        def snapshot():
            """the global cloud + history on rank 0 (None elsewhere)"""
            P = grp.gather_rows(eng.download(), "particles")
            w = W = None
            if weight_history:
                w = grp.gather_rows(np.column_stack(w_hist), "w")
                W = grp.gather_rows(np.column_stack(W_hist), "W")
            if lead:
                cloud.particles = np.asfortranarray(P)
                cloud.ESS = np.array(ess_list)
                if not use_fixed_schedule:
                    cloud.tempering_schedule = np.array(sched_list)
            return w, W

        # ---- recursion (smc_main.jl:377-508) ---------------------------------------------------------------
        # Stages run in batches through smcb200_run_stages: on a fixed schedule a whole batch is enqueued without the host in
        # the loop; a batch ends where the reference would look at the cloud (checkpoint, run_test, phi_n = 1).
        while phi_n < 1.0:
            t0 = time.perf_counter()
            n_batch = 1
            if use_fixed_schedule and verbose == "none":
                n_batch = n_Φ - i
                if save_intermediate:
                    n_batch = min(n_batch, intermediate_stage_increment - (i % intermediate_stage_increment))
                if run_test:
                    n_batch = min(n_batch, max(1, 3 - i))
            cfg = StageConfig(phi_n1=sched_list[-1], phi_n=0.0,
                              threshold_ratio=threshold_ratio, target=target, alpha=α, tempering_target=tempering_target,
                              prior_weight=tempered_update_prior_weight, log_prob_old_data=log_prob_old_data,
                              n_mh_steps=n_mh_steps, n_blocks=n_blocks, resample_method=RESAMPLERS[resampling_method],
                              adaptive=0 if use_fixed_schedule else 1, has_old_data=1 if tempered_update else 0,
                              seed=seed, stage=i + 1)
            inc_h = np.zeros((n_batch, eng.count)) if weight_history else None
            nw_h = np.zeros((n_batch, eng.count)) if weight_history else None
            try:
                results = eng.run_stages(cfg, state, schedule, i + 1, n_batch, inc_hist=inc_h, normw_hist=nw_h)
            except AssertionError:
                # NaN ESS (smc_main.jl:430 -> check_nan_ess): the failing stage left its normalised weights in the cloud and,
                # when the history is on, its incremental weights in the stream
                k_bad = len(eng.last_results)
                nw_bad = grp.gather_rows(eng.read_column(-1)[:, None], "nan_W")
                inc_bad = grp.gather_rows((inc_h[k_bad] if inc_h is not None and k_bad < len(inc_h) else np.full(eng.count, np.nan))[:, None],
                                          "nan_w")
                P_bad = grp.gather_rows(eng.download(), "nan_particles")
                if lead:
                    cloud.particles = np.asfortranarray(P_bad)
                    cloud.ESS = np.array(ess_list + [np.nan])
                    check_nan_ess(cloud, inc_bad[:, 0], nw_bad[:, 0], savepath, debug_assertion)
                raise
            dt = (time.perf_counter() - t0) / max(len(results), 1)
            for k, res in enumerate(results):
                i += 1
                phi_n = res.phi_n
                sched_list.append(phi_n)
                ess_list.append(res.ess)
                cloud.resamples += res.resampled
                cloud.c, cloud.accept, cloud.stage_index = res.c, res.accept, i
                if weight_history:
                    w_hist.append(inc_h[k])
                    W_hist.append(nw_h[k])
                cloud.total_sampling_time += dt
                _println(verbose, "low", " stage %4d  phi %.6g  c %.4f  accept %.4f  ESS %.1f  (%d resamples)"
                         % (i, phi_n, res.c, res.accept, res.ess, cloud.resamples))
            if run_test and i >= 3:
                break
            if save_intermediate and i % intermediate_stage_increment == 0:       # smc_main.jl:499-507 (also when testing)
                w, W = snapshot()
                if lead:
                    _save(_stage_path(savepath, i), cloud, w, W, int(state.j))
        w, W = snapshot()
        if lead and not testing:
            _save(savepath, cloud, w, W)
            if particle_store_path.endswith(".npz"):
                np.savez(particle_store_path, smcparams=cloud.particles[:, :n_para])
            else:
                from .jld2 import write_h5_matrix
                write_h5_matrix(particle_store_path, "smcparams", cloud.particles[:, :n_para])
        return (cloud, w, W) if lead else (None, None, None)
    finally:
        if own:
            eng.close()
