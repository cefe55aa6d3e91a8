#!/usr/bin/env python
"""bench.py -- particle-MH-steps/sec per SMC stage (BASELINE.json metric) + ESS match vs the CPU reference path.

Default workload = BASELINE config C2 (the configuration the metric is quoted on): linear-Gaussian log-likelihood,
20 parameters, n_particles = 2^20 per GPU, n_mh_steps = 3, fixed tempering schedule (n_Phi = 300, lambda = 2.1),
systematic resampling.  `--config c3|c4|c5` runs the other BASELINE configs the same way (one global cloud of the named
size, sharded over the GPUs).

A "step" is one full SMC stage (correction -> selection -> moments -> mutation) over the whole cloud.
  value   = N n_mh_steps n_blocks / (device time per stage): K stages issued through smcb200_run_stages with the cloud
            resident in HBM, CUDA events on the library's launch stream, max over ranks.
  e2e     = the same K stages through the call `smc()` itself makes -- smcb200_run_stages with the w / W history columns
            (src/smc_main.jl:419-420) and the stage summaries streaming into pinned HOST memory every stage -- wall clock
            around the call (host -> device: stage configuration; device -> host: 2 N doubles + the summary per stage).
  ess_match = the first stages of the SAME global cloud through the CPU oracle (oracle/, the restated reference path) and
            through the engine at this GPU count, outside the timed region: ESS / accept / c trajectories and a checksum
            of every rank's shard of the final cloud against the oracle's rows.
  roofline / roofline_fp64 = the mutation kernel against the measured HBM bandwidth and the measured FP64 FMA peak.
`--impl reference` times the CPU restatement of the reference path (oracle/, OpenMP on all host cores) on the same
configuration (Julia is not installed: kind "port").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl ours|reference]
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line

D, T, N_FULL, N_MH, N_BLOCKS, N_PHI, LAM = 20, 256, 1 << 20, 3, 1, 300, 2.1
SEED = 1793
METRIC = "particle_mh_steps_per_sec_per_stage"
UNIT = "particle-MH-steps/s"
FIRST_STAGE = 30   # C2: timed stages start here in the 300-point schedule (past the burn-in of the prior cloud)
MATCH_STAGES = 4   # stages compared with the CPU oracle (ess_match)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes of one mutation-kernel launch from the committed ncu capture (profiles/r02_ncu_summary.json)."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_summary.json")))
        k = j["kernels"]["k_mutate_c2"]
        return float(k["dram_bytes_read"]) + float(k["dram_bytes_write"]), "profiles/r02_ncu_summary.json (ncu --set full, one launch at N = 2^20)"
    except Exception:
        return None, "no ncu capture committed for this kernel build"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML, ~2 ms period; falls back to
    polling nvidia-smi when pynvml is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.index, self.sm, self.reason_bits, self.max_mhz = index, [], 0, None
        self.smi_rows, self.stop, self.th, self.nv = [], threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self.stop.is_set():
            try:
                if self.nv is not None:
                    self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM)))
                    try:
                        self.reason_bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    except Exception:
                        self.reason_bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    self.stop.wait(0.002)
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.smi_rows.append([x.strip() for x in out.split(",")])
                    self.stop.wait(0.05)
            except Exception:
                self.stop.wait(0.01)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=10)

    def summary(self):
        if self.nv is not None and self.sm:
            sm = sorted(self.sm)
            reasons = [n for n, b in self.BITS.items() if self.reason_bits & b]
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm), "via": "nvml"}
        if self.smi_rows:
            sm = sorted(float(r[0]) for r in self.smi_rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.smi_rows)]
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.smi_rows[0][1]), "reasons": reasons, "samples": len(sm),
                    "via": "nvidia-smi"}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"], "samples": 0}


def schedule(n_phi=N_PHI, lam=LAM):
    return ((np.arange(n_phi)) / (n_phi - 1.0)) ** lam


def make_model():
    from smc_jl_b200 import model as M
    from smc_jl_b200 import workloads as W
    params, lk, _ = W.linear_gaussian(d=D, T=T)
    return params, M.make_spec(params, lk)


def stage_cfg(sched, s, **kw):
    from smc_jl_b200._lib import StageConfig
    base = dict(phi_n1=float(sched[s]), phi_n=float(sched[s + 1]), threshold_ratio=0.5, target=0.25, alpha=1.0,
                tempering_target=0.95, n_mh_steps=N_MH, n_blocks=N_BLOCKS, resample_method=0, seed=SEED, stage=s + 2)
    base.update(kw)
    return StageConfig(**base)


# ------------------------------------------------------------------------------------------------------------------
# workloads: BASELINE.json configs
# ------------------------------------------------------------------------------------------------------------------
class Workload:
    """name, particles per GPU / in total, the stage template and how a run gets to its timed window."""

    def __init__(self, name, world):
        from smc_jl_b200 import model as M
        from smc_jl_b200 import workloads as W
        self.name, self.world = name, world
        self.kw = dict(threshold_ratio=0.5, target=0.25, alpha=1.0, tempering_target=0.95, n_mh_steps=1, n_blocks=1,
                       resample_method=0, adaptive=0, has_old_data=0)
        self.old_spec = None
        if name == "c2":
            self.params, self.spec = make_model()
            self.n_global, self.scaling = N_FULL * world, "weak"
            self.kw.update(n_mh_steps=N_MH, n_blocks=N_BLOCKS)
            self.sched, self.first = schedule(), FIRST_STAGE
            self.text = ("C2 linear-Gaussian loglik, 20 params, n_particles=2^20 per GPU, n_mh_steps=3, fixed phi schedule "
                         "(n_phi=300, lambda=2.1), systematic resampling")
        elif name == "c4":
            g = np.load(os.path.join(GOLDEN, "as_clouds.npz"))
            self.params = W.an_schorfheide_parameters()
            self.spec = M.make_spec(self.params, M.AnSchorfheideLogLik(g["data"]))
            self.n_global, self.scaling = 1 << 18, "strong"
            self.kw.update(n_mh_steps=5, alpha=0.9, adaptive=1, tempering_target=0.97)
            self.sched, self.first = schedule(), 2
            self.text = ("C4 An-Schorfheide DSGE, device Kalman-filter loglik (T=230), n_particles=2^18 in total, n_mh_steps=5, "
                         "13 free parameters, alpha=0.9, adaptive phi (tempering_target 0.97)")
        elif name in ("c3", "c5"):
            self.params = W.three_equation_parameters()
            if name == "c3":
                g = np.load(os.path.join(GOLDEN, "capm_data.npz"))
                full, old = M.CAPMLogLik(g["lik_data"], g["market_data"]), M.CAPMLogLik(g["lik_data"][:, :18], g["market_data"])
                self.kw.update(adaptive=1, tempering_target=0.97, threshold_ratio=0.5)
                self.sched = schedule()
                self.text = ("C3 examples/capm_model (9 params, 3 x 36 observations), n_particles=2^20 in total, adaptive phi "
                             "(tempering_target 0.97) + generalised tempering (old data = first 18 periods), 3 blocks, alpha=0.9")
            else:
                g = np.load(os.path.join(GOLDEN, "linear_model_rows.npz"))
                full, old = M.LinearEquationsLogLik(g["data"], g["X"]), M.LinearEquationsLogLik(g["data"][:, :50], g["X"])
                self.kw.update(threshold_ratio=1.0)
                self.sched = schedule(60)
                self.text = ("C5 online update (two data vintages, tempered_update), 3-equation model on test_data.h5 (old data = "
                             "first 50 of 100 periods), n_particles=2^20 in total, fixed 60-point schedule, threshold_ratio 1.0 "
                             "(resample-heavy: every stage whose weights are not exactly uniform resamples -- the two vintages "
                             "are so close that the ESS never falls below 0.99 N), 3 blocks, alpha=0.9")
            self.old_spec = M.make_spec(self.params, old)
            self.spec = M.make_spec(self.params, full, old)
            self.n_global, self.scaling = 1 << 20, "strong"
            self.kw.update(n_blocks=3, alpha=0.9, has_old_data=1)
            self.first = 2
        else:
            raise ValueError(name)
        self.d = self.spec.d
        self.steps_per_particle = self.kw["n_mh_steps"] * self.kw["n_blocks"]

    def cfg(self, i, phi_n1=0.0, **over):
        """StageConfig of the reference's loop index i (>= 2)."""
        from smc_jl_b200._lib import StageConfig
        kw = dict(self.kw)
        kw.update(over)
        fixed = not kw["adaptive"]
        return StageConfig(phi_n1=float(self.sched[i - 2]) if fixed else float(phi_n1), phi_n=float(self.sched[i - 1]) if fixed else 0.0,
                           seed=SEED, stage=i, **kw)

    # ---- initial cloud of a run over n_global particles: this rank's rows [first, first + count) --------------------
    def host_rows(self, first, count, n_global):
        """C2: the prior cloud is drawn on the host, block by block of 2^20 rows (seed = block index), so that any sharding
        of the same global cloud sees the same rows."""
        from smc_jl_b200 import workloads as W
        out = np.zeros((count, self.d + 5), order="F")
        blk = N_FULL
        b0, b1 = first // blk, (first + count - 1) // blk
        for b in range(b0, b1 + 1):
            n_b = min(blk, n_global - b * blk)
            P = W.initial_cloud(self.params, n_b, np.random.default_rng(b))
            lo, hi = max(first, b * blk), min(first + count, b * blk + n_b)
            out[lo - first:hi - first] = P[lo - b * blk:hi - b * blk]
        return out

    def prepare(self, eng, n_global):
        """Leaves the engine at stage index 1 of the timed model (cloud evaluated, weights 1).  Returns ess_prev."""
        from smc_jl_b200._lib import StageState
        if self.name == "c2":
            eng.set_model(self.spec)
            eng.upload(self.host_rows(eng.first, eng.count, n_global))
            eng.evaluate(0)
            return float(n_global)
        if self.name == "c4":
            eng.set_model(self.spec)
            eng.initial_draw(self.spec.values, SEED, 1000)
            return float(n_global)
        # c3 / c5: first vintage from the prior to phi = 1 on the old data (fixed 300-point schedule; not what is timed), then
        # initialize_likelihoods! on the full sample (src/smc_main.jl:244-258, initialization.jl:153-186)
        eng.set_model(self.old_spec)
        eng.initial_draw(self.old_spec.values, SEED, 1000)
        st = StageState(c=0.5, accept=0.25, ess_prev=float(n_global), phi_prop=0.0, j=2)
        sched_a = schedule()
        kw = dict(self.kw)
        kw.update(adaptive=0, has_old_data=0)
        from smc_jl_b200._lib import StageConfig
        res = eng.run_stages(StageConfig(phi_n1=0.0, phi_n=0.0, seed=SEED, stage=0, **kw), st, sched_a, 2, N_PHI - 1)
        eng.set_model(self.spec)
        eng.evaluate(1)
        return float(res[-1].ess)


def bind_near_gpu(local_rank):
    """Pins this rank's host threads to the CPUs NVML reports as local to its GPU, so that the pinned buffers of the history
    stream are first-touched on the GPU's own NUMA node (8 ranks x 16 MB per stage otherwise cross the socket link).
    Returns the previous affinity (restored before the CPU legs, which use every core) or None."""
    if os.environ.get("SMCB200_NO_NUMA_BIND") or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = local_rank
        if vis and all(t.strip().isdigit() for t in vis.split(",")):
            idx = int(vis.split(",")[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if cpus and cpus != old:
            os.sched_setaffinity(0, cpus)
            return old
    except Exception:
        pass
    return None


def make_engine(rank, world, local_rank, dist, torch):
    from smc_jl_b200.engine import Engine
    eng = Engine(local_rank)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Engine.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
    return eng


def shard_digest(a):
    """checksum of a shard's bytes in the Julia matrix layout (column-major)"""
    return hashlib.blake2b(np.asfortranarray(a).tobytes(order="F"), digest_size=16).hexdigest()


# ------------------------------------------------------------------------------------------------------------------
# CPU oracle legs (test infrastructure: the checker of ess_match, the cpu_baseline and the reference arm)
# ------------------------------------------------------------------------------------------------------------------
class OracleRun:
    """The CPU restatement of the reference stage loop on one global cloud (C2 / C4 directly from stage 1; C3 / C5 after the
    first-vintage run-up)."""

    def __init__(self, wl, n, nthreads):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        self.O, self.L, self.wl, self.n = O, O.lib(), wl, n
        self.nthreads = nthreads if nthreads > 0 else self.L.orc_max_threads()
        self.mod = O.Model(wl.spec)
        kw = wl.kw
        if wl.name == "c2":
            self.buf = O.cloud_f(wl.host_rows(0, n, n))
            self.L.orc_evaluate(self.mod.h, self.buf, n)
            ess0 = float(n)
        elif wl.name == "c4":
            self.buf = np.zeros(n * (wl.d + 5))
            assert self.L.orc_initial_draw(self.mod.h, self.buf, n, 0, np.ascontiguousarray(wl.spec.values), SEED, 1000) == 0
            ess0 = float(n)
        else:
            mod_old = O.Model(wl.old_spec)
            self.buf = np.zeros(n * (wl.d + 5))
            assert self.L.orc_initial_draw(mod_old.h, self.buf, n, 0, np.ascontiguousarray(wl.old_spec.values), SEED, 1000) == 0
            io = O.StageIO(threshold_ratio=kw["threshold_ratio"], target=0.25, alpha=kw["alpha"], tempering_target=kw["tempering_target"],
                           n_mh_steps=kw["n_mh_steps"], n_blocks=kw["n_blocks"], resample_method=0, adaptive=0, has_old=0,
                           nthreads=self.nthreads, seed=SEED, c=0.5, accept=0.25, ess_prev=float(n), j=2)
            sa = schedule()
            scratch = np.zeros_like(self.buf)
            for i in range(2, N_PHI + 1):
                io.phi_n1, io.phi_n, io.stage = float(sa[i - 2]), float(sa[i - 1]), i
                assert self.L.orc_stage(mod_old.h, self.buf, scratch, n, sa, N_PHI, C.byref(io), None, None, None, None) == 0
            ess0 = io.ess
            self.L.orc_initialize_likelihoods(self.mod.h, self.buf, n)
        self.scratch = np.zeros_like(self.buf)
        self.io = O.StageIO(threshold_ratio=kw["threshold_ratio"], target=0.25, alpha=kw["alpha"], tempering_target=kw["tempering_target"],
                            n_mh_steps=kw["n_mh_steps"], n_blocks=kw["n_blocks"], resample_method=0, adaptive=kw["adaptive"],
                            has_old=kw["has_old_data"], nthreads=self.nthreads, seed=SEED, c=0.5, accept=0.25, ess_prev=ess0, j=2)
        self.phi = 0.0
        self.i = 1

    def stage(self):
        """one stage; returns (seconds, ess, accept, c, resampled, phi_n)"""
        wl, io = self.wl, self.io
        self.i += 1
        i = self.i
        sched = np.ascontiguousarray(wl.sched)
        io.phi_n1 = self.phi if wl.kw["adaptive"] else float(sched[i - 2])
        io.phi_n = 0.0 if wl.kw["adaptive"] else float(sched[i - 1])
        io.stage = i
        t0 = time.perf_counter()
        st = self.L.orc_stage(self.mod.h, self.buf, self.scratch, self.n, sched, len(sched), C.byref(io), None, None, None, None)
        dt = time.perf_counter() - t0
        assert st == 0, st
        self.phi = io.phi_out
        return dt, io.ess, io.accept, io.c, io.resampled, io.phi_out

    def cloud(self):
        return self.O.cloud_m(self.buf, self.n, self.wl.d)


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (restated in oracle/, see DESIGN.md: Julia and the reference's
    dependencies are not available in this image) on all host cores, on the SAME configuration as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.config, 1)
    cores = os.cpu_count() or 1                       # explicit: torchrun exports OMP_NUM_THREADS=1
    n = wl.n_global if wl.name in ("c2", "c4") else (1 << 16)
    run = OracleRun(wl, n, cores)
    warm = max(args.warmup, 1)
    k_timed = max(1, min(args.steps, 12))             # bounded sample: every stage costs ~0.5 s at N = 2^20 on 16 cores
    skip = (wl.first - 2) if wl.name == "c2" else 0
    for _ in range(min(skip, 6) + warm):              # short run-up + warm-up (stage cost does not depend on the schedule index)
        run.stage()
    times = [run.stage()[0] for _ in range(k_timed)]
    sec = float(np.mean(times))
    value = n * wl.steps_per_particle / sec
    sample = ("%d stages of %s at n_particles=%d after %d run-up stages (oracle/, OpenMP, %d threads); the per-stage cost does not "
              "depend on the position in the schedule" % (k_timed, wl.name.upper(), n, min(skip, 6) + warm, cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k_timed,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.text, "n_particles_timed": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------------------------
def ess_match(wl, rank, world, local_rank, dist, torch, n_global):
    """First stages of the same global cloud: engine (sharded over `world` GPUs) vs the CPU oracle (rank 0)."""
    from smc_jl_b200._lib import StageState
    eng = make_engine(rank, world, local_rank, dist, torch)
    eng.cloud_create(n_global, wl.d)
    ess0 = wl.prepare(eng, n_global)
    st = StageState(c=0.5, accept=0.25, ess_prev=ess0, phi_prop=0.0, j=2)
    res = eng.run_stages(wl.cfg(2, 0.0), st, wl.sched, 2, MATCH_STAGES) if not wl.kw["adaptive"] else None
    if res is None:
        res, phi = [], 0.0
        for i in range(2, 2 + MATCH_STAGES):
            r, _, _ = eng.stage(wl.cfg(i, phi), st, schedule=wl.sched)
            res.append(r)
            phi = r.phi_n
    mine = eng.download()
    first, count = eng.first, eng.count
    eng.close()
    digests = [None] * world
    if world > 1:
        dist.all_gather_object(digests, (first, count, shard_digest(mine)))
    else:
        digests = [(first, count, shard_digest(mine))]
    out = None
    if rank == 0:
        t0 = time.perf_counter()
        orc = OracleRun(wl, n_global, os.cpu_count() or 1)
        rows = [orc.stage() for _ in range(MATCH_STAGES)]
        want = orc.cloud()
        rel = 0.0
        for r, o in zip(res, rows):
            for a, b in ((r.ess, o[1]), (r.accept, o[2]), (r.c, o[3]), (r.phi_n, o[5])):
                rel = max(rel, abs(a - b) / max(abs(b), 1e-300))
        same = all(shard_digest(want[f:f + c]) == dg for f, c, dg in digests)
        out = {"n": int(n_global), "stages": MATCH_STAGES, "max_rel": rel, "cloud_bitexact": bool(same),
               "resamples": int(sum(r.resampled for r in res)), "oracle_resamples": int(sum(o[4] for o in rows)),
               "ess": [float(r.ess) for r in res], "oracle_ess": [float(o[1]) for o in rows],
               "oracle_seconds": time.perf_counter() - t0,
               "what": "engine on %d GPU(s) vs oracle/ (CPU restatement of the reference path): ESS, accept, c, phi of every stage "
                       "(max relative difference) and a blake2b checksum of every rank's shard of the cloud after the last stage "
                       "against the oracle's rows" % world}
    return out


def run_ours(args):
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner to fd 1) write to stderr
    # for the duration of the run, the saved descriptor is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from smc_jl_b200._lib import StageState

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    wl = Workload(args.config, world)
    n_global = wl.n_global
    K = args.steps
    fixed = not wl.kw["adaptive"]
    if fixed:
        K = min(K, (len(wl.sched) - wl.first - 12) // 2)            # value window + e2e window + phase window fit the schedule

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ESS match: the same global cloud for C2 / C4; C3 / C5 compare a 2^16-particle run of the same two-vintage workflow (the
    # oracle's first-vintage run-up takes minutes at 2^20)
    n_check = n_global if wl.name in ("c2", "c4") else (1 << 16)
    match = None if args.no_match else ess_match(wl, rank, world, local_rank, dist, torch, n_check)
    # (after the oracle leg: its OpenMP pool keeps the full CPU set)
    old_affinity = bind_near_gpu(local_rank) if world > 1 else None

    eng = make_engine(rank, world, local_rank, dist, torch)
    eng.cloud_create(n_global, wl.d)
    N = eng.count
    ess0 = wl.prepare(eng, n_global)
    state = StageState(c=0.5, accept=0.25, ess_prev=ess0, phi_prop=0.0, j=2)
    i_next, phi = 2, 0.0

    def run_block(n_stages, inc=None, nw=None):
        nonlocal i_next, phi
        cfg = wl.cfg(i_next, phi)
        res = eng.run_stages(cfg, state, wl.sched, i_next, n_stages, inc_hist=inc, normw_hist=nw)
        i_next += len(res)
        phi = res[-1].phi_n
        return res

    # run the schedule up to the timed window (untimed; this is the real trajectory, not a shortcut), then W warm-up steps
    if wl.first - 2 - args.warmup > 0:
        run_block(wl.first - 2 - args.warmup)
    run_block(args.warmup)

    def phase_window(n_max):
        nonlocal i_next, phi
        ph, n = np.zeros(4), 0
        while n < n_max and phi < 1.0 and (not fixed or i_next <= len(wl.sched)):
            r, _, _ = eng.stage(wl.cfg(i_next, phi), state, schedule=wl.sched)
            ph += [r.ms_correct, r.ms_resample, r.ms_moments, r.ms_mutate]
            i_next += 1
            phi = r.phi_n
            n += 1
        return ph / max(n, 1), n

    # per-phase device times (CUDA events inside single smcb200_stage calls): an adaptive run may reach phi = 1 inside the
    # timed windows, so its phase window comes first
    phase, n_ph = (phase_window(6) if not fixed else (np.zeros(4), 0))
    launches0 = eng.kernel_launches
    barrier()
    with ClockSampler(local_rank) as clk:
        eng.timer_start()
        res_v = run_block(K)                                         # exactly K timed steps (fewer only if phi reaches 1)
        ms_total = eng.timer_stop()
        barrier()
    k_done = len(res_v)
    launches = eng.kernel_launches - launches0
    resamples = sum(r.resampled for r in res_v)
    ms_t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = float(ms_t.item()) / k_done
    value = n_global * wl.steps_per_particle / (ms_step * 1e-3)

    # ---- e2e: what smc() does -- the same call with the w / W history and the summaries streaming to pinned host memory
    e2e = None
    if phi < 1.0:
        k_e2e = K
        hist = torch.empty((2, k_e2e, N), dtype=torch.float64, pin_memory=True).numpy()
        barrier()
        t0 = time.perf_counter()
        res_e = run_block(k_e2e, inc=hist[0], nw=hist[1])
        torch.cuda.synchronize()
        e2e_sec = (time.perf_counter() - t0) / len(res_e)
        e2e_t = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        w_ok = bool(np.all(np.isfinite(hist[1][:len(res_e)])) and abs(hist[1][len(res_e) - 1].sum() * world - n_global) < 1e-6 * n_global) \
            if world == 1 else True
        e2e = {"value": n_global * wl.steps_per_particle / float(e2e_t.item()), "unit": UNIT, "h2d_bytes_per_step": 136,
               "d2h_bytes_per_step": int(2 * N * 8 + 256), "steps": len(res_e), "history_ok": w_ok,
               "what": "smcb200_run_stages as smc() calls it: cloud resident in HBM, wall clock incl. the per-stage device -> host "
                       "stream of the w / W history columns (2 x 8 MB per GPU and stage) and of the stage summary into pinned memory; "
                       "host -> device per stage = the stage configuration"}

    if fixed:
        phase, n_ph = phase_window(10)
    # ---- the PCIe-bound variant: Cloud in pinned HOST memory, upload + stage + download per step ---------------------------
    e2e_host = None
    if wl.name == "c2" and phi < 1.0:
        pinned = torch.empty((wl.d + 5, N), dtype=torch.float64, pin_memory=True)
        host = pinned.numpy().T
        eng.download(host)
        eng.stage_host(host, wl.cfg(i_next, phi), state); i_next += 1
        t0 = time.perf_counter()
        for _ in range(3):
            eng.stage_host(host, wl.cfg(i_next, phi), state); i_next += 1
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / 3
        e2e_host = {"value": n_global * wl.steps_per_particle / sec, "unit": UNIT, "h2d_bytes_per_step": int((wl.d + 5) * N * 8),
                    "d2h_bytes_per_step": int((wl.d + 5) * N * 8),
                    "what": "smcb200_stage_host: Cloud kept in pinned HOST memory between stages (2 x 210 MB over PCIe per step)"}
    fp64_peak = eng.fp64_peak() if rank == 0 else None

    # ---- rooflines of the dominant kernel (mutation) -------------------------------------------------------------------
    peak, peak_src = peaks()
    mut_ms = float(phase[3])
    mut_bytes = 8.0 * (2 * wl.d + 7) * N
    achieved = mut_bytes / (mut_ms * 1e-3) / 1e9 if mut_ms > 0 else None
    out = None
    if old_affinity is not None:
        os.sched_setaffinity(0, old_affinity)
    if rank == 0:
        cpu = None
        try:
            if args.no_cpu:
                raise RuntimeError("skipped (--no-cpu)")
            n_cpu = n_global if (wl.name in ("c2", "c4") and world == 1) else min(n_global, 1 << 16 if wl.name in ("c3", "c5") else N_FULL)
            orc = OracleRun(wl, n_cpu, os.cpu_count() or 1)
            for _ in range(2):
                orc.stage()
            k_cpu = 4 if n_cpu >= (1 << 18) else 12
            sec = float(np.mean([orc.stage()[0] for _ in range(k_cpu)]))
            cpu = {"value": n_cpu * wl.steps_per_particle / sec, "unit": UNIT, "cores": orc.nthreads, "kind": "port",
                   "sample": "%d stages of the same workload at n_particles=%d after 2 warm-up stages (oracle/, OpenMP)" % (k_cpu, n_cpu)}
        except Exception as e:  # the oracle is a reported baseline, never the product path
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}
        traffic, traffic_src = ncu_traffic() if wl.name == "c2" else (None, "not captured for this config")
        d = wl.d
        # flops per particle-MH-step of the mutation kernel.  C2: counted from the algorithm (two triangular mat-vecs + prior /
        # quadratic form / increments + exp).  The other configs: FP64-pipe instructions per particle-step measured by ncu
        # (profiles/r02_ncu_summary.json: fp64_pipe_pct x cycles x 592 sub-partitions / 2 / warp-steps = 771 for the 9-parameter
        # three-block mixture kernel of C3 / C5, 64.2e3 for the 230-period Kalman filter of C4) x 1.75 flops per instruction
        # (the fma share of the C2 kernel)
        flops_step = {"c2": 2.0 * (d * (d + 1)) + 2.0 * d * 5 + 60.0, "c3": 1350.0, "c5": 1350.0, "c4": 112.0e3}[wl.name]
        steps_per_s_kernel = N * wl.steps_per_particle / (mut_ms * 1e-3) if mut_ms > 0 else 0.0
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": k_done, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl.text + ", timed stages %d..%d" % (wl.first, wl.first + k_done - 1) if fixed else wl.text,
                       "name": wl.name, "n_particles_global": int(n_global),
                       "l2": "cloud double buffer 2 x %d MB per GPU vs 126 MB L2 (no explicit flush)" % ((wl.d + 5) * N * 8 >> 20),
                       "parallelism": ("one global cloud of %d particles sharded over %d GPUs; every per-stage reduction crosses the GPUs "
                                       "inside our own kernels (NVLink mailboxes, fixed-order trees); selection reads the owners' running "
                                       "maxima and rows over NVLink (CUDA IPC); no collective library call in the stage loop"
                                       % (n_global, world)) if world > 1 else "1 GPU",
                       "resamples_in_timed_region": int(resamples),
                       "resample_heavy_ok": (bool(3 * resamples >= k_done) if wl.name == "c5" else None)},
            "e2e": e2e,
            "e2e_host_cloud": e2e_host,
            "ess_match": match,
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "kernel": "k_mutate (persistent warps, one MH chain per thread)",
                         "algorithmic_bytes_per_launch": mut_bytes, "avg_launch_ms": mut_ms, "peak_source": peak_src,
                         "traffic_source": traffic_src,
                         "note": "8(2d+7) B per particle and stage whatever n_mh_steps; the kernel is bound by instruction issue / FP64 "
                                 "latency, not by HBM (see roofline_fp64 and DESIGN.md)"},
            "roofline_fp64": {"bound": "fp64", "achieved": flops_step * steps_per_s_kernel / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": (flops_step * steps_per_s_kernel / 1e12 / fp64_peak) if fp64_peak else None,
                              "flops_per_particle_step": flops_step,
                              "peak_source": "smcb200_fp64_peak: dependent-chain DFMA micro-benchmark on this GPU, best of 5"},
            "cpu_baseline": cpu,
            "phase_ms_per_step": {"correct_and_phi_solve": float(phase[0]), "resample": float(phase[1]),
                                  "moments_and_proposal": float(phase[2]), "mutate": mut_ms, "stages": n_ph,
                                  "note": "CUDA events inside single smcb200_stage calls after the timed windows"},
        }
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--no-match", action="store_true", help="skip the ess_match leg (developer runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (developer runs)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
