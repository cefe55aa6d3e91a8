#!/usr/bin/env python
"""bench.py -- particle-MH-steps/sec per SMC stage (BASELINE.json metric) on config C2:
linear-Gaussian log-likelihood, 20 parameters, n_particles = 2^20, n_mh_steps = 3, fixed
tempering schedule (n_Phi = 300, lambda = 2.1), systematic resampling, one B200.

A "step" is one full SMC stage (correction -> selection -> moments -> mutation) over the whole
cloud.  `value` = N * n_mh_steps * n_blocks / (device time per stage) with the cloud resident in
HBM; `e2e` = the same through smcb200_stage_host with the Cloud in pinned HOST memory (upload +
stage + download inside the timed region).  `--impl reference` times the CPU restatement of the
reference path (oracle/, OpenMP on all host cores) on a bounded sample of the same workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import ctypes as C
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
MUTATE_DRAM_BYTES_NCU = 335.8e6   # measured once per kernel change with ncu (profiles/r01_final_summary.md)
FIRST_STAGE = 30   # timed stages start here in the 300-point schedule (past the burn-in of the prior cloud)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def schedule():
    return ((np.arange(N_PHI)) / (N_PHI - 1.0)) ** LAM


def make_model():
    from smc_jl_b200 import model as M
    from smc_jl_b200 import workloads as W
    params, lk, _ = W.linear_gaussian(d=D, T=T)
    return params, M.make_spec(params, lk)


def stage_cfg(sched, s):
    from smc_jl_b200._lib import StageConfig
    return StageConfig(phi_n1=float(sched[s]), phi_n=float(sched[s + 1]), threshold_ratio=0.5, target=0.25, alpha=1.0,
                       tempering_target=0.95, n_mh_steps=N_MH, n_blocks=N_BLOCKS, resample_method=0, seed=SEED, stage=s + 2)


def oracle_run(spec, params, n, first_stage, n_stages, nthreads):
    """CPU restatement of the reference stage loop on `n` particles: runs the real trajectory from the prior
    cloud (stages before `first_stage` untimed); returns (seconds per timed stage, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from smc_jl_b200 import workloads as W
    L = O.lib()
    mod = O.Model(spec)
    P = W.initial_cloud(params, n, np.random.default_rng(0))
    buf = O.cloud_f(P)
    L.orc_evaluate(mod.h, buf, n)
    scratch = np.zeros_like(buf)
    sched = schedule()
    io = O.StageIO(threshold_ratio=0.5, target=0.25, alpha=1.0, tempering_target=0.95, n_mh_steps=N_MH, n_blocks=N_BLOCKS,
                   resample_method=0, nthreads=nthreads, seed=SEED, c=0.5, accept=0.25, ess_prev=float(n), j=2)
    times = []
    for s in range(0, first_stage + n_stages):
        io.phi_n1, io.phi_n, io.stage = float(sched[s]), float(sched[s + 1]), s + 2
        t0 = time.perf_counter()
        st = L.orc_stage(mod.h, buf, scratch, n, sched, N_PHI, C.byref(io), None, None, None, None)
        dt = time.perf_counter() - t0
        assert st == 0, st
        if s >= first_stage:
            times.append(dt)
    return float(np.mean(times)), (nthreads if nthreads > 0 else L.orc_max_threads())


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (restated in oracle/, see DESIGN.md: Julia and the
    reference's dependencies are not available in this image) on all host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params, spec = make_model()
    n = 1 << 16
    warm = max(args.warmup, 1)
    sec, cores = oracle_run(spec, params, n, FIRST_STAGE, args.steps, os.cpu_count() or 1)   # explicit: torchrun exports OMP_NUM_THREADS=1
    value = n * N_MH * N_BLOCKS / sec
    sample = "N=2^16 particles of the same d=20/T=256 model, %d stages from schedule index %d, n_mh_steps=3" % (args.steps, FIRST_STAGE)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 linear-Gaussian loglik, 20 params, n_mh_steps=3, fixed phi schedule (n_phi=300, lambda=2.1); "
                               "CPU sample N=2^16 (throughput per particle-step is size-independent)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner to fd 1) write to stderr
    # for the duration of the run, the saved descriptor is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from smc_jl_b200 import workloads as W
    from smc_jl_b200._lib import StageState
    from smc_jl_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    params, spec = make_model()
    sched = schedule()
    # weak scaling: ONE global cloud of world x 2^20 particles, sharded over the ranks (2^20 per GPU); the
    # weight normaliser / ESS / moments / accept reductions and the post-resample row exchange cross GPUs
    N = N_FULL                     # particles per GPU
    N_global = N * world
    eng = Engine(local_rank)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Engine.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
    eng.cloud_create(N_global, D)
    assert eng.count == N
    eng.set_model(spec)
    P0 = W.initial_cloud(params, N, np.random.default_rng(rank))     # this rank's shard of the prior cloud
    eng.upload(P0)
    eng.evaluate(0)
    state = StageState(c=0.5, accept=0.25, ess_prev=float(N_global), phi_prop=0.0, j=2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # run the schedule up to the timed window (untimed; this is the real trajectory, not a shortcut)
    prep = FIRST_STAGE - args.warmup
    for s in range(prep):
        eng.stage(stage_cfg(sched, s), state)
    for s in range(prep, FIRST_STAGE):
        eng.stage(stage_cfg(sched, s), state)                       # W warm-up steps
    launches0 = eng.kernel_launches
    phase = np.zeros(4)
    resamples = 0
    barrier()
    with ClockSampler(local_rank) as clk:
        eng.timer_start()
        for s in range(FIRST_STAGE, FIRST_STAGE + args.steps):      # exactly K timed steps
            res, _, _ = eng.stage(stage_cfg(sched, s), state)
            phase += [res.ms_correct, res.ms_resample, res.ms_moments, res.ms_mutate]
            resamples += res.resampled
        ms_total = eng.timer_stop()
        barrier()
    launches = eng.kernel_launches - launches0
    ms_t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = float(ms_t.item()) / args.steps
    value = world * N * N_MH * N_BLOCKS / (ms_step * 1e-3)

    # ---- e2e: Cloud in pinned host memory, upload + stage + download per step ---------------------
    cols = D + 5
    pinned = torch.empty((cols, N), dtype=torch.float64, pin_memory=True)
    host = pinned.numpy().T                                          # N x cols, Fortran-ordered view
    eng.download(host)
    st2 = StageState(c=state.c, accept=state.accept, ess_prev=state.ess_prev, phi_prop=0.0, j=2)
    s0 = FIRST_STAGE + args.steps
    eng.stage_host(host, stage_cfg(sched, s0), st2)                  # warm-up
    barrier()
    t0 = time.perf_counter()
    k_e2e = max(2, min(args.steps, 5))
    for s in range(s0 + 1, s0 + 1 + k_e2e):
        eng.stage_host(host, stage_cfg(sched, s), st2)
    torch.cuda.synchronize()
    e2e_sec = (time.perf_counter() - t0) / k_e2e
    e2e_t = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * N * N_MH * N_BLOCKS / float(e2e_t.item())

    # ---- the per-stage call smc() itself makes: cloud resident, stage summary + the two weight-history columns
    # (w_matrix / W_matrix, src/smc_main.jl:419-420) read back into pinned host memory every stage ------------
    hist = torch.empty((2, N), dtype=torch.float64, pin_memory=True).numpy()
    eng.upload(host)
    st3 = StageState(c=st2.c, accept=st2.accept, ess_prev=st2.ess_prev, phi_prop=0.0, j=2)
    s1 = s0 + 1 + k_e2e
    eng.stage(stage_cfg(sched, s1), st3, inc_out=hist[0], normw_out=hist[1])
    barrier()
    t0 = time.perf_counter()
    for s in range(s1 + 1, s1 + 1 + 10):
        eng.stage(stage_cfg(sched, s), st3, inc_out=hist[0], normw_out=hist[1])
    torch.cuda.synchronize()
    res_sec = (time.perf_counter() - t0) / 10
    res_t = torch.tensor([res_sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(res_t, op=dist.ReduceOp.MAX)
    resident_value = world * N * N_MH * N_BLOCKS / float(res_t.item())

    # ---- roofline of the dominant kernel (mutation): algorithmic bytes / CUDA-event duration ---------
    peak, peak_src = peaks()
    mut_ms = phase[3] / args.steps
    mut_bytes = 8.0 * (2 * D + 7) * N
    achieved = mut_bytes / (mut_ms * 1e-3) / 1e9

    out = None
    if rank == 0:
        cpu = None
        try:
            sec, cores = oracle_run(spec, params, 1 << 15, FIRST_STAGE, 3, os.cpu_count() or 1)
            cpu = {"value": (1 << 15) * N_MH * N_BLOCKS / sec, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "N=2^15 particles of the same model, 3 stages from schedule index %d (oracle/, OpenMP)" % FIRST_STAGE}
        except Exception as e:  # the oracle is a reported baseline, never the product path
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C2 linear-Gaussian loglik, 20 params, n_particles=2^20 per GPU, n_mh_steps=3, fixed phi "
                                   "schedule (n_phi=300, lambda=2.1), systematic resampling, timed stages %d..%d"
                                   % (FIRST_STAGE + 2, FIRST_STAGE + 1 + args.steps),
                       "l2": "cloud double buffer 2 x 210 MB > 126 MB L2 (inputs larger than L2, no explicit flush)",
                       "parallelism": ("one global cloud of %d x 2^20 particles sharded over %d GPUs: NCCL all-gather + fixed-order "
                                       "cross-rank trees for the reductions, NVLink peer reads (CUDA IPC) for the post-resample "
                                       "row exchange" % (world, world)) if world > 1 else "1 GPU",
                       "resamples_in_timed_region": int(resamples)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(cols * N * 8), "d2h_bytes_per_step": int(cols * N * 8),
                    "what": "smcb200_stage_host: Cloud in pinned host memory, upload + stage + download each step "
                            "(PCIe-bound: 2 x 210 MB per step)"},
            "e2e_resident": {"value": resident_value, "unit": UNIT, "h2d_bytes_per_step": 128, "d2h_bytes_per_step": int(2 * N * 8 + 72),
                             "what": "the per-stage call smc() makes: cloud resident in HBM, wall clock around smcb200_stage incl. the "
                                     "D2H of the stage summary and of the w/W history columns into pinned host memory"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": MUTATE_DRAM_BYTES_NCU, "kernel": "k_mutate<GaussReg<1,20,20,0,-1>, HAS_OLD=false, BLK=2, MIX=false>",
                         "algorithmic_bytes_per_launch": mut_bytes, "avg_launch_ms": mut_ms, "peak_source": peak_src,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                           "(profiles/r01_final_summary.md); per launch at N = 2^20",
                         "note": "instruction-issue bound (smsp__issue_active 73 %, FP64 pipe 37 %), see DESIGN.md"},
            "cpu_baseline": cpu,
            "phase_ms_per_step": {"correct": phase[0] / args.steps, "resample": phase[1] / args.steps,
                                  "moments_and_proposal": phase[2] / args.steps, "mutate": mut_ms},
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
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
