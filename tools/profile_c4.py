#!/usr/bin/env python
"""Config C4 (BASELINE.json configs[3]) on one GPU: An-Schorfheide DSGE with the device Kalman-filter
log-likelihood, n_particles = 2^18, n_mh_steps = 5, 13 free parameters, adaptive phi.  Prints one JSON line
with the per-stage throughput (CUDA-event device time from the library's stopwatch) and, optionally,
times the CPU oracle on a bounded sample.  Also used as the ncu driver for the C4 mutation kernel.

    python tools/profile_c4.py [--stages 12] [--n 262144] [--cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smc_jl_b200 import model as M  # noqa: E402
from smc_jl_b200 import workloads as W  # noqa: E402
from smc_jl_b200._lib import StageConfig, StageState  # noqa: E402
from smc_jl_b200.engine import Engine  # noqa: E402

# one process per GPU under torchrun: the SAME global cloud is sharded over the ranks (strong scaling)
RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
dist = None
if WORLD > 1:
    import torch
    import torch.distributed as dist
    _fd1 = os.dup(1); os.dup2(2, 1)                      # NCCL prints its banner to stdout
    torch.cuda.set_device(LOCAL)
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))


def make_engine():
    eng = Engine(LOCAL)
    if WORLD > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if RANK == 0:
            idt.copy_(torch.frombuffer(bytearray(Engine.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init(RANK, WORLD, bytes(idt.cpu().numpy().tobytes()))
    return eng


def max_over_ranks(x):
    if WORLD == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def emit(obj):
    if WORLD > 1:
        sys.stdout.flush(); os.dup2(_fd1, 1)
        dist.barrier(); dist.destroy_process_group()
    if RANK == 0:
        print(json.dumps(obj), flush=True)

ap = argparse.ArgumentParser()
ap.add_argument("--stages", type=int, default=12)
ap.add_argument("--n", type=int, default=1 << 18)
ap.add_argument("--n-mh", type=int, default=5)
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()

g = np.load(os.path.join(ROOT, "tests", "golden", "as_clouds.npz"))
params = W.an_schorfheide_parameters()
spec = M.make_spec(params, M.AnSchorfheideLogLik(g["data"]))
N = args.n
eng = make_engine()
eng.cloud_create(N, 16)
eng.set_model(spec)
eng.timer_start()
eng.initial_draw(spec.values, 1793, 1000)
ms_init = eng.timer_stop()
sched = (np.arange(300) / 299.0) ** 2.1
state = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2)
phi_prev, rows = 0.0, []
for s in range(args.stages):
    cfg = StageConfig(phi_n1=phi_prev, phi_n=0.0, threshold_ratio=0.5, target=0.25, alpha=0.9, tempering_target=0.97,
                      n_mh_steps=args.n_mh, n_blocks=1, resample_method=0, adaptive=1, seed=1793, stage=s + 2)
    eng.timer_start()
    res, _, _ = eng.stage(cfg, state, schedule=sched)
    ms = max_over_ranks(eng.timer_stop())
    phi_prev = res.phi_n
    rows.append((ms, res.ms_correct, res.ms_resample, res.ms_moments, res.ms_mutate, res.ess, res.accept, res.phi_n))
    if phi_prev >= 1.0:
        break
rows = np.array(rows)
timed = rows[2:] if len(rows) > 4 else rows
ms_stage, ms_mut = timed[:, 0].mean(), timed[:, 4].mean()
out = {
    "workload": "C4 An-Schorfheide DSGE, device Kalman loglik (T=230), n_particles=%d, n_mh_steps=%d, 13 free parameters, "
                "alpha=0.9, adaptive phi (tempering_target 0.97), %d GPU(s), one global cloud sharded" % (N, args.n_mh, WORLD),
    "n_gpus": WORLD,
    "stages_timed": int(len(timed)), "ms_per_stage": float(ms_stage), "ms_mutate": float(ms_mut),
    "ms_initial_draw": float(ms_init),
    "particle_mh_steps_per_sec_per_stage": float(N * args.n_mh / (ms_stage * 1e-3)),
    "particle_mh_steps_per_sec_mutation_kernel": float(N * args.n_mh / (ms_mut * 1e-3)),
    "loglik_evaluations_per_sec": float(N * args.n_mh / (ms_mut * 1e-3)),
    "phi": [float(v) for v in rows[:, 7]], "ess": [float(v) for v in rows[:, 5]], "accept": [float(v) for v in rows[:, 6]],
    "phase_ms": {"correct_and_phi_solve": float(timed[:, 1].mean()), "resample": float(timed[:, 2].mean()),
                 "moments": float(timed[:, 3].mean()), "mutate": float(ms_mut)},
}
eng.close()
if args.cpu and RANK == 0:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C

    import oracle_lib as O
    n = 1 << 12
    mod = O.Model(spec)
    buf = np.zeros(n * 21)
    O.lib().orc_initial_draw(mod.h, buf, n, 0, np.ascontiguousarray(spec.values), 1793, 1000)
    scratch = np.zeros_like(buf)
    io = O.StageIO(threshold_ratio=0.5, target=0.25, alpha=0.9, tempering_target=0.97, n_mh_steps=args.n_mh, n_blocks=1,
                   resample_method=0, adaptive=1, nthreads=os.cpu_count() or 1, seed=1793, c=0.5, accept=0.25, ess_prev=float(n), j=2)
    t = []
    for s in range(3):
        io.phi_n1, io.stage = (0.0 if s == 0 else io.phi_out), s + 2
        t0 = time.perf_counter()
        assert O.lib().orc_stage(mod.h, buf, scratch, n, sched, 300, C.byref(io), None, None, None, None) == 0
        t.append(time.perf_counter() - t0)
    out["cpu_oracle"] = {"particle_mh_steps_per_sec_per_stage": float(n * args.n_mh / np.mean(t[1:])), "cores": os.cpu_count(),
                         "sample": "N=2^12, 2 timed stages (oracle/, OpenMP)"}
emit(out)
