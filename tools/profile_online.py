#!/usr/bin/env python
"""Configs C3 and C5 of BASELINE.json on one GPU (or sharded under torchrun): an online update with generalised
tempering -- estimate on the first data vintage, then `tempered_update` on the full sample starting from that cloud
(SURVEY 3.5 branch (a): same n_parts, prior weight 0; src/smc_main.jl:244-258, src/initialization.jl:153-186).

  --config c3 : examples/capm_model (9 parameters, lik_data 3 x 36, old data = first 18 periods), adaptive phi
                (tempering_target 0.97), n_particles = 2^20
  --config c5 : test/modelsetup.jl 3-equation model on the reference's test_data.h5 (3 x 100, old data = first 50),
                fixed schedule, threshold_ratio 0.9 (resample-heavy), n_particles = 2^20

Prints one JSON line: per-stage device time (library stopwatch, CUDA events) of the SECOND-vintage stages, whose
correction and Metropolis ratio carry the old-data likelihood (two likelihood evaluations per MH step).

    python tools/profile_online.py --config c5 [--n 1048576] [--stages 20]
"""
import argparse
import json
import os
import sys

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
ap.add_argument("--config", default="c5", choices=["c3", "c5"])
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--stages", type=int, default=20)
ap.add_argument("--n-mh", type=int, default=1)
args = ap.parse_args()

G = os.path.join(ROOT, "tests", "golden")
params = W.three_equation_parameters()
if args.config == "c3":
    g = np.load(os.path.join(G, "capm_data.npz"))
    full = M.CAPMLogLik(g["lik_data"], g["market_data"])
    old = M.CAPMLogLik(g["lik_data"][:, :18], g["market_data"])
    adaptive, thr, what = 1, 0.5, "C3 examples/capm_model, adaptive phi (target 0.97) + generalised tempering (old data = 18 of 36 periods)"
else:
    g = np.load(os.path.join(G, "linear_model_rows.npz"))
    full = M.LinearEquationsLogLik(g["data"], g["X"])
    old = M.LinearEquationsLogLik(g["data"][:, :50], g["X"])
    adaptive, thr, what = 0, 0.9, "C5 online update, 3-equation model (old data = 50 of 100 periods), threshold_ratio 0.9 (resample-heavy)"
N = args.n
sched = (np.arange(300) / 299.0) ** 2.1
eng = make_engine()
eng.cloud_create(N, 9)


def run(spec, has_old, n_stage, ess0, timed):
    state = StageState(c=0.5, accept=0.25, ess_prev=ess0, phi_prop=0.0, j=2)
    phi_prev, rows = 0.0, []
    for s in range(n_stage):
        cfg = StageConfig(phi_n1=phi_prev, phi_n=float(sched[min(s + 1, 299)]), threshold_ratio=thr, target=0.25, alpha=0.9,
                          tempering_target=0.97, n_mh_steps=args.n_mh, n_blocks=3, resample_method=0, adaptive=adaptive,
                          has_old_data=has_old, seed=1793, stage=s + 2)
        if timed:
            eng.timer_start()
        res, _, _ = eng.stage(cfg, state, schedule=sched)
        ms = max_over_ranks(eng.timer_stop()) if timed else 0.0
        rows.append((ms, res.ms_correct, res.ms_resample, res.ms_moments, res.ms_mutate, res.ess, res.resampled, res.phi_n))
        phi_prev = res.phi_n
        if phi_prev >= 1.0:
            break
    return np.array(rows), state


# first vintage: from the prior to phi = 1 on the old data (fixed 300-point schedule: the run-up is not what is timed)
spec_old = M.make_spec(params, old)
eng.set_model(spec_old)
eng.initial_draw(spec_old.values, 1793, 1000)
adaptive_saved, adaptive = adaptive, 0
rows_a, st_a = run(spec_old, 0, 299, float(N), False)
adaptive = adaptive_saved
# second vintage: initialize_likelihoods! (old_loglh <- loglh, re-evaluate on the full sample), then the tempered update
spec_new = M.make_spec(params, full, old)
eng.set_model(spec_new)
eng.evaluate(1)
rows_b, _ = run(spec_new, 1, args.stages, float(rows_a[-1, 5]), True)
timed = rows_b[2:] if len(rows_b) > 4 else rows_b
ms = float(timed[:, 0].mean())
out = {
    "workload": what + ", n_particles=%d, n_mh_steps=%d, 3 blocks, alpha=0.9, %d GPU(s), one global cloud sharded" % (N, args.n_mh, WORLD),
    "n_gpus": WORLD,
    "stages_timed": int(len(timed)), "ms_per_stage": ms,
    "particle_mh_steps_per_sec_per_stage": float(N * args.n_mh * 3 / (ms * 1e-3)),
    "resamples_in_timed_stages": int(timed[:, 6].sum()),
    "phase_ms": {"correct_and_phi_solve": float(timed[:, 1].mean()), "resample": float(timed[:, 2].mean()),
                 "moments": float(timed[:, 3].mean()), "mutate": float(timed[:, 4].mean())},
    "phi": [float(v) for v in rows_b[:, 7]], "ess": [float(v) for v in rows_b[:, 5]],
    "first_vintage": {"stages": int(len(rows_a)), "final_ess": float(rows_a[-1, 5])},
}
eng.close()
emit(out)
