#!/usr/bin/env python
"""Short driver for ncu: a few fused stages of config C2 (d=20, N=2^20, n_mh_steps=3) from the prior cloud.
Kernel durations do not depend on the tempering stage, so this is what bench.py times, minus the run-up."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from smc_jl_b200 import workloads as W  # noqa: E402
from smc_jl_b200._lib import StageState  # noqa: E402
from smc_jl_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--stages", type=int, default=6)
ap.add_argument("--n", type=int, default=bench.N_FULL)
args = ap.parse_args()
params, spec = bench.make_model()
sched = bench.schedule()
eng = Engine(0)
eng.cloud_create(args.n, bench.D)
eng.set_model(spec)
eng.upload(W.initial_cloud(params, args.n, np.random.default_rng(0)))
eng.evaluate(0)
state = StageState(c=0.5, accept=0.25, ess_prev=float(args.n), phi_prop=0.0, j=2)
for s in range(args.stages):
    res, _, _ = eng.stage(bench.stage_cfg(sched, s), state)
    print("stage %d ess %.1f resampled %d accept %.3f ms: correct %.3f resample %.3f moments %.3f mutate %.3f" % (
        s + 2, res.ess, res.resampled, res.accept, res.ms_correct, res.ms_resample, res.ms_moments, res.ms_mutate))
eng.close()
