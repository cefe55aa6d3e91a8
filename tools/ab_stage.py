#!/usr/bin/env python
"""Developer A/B harness: per-phase device times of the C2 stage for the library named by SMCB200_LIB."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from smc_jl_b200 import workloads as W  # noqa: E402
from smc_jl_b200._lib import StageState  # noqa: E402
from smc_jl_b200.engine import Engine  # noqa: E402


def main():
    params, spec = B.make_model()
    sched = B.schedule()
    N = B.N_FULL
    eng = Engine(0)
    eng.cloud_create(N, B.D)
    eng.set_model(spec)
    eng.upload(W.initial_cloud(params, N, np.random.default_rng(0)))
    eng.evaluate(0)
    state = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2)
    for s in range(30):
        eng.stage(B.stage_cfg(sched, s), state)
    ph = np.zeros(4)
    K = 20
    eng.timer_start()
    for s in range(30, 30 + K):
        res, _, _ = eng.stage(B.stage_cfg(sched, s), state)
        ph += [res.ms_correct, res.ms_resample, res.ms_moments, res.ms_mutate]
    tot = eng.timer_stop() / K
    print(json.dumps({"lib": os.path.basename(os.environ.get("SMCB200_LIB", "default")), "ms_stage": tot, "correct": ph[0] / K,
                      "resample": ph[1] / K, "moments": ph[2] / K, "mutate": ph[3] / K, "accept": res.accept, "ess": res.ess}))
    eng.close()


if __name__ == "__main__":
    main()
