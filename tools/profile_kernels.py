#!/usr/bin/env python
"""Driver for `ncu --set full --profile-from-start off`: brings a BASELINE config to its timed window, then brackets a
few stages with cudaProfilerStart / cudaProfilerStop so that every kernel of the stage is captured exactly once per
stage (correction / adaptive solve, the selection chain with a stage that really resamples, moments, mutation).

usage: ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_c2 \
           python tools/profile_kernels.py --config c2 [--stages 2]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from smc_jl_b200._lib import StageState  # noqa: E402
from smc_jl_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--stages", type=int, default=2)
    ap.add_argument("--runup", type=int, default=-1, help="untimed stages before the capture (default: bench.py's)")
    args = ap.parse_args()
    try:
        rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
    except OSError:
        import torch
        rt = torch.cuda.cudart()
    start, stop = rt.cudaProfilerStart, rt.cudaProfilerStop
    wl = B.Workload(args.config, 1)
    eng = Engine(0)
    eng.cloud_create(wl.n_global, wl.d)
    ess_prev = wl.prepare(eng, wl.n_global)
    st = StageState(c=0.5, accept=0.25, ess_prev=ess_prev, phi_prop=0.0, j=2)
    runup = args.runup if args.runup >= 0 else (wl.first - 2 if wl.name == "c2" else 4)
    i = 2
    phi = 0.0
    if runup > 0:
        res = eng.run_stages(wl.cfg(i, phi_n1=phi), st, wl.sched, i, runup)
        i += len(res)
        phi = res[-1].phi_n
    start()
    for k in range(args.stages):
        # the second captured stage is forced to resample so that the selection chain shows its real cost
        over = dict(threshold_ratio=2.0) if k == 1 else {}
        cfg = wl.cfg(i, phi_n1=phi, **over)
        res, _, _ = eng.stage(cfg, st, schedule=wl.sched if cfg.adaptive else None)
        print("captured stage %d: phi %.6g ess %.1f resampled %d accept %.3f | ms correct %.3f resample %.3f moments %.3f mutate %.3f"
              % (i, res.phi_n, res.ess, res.resampled, res.accept, res.ms_correct, res.ms_resample, res.ms_moments, res.ms_mutate))
        phi = res.phi_n
        i += 1
    stop()
    eng.close()


if __name__ == "__main__":
    main()
