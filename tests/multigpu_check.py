#!/usr/bin/env python
"""Multi-GPU invariance check (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank runs the SAME global cloud (a) sharded over the job's GPUs through the communicator and (b) alone on
its own GPU, stage by stage, and asserts that its shard of (a) equals the corresponding rows of (b) BIT-FOR-BIT,
together with phi / ESS / c / accept and the resample decisions (results must not depend on the GPU count)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smc_jl_b200 import model as M  # noqa: E402
from smc_jl_b200 import workloads as W  # noqa: E402
from smc_jl_b200._lib import StageConfig, StageState  # noqa: E402
from smc_jl_b200.engine import Engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_comm_id():                      # a communicator id is single-use: one per sharded engine
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Engine.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    cases = [("linreg20", 1 << 17, 12, dict(n_mh_steps=2, n_blocks=1, adaptive=0)),
             ("threeeq_blocks_adaptive", 65000, 16, dict(n_mh_steps=1, n_blocks=3, adaptive=1)),
             ("an_schorfheide_mixture", 65000, 5, dict(n_mh_steps=2, n_blocks=1, adaptive=1, alpha=0.9, device_draw=True))]
    for name, N, n_stage, kw in cases:
        if name == "an_schorfheide_mixture":
            g = np.load(os.path.join(ROOT, "tests", "golden", "as_clouds.npz"))
            params = W.an_schorfheide_parameters()
            spec = M.make_spec(params, M.AnSchorfheideLogLik(g["data"]))
            sched = (np.arange(60) / 59.0) ** 3.0
        elif name == "linreg20":
            params, lk, _ = W.linear_gaussian(d=20, T=256, prior_sd=1.0)
            spec = M.make_spec(params, lk)
            sched = (np.arange(40) / 39.0) ** 2.1
        else:
            data, X = W.synthetic_three_equation(T=100)
            params = W.three_equation_parameters(prior_para=10.0)
            spec = M.make_spec(params, M.LinearEquationsLogLik(data, X))
            sched = (np.arange(60) / 59.0) ** 2.1
        d = spec.d
        single = Engine(local)
        single.cloud_create(N, d); single.set_model(spec)
        shard = Engine(local)
        shard.comm_init(rank, world, fresh_comm_id())
        shard.cloud_create(N, d); shard.set_model(spec)
        lo, hi = shard.first, shard.first + shard.count
        if kw.get("device_draw"):                                         # initial_draw! on the device: global-index RNG
            single.initial_draw(spec.values, 77, 1000); shard.initial_draw(spec.values, 77, 1000)
            assert np.array_equal(single.download()[lo:hi], shard.download()), "initial_draw! depends on the sharding"
        else:
            P0 = W.initial_cloud(params, N, np.random.default_rng(123))  # same global cloud on every rank
            single.upload(P0); single.evaluate(0)
            shard.upload(P0); shard.evaluate(0)
        s1 = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2)
        s2 = StageState(c=0.5, accept=0.25, ess_prev=float(N), phi_prop=0.0, j=2)
        phi_prev, nres = 0.0, 0
        for s in range(n_stage):
            cfg = StageConfig(phi_n1=phi_prev, phi_n=float(sched[s + 1]), threshold_ratio=0.5 if name != "an_schorfheide_mixture" else 0.9, target=0.25, alpha=kw.get("alpha", 1.0),
                              tempering_target=0.8, n_mh_steps=kw["n_mh_steps"], n_blocks=kw["n_blocks"], resample_method=s % 2 if name != "linreg20" else 0,
                              adaptive=kw["adaptive"], seed=1793, stage=s + 2)
            r1, inc1, nw1 = single.stage(cfg, s1, schedule=sched, want_inc=True, want_normw=True)
            r2, inc2, nw2 = shard.stage(cfg, s2, schedule=sched, want_inc=True, want_normw=True)
            assert (r1.phi_n, r1.ess, r1.sum_weights, r1.c, r1.accept, r1.resampled) == \
                   (r2.phi_n, r2.ess, r2.sum_weights, r2.c, r2.accept, r2.resampled), (name, s, r1.ess, r2.ess, r1.accept, r2.accept)
            assert np.array_equal(inc1[lo:hi], inc2) and np.array_equal(nw1[lo:hi], nw2)
            full = single.download()
            mine = shard.download()
            assert np.array_equal(full[lo:hi], mine), "%s: shard differs from the single-GPU run at stage %d" % (name, s + 2)
            nres += r1.resampled
            if rank == 0 and os.environ.get("MG_VERBOSE"):
                print(name, s, "phi", r1.phi_n, "ess", r1.ess, "res", r1.resampled, "acc", r1.accept, flush=True)
            phi_prev = r1.phi_n
        m1, c1 = single.moments()
        m2, c2 = shard.moments()
        assert np.array_equal(m1, m2) and np.array_equal(c1, c2)
        if name != "linreg20" or phi_prev < 1.0:
            # the rest of the recursion in ONE call per engine (smcb200_run_stages: no host in the loop on a fixed schedule)
            i_first = n_stage + 2
            n_left = len(sched) - i_first + 1
            if n_left > 0 and phi_prev < 1.0:
                cfg = StageConfig(phi_n1=phi_prev, phi_n=0.0, threshold_ratio=0.5, target=0.25, alpha=kw.get("alpha", 1.0), tempering_target=0.8,
                                  n_mh_steps=kw["n_mh_steps"], n_blocks=kw["n_blocks"], resample_method=0, adaptive=kw["adaptive"], seed=1793, stage=0)
                h1, h2 = np.zeros((n_left, N)), np.zeros((n_left, hi - lo))
                q1 = single.run_stages(cfg, s1, sched, i_first, n_left, normw_hist=h1)
                q2 = shard.run_stages(cfg, s2, sched, i_first, n_left, normw_hist=h2)
                assert len(q1) == len(q2) and all((a.phi_n, a.ess, a.c, a.accept, a.resampled) == (b.phi_n, b.ess, b.c, b.accept, b.resampled)
                                                  for a, b in zip(q1, q2)), name
                assert np.array_equal(h1[:len(q1), lo:hi], h2[:len(q2)])
                assert np.array_equal(single.download()[lo:hi], shard.download()), "%s: run_stages differs on the sharded engine" % name
                nres += sum(a.resampled for a in q1)
        assert nres >= (2 if name != "an_schorfheide_mixture" else 1), nres
        single.close(); shard.close()
        dist.barrier()
        if rank == 0:
            print("multigpu_check %-26s world=%d N=%d stages=%d resamples=%d: shard == single-GPU (bit-exact)" % (name, world, N, n_stage, nres))
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "mg_rank%s.log" % os.environ.get("RANK", "x")), "w") as f:
            traceback.print_exc(file=f)
        traceback.print_exc()
        raise
