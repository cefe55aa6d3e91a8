"""world_size-2 `gloo` tests (CPU) of the host-side multi-GPU logic: shard geometry, the cross-rank
combination tree, and shard-invariance of correction and mutation when every rank works on its own shard
and exchanges only per-shard roots -- checked against the single-process oracle, bit-for-bit."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from smc_jl_b200 import model as M
from smc_jl_b200 import workloads as W
from sharding_mirror import combine_ranks, shard_range


def test_shard_geometry():
    assert shard_range(1 << 20, 1, 0) == (0, 1 << 20, 1 << 20)
    assert [shard_range(1 << 20, 8, r)[:2] for r in range(8)] == [(r << 17, 1 << 17) for r in range(8)]
    assert [shard_range(40000, 2, r)[:2] for r in range(2)] == [(0, 32768), (32768, 7232)]      # uneven tail
    with pytest.raises(ValueError):
        shard_range(40000, 8, 5)                     # 40000 < 5 * 8192: empty shard
    with pytest.raises(NotImplementedError):
        shard_range(5000, 4, 0)                      # 8192 / 4 = 2048 < 4096
    with pytest.raises(ValueError):
        shard_range(1 << 20, 3, 0)


def _worker(rank, world, port, N, d, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = O.lib()
        params, lk, _ = W.linear_gaussian(d=d, T=64, prior_sd=1.0)
        spec = M.make_spec(params, lk)
        mod = O.Model(spec)
        P = W.initial_cloud(params, N, np.random.default_rng(5))          # the same global cloud on every rank
        buf = O.cloud_f(P)
        L.orc_evaluate(mod.h, buf, N)
        full = O.cloud_m(buf, N, d)
        first, count, per = shard_range(N, world, rank)
        mine = np.asfortranarray(full[first:first + count])

        def allgather(x):
            t = torch.tensor(np.atleast_1d(x), dtype=torch.float64)
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [o.numpy() for o in out]

        # ---- correction on shards: local roots -> all_gather -> rank tree ------------------------------
        phi_n1, phi_n = 0.0, 0.004
        ll, w = mine[:, d], mine[:, d + 4].copy()
        inc = np.array([L.orc_exp((phi_n1 - phi_n) * 0.0 + (phi_n - phi_n1) * v) for v in ll])
        wt = w * inc
        S = float(combine_ranks(allgather(L.orc_canon_sum(np.ascontiguousarray(wt), count)))[0])
        Wn = (wt * N) / S
        Q = float(combine_ranks(allgather(L.orc_canon_sumsq(np.ascontiguousarray(Wn), count)))[0])
        ess = (float(N) * float(N)) / Q
        ref = O.cloud_f(full)
        out = np.zeros(3)
        assert L.orc_correct(ref, N, d, phi_n1, phi_n, 0.0, 0.0, None, None, out) == 0
        refm = O.cloud_m(ref, N, d)
        ok = (S == out[0]) and (ess == out[1]) and np.array_equal(Wn, refm[first:first + count, d + 4])

        # ---- mutation on shards: RNG keyed on the global particle index -------------------------------------
        mean = np.average(full[:, :d], axis=0, weights=refm[:, d + 4])
        cov = np.cov(full[:, :d].T, aweights=refm[:, d + 4], bias=True).reshape(d, d)
        cov = (cov + cov.T) / 2
        st = C.c_int()
        blk = np.arange(d, dtype=np.int32)
        pr = L.orc_proposal_create(d, d, np.ascontiguousarray(mean), np.ascontiguousarray(cov), 1, np.array([d], np.int32), blk,
                                   blk, 0.4, C.byref(st))
        assert pr and st.value == 0
        shard_buf = O.cloud_f(refm[first:first + count])
        L.orc_mutate(mod.h, pr, shard_buf, count, first, phi_n, phi_n1, 1.0, 2, d, 0, 99, 7, 1)
        whole = O.cloud_f(refm)
        L.orc_mutate(mod.h, pr, whole, N, 0, phi_n, phi_n1, 1.0, 2, d, 0, 99, 7, 1)
        L.orc_proposal_free(pr)
        ok = ok and np.array_equal(O.cloud_m(shard_buf, count, d), O.cloud_m(whole, N, d)[first:first + count])
        # accept mean: the shards' exact integer totals (accept column = count / n_free), combined across ranks
        tot = np.rint(O.cloud_m(shard_buf, count, d)[:, d + 3] * d).sum()
        acc = (float(combine_ranks(allgather(float(tot)))[0]) / d) / N
        ok = ok and acc == L.orc_mean_accept(whole, N, d, d)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [1 << 14, 12000])
def test_two_rank_shards_match_single_process(N):
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + (N % 97)
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, 4, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True


def _gather_worker(rank, world, port, tmpdir, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smc_jl_b200.driver import _Group
        grp = _Group(rank, world, dist, tmpdir if rank == 0 else None)
        first, count, _ = shard_range(20000, world, rank)
        full = np.arange(20000 * 3, dtype=np.float64).reshape(20000, 3)
        ok = True
        for tag in ("particles", "w"):                                     # two gathers in a row reuse the directory
            got = grp.gather_rows(full[first:first + count] + (tag == "w"), tag)
            ok = ok and ((got is None) if rank else np.array_equal(got, full + (tag == "w")))
        ok = ok and not [f for f in os.listdir(tmpdir) if f.endswith("_%d.npy" % rank)]   # shard files are removed
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_smc_result_gathering(tmp_path):
    """Host side of `smc(...; n_gpus = G)`: every rank's rows of the cloud / history reach rank 0 in rank order."""
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, 29731, str(tmp_path), ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True
