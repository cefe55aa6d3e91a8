"""TEST MIRROR of the multi-GPU engine's sharding geometry (smcb200_cloud_create) and cross-rank tree (peer_exchange_block);
used by tests/test_sharding_gloo.py only -- the product computes both inside libsmcb200.

Particles are split into contiguous ranges of the zero-padded power-of-two index space so that every
canonical reduction tree is shard-aligned: per-rank roots combined by an adjacent-pair tree in rank order
reproduce the single-GPU result bit-for-bit.
"""
import numpy as np

MIN_PER_RANK = 4096   # SCAN_TILE: smallest aligned shard of the canonical cumsum


def next_pow2(n):
    p = 1
    while p < n:
        p <<= 1
    return p


def shard_range(n_parts, world, rank):
    """(first, count, per) of `rank`'s shard; raises like the library for unusable geometries."""
    if world < 1 or world & (world - 1) or world > 16 or not (0 <= rank < world):
        raise ValueError("world must be a power of two <= 16 and 0 <= rank < world")
    per = next_pow2(n_parts) // world
    if world > 1 and per % MIN_PER_RANK:
        raise NotImplementedError("multi-GPU needs at least 4096 (padded) particles per rank")
    first = min(per * rank, n_parts)
    last = min(first + per, n_parts)
    if last - first < 1:
        raise ValueError("empty shard: fewer particles than ranks")
    return first, last - first, per


def combine_ranks(values):
    """Adjacent-pair tree over per-rank values (rank order); values: sequence of floats / arrays."""
    v = [np.asarray(x, dtype=np.float64) for x in values]
    assert len(v) & (len(v) - 1) == 0
    while len(v) > 1:
        v = [v[i] + v[i + 1] for i in range(0, len(v), 2)]
    return v[0]
