"""The reference's output files (src/smc_main.jl:499-526) as written by smc_jl_b200/jld2.py: JLD2 `cloud` / `w` / `W` (/ `j`) with
the committed `SMC.Cloud` type and the HDF5 `smcparams` matrix.  The golden `jld2_structure.npz` holds every metadata byte of the
reference's own `test/reference/smc_cloud_fix=true_version=150.jld2` (tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from smc_jl_b200 import jld2  # noqa: E402
from smc_jl_b200.cloud import Cloud  # noqa: E402


def _cloud_like_the_fixture(g, rng):
    n, ncol, nst = (int(v) for v in g["shape"])
    si, nphi, nres = (int(v) for v in g["scalars"])
    c, acc, tst = (float(v) for v in g["fscalars"])
    cl = Cloud(np.asfortranarray(rng.normal(size=(n, ncol))), g["tempering_schedule"], g["ESS"], si, nphi, nres, c, acc, tst)
    return cl, rng.uniform(size=(n, nst)), rng.uniform(size=(n, nst))


def test_jld2_writer_reproduces_the_reference_file_byte_for_byte(golden, tmp_path):
    """Same shapes and scalars as the reference's fixture, random payloads: every byte that is not one of the three big
    Float64 payloads -- user block, superblock + checksum, `_types/00000001` (DataType), the global heap with the type names,
    `_types/00000002` (the SMC.Cloud compound + its julia_type attribute), the `cloud` struct with its object references, all
    array headers, the compact schedule / ESS datasets, the `_types` and root groups -- equals what JLD2.jl wrote."""
    g = golden("jld2_structure.npz")
    cl, w, W = _cloud_like_the_fixture(g, np.random.default_rng(0))
    path = str(tmp_path / "smc_cloud.jld2")
    jld2.write_jld2(path, cl, w, W)
    raw = open(path, "rb").read()
    assert len(raw) == int(g["file_size"])
    n, ncol, nst = (int(v) for v in g["shape"])
    head, mid, whdr, tail = (g[k].tobytes() for k in ("head", "mid", "whdr", "tail"))
    p0 = len(head); p1 = p0 + n * ncol * 8
    w0 = p1 + len(mid); w1 = w0 + n * nst * 8
    W0 = w1 + len(whdr); W1 = W0 + n * nst * 8
    assert raw[:p0] == head and raw[p1:w0] == mid and raw[w1:W0] == whdr and raw[W1:] == tail
    # ... and the payloads are the Julia (column-major) bytes of the matrices
    assert raw[p0:p1] == np.asfortranarray(cl.particles).tobytes(order="F")
    assert raw[w0:w1] == np.asfortranarray(w).tobytes(order="F") and raw[W0:W1] == np.asfortranarray(W).tobytes(order="F")


def test_jld2_round_trip_and_fixture_reader(golden, tmp_path):
    """read_jld2(write_jld2(...)) incl. the checkpoint key `j`, ragged shapes, and the independent survey reader."""
    from fixture_reader import JLD2
    rng = np.random.default_rng(1)
    for n, d, nst, j in ((7, 2, 1, None), (1000, 16, 33, 12), (4096, 9, 300, 2)):
        cl = Cloud(np.asfortranarray(rng.normal(size=(n, d + 5))), np.sort(rng.uniform(size=nst)), rng.uniform(1, n, size=nst), nst, 300, 3,
                   0.41, 0.27, 12.5)
        w, W = rng.uniform(size=(n, nst)), rng.uniform(size=(n, nst))
        path = str(tmp_path / ("c%d.jld2" % n))
        jld2.write_jld2(path, cl, w, W, j)
        r = jld2.read_jld2(path)
        assert sorted(r) == sorted(["cloud", "w", "W"] + (["j"] if j is not None else []))
        c2 = r["cloud"]
        assert np.array_equal(c2.particles, cl.particles) and c2.particles.flags.f_contiguous
        assert np.array_equal(c2.tempering_schedule, cl.tempering_schedule) and np.array_equal(c2.ESS, cl.ESS)
        assert (c2.stage_index, c2.n_Φ, c2.resamples, c2.c, c2.accept, c2.total_sampling_time) == (nst, 300, 3, 0.41, 0.27, 12.5)
        assert np.array_equal(r["w"], w) and np.array_equal(r["W"], W) and (j is None or int(r["j"]) == j)
        ref = JLD2(path)                                   # the survey's reader (written against the reference's files)
        assert set(ref.keys()) == set(r)
        c3 = ref["cloud"]
        assert list(c3) == list(jld2.CLOUD_FIELDS) and np.array_equal(c3["particles"], cl.particles) and c3["n_Φ"] == 300
        assert np.array_equal(ref["W"], W)
        t2 = ref._obj(ref._obj(ref._obj(ref.root)["links"]["_types"])["links"]["00000002"])["dt"]
        assert t2.cls == 6 and t2.size == 72 and [m[0] for m in t2.members] == list(jld2.CLOUD_FIELDS)


def test_lookup3_known_answers():
    """Jenkins lookup3 hashlittle (HDF5 metadata checksums): published self-test values."""
    assert jld2.lookup3(b"") == 0xDEADBEEF
    assert jld2.lookup3(b"Four score and seven years ago") == 0x17770551
    assert jld2.lookup3(b"Four score and seven years ago", 1) == 0xCD628161


def test_smcparams_hdf5_round_trip(tmp_path):
    from fixture_reader import JLD2
    A = np.random.default_rng(2).normal(size=(600, 16))
    for arr in (A, A[:5, :3]):
        path = str(tmp_path / "smcsave.h5")
        jld2.write_h5_matrix(path, "smcparams", arr)
        assert np.array_equal(jld2.read_h5_matrix(path, "smcparams"), arr)
        assert np.array_equal(JLD2(path)["smcparams"], arr)
        raw = open(path, "rb").read()
        assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 2
        assert jld2.lookup3(raw[:44]) == int.from_bytes(raw[44:48], "little")


def test_driver_containers_round_trip(tmp_path):
    """driver._save / load_cloud for both containers (keys of src/smc_main.jl:499-507: cloud, w, W, j)."""
    pytest.importorskip("smc_jl_b200.driver")
    from smc_jl_b200.driver import _save, _stage_path, load_cloud
    rng = np.random.default_rng(3)
    c = Cloud(np.asfortranarray(rng.normal(size=(50, 8))), tempering_schedule=np.linspace(0, 1, 7), ESS=np.array([50.0, 31.5]),
              stage_index=4, n_Φ=7, resamples=1, c=0.37, accept=0.22, total_sampling_time=1.5)
    w, W = rng.uniform(size=(50, 4)), rng.uniform(size=(50, 4))
    for ext in (".jld2", ".npz"):
        path = _stage_path(str(tmp_path / ("ck" + ext)), 10)
        assert path.endswith("ck_stage=10" + ext)
        _save(path, c, w, W, 5)
        c2, w2, W2, j2 = load_cloud(path)
        assert np.array_equal(c2.particles, c.particles) and np.array_equal(w2, w) and np.array_equal(W2, W) and j2 == 5
        assert (c2.stage_index, c2.n_Φ, c2.resamples, c2.c, c2.accept) == (4, 7, 1, 0.37, 0.22)


def test_check_nan_ess_message_and_debug_dump(tmp_path):
    """check_nan_ess (src/helpers.jl:270-305): the assertion text names what went wrong, and `debug_assertion` writes
    incremental_weights / normalized_weights / cloud to <savepath>_debug_assertion.jld2 before raising."""
    pytest.importorskip("smc_jl_b200.driver")
    from smc_jl_b200.driver import check_nan_ess, nan_ess_message
    from smc_jl_b200.jld2 import read_jld2
    inc, nw = np.zeros(6), np.full(6, np.nan)
    assert nan_ess_message(inc, nw) == ("No particles have non-zero weight. The squared sum of the normalized weights is returning a NaN."
                                        " Part of the reason is that one of the normalized weights is a NaN")
    assert "infinite log-likelihoods" in nan_ess_message(np.array([np.inf, 1.0]), np.array([0.0, 0.0]))
    assert "at machine-error" in nan_ess_message(np.array([0.0, 0.0]), np.array([0.0, 0.0]))
    c = Cloud(np.asfortranarray(np.arange(48.0).reshape(6, 8)), tempering_schedule=np.linspace(0, 1, 3), ESS=np.array([6.0, np.nan]),
              stage_index=2, n_Φ=3, resamples=0, c=0.5, accept=0.25, total_sampling_time=0.0)
    save = str(tmp_path / "smc_cloud.jld2")
    with pytest.raises(AssertionError, match="No particles have non-zero weight"):
        check_nan_ess(c, inc, nw, save, True)
    d = read_jld2(str(tmp_path / "smc_cloud_debug_assertion.jld2"))
    assert set(d) == {"cloud", "incremental_weights", "normalized_weights"}
    assert np.array_equal(d["incremental_weights"], inc) and np.all(np.isnan(d["normalized_weights"]))
    assert np.array_equal(d["cloud"].particles, c.particles)
    with pytest.raises(AssertionError):
        check_nan_ess(c, inc, nw, str(tmp_path / "other.jld2"), False)
    assert not (tmp_path / "other_debug_assertion.jld2").exists()
