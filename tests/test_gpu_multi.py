"""Multi-GPU parity (-m gpu, needs >= 2 devices; skipped otherwise): chains sharded over two contexts / two
GPUs with the NCCL all-gather of appended rows every K generations must reproduce the single-GPU run bit for
bit in strict sync mode (global chain id = Philox counter, rank order = chain order).  Mirrors the sync mode of
demcz_sample_par (src/demcz.jl:129-137 upstream)."""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

import demc_jl_b200 as DEMC
from demc_jl_b200 import _lib, dist as ddist
from tests.parity import example_normpdf_target

pytestmark = pytest.mark.gpu


def _run_sharded(logobj_factory, d, Z0, N, world, p2p=False, **kw):
    lib = _lib.load()
    buf = (C.c_uint8 * 128)()
    assert lib.demcz_comm_unique_id(buf) == 0
    uid = bytes(buf)
    out = [None] * world
    infos = [None] * world
    bar = threading.Barrier(world)
    err = []

    def work(r):
        try:
            n_loc, off = ddist.shard(N, r, world)
            s = DEMC.Sampler(logobj_factory(), d, N=n_loc, rank=r, world=world, N_global=N, chain_offset=off, device=r, **kw)
            s.set_archive(Z0)
            s.comm_init(uid)
            if p2p:                                         # push kernel over peer-mapped archives instead of ncclAllGather
                infos[r] = s.peer_info()
                bar.wait(60)
                s.open_peers(b"".join(infos))
                bar.wait(60)
            info = s.run()
            X, lp = s.get_state()
            out[r] = (X, lp, s.get_archive(), info, s.get_rhat() if info["rhat_valid"] else None)
            s.close()
        except Exception as e:  # noqa: BLE001
            err.append(e)
            bar.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert not err, err
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("sync_mode,p2p", [("strict", False), ("strict", True), ("overlap", False), ("overlap", True)])
def test_two_gpus_reproduce_one_gpu_bitwise(sync_mode, p2p):
    d, N, K, G = 5, 256, 5, 40
    mu, S = example_normpdf_target(d)
    Z0 = np.random.default_rng(0).normal(size=(3 * N, d))
    kw = dict(K=K, Ngeneration=G, blockindex=[list(range(d))], eps_scale=1e-5 * np.ones(d), γ=2.38, seed=7,
              store_chain=False, autostop=True, autostop_every=20, autostop_Rhat=0.0, sync_mode=sync_mode)
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, device=0, **kw) as s:
        s.set_archive(Z0)
        info1 = s.run()
        X1, lp1 = s.get_state()
        Z1 = s.get_archive()
        rhat1 = s.get_rhat()
    out = _run_sharded(lambda: DEMC.MvNormalLogpdf(mu, S), d, Z0, N, 2, p2p=p2p, **kw)
    X2 = np.concatenate([o[0] for o in out])
    lp2 = np.concatenate([o[1] for o in out])
    np.testing.assert_array_equal(X2, X1)
    np.testing.assert_array_equal(lp2, lp1)
    for o in out:
        np.testing.assert_array_equal(o[2], Z1)            # replicated archive, identical on every rank
        np.testing.assert_allclose(o[4], rhat1, rtol=1e-6)  # all-reduced Rhat (summation order differs)
    assert sum(o[3]["accepted"] for o in out) == info1["accepted"]


_one_gpu_cache = {}


def _two_rank_run(make_logobj, d, Z0, N, kw, chunks, **rank_kw):
    """Two contexts on two GPUs (threads), peer-push transport; returns [(X, lp, Z, info)] per rank."""
    lib = _lib.load()
    buf = (C.c_uint8 * 128)()
    assert lib.demcz_comm_unique_id(buf) == 0
    uid, world = bytes(buf), 2
    out, infos, err = [None] * world, [None] * world, []
    bar = threading.Barrier(world)

    def work(r):
        try:
            n_loc, off = ddist.shard(N, r, world)
            s = DEMC.Sampler(make_logobj(), d, N=n_loc, rank=r, world=world, N_global=N, chain_offset=off, device=r, **rank_kw, **kw)
            s.set_archive(Z0)
            s.comm_init(uid)
            infos[r] = s.peer_info()
            bar.wait(60)
            s.open_peers(b"".join(infos))
            bar.wait(60)
            for c in chunks:
                s.run(c)
            info = s.run()
            X, lp = s.get_state()
            out[r] = (X, lp, s.get_archive(), info, s.get_archive_checksum())
            bar.wait(60)
            s.close()
        except Exception as e:  # noqa: BLE001
            err.append(e)
            bar.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert not err, err
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("d,sync_mode,layout,engine", [
    (10, "strict", "in_place", "kernel"), (10, "overlap", "in_place", "copy_engine"), (13, "overlap", "in_place", "kernel"),
    (20, "strict", "dense_slots", "kernel"), (20, "overlap", "dense_slots", "kernel"), (20, "overlap", "dense_slots", "copy_engine"),
    (20, "strict", "dense_slots", "copy_engine"), (20, "overlap", "in_place", "copy_engine"), (20, "overlap", "in_place", "kernel")])
def test_two_gpus_exchange_layouts_bitwise(d, sync_mode, layout, engine):
    """The ways a rank's appended rows reach the peers' replicas — padded slots or an archive without slot padding
    (dense_slots), moved by our push kernel or by the copy engines, the step kernels polling the segment flag words in
    their prologue — all give the state / archive / accept counts of one GPU bit for bit.  Two run calls, so that the
    drain between them is exercised."""
    N, K, G = 512, 3, 36
    S = 0.3 * np.ones((d, d)) + 0.7 * np.eye(d)
    mu = np.linspace(-1, 1, d)
    Z0 = np.random.default_rng(d).normal(size=(2 * N, d))
    kw = dict(K=K, Ngeneration=G, blockindex=[list(range(d))], eps_scale=1e-5 * np.ones(d), γ=2.38, seed=11,
              store_chain=False, sync_mode=sync_mode)
    if (d, sync_mode) not in _one_gpu_cache:
        with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, device=0, **kw) as s:
            s.set_archive(Z0)
            s.run(G // 2)
            info1 = s.run()
            _one_gpu_cache[(d, sync_mode)] = (s.get_state(), s.get_archive(), info1["accepted"], s.get_archive_checksum())
    (X1, lp1), Z1, acc1, h1 = _one_gpu_cache[(d, sync_mode)]
    out = _two_rank_run(lambda: DEMC.MvNormalLogpdf(mu, S), d, Z0, N, kw, [G // 2], exchange_layout=layout, push_engine=engine)
    np.testing.assert_array_equal(np.concatenate([o[0] for o in out]), X1)
    np.testing.assert_array_equal(np.concatenate([o[1] for o in out]), lp1)
    for o in out:
        np.testing.assert_array_equal(o[2], Z1)
        assert o[4] == h1                                  # the device-side checksum agrees with the row-by-row comparison
    assert sum(o[3]["accepted"] for o in out) == acc1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("sync_mode,layout,blocks", [("overlap", "in_place", 1), ("overlap", "dense_slots", 1), ("strict", "in_place", 1),
                                                     ("overlap", "in_place", 2)])
def test_two_gpus_ring_wrap_bitwise(sync_mode, layout, blocks):
    """The archive ring wraps several times during the run.  In overlap mode a peer may push its NEXT segment into this
    replica while this rank is still drawing, so the view keeps off the 2 N oldest rows (demcz_archive_view) — for every
    world size, hence still bit-identical to one GPU."""
    d, N, K, G = 20, 512, 2, 60
    S = 0.3 * np.ones((d, d)) + 0.7 * np.eye(d)
    mu = np.linspace(-1, 1, d)
    Z0 = np.random.default_rng(5).normal(size=(2 * N, d))
    bi = [list(range(d))] if blocks == 1 else [list(range(0, 7)), list(range(7, d))]
    kw = dict(K=K, Ngeneration=G, blockindex=bi, eps_scale=1e-5 * np.ones(d), γ=2.38, seed=13,
              store_chain=False, sync_mode=sync_mode, archive_capacity=6 * N)          # 30 appends into a 6-segment ring
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, device=0, **kw) as s:
        s.set_archive(Z0)
        info1 = s.run()
        (X1, lp1), Z1 = s.get_state(), s.get_archive()
    assert Z1.shape[0] == 6 * N
    out = _two_rank_run(lambda: DEMC.MvNormalLogpdf(mu, S), d, Z0, N, kw, [], exchange_layout=layout)
    np.testing.assert_array_equal(np.concatenate([o[0] for o in out]), X1)
    np.testing.assert_array_equal(np.concatenate([o[1] for o in out]), lp1)
    for o in out:
        np.testing.assert_array_equal(o[2], Z1)
    assert sum(o[3]["accepted"] for o in out) == info1["accepted"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_wide_rows_and_linreg_bitwise():
    # (a) d = 24: > 48 KB of dynamic shared memory per CTA — kernel attributes must be set on every device
    d, N = 24, 256
    S = 0.5 * np.ones((d, d)) + 0.5 * np.eye(d)
    mu = np.linspace(0, 1, d)
    Z0 = np.random.default_rng(1).normal(size=(2 * N, d))
    kw = dict(K=4, Ngeneration=16, blockindex=[list(range(d))], eps_scale=1e-5 * np.ones(d), γ=2.38, seed=3, store_chain=False)
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, device=0, **kw) as s:
        s.set_archive(Z0); s.run(); X1, lp1 = s.get_state(); Z1 = s.get_archive()
    out = _run_sharded(lambda: DEMC.MvNormalLogpdf(mu, S), d, Z0, N, 2, p2p=True, **kw)
    np.testing.assert_array_equal(np.concatenate([o[0] for o in out]), X1)
    np.testing.assert_array_equal(out[1][2], Z1)
    # (b) the tcgen05 linreg path, chains sharded, X replicated
    rng = np.random.default_rng(2)
    n, p = 3000, 9
    X = np.ones((n, p)); X[:, 1:] = rng.normal(size=(n, p - 1)); beta = 1 + 3 * rng.random(p); y = X @ beta + rng.normal(size=n)
    bols = np.linalg.solve(X.T @ X, X.T @ y)
    Z0 = bols + 0.05 * rng.normal(size=(2 * N, p))
    kw = dict(K=5, Ngeneration=15, blockindex=[list(range(p))], eps_scale=1e-5 * np.ones(p), γ=2.0, seed=5, store_chain=False)
    with DEMC.Sampler(DEMC.LinRegLoglik(X, y), p, N=N, device=0, **kw) as s:
        s.set_archive(Z0); s.run(); X1, lp1 = s.get_state(); Z1 = s.get_archive()
    out = _run_sharded(lambda: DEMC.LinRegLoglik(X, y), p, Z0, N, 2, p2p=True, sync_mode="strict", **kw)
    np.testing.assert_array_equal(np.concatenate([o[0] for o in out]), X1)
    np.testing.assert_array_equal(np.concatenate([o[1] for o in out]), lp1)
    np.testing.assert_array_equal(out[0][2], Z1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_par_entry_points_under_torchrun(tmp_path):
    """demcz_sample_par / demcz_anneal_par (src/demcz.jl:101-165, src/demcz_anneal.jl:82-140) launched under torchrun with
    two ranks shard the chains over two GPUs (strict sync) and return, on every rank, the whole MC record and archive —
    bit-identical to the same call in one process on one GPU (R-hat to 1e-6: all-reduced sums)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    worker = os.path.join(root, "tests", "par_worker.py")
    one, two = str(tmp_path / "one.npz"), str(tmp_path / "two.npz")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, worker, one], capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", worker, two], capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    a, b = np.load(one), np.load(two)
    for k in ("chain", "log_obj", "X", "lp", "Z", "a_chain", "a_X", "a_lp", "a_Z"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    np.testing.assert_allclose(a["rhat"], b["rhat"], rtol=1e-6)
    assert int(a["accepted"]) == int(b["accepted"]) and int(a["a_accepted"]) == int(b["a_accepted"])
    assert float(a["a_gamma"]) == float(b["a_gamma"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind,sync_mode", [("mvn", "overlap"), ("mvn", "strict"), ("logistic", "overlap")])
def test_two_gpus_snooker_and_logistic_bitwise(kind, sync_mode):
    """The round-2 extensions on a sharded run: snooker generations (own kernel, polls the segment flags itself), R-hat
    accumulators and history moments across ranks, the logistic likelihood — bit-identical to one GPU."""
    rng = np.random.default_rng(4)
    N, K, G = 512, 4, 32
    if kind == "mvn":
        d = 9
        S = 0.4 * np.ones((d, d)) + 0.6 * np.eye(d)
        mu = np.linspace(-1, 1, d)
        make = lambda: DEMC.MvNormalLogpdf(mu, S)
        Z0 = rng.normal(size=(2 * N, d))
    else:
        n, d = 3000, 7
        X = np.ones((n, d)); X[:, 1:] = rng.normal(size=(n, d - 1))
        beta = rng.normal(size=d) / 3
        y = (rng.random(n) < 1 / (1 + np.exp(-X @ beta))).astype(float)
        make = lambda: DEMC.LogisticLoglik(X, y)
        Z0 = beta + 0.2 * rng.normal(size=(2 * N, d))
    kw = dict(K=K, Ngeneration=G, blockindex=[list(range(d))], eps_scale=1e-5 * np.ones(d), γ=2.2, seed=21, store_chain=False,
              sync_mode=sync_mode, snooker_prob=0.3, autostop=True, autostop_every=16, autostop_Rhat=0.0, moments=(kind == "mvn"))
    with DEMC.Sampler(make(), d, N=N, device=0, **kw) as s:
        s.set_archive(Z0)
        info1 = s.run()
        (X1, lp1), Z1, rhat1 = s.get_state(), s.get_archive(), s.get_rhat()
        mom1 = s.get_history_moments() if kind == "mvn" else None
    lib = _lib.load()
    buf = (C.c_uint8 * 128)()
    assert lib.demcz_comm_unique_id(buf) == 0
    uid, world = bytes(buf), 2
    out, infos, err = [None] * world, [None] * world, []
    bar = threading.Barrier(world)

    def work(r):
        try:
            n_loc, off = ddist.shard(N, r, world)
            s = DEMC.Sampler(make(), d, N=n_loc, rank=r, world=world, N_global=N, chain_offset=off, device=r, **kw)
            s.set_archive(Z0)
            s.comm_init(uid)
            infos[r] = s.peer_info()
            bar.wait(60)
            s.open_peers(b"".join(infos))
            bar.wait(60)
            info = s.run()
            mom = s.get_history_moments() if kind == "mvn" else None        # collective (all-reduce): every rank calls it
            out[r] = (s.get_state(), s.get_archive(), info, s.get_rhat(), mom)
            bar.wait(60)
            s.close()
        except Exception as e:  # noqa: BLE001
            err.append(e)
            bar.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert not err, err
    np.testing.assert_array_equal(np.concatenate([o[0][0] for o in out]), X1)
    np.testing.assert_array_equal(np.concatenate([o[0][1] for o in out]), lp1)
    for o in out:
        np.testing.assert_array_equal(o[1], Z1)
        np.testing.assert_allclose(o[3], rhat1, rtol=1e-6)
        if mom1 is not None:
            assert o[4][2] == mom1[2]
            # FP32 partial sums per warp, FP32 shared-memory atomics per CTA (order varies), FP64 across CTAs and ranks
            np.testing.assert_allclose(o[4][0], mom1[0], rtol=1e-6, atol=1e-9)
            np.testing.assert_allclose(o[4][1], mom1[1], rtol=1e-5, atol=1e-9)
    assert sum(o[2]["accepted"] for o in out) == info1["accepted"]
