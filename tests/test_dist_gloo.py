"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): the chain partition, the NCCL
unique-id broadcast, and the exchange protocol itself — chains sharded contiguously with global
chain ids as Philox counters, archive replicated, every K generations an all-gather of the appended
rows in rank order — emulated with the CPU oracle as the per-rank compute.  The sharded run must
reproduce the single-rank run bit for bit (the property the NCCL path relies on; mirrors the sync
mode of demcz_sample_par, src/demcz.jl:129-137 upstream)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from demc_jl_b200 import dist as ddist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_partition():
    assert ddist.shard(1 << 20, 3, 8) == (131072, 393216)
    assert ddist.shard(100, 0, 1) == (100, 0)
    offs = [ddist.shard(64, r, 4) for r in range(4)]
    assert [o for _, o in offs] == [0, 16, 32, 48] and all(n == 16 for n, _ in offs)
    with pytest.raises(ValueError):
        ddist.shard(10, 0, 4)
    with pytest.raises(ValueError):
        ddist.shard(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import demcz_oracle as orc
        co = orc.COracle(build=False)
        # --- unique id broadcast: only rank 0's bytes count ---------------------------------------
        uid = ddist.broadcast_unique_id(lambda: bytes(range(128)) if rank == 0 else b"\xff" * 128, rank)
        assert uid == bytes(range(128))
        assert ddist.env_rank_world()[:2] == (rank, world)

        # --- sharded generation loop with an all-gather of the appended rows every K generations ----
        d, N, K, G, M0, seed = 4, 24, 3, 12, 40, 77
        rng = np.random.default_rng(0)                       # same on every rank: replicated archive
        mu = rng.random(d)
        A = rng.random((d, d))
        lik = orc.MvNormalLik(mu, A @ A.T + np.eye(d))
        Z0 = rng.normal(size=(M0, d))
        blocks = [[0, 1], [2, 3]]
        eps = 1e-3 * np.ones(d)
        n_loc, off = ddist.shard(N, rank, world)
        Z = np.zeros((M0 + N * (G // K), d), order="F")
        Z[:M0] = Z0
        M = M0
        X = np.asfortranarray(Z0[M0 - N + off: M0 - N + off + n_loc])      # last N rows, this rank's slice
        lp = lik.logdensity(X)
        for g in range(1, G + 1):
            X, lp, _, _, _ = co.generation(X, lp, Z, M, lik, blocks, eps, 2.38, seed, g, chain_offset=off)
            if g % K == 0:
                rows = [torch.zeros(n_loc, d, dtype=torch.float64) for _ in range(world)]
                dist.all_gather(rows, torch.from_numpy(np.ascontiguousarray(X)))
                for r, t in enumerate(rows):                  # rank order == global chain order
                    Z[M + r * n_loc: M + (r + 1) * n_loc] = t.numpy()
                M += N
        allX = [torch.zeros(n_loc, d, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allX, torch.from_numpy(np.ascontiguousarray(X)))
        if rank == 0:
            ref = co.run(lik, Z0, N, K, G, blocks, eps, 2.38, seed)
            Xall = np.concatenate([t.numpy() for t in allX])
            q.put((bool(np.array_equal(Xall, ref["X"])), bool(np.array_equal(Z[:M], ref["Z"])), int(M), int(ref["M"])))
    finally:
        dist.destroy_process_group()


def test_world2_sharded_protocol_matches_single_rank(coracle):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same_X, same_Z, M, Mref = q.get()
    assert M == Mref
    assert same_X and same_Z


def _helpers_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    try:
        ddist.ensure_process_group(rank)                     # gloo here (no CUDA): the torchrun environment is enough
        assert dist.is_initialized() and dist.get_world_size() == world
        a = np.asfortranarray(np.arange(12, dtype=np.float64).reshape(3, 2, 2) + 100 * rank)
        g = ddist.allgather_rows(a, rank)                    # rank order = global chain order
        assert g.shape == (3 * world, 2, 2) and np.array_equal(g[:3], a - 100 * rank) and np.array_equal(g[3:6], a - 100 * rank + 100)
        assert ddist.allreduce_sum(rank + 1.0, rank) == sum(range(1, world + 1))
        ddist.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_sharded_entry_point_helpers_gloo():
    """Host plumbing of the _par entry points (api._run, sharded): process group from the torchrun environment, rank-ordered
    all-gather of the per-rank results, reductions — on CPU with gloo, world size 2."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_helpers_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(60)
    assert res == {0: "ok", 1: "ok"}, res


def test_pad_archive_for_sharding():
    Z = np.arange(14, dtype=np.float64).reshape(7, 2)
    Zp, cap = ddist.pad_archive_for_sharding(Z, 4, 25, 10)          # 7 rows -> 8 (one oldest row repeated in front), 3 appends of 4
    assert Zp.shape == (8, 2) and np.array_equal(Zp[1:], Z) and np.array_equal(Zp[0], Z[0]) and cap == 8 + 4 * 3
    Zq, cap = ddist.pad_archive_for_sharding(Z[:4], 4, 20, 10)
    assert Zq.shape == (4, 2) and cap == 4 + 8
