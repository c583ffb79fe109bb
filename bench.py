#!/usr/bin/env python
"""bench.py — chain-updates/s of the DEMCz generation loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c3|c1|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = `--gens-per-step` generations of every chain (linreg: 10 = one K-block; closed-form likelihoods: 100 =
10 K-blocks, 10 archive appends).  The default workload is BASELINE configs[1] (C2: linreg 10^6 x 32, 4096 chains per
GPU, the tcgen05 likelihood); without --workload the line also carries short runs of the other BASELINE configs under
the key `also` (C4: the 1M-chain MvNormal scaling config at 131072 chains per GPU; C3; C1; C5 on 8 GPUs).

Rank 0 prints ONE JSON line.
  value          chain-updates/s, CUDA events on the library's stream around exactly K steps, inputs resident in HBM
  e2e            the same through the public Sampler API with HOST buffers (pinned) in and out every step
  roofline       the dominant kernel against the measured peak (MEASURED_PEAKS.json), per-launch CUDA-event durations
  cpu_baseline   the CPU restatement of the reference (oracle/, all host cores) on a bounded sample + a batched-DGEMM arm
  parity_probe   AFTER the timed run, at the benchmark's full size: (linreg) log-objective of sampled chains re-evaluated
                 in FP64 numpy; (all) one traced K-block of sampled chains replayed proposal by proposal through the CPU
                 oracle — indices / uniforms bit-exact, proposals, log-densities to 1e-5 relative, decisions exact
                 outside the tolerance band.  The oracle is the CHECKER here, never on the measured path.
  consistency    (N > 1) archive replicas identical on every rank (64-bit checksum) and a small strict-mode config
                 sharded over all ranks bit-identical to a world=1 context
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# rank 0 prints exactly ONE line on stdout: keep NCCL's banner / debug output off it
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

LP_RTOL = 1e-5          # north star: log-densities within 1e-5 relative, decisions exact outside that band


# ------------------------------------------------------------------------------------------------
# workloads (synthetic data of each config's shape, fixed seeds)
# ------------------------------------------------------------------------------------------------
def workload(name, chains_per_gpu=0, archive_rows=0):
    if name == "c4":      # 20-dim highly correlated MvNormal, K=10, Rhat autostop machinery on
        d = 20
        S = 0.95 * np.ones((d, d)) + 0.05 * np.eye(d)
        mu = np.linspace(-1.0, 1.0, d)
        return dict(key="c4", name="C4 d=20 equicorrelated(0.95) MvNormal, K=10, Rhat accumulators on", d=d, mu=mu, Sigma=S,
                    kind="mvn", K=10, blocks=[list(range(d))], anneal=False, autostop=True, gamma=2.38,
                    chains_per_gpu=chains_per_gpu or 131072, archive_rows=archive_rows or (8 << 20))
    if name == "c3":      # demcz_anneal on 10-dim quadratic objective, Nblocks=2, 65536 chains
        d = 10
        mu = np.random.default_rng(31953150).random(d)
        return dict(key="c3", name="C3 demcz_anneal d=10 quadratic, Nblocks=2, T0=2->TN=1e-4", d=d, mu=mu, Sigma=0.5 * np.eye(d),
                    kind="quad", K=10, blocks=[list(range(5)), list(range(5, 10))], anneal=True, autostop=False,
                    gamma=2.38, chains_per_gpu=chains_per_gpu or 65536, archive_rows=archive_rows or (8 << 20))
    if name == "c1":      # test/example_normpdf.jl: 5-dim MvNormal, N=100, K=10
        d = 5
        rng = np.random.default_rng(31953150)
        mu = rng.random(d)
        A = rng.random((d, d))
        S = A.T @ A + 2 * np.eye(d)
        S = S / S.max() / 100
        return dict(key="c1", name="C1 example_normpdf d=5 MvNormal N=100", d=d, mu=mu, Sigma=S, kind="mvn", K=10,
                    blocks=[list(range(d))], anneal=False, autostop=False, gamma=2.38,
                    chains_per_gpu=chains_per_gpu or 100, archive_rows=archive_rows or 150)
    if name in ("c2", "c5"):   # test/example_linreg.jl scaled up: the tensor-core GEMM likelihood
        n, p, N = (1_000_000, 32, 4096) if name == "c2" else (10_000_000, 64, 32768)
        return dict(key=name, name=f"{name.upper()} linreg n={n} x p={p} (example_linreg.jl generator), K=10, gamma=2.0", d=p, kind="linreg",
                    n=n, K=10, blocks=[list(range(p))], anneal=False, autostop=False, gamma=2.0,
                    chains_per_gpu=chains_per_gpu or N, archive_rows=archive_rows or 2 * (chains_per_gpu or N) * 8)
    raise SystemExit(f"unknown workload {name}")


def linreg_data(n, p, seed=319531501):
    """test/example_linreg.jl:9-29: intercept + (p-1) unit-variance regressors with pairwise correlation 0.25,
    beta = 1 + 3 U(0,1), y = X beta + N(0,1)."""
    rng = np.random.default_rng(seed)
    X = np.empty((n, p), order="F")
    X[:, 0] = 1.0
    z0 = rng.standard_normal(n, dtype=np.float32).astype(np.float64)
    for k in range(1, p):
        X[:, k] = 0.5 * z0 + np.sqrt(0.75) * rng.standard_normal(n, dtype=np.float32)
    beta = 1 + 3 * rng.random(p)
    y = X @ beta + rng.standard_normal(n)
    return X, y


def make_archive(w, rows, seed=7):
    """Z0 ~ N(mu, 4*Sigma): over-dispersed start, identical on every rank (replicated archive).  Values are rounded to
    FP32 (the device archive's type) so that a host-side checker sees exactly the rows the kernels gather."""
    rng = np.random.default_rng(seed)
    if w["kind"] == "linreg":      # around the OLS point, 3x the posterior scale
        Z = w["bols"] + 3.0 * w["post_sd"] * rng.standard_normal((rows, w["d"]))
        return np.asfortranarray(Z.astype(np.float32).astype(np.float64))
    L = np.linalg.cholesky(w["Sigma"])
    Z = np.empty((rows, w["d"]), order="F")
    step = 1 << 20
    for i in range(0, rows, step):
        n = min(step, rows - i)
        Z[i:i + n] = (w["mu"] + 2.0 * (rng.standard_normal((n, w["d"]), dtype=np.float32).astype(np.float64) @ L.T)).astype(np.float32)
    return Z


def algorithmic_bytes_per_update(d, K, G_f):
    """B(G_f) = 8d [two archive rows] + (8d+8)/G_f [X, logp read+write] + 4d/K [append]."""
    return 8 * d + (8 * d + 8) / G_f + 4 * d / K


def prepare_data(w):
    """Synthetic likelihood data of the workload's shape (identical on every rank)."""
    if w["kind"] == "linreg" and "X" not in w:
        w["X"], w["y"] = linreg_data(w["n"], w["d"])
        XtX = w["X"].T @ w["X"]
        w["bols"] = np.linalg.solve(XtX, w["X"].T @ w["y"])
        w["post_sd"] = np.sqrt(np.diag(np.linalg.inv(XtX)))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        hot = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return dict(sm_mhz=float(np.median(hot)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), samples=len(sm),
                    reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------
# CPU arms: the restatement of the reference loop (oracle/), all host threads
# ------------------------------------------------------------------------------------------------
def _oracle():
    from oracle import demcz_oracle as orc      # the checker, timed as the CPU baseline (never on the product path)
    return orc


def cpu_run(w, *, chains, gens, archive_rows, threads, faithful=False, seed=31953150):
    """The C restatement of the reference's loop (per-chain closure evaluation, OpenMP over chains)."""
    orc = _oracle()
    co = orc.COracle(build=not os.path.exists(os.path.join(ROOT, "oracle", "libdemcz_oracle.so")))
    if w["kind"] == "linreg":
        lik = orc.LinRegLik(w["X"], w["y"])
    else:
        lik = orc.MvNormalLik(w["mu"], w["Sigma"], drop_constant=(w["kind"] == "quad"))
    Z0 = make_archive(w, max(archive_rows, chains), seed=7)
    kw = dict(anneal=w["anneal"], store_chain=False, nthreads=threads, z_rowmajor=0 if faithful else 1,
              index_mode=1 if faithful else 0)
    if w["anneal"]:
        kw["Ttable"] = np.array([orc.tempbaseline(ig, gens, 2.0, 1e-4) for ig in range(1, gens + 1)])
    r = co.run(lik, Z0, chains, w["K"], gens, w["blocks"], 1e-5 * np.ones(w["d"]), w["gamma"], seed, **kw)
    dt = r["seconds"]                            # the C generation loop only (archive set-up excluded)
    return chains * gens / dt, dt


def cpu_run_batched(w, *, chains, gens, archive_rows, seed=31953150):
    """A CPU arm that is NOT a strawman: the numpy restatement, generation-synchronous, proposals for ALL sampled chains
    at once and ONE X @ B^T (multi-threaded DGEMM, all cores) per generation and block — the batching the GPU path uses.
    (The reference itself evaluates the closure chain by chain: src/demcz.jl:189.)"""
    orc = _oracle()
    if w["kind"] == "linreg":
        class _ChunkedLinReg:                      # same formula as orc.LinRegLik, rows streamed in cache-sized chunks: DGEMM-bound
            def logdensity(self, B):
                return _linreg_logdensity_chunked(w, np.atleast_2d(B), rows_per_chunk=512)      # fastest chunk measured (L2-resident residual block)
        lik = _ChunkedLinReg()
    else:
        lik = orc.MvNormalLik(w["mu"], w["Sigma"], drop_constant=(w["kind"] == "quad"))
    Z = make_archive(w, max(archive_rows, chains), seed=7)
    M = Z.shape[0]
    X = np.array(Z[M - chains:])
    lp = lik.logdensity(X)
    eps = 1e-5 * np.ones(w["d"])
    t0 = time.perf_counter()
    for g in range(1, gens + 1):
        T = orc.tempbaseline(g, gens, 2.0, 1e-4) if w["anneal"] else None
        X, lp, _ = orc.generation(X, lp, Z, M, lik, w["blocks"], eps, w["gamma"], seed, g, T=T)
    dt = time.perf_counter() - t0
    return chains * gens / dt, dt


def cpu_sample(w, threads):
    """The bounded sample both CPU lines (cpu_baseline of our arm, --impl reference) time: (chains, calibration gens)."""
    if w["kind"] == "linreg":      # one proposal streams the whole data set: a few chains per thread is already seconds
        return threads * 2, 2
    chains = min(w["chains_per_gpu"], 16384)
    return max(chains // threads * threads, min(w["chains_per_gpu"], threads)), 20


def cpu_plan(w, threads, budget_s):
    """(chains, generations, archive rows) of the timed CPU sample: ~budget_s seconds of work, calibrated by a short run."""
    arch = min(w["archive_rows"], 1 << 20)
    chains, g0 = cpu_sample(w, threads)
    v, _ = cpu_run(w, chains=chains, gens=g0, archive_rows=arch, threads=threads)            # calibrate
    gens = int(max(g0, min(2000, budget_s * v / chains)))
    return chains, gens, arch


def cpu_baseline(w, budget_s=2.0):
    threads = os.cpu_count() or 1
    chains, gens, arch = cpu_plan(w, threads, budget_s)
    v, dt = cpu_run(w, chains=chains, gens=gens, archive_rows=arch, threads=threads)
    out = dict(value=v, unit="chain-updates/s", cores=threads, kind="port",
               sample=f"{chains} chains x {gens} generations of the same target, archive {max(arch, chains)} rows, FP64, O(1) index draw, "
                      f"row-major archive, OpenMP over chains ({dt:.1f} s) - the same sample as one step of --impl reference")
    # batched arm: all sampled chains per DGEMM
    bch = 256 if w["kind"] == "linreg" else min(w["chains_per_gpu"], 16384)
    if w["kind"] == "linreg" and w["n"] * bch * 8 > 8e9:
        bch = max(32, int(8e9 / (w["n"] * 8)) // 32 * 32)
    bg = 2
    bv, bdt = cpu_run_batched(w, chains=bch, gens=bg, archive_rows=arch)
    if bdt < 2.0:
        bg = int(min(200, max(2, 4.0 / (bdt / bg))))
        bv, bdt = cpu_run_batched(w, chains=bch, gens=bg, archive_rows=arch)
    out["batched"] = dict(value=bv, unit="chain-updates/s", cores=threads, kind="port (numpy restatement, generation-synchronous)",
                          sample=f"{bch} chains x {bg} generations, one X @ B^T DGEMM per generation over all sampled chains "
                                 f"(multi-threaded BLAS), archive {max(arch, bch)} rows ({bdt:.1f} s)")
    # reference-faithful serial mode: O(M) index-set construction per proposal (src/demcz.jl:176-178), 1 thread
    fch = min(chains, 64) if w["kind"] != "linreg" else 4
    fg = 10 if w["kind"] != "linreg" else 2
    fv, fdt = cpu_run(w, chains=fch, gens=fg, archive_rows=min(arch, 1 << 17), threads=1, faithful=True)
    out["faithful_serial"] = dict(value=fv, cores=1, sample=f"{fch} chains x {fg} generations, archive {max(min(arch, 1 << 17), fch)} rows "
                                                            f"({fdt:.1f} s), collect(1:M)+deleteat! per proposal")
    return out


def measure_tf32_peak():
    """Dense TF32 tensor throughput of this GPU (cuBLAS via torch.matmul, 8192^3, best of 10): one of the two
    denominators of the linreg roofline.  Library call used for the PEAK only."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12


# ------------------------------------------------------------------------------------------------
# parity probes (the oracle as CHECKER, after the timed region, at the benchmark's size)
# ------------------------------------------------------------------------------------------------
def _oracle_lik(w):
    orc = _oracle()
    if w["kind"] == "linreg":
        return orc.LinRegLik(w["X"], w["y"])
    return orc.MvNormalLik(w["mu"], w["Sigma"], drop_constant=(w["kind"] == "quad"))


def _linreg_logdensity_chunked(w, B, rows_per_chunk=1 << 20):
    """-0.5 |y - X b|^2 for the rows of B in FP64, streaming X in row chunks (bounded memory)."""
    X, y = w["X"], w["y"]
    out = np.zeros(B.shape[0])
    for i in range(0, X.shape[0], rows_per_chunk):
        R = y[i:i + rows_per_chunk, None] - X[i:i + rows_per_chunk] @ B.T
        out += np.einsum("ij,ij->j", R, R)
    return -0.5 * out


def replay_traced_kblock(w, make_sampler, Z0, N, K, seed, sample, anneal_T=None, sync_mode="strict"):
    """Runs ONE K-block (K generations, all N chains, full-size archive) on the GPU with the per-proposal trace on and
    replays `sample` chains proposal by proposal through the numpy oracle (teacher-forced: the oracle is fed the state
    the GPU entered each proposal with).  Same rules as tests/parity.py."""
    orc = _oracle()
    lik = _oracle_lik(w)
    d, blocks = w["d"], w["blocks"]
    eps = 1e-5 * np.ones(d)
    s = make_sampler(Ngeneration=K, trace=True, store_chain=False, seed=seed)
    try:
        s.set_archive(Z0)
        X_all, lp_all = s.get_state()
        s.run(0)
        tr = s.get_trace()
    finally:
        s.close()
    M0 = Z0.shape[0]
    rng = np.random.default_rng(12345)
    ids = np.sort(rng.choice(N, size=min(sample, N), replace=False))
    X = X_all[ids].copy()
    lp = lp_all[ids].copy()
    if w["kind"] == "linreg":
        ref0 = _linreg_logdensity_chunked(w, X)
    else:
        ref0 = lik.logdensity(X)
    st = dict(chains=int(ids.size), chains_total=int(N), generations=int(K), proposals=0, accepted=0, index_or_uniform_mismatch=0,
              max_xprop_err=0.0, xprop_out_of_tol=0, max_lp_relerr=float(np.max(np.abs(lp - ref0) / np.maximum(np.abs(ref0), 1.0))),
              lp_out_of_tol=0, decisions_outside_band=0, decision_mismatch_outside_band=0, decisions_in_band=0, in_band_disagree=0)
    eps32 = eps.astype(np.float32).astype(np.float64)
    for g in range(1, K + 1):
        T = None if anneal_T is None else float(np.float32(anneal_T[g - 1]))
        for ib, block in enumerate(blocks):
            c = w["gamma"] if len(block) == 1 else w["gamma"] / np.sqrt(2.0 * len(block))
            c32 = float(np.float32(c))
            g_eff = c32 if len(block) == 1 else c32 * np.sqrt(2.0 * len(block))          # the device applies fl32(c)
            o = orc.block_update(X, lp, Z0, M0, lik if w["kind"] != "linreg" else _NoLik(), block, ib, eps32, g_eff, seed, g, T=T,
                                 chain_ids=ids)
            g_r1, g_r2, g_u = tr["r1"][g - 1, ib][ids], tr["r2"][g - 1, ib][ids], tr["u"][g - 1, ib][ids].astype(np.float64)
            st["index_or_uniform_mismatch"] += int(np.sum(g_r1 != o["r1"]) + np.sum(g_r2 != o["r2"]) + np.sum(g_u != o["u"]))
            g_xp = tr["xprop"][g - 1, ib][:, ids].T.astype(np.float64)
            err = np.abs(g_xp - o["xprop"])
            tol = 1e-6 + 2e-6 * np.abs(o["xprop"]) + 1e-4 * np.abs(eps)[None, :] \
                + 4e-7 * abs(w["gamma"]) * (np.abs(Z0[o["r1"]]) + np.abs(Z0[o["r2"]]) + np.abs(X))
            st["max_xprop_err"] = max(st["max_xprop_err"], float(err.max()))
            st["xprop_out_of_tol"] += int(np.sum(err > tol))
            ref_lpp = _linreg_logdensity_chunked(w, g_xp) if w["kind"] == "linreg" else lik.logdensity(g_xp)
            g_lpp = tr["lp_prop"][g - 1, ib][ids].astype(np.float64)
            rel = np.abs(g_lpp - ref_lpp) / np.maximum(np.abs(ref_lpp), 1.0)
            st["max_lp_relerr"] = max(st["max_lp_relerr"], float(rel.max()))
            st["lp_out_of_tol"] += int(np.sum(rel > LP_RTOL))
            g_acc = tr["accepted"][g - 1, ib][ids].astype(bool)
            logu = np.log(o["u"])
            margin = (ref_lpp - lp) - (logu if T is None else T * logu)
            band = LP_RTOL * np.maximum(np.maximum(np.abs(ref_lpp), np.abs(lp)), 1.0) * 2.0
            ref_acc = margin > 0
            outside = np.abs(margin) > band
            st["decisions_outside_band"] += int(outside.sum())
            st["decision_mismatch_outside_band"] += int(np.sum(g_acc[outside] != ref_acc[outside]))
            st["decisions_in_band"] += int((~outside).sum())
            st["in_band_disagree"] += int(np.sum(g_acc[~outside] != ref_acc[~outside]))
            st["proposals"] += int(ids.size)
            st["accepted"] += int(g_acc.sum())
            X[g_acc] = g_xp[g_acc]                       # carry the GPU's decision forward (teacher forcing)
            lp[g_acc] = g_lpp[g_acc] if w["kind"] != "linreg" else ref_lpp[g_acc]      # (the linreg trace rounds lp' to FP32)
    st["ok"] = bool(st["index_or_uniform_mismatch"] == 0 and st["xprop_out_of_tol"] == 0 and st["lp_out_of_tol"] == 0
                    and st["decision_mismatch_outside_band"] == 0)
    st["rule"] = ("indices, uniforms bit-exact; x' within FP32 rounding; log-density within 1e-5 relative of FP64; accept decisions "
                  "exact wherever |margin| > 2e-5 * max(|lp'|, |lp|, 1)")
    return st


class _NoLik:
    """block_update computes lp' through `lik`; for the regression the probe recomputes it chunked instead."""

    def logdensity(self, X):
        return np.zeros(np.atleast_2d(X).shape[0])


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None,
                    help="c2 (default: BASELINE configs[1], linreg on tcgen05), c4, c3, c1, c5; without this flag the line also "
                         "carries short runs of the other configs under the key 'also'")
    ap.add_argument("--also", default=None, help="comma list of extra workloads for the 'also' block (default: c4,c3,c1 and c5 on 8 GPUs)")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--gens-per-step", type=int, default=0)
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--archive-rows", type=int, default=0)
    ap.add_argument("--sync", default="overlap", choices=["strict", "overlap"],
                    help="strict: appended rows drawable at once (bit-identical to 1 GPU); overlap: one K-block later, exchange hidden under compute")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="p2p: own transport over peer-mapped archives; nccl: ncclAllGather")
    ap.add_argument("--exchange-layout", default="auto", choices=["auto", "in_place", "dense_slots"],
                    help="rows in padded power-of-two slots, or back to back (dense_slots); auto = dense_slots for more than 4 GPUs")
    ap.add_argument("--push-engine", default="auto", choices=["auto", "kernel", "copy_engine"],
                    help="who moves a rank's appended rows into the peers' replicas: our push kernel or the copy engines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--moments", action="store_true", help="opts.moments on (streaming mean / covariance of every generation inside the step kernel): what it costs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: --warmup {args.warmup} < 3 breaks the timing rules; use for profiling only", file=sys.stderr)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    primary = args.workload or "c2"

    def make_cfg(wname):
        w = workload(wname, args.chains_per_gpu, args.archive_rows)
        gps = args.gens_per_step or (10 if w["kind"] == "linreg" else 100)
        return w, gps, _cfg(w, gps)

    # what DEMCZ_XCHG_AUTO resolves to (include/demcz.h): rows without slot padding above 4 GPUs
    eff_layout = args.exchange_layout if args.exchange_layout != "auto" else ("dense_slots" if world >= 4 else "in_place")

    def _cfg(w, gps):
        K = w["K"]
        return dict(workload=w["name"], chains_per_gpu=w["chains_per_gpu"], chains_total=w["chains_per_gpu"] * max(world, 1),
               generations_per_step=gps, K=K, Nblocks=len(w["blocks"]), archive_rows_initial=w["archive_rows"],
               archive="replicated, ring buffer",
               arithmetic=("FP32-equivalent split tensor-core products (see roofline.kernel), FP32 accumulate per row tile, FP64 sums, FP64 log-objective"
                           if w["kind"] == "linreg" else "FP32 state, archive and log-density"),
               l2=("the split X operand (>= 128 MB) > L2 (126 MB) and re-streamed every generation; no flush needed" if w["kind"] == "linreg"
                   else "archive > L2 (126 MB): random gathers miss L2; no flush needed"),
               sync_mode=args.sync,
               parallelism=(f"chains sharded x{world}; appended rows exchanged every K generations by "
                            + (f"our own transport ({args.push_engine}) into peer-mapped archive replicas (NVLink; archive layout {eff_layout}), "
                               f"published by flag words that the step kernels poll" if args.exchange == "p2p" else "ncclAllGather")
                            + (", on a side stream under the next K-block" if args.sync == "overlap" else ", on the compute stream")))

    # ---------------------------------------------------------------- reference arm (CPU) -------
    if args.impl == "reference":
        if rank != 0:
            return 0
        w, gps, cfg = make_cfg(primary)
        threads = os.cpu_count() or 1
        prepare_data(w)
        chains, gens, arch = cpu_plan(w, threads, 2.0)                       # ~2 s per step, the sample cpu_baseline times
        for _ in range(min(args.warmup, 1)):
            cpu_run(w, chains=chains, gens=gens, archive_rows=arch, threads=threads)
        t = 0.0
        for _ in range(args.steps):
            _, dt = cpu_run(w, chains=chains, gens=gens, archive_rows=arch, threads=threads)
            t += dt
        val = chains * gens * args.steps / t
        sample = (f"each step = {chains} chains x {gens} generations of the same target, archive {max(arch, chains)} rows; "
                  f"CPU restatement of DEMC.jl (oracle/demcz_oracle.c, FP64, per-chain closure evaluation as src/demcz.jl:189, OpenMP over chains) - Julia is not installed")
        print(json.dumps(dict(impl="reference", metric="chain-updates/s", value=val, unit="chain-updates/s", n_gpus=args.gpus,
                              steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * t / args.steps, higher_is_better=True,
                              scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=cfg,
                              cpu_baseline=dict(value=val, unit="chain-updates/s", cores=threads, kind="port", sample=sample),
                              e2e=dict(value=val, unit="chain-updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return 0

    # ---------------------------------------------------------------- our arm (GPU) -------------
    import torch
    import demc_jl_b200 as DEMC
    from demc_jl_b200 import dist as ddist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        tdist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, tdist.ReduceOp.MAX) if world > 1 else float(x)

    def min_over_ranks(x):
        return reduce_ranks(x, tdist.ReduceOp.MIN) if world > 1 else float(x)

    def sum_over_ranks(x):
        return reduce_ranks(x, tdist.ReduceOp.SUM) if world > 1 else float(x)

    def make_logobj(w):
        if w["kind"] == "linreg":
            return DEMC.LinRegLoglik(w["X"], w["y"])
        return DEMC.QuadraticLogobj(w["mu"]) if w["kind"] == "quad" else DEMC.MvNormalLogpdf(w["mu"], w["Sigma"])

    def sampler_kwargs(w):
        return dict(K=w["K"], blockindex=w["blocks"], eps_scale=1e-5 * np.ones(w["d"]), γ=w["gamma"], anneal=w["anneal"], T0=2.0, TN=1e-4,
                    adapt=False, moments=args.moments)

    # ---- parity probe of one workload at its benchmark size (this rank's GPU; a world=1 context) ----------------
    def parity_probe(w, s_timed, Z0):
        N = w["chains_per_gpu"]
        out = {}
        t0 = time.perf_counter()
        if w["kind"] == "linreg":
            # (1) the log-objective the timed run ended with, sampled chains, against FP64 numpy
            X, lp = s_timed.get_state()
            ids = np.sort(np.random.default_rng(99).choice(N, size=min(256 if w["n"] <= 2_000_000 else 64, N), replace=False))
            ref = _linreg_logdensity_chunked(w, X[ids])
            rel = np.abs(lp[ids] - ref) / np.maximum(np.abs(ref), 1.0)
            out["final_state_logdensity"] = dict(chains=int(ids.size), max_relerr=float(rel.max()), max_abserr=float(np.abs(lp[ids] - ref).max()),
                                                 mean_abs_logp=float(np.abs(ref).mean()), ok=bool(rel.max() <= LP_RTOL),
                                                 rule="|lp_gpu - lp_fp64| <= 1e-5 * |lp_fp64| (FP64 numpy over all data rows)")
        sample = 1024 if w["kind"] != "linreg" else (256 if w["n"] <= 2_000_000 else 64)
        Tt = np.array([2.0 * (1e-4 / 2.0) ** (ig / w["K"]) for ig in range(1, w["K"] + 1)]) if w["anneal"] else None

        def mk(**kw):
            return DEMC.Sampler(make_logobj(w), w["d"], N=N, device=local_rank, autostop=False, sync_mode="strict",
                                exchange_layout=eff_layout, **sampler_kwargs(w), **kw)
        out["traced_kblock"] = replay_traced_kblock(w, mk, Z0, N, w["K"], 424242 + rank, sample, anneal_T=Tt)
        out["ok"] = bool(out["traced_kblock"]["ok"] and out.get("final_state_logdensity", {"ok": True})["ok"])
        out["seconds"] = time.perf_counter() - t0
        # every rank checked its own GPU: the line carries the worst case
        out["ok_all_ranks"] = bool(min_over_ranks(1.0 if out["ok"] else 0.0) == 1.0)
        out["max_lp_relerr_all_ranks"] = max_over_ranks(out["traced_kblock"]["max_lp_relerr"])
        return out

    # ---- N > 1: replicas identical, sharded == one GPU bit for bit --------------------------------------------
    def consistency(s_timed):
        out = {}
        h = s_timed.get_archive_checksum()
        t = torch.tensor([h & 0x7FFFFFFFFFFFFFFF], device="cuda", dtype=torch.int64)
        hs = [torch.zeros_like(t) for _ in range(world)]
        tdist.all_gather(hs, t)
        hs = [int(x.item()) for x in hs]
        out["archive_checksum_equal"] = bool(all(x == hs[0] for x in hs))
        res = {}
        for mode in ("strict", "overlap"):
            for kind in ("mvn", "linreg"):
                rng = np.random.default_rng(5)
                Ng, K, G = 1024 * world, 4, 24
                if kind == "mvn":
                    d = 20
                    S = 0.6 * np.ones((d, d)) + 0.4 * np.eye(d)
                    lo = lambda: DEMC.MvNormalLogpdf(np.linspace(-1, 1, d), S)
                    Z0 = rng.normal(size=(2 * Ng, d))
                else:
                    n, d = 4000, 12
                    Xd = np.ones((n, d)); Xd[:, 1:] = rng.normal(size=(n, d - 1))
                    yd = Xd @ (1 + 3 * rng.random(d)) + rng.normal(size=n)
                    lo = lambda: DEMC.LinRegLoglik(Xd, yd)
                    Z0 = np.linalg.solve(Xd.T @ Xd, Xd.T @ yd) + 0.05 * rng.normal(size=(2 * Ng, d))
                kw = dict(K=K, Ngeneration=G, blockindex=[list(range(d))], eps_scale=1e-5 * np.ones(d), γ=2.2, seed=17, store_chain=False,
                          sync_mode=mode, exchange_layout=eff_layout, archive_capacity=5 * Ng)      # the ring wraps
                s = ddist.make_sharded_sampler(lo(), d, N_global=Ng, rank=rank, world=world, device=local_rank,
                                               push_engine=args.push_engine, **kw)
                s.set_archive(Z0)
                if args.exchange == "p2p":
                    ddist.open_peer_archives(s, rank=rank, world=world, device=local_rank)
                s.run(G // 2)
                s.run()
                X, lp = s.get_state()
                hz = s.get_archive_checksum()
                s.close()
                Xg = ddist.allgather_rows(X, local_rank)
                lpg = ddist.allgather_rows(lp, local_rank)
                ok = True
                if rank == 0:
                    with DEMC.Sampler(lo(), d, N=Ng, device=local_rank, **kw) as s1:
                        s1.set_archive(Z0)
                        s1.run(G // 2)
                        s1.run()
                        X1, lp1 = s1.get_state()
                        h1 = s1.get_archive_checksum()
                    ok = bool(np.array_equal(X1, Xg) and np.array_equal(lp1, lpg) and h1 == hz)
                ok = bool(min_over_ranks(1.0 if ok else 0.0) == 1.0)
                t = torch.tensor([hz & 0x7FFFFFFFFFFFFFFF], device="cuda", dtype=torch.int64)
                hs = [torch.zeros_like(t) for _ in range(world)]
                tdist.all_gather(hs, t)
                res[f"{kind}_{mode}"] = bool(ok and all(int(x.item()) == int(hs[0].item()) for x in hs))
        out["cases"] = res
        out["multi_gpu_bitwise"] = bool(all(res.values()))
        out["what"] = (f"{1024 * world} chains sharded over {world} GPUs (ring-buffer archive wrapping, two run calls) vs one world=1 context on rank 0: "
                       "state, log-objective bit-identical, archive checksums equal on every rank")
        return out

    def measure(wname, steps, warmup, do_e2e, do_cpu, do_probe):
        w, gps, cfg = make_cfg(wname)
        d, K = w["d"], w["K"]
        N_local = w["chains_per_gpu"]
        N_global = N_local * world
        total_steps = warmup + steps + (warmup + steps if do_e2e else 0) + 1
        G_total = total_steps * gps
        M0 = w["archive_rows"]
        if world > 1:
            M0 = (M0 + N_global - 1) // N_global * N_global                 # append segments never straddle the ring end
        cap = max(M0 + N_global * 16, 64 << 20) if wname != "c1" else 0
        if cap:
            cap = (cap + N_global - 1) // N_global * N_global
        prepare_data(w)
        logobj = make_logobj(w)
        if w["kind"] == "linreg":
            cap = 0
        Z0 = make_archive(w, M0)
        s = ddist.make_sharded_sampler(logobj, d, N_global=N_global, rank=rank, world=world, device=local_rank,
                                       Ngeneration=G_total, autostop=w["autostop"], autostop_every=1000,
                                       autostop_Rhat=0.0, seed=31953150, store_chain=False, archive_capacity=cap,
                                       sync_mode=args.sync, exchange_layout=args.exchange_layout, push_engine=args.push_engine,
                                       **sampler_kwargs(w))
        s.set_archive(Z0)
        if world > 1 and args.exchange == "p2p":
            ddist.open_peer_archives(s, rank=rank, world=world, device=local_rank)

        for _ in range(warmup):
            s.run(gps)

        clocks = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            clocks.start()
        t0 = time.perf_counter()
        info = s.run(gps * steps)              # one call: CUDA events on the library's stream bracket exactly K steps
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        clk = clocks.stop() if rank == 0 else None
        dev_ms = max_over_ranks(info["device_ms"])
        wall_ms = max_over_ranks(wall_ms)
        launches = int(sum_over_ranks(info["step_launches"]))
        updates = N_global * gps * steps
        value = updates / (dev_ms * 1e-3)
        acc_frac = None

        # ---- e2e: public Sampler API, host (pinned) population in and out every step ------------------
        e2e = None
        if do_e2e:
            Xh = torch.empty((d, N_local), dtype=torch.float64).pin_memory().numpy().T      # N x d, Fortran order
            lph = torch.empty(N_local, dtype=torch.float64).pin_memory().numpy()
            Xo = torch.empty((d, N_local), dtype=torch.float64).pin_memory().numpy().T
            lpo = torch.empty(N_local, dtype=torch.float64).pin_memory().numpy()
            dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
            s._check(s.lib.demcz_get_state(s.ctx, dp(Xh), dp(lph)))

            bufs = [(Xh, lph), (Xo, lpo)]

            def e2e_step():
                (xi, li), (xo, lo) = bufs                                                     # this step's input / output host arrays
                s._check(s.lib.demcz_set_state(s.ctx, dp(xi), dp(li)))                       # H2D: population + log-objective
                s.run(gps)
                s._check(s.lib.demcz_get_state(s.ctx, dp(xo), dp(lo)))                       # D2H: new population + log-objective
                bufs.reverse()                                                                # the result is the next step's input

            for _ in range(warmup):
                e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step()
            barrier()
            e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
            nbytes = N_global * (d + 1) * 8
            e2e = dict(value=updates / (e2e_ms * 1e-3), unit="chain-updates/s", h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes,
                       ms_per_step=e2e_ms / steps,
                       api="Sampler.set_state(host X, logp) -> run(generations_per_step) -> get_state() each step; archive resident")

        # ---- roofline of the step kernel: per-launch CUDA-event durations, live ---------------------------
        s.set_profile(True)
        pinfo = s.run(gps)
        s.set_profile(False)
        roof = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wname)
        except Exception:
            pass
        if pinfo["step_launches"] > 0 and w["kind"] == "linreg":
            n_l = pinfo["step_launches"]
            avg_ms = pinfo["step_kernel_ms"] / n_l
            flops = 2.0 * w["n"] * d * N_local                       # FP32-equivalent flops of X (n x p) . B^T (p x N)
            achieved = flops / (avg_ms * 1e-3) / 1e12
            kinfo = s.kernel_info() if hasattr(s, "kernel_info") else {}
            terms = kinfo.get("linreg_terms", 3)
            kind = kinfo.get("linreg_kind", "tf32")
            tf32_live = measure_tf32_peak()
            bf16 = float(peaks["bf16_tflops"]) if "bf16_tflops" in peaks else 1590.0
            bf16_src = "MEASURED_PEAKS.json bf16_tflops (measured)" if "bf16_tflops" in peaks else "fallback 1590"
            # dense rate of the MMA kind the kernel issues, from the driver-written bf16 figure: kind::f16 = bf16 rate, kind::tf32 = half
            dense = bf16 if kind == "f16" else bf16 / 2.0
            peak = dense / terms
            roof = dict(bound="tensor", kernel=kinfo.get("linreg_kernel", "linreg_ssr_kernel (tcgen05 kind::tf32, 3xTF32)"),
                        achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic,
                        peak_source=f"{bf16_src} = {bf16:.0f} TFLOP/s dense bf16; kind::{kind} dense rate = {dense:.0f}; / {terms} MMA terms per FP32-equivalent product",
                        frac_vs_sustained=(achieved / ((float(peaks["bf16_tflops_sustained"]) if kind == "f16" else float(peaks["bf16_tflops_sustained"]) / 2.0) / terms))
                        if "bf16_tflops_sustained" in peaks else None,
                        frac_vs_live_tf32=(achieved / (tf32_live / 3.0)) if kind == "tf32" else None,
                        tf32_dense_live_tflops=tf32_live,
                        tensor_issued_tflops=terms * achieved, launch_us=avg_ms * 1e3,
                        flops_per_launch=flops, kernel_share_of_step=pinfo["step_kernel_ms"] / pinfo["device_ms"])
            acc_frac = pinfo["accepted"] / max(pinfo["proposals"], 1)
        elif pinfo["step_launches"] > 0:
            n_l = pinfo["step_launches"]
            avg_ms = pinfo["step_kernel_ms"] / n_l
            G_f = gps / n_l
            bpu = algorithmic_bytes_per_update(d, K, G_f)
            bytes_per_launch = N_local * G_f * bpu
            achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
            peak, which = 6650.0, "fallback"
            if "hbm_gbs" in peaks:
                peak, which = float(peaks["hbm_gbs"]), "measured"
            # ceiling: every gathered row costs one whole 128-byte DRAM line (measured granularity, profiles/), so the algorithmic
            # fraction of a perfect gather is (algorithmic bytes) / (line bytes actually moved)
            slot = 32 if d <= 8 else 64 if d <= 16 else 128
            line_bytes = 2 * 128 * len(w["blocks"]) + (8 * d + 8) / G_f + slot / K
            ceiling = bpu / line_bytes
            roof = dict(bound="hbm", kernel="demcz_step_mvn_kernel", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                        peak_source=f"MEASURED_PEAKS.json hbm_gbs ({which})", traffic=traffic, launch_us=avg_ms * 1e3,
                        generations_per_launch=G_f, algorithmic_bytes_per_update=bpu, bytes_per_launch=bytes_per_launch,
                        ceiling_frac=ceiling, frac_of_ceiling=(achieved / peak) / ceiling,
                        ceiling_note="a random archive row costs one 128-byte DRAM line whatever its payload (ncu sector counters, profiles/): "
                                     "ceiling_frac = algorithmic bytes / line-granular bytes of a perfect gather",
                        kernel_share_of_step=pinfo["step_kernel_ms"] / pinfo["device_ms"])
            acc_frac = pinfo["accepted"] / max(pinfo["proposals"], 1)
        if roof is not None and world > 1:
            roof["exchange"] = {k: v for k, v in s.kernel_info().items() if k in ("exchange", "memop_wait", "memop_write")}
        if roof is not None and pinfo.get("exchanges"):
            roof["exchange_us"] = 1e3 * pinfo["exchange_ms"] / pinfo["exchanges"]      # per K-block, this rank (device time of the push / all-gather)

        probe = None
        if do_probe:
            try:
                probe = parity_probe(w, s, Z0)
            except Exception as e:  # noqa: BLE001 — a failing probe must show up in the line, not kill the measurement
                probe = dict(ok=False, ok_all_ranks=False, error=f"{type(e).__name__}: {e}")
        cons = None
        if world > 1 and do_probe and wname == primary:
            try:
                cons = consistency(s)
            except Exception as e:  # noqa: BLE001
                cons = dict(multi_gpu_bitwise=False, error=f"{type(e).__name__}: {e}")
        elif world > 1 and do_probe:
            h = s.get_archive_checksum()
            t = torch.tensor([h & 0x7FFFFFFFFFFFFFFF], device="cuda", dtype=torch.int64)
            hs = [torch.zeros_like(t) for _ in range(world)]
            tdist.all_gather(hs, t)
            cons = dict(archive_checksum_equal=bool(all(int(x.item()) == int(hs[0].item()) for x in hs)))
        s.close()
        del Z0

        cpu = None
        if rank == 0 and world == 1 and do_cpu:
            cpu = cpu_baseline(w)
        return dict(metric="chain-updates/s", value=value, unit="chain-updates/s", n_gpus=world, steps=steps, warmup=warmup,
                    ms_per_step=dev_ms / steps, wall_ms_per_step=wall_ms / steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None,
                    dtype="f32",
                    data="synthetic", config=cfg, clocks=clk, e2e=e2e, gpu_launches=launches, roofline=roof, cpu_baseline=cpu,
                    parity_probe=probe, consistency=cons,
                    accepted_fraction=acc_frac, proposals_per_s=value * len(w["blocks"]))

    res = measure(primary, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline, not args.no_probe)
    if args.workload is None and not args.no_also:
        # the other BASELINE configs, short runs, same rules (C4 = the 1M-chain MvNormal configuration at 131072 chains per GPU)
        names = args.also.split(",") if args.also else (["c4", "c3", "c1"] + (["c5"] if world == 8 else []))
        res["also"] = {}
        for nm in names:
            saved = args.gens_per_step
            if nm == "c5" and not saved:
                args.gens_per_step = 2                     # 189 ms per generation and GPU: keep the line short
            try:
                a = measure(nm, max(3, args.steps // 2) if nm != "c5" else 3, 3, nm in ("c4", "c3"), False, not args.no_probe)
                res["also"][nm] = {k: a[k] for k in ("value", "unit", "ms_per_step", "config", "roofline", "e2e", "gpu_launches",
                                                      "accepted_fraction", "parity_probe", "consistency")}
            except Exception as e:  # noqa: BLE001
                res["also"][nm] = dict(error=f"{type(e).__name__}: {e}")
            args.gens_per_step = saved
    if world > 1:
        barrier()
        tdist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
