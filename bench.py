#!/usr/bin/env python
"""bench.py — chain-updates/s of the DEMCz generation loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3|c1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = `--gens-per-step` generations (default 100 = 10 K-blocks, 10 archive appends) of every
chain.  Workload c4 (default) is BASELINE configs[3] — 20-dim equicorrelated MvNormal, K=10, R-hat
accumulators on — at 131072 chains per GPU, so that 8 GPUs run the 1M-chain configuration exactly
(weak scaling); the archive starts at 8 Mi rows (671 MB, > L2) on every GPU count.

Rank 0 prints ONE JSON line.  `value` is timed with CUDA events on the library's stream with the
inputs resident in HBM; `e2e` goes through the public Sampler API with host buffers (pinned) in and
out every step; `roofline` is the step kernel against the measured HBM peak; `cpu_baseline` is the
CPU restatement of the reference (oracle/, all host cores) on a bounded sample of the workload.
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


# ------------------------------------------------------------------------------------------------
# workloads (synthetic data of each config's shape, fixed seeds)
# ------------------------------------------------------------------------------------------------
def workload(name, chains_per_gpu, archive_rows):
    if name == "c4":      # 20-dim highly correlated MvNormal, K=10, Rhat autostop machinery on
        d = 20
        S = 0.95 * np.ones((d, d)) + 0.05 * np.eye(d)
        mu = np.linspace(-1.0, 1.0, d)
        return dict(name="C4 d=20 equicorrelated(0.95) MvNormal, K=10, Rhat accumulators on", d=d, mu=mu, Sigma=S,
                    kind="mvn", K=10, blocks=[list(range(d))], anneal=False, autostop=True, gamma=2.38,
                    chains_per_gpu=chains_per_gpu or 131072, archive_rows=archive_rows or (8 << 20))
    if name == "c3":      # demcz_anneal on 10-dim quadratic objective, Nblocks=2, 65536 chains
        d = 10
        mu = np.random.default_rng(31953150).random(d)
        return dict(name="C3 demcz_anneal d=10 quadratic, Nblocks=2, T0=2->TN=1e-4", d=d, mu=mu, Sigma=0.5 * np.eye(d),
                    kind="quad", K=10, blocks=[list(range(5)), list(range(5, 10))], anneal=True, autostop=False,
                    gamma=2.38, chains_per_gpu=chains_per_gpu or 65536, archive_rows=archive_rows or (8 << 20))
    if name == "c1":      # test/example_normpdf.jl: 5-dim MvNormal, N=100, K=10
        d = 5
        rng = np.random.default_rng(31953150)
        mu = rng.random(d)
        A = rng.random((d, d))
        S = A.T @ A + 2 * np.eye(d)
        S = S / S.max() / 100
        return dict(name="C1 example_normpdf d=5 MvNormal N=100", d=d, mu=mu, Sigma=S, kind="mvn", K=10,
                    blocks=[list(range(d))], anneal=False, autostop=False, gamma=2.38,
                    chains_per_gpu=chains_per_gpu or 100, archive_rows=archive_rows or 150)
    if name in ("c2", "c5"):   # test/example_linreg.jl scaled up: the tensor-core GEMM likelihood
        n, p, N = (1_000_000, 32, 4096) if name == "c2" else (10_000_000, 64, 32768)
        return dict(name=f"{name.upper()} linreg n={n} x p={p} (example_linreg.jl generator), K=10, gamma=2.0", d=p, kind="linreg",
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
    """Z0 ~ N(mu, 4*Sigma): over-dispersed start, identical on every rank (replicated archive)."""
    rng = np.random.default_rng(seed)
    if w["kind"] == "linreg":      # around the OLS point, 3x the posterior scale
        return w["bols"] + 3.0 * w["post_sd"] * rng.standard_normal((rows, w["d"]))
    L = np.linalg.cholesky(w["Sigma"])
    Z = np.empty((rows, w["d"]), order="F")
    step = 1 << 20
    for i in range(0, rows, step):
        n = min(step, rows - i)
        Z[i:i + n] = w["mu"] + 2.0 * (rng.standard_normal((n, w["d"]), dtype=np.float32).astype(np.float64) @ L.T)
    return Z


def algorithmic_bytes_per_update(d, K, G_f):
    """B(G_f) = 8d [two archive rows] + (8d+8)/G_f [X, logp read+write] + 4d/K [append]."""
    return 8 * d + (8 * d + 8) / G_f + 4 * d / K


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
# CPU arm: the restatement of the reference loop (oracle/), all host threads
# ------------------------------------------------------------------------------------------------
def cpu_run(w, *, chains, gens, archive_rows, threads, faithful=False, seed=31953150):
    from oracle import demcz_oracle as orc      # the checker, timed as the CPU baseline (never on the product path)
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


def cpu_sample_size(w, threads):
    if w["kind"] == "linreg":      # one proposal streams the whole data set: a few chains per thread is already seconds
        return threads * 2, 2
    chains = min(w["chains_per_gpu"], 16384)
    return max(chains // threads * threads, min(w["chains_per_gpu"], threads)), 20


def cpu_baseline(w, budget_s=12.0):
    threads = os.cpu_count() or 1
    arch = min(w["archive_rows"], 1 << 20)
    chains, g0 = cpu_sample_size(w, threads)
    v, dt = cpu_run(w, chains=chains, gens=g0, archive_rows=arch, threads=threads)            # calibrate
    gens = int(max(g0, min(5000, budget_s * v / chains)))
    v, dt = cpu_run(w, chains=chains, gens=gens, archive_rows=arch, threads=threads)
    out = dict(value=v, unit="chain-updates/s", cores=threads, kind="port",
               sample=f"{chains} chains x {gens} generations, archive {max(arch, chains)} rows, FP64, O(1) index draw, "
                      f"row-major archive, OpenMP over chains ({dt:.1f} s)")
    # reference-faithful serial mode: O(M) index-set construction per proposal (src/demcz.jl:176-178), 1 thread
    fch = min(chains, 64) if w["kind"] != "linreg" else 4
    fg = 10 if w["kind"] != "linreg" else 2
    fv, fdt = cpu_run(w, chains=fch, gens=fg, archive_rows=min(arch, 1 << 17), threads=1, faithful=True)
    out["faithful_serial"] = dict(value=fv, cores=1, sample=f"{fch} chains x {fg} generations, archive {max(min(arch, 1 << 17), fch)} rows "
                                                            f"({fdt:.1f} s), collect(1:M)+deleteat! per proposal")
    return out


def prepare_data(w):
    """Synthetic likelihood data of the workload's shape (identical on every rank)."""
    if w["kind"] == "linreg" and "X" not in w:
        w["X"], w["y"] = linreg_data(w["n"], w["d"])
        XtX = w["X"].T @ w["X"]
        w["bols"] = np.linalg.solve(XtX, w["X"].T @ w["y"])
        w["post_sd"] = np.sqrt(np.diag(np.linalg.inv(XtX)))


def measure_tf32_peak():
    """Dense TF32 tensor throughput of this GPU (cuBLAS via torch.matmul, 8192^3, best of 10): the denominator of
    the linreg roofline.  Library call used for the PEAK only."""
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
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None,
                    help="c2 (default: BASELINE configs[1], linreg on tcgen05), c4, c3, c1, c5; without this flag the line also "
                         "carries a short run of c4 (the 1M-chain MvNormal scaling config) under the key 'also'")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--gens-per-step", type=int, default=0)
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--archive-rows", type=int, default=0)
    ap.add_argument("--sync", default="overlap", choices=["strict", "overlap"],
                    help="strict: appended rows drawable at once (bit-identical to 1 GPU); overlap: one K-block later, exchange hidden under compute")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="p2p: own push kernel over peer-mapped archives; nccl: ncclAllGather")
    ap.add_argument("--exchange-layout", default="auto", choices=["auto", "in_place", "packed", "dense_slots"],
                    help="p2p push: rows stored straight into the peers' padded slots, or packed through a staging ring and scattered "
                         "by the receiver (fewer NVLink lines), or the archive keeps rows without slot padding (dense_slots); auto = dense_slots for more than 4 GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    eff_layout = args.exchange_layout if args.exchange_layout != "auto" else ("dense_slots (auto)" if world > 4 else "in_place (auto)")

    def _cfg(w, gps):
        K = w["K"]
        return dict(workload=w["name"], chains_per_gpu=w["chains_per_gpu"], chains_total=w["chains_per_gpu"] * max(world, 1),
               generations_per_step=gps, K=K, Nblocks=len(w["blocks"]), archive_rows_initial=w["archive_rows"],
               archive="replicated, ring buffer",
               arithmetic=("FP32-equivalent: 3xTF32 tensor-core products, FP32 accumulate per 192-row tile, FP64 sums, FP64 log-objective"
                           if w["kind"] == "linreg" else "FP32 state, archive and log-density"),
               l2=("X_hi + X_lo (2 x 4 n p bytes) > L2 (126 MB) and re-streamed every generation; no flush needed" if w["kind"] == "linreg"
                   else "archive > L2 (126 MB): random gathers miss L2; no flush needed"),
               sync_mode=args.sync,
               parallelism=(f"chains sharded x{world}; appended rows exchanged every K generations by "
                            + (f"our push kernel into peer-mapped archive replicas (NVLink; archive layout {eff_layout}), published by flag words" if args.exchange == "p2p" else "ncclAllGather")
                            + (", on a side stream under the next K-block" if args.sync == "overlap" else ", on the compute stream")))

    # ---------------------------------------------------------------- reference arm (CPU) -------
    if args.impl == "reference":
        if rank != 0:
            return 0
        w, gps, cfg = make_cfg(primary)
        threads = os.cpu_count() or 1
        prepare_data(w)
        chains, g0 = cpu_sample_size(w, threads)
        arch = min(w["archive_rows"], 1 << 20)
        v, _ = cpu_run(w, chains=chains, gens=g0, archive_rows=arch, threads=threads)
        gens = int(max(g0, min(2000, 4.0 * v / chains)))                     # ~4 s per step
        for _ in range(min(args.warmup, 1)):
            cpu_run(w, chains=chains, gens=gens, archive_rows=arch, threads=threads)
        t = 0.0
        for _ in range(args.steps):
            _, dt = cpu_run(w, chains=chains, gens=gens, archive_rows=arch, threads=threads)
            t += dt
        val = chains * gens * args.steps / t
        sample = (f"each step = {chains} chains x {gens} generations of the same target, archive {max(arch, chains)} rows; "
                  f"CPU restatement of DEMC.jl (oracle/demcz_oracle.c, FP64, OpenMP over chains) - Julia is not installed")
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

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        return float(t.item())

    def measure(wname, steps, warmup, do_e2e, do_cpu):
        w, gps, cfg = make_cfg(wname)
        d, K = w["d"], w["K"]
        N_local = w["chains_per_gpu"]
        N_global = N_local * world
        total_steps = warmup + steps + (warmup + steps if do_e2e else 0) + 1
        G_total = total_steps * gps
        M0 = w["archive_rows"]
        cap = max(M0 + N_global * 16, 64 << 20) if wname != "c1" else 0
        if cap:
            cap = (cap + N_global - 1) // N_global * N_global            # append segments never straddle the ring end
        prepare_data(w)
        if w["kind"] == "linreg":
            logobj = DEMC.LinRegLoglik(w["X"], w["y"])
            cap = 0
        else:
            logobj = (DEMC.QuadraticLogobj(w["mu"]) if w["kind"] == "quad" else DEMC.MvNormalLogpdf(w["mu"], w["Sigma"]))
        Z0 = make_archive(w, M0)
        s = ddist.make_sharded_sampler(logobj, d, N_global=N_global, rank=rank, world=world, device=local_rank, K=K,
                                       Ngeneration=G_total, blockindex=w["blocks"], eps_scale=1e-5 * np.ones(d), γ=w["gamma"],
                                       anneal=w["anneal"], T0=2.0, TN=1e-4, autostop=w["autostop"], autostop_every=1000,
                                       autostop_Rhat=0.0, adapt=False, seed=31953150, store_chain=False, archive_capacity=cap,
                                       sync_mode=args.sync, exchange_layout=args.exchange_layout)
        s.set_archive(Z0)
        del Z0
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
        if pinfo["step_launches"] > 0 and w["kind"] == "linreg":
            n_l = pinfo["step_launches"]
            avg_ms = pinfo["step_kernel_ms"] / n_l
            flops = 2.0 * w["n"] * d * N_local                       # FP32-equivalent flops of X (n x p) . B^T (p x N)
            achieved = flops / (avg_ms * 1e-3) / 1e12
            tf32 = measure_tf32_peak()
            bf16 = None
            try:
                bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
            except Exception:
                pass
            roof = dict(bound="tensor", kernel="linreg_ssr_kernel (tcgen05 kind::tf32, 3xTF32)", achieved=achieved, peak=tf32 / 3.0,
                        unit="TFLOP/s", frac=achieved / (tf32 / 3.0), traffic=None,
                        peak_source=f"TF32 dense measured live with torch.matmul 8192^3 = {tf32:.0f} TFLOP/s, / 3 for the 3xTF32 split "
                                    f"(MEASURED_PEAKS.json bf16_tflops = {bf16})",
                        tensor_issued_tflops=3 * achieved, tensor_frac_of_tf32_peak=3 * achieved / tf32, launch_us=avg_ms * 1e3,
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
            try:
                peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); which = "measured"
            except Exception:
                pass
            traffic = None
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wname)
            except Exception:
                pass
            roof = dict(bound="hbm", kernel="demcz_step_mvn_kernel", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                        peak_source=f"MEASURED_PEAKS.json hbm_gbs ({which})", traffic=traffic, launch_us=avg_ms * 1e3,
                        generations_per_launch=G_f, algorithmic_bytes_per_update=bpu, bytes_per_launch=bytes_per_launch,
                        kernel_share_of_step=pinfo["step_kernel_ms"] / pinfo["device_ms"])
            acc_frac = pinfo["accepted"] / max(pinfo["proposals"], 1)
        if roof is not None and pinfo.get("exchanges"):
            roof["exchange_us"] = 1e3 * pinfo["exchange_ms"] / pinfo["exchanges"]      # per K-block, this rank (device time of the push / all-gather)
        s.close()

        cpu = None
        if rank == 0 and world == 1 and do_cpu:
            cpu = cpu_baseline(w)
        return dict(metric="chain-updates/s", value=value, unit="chain-updates/s", n_gpus=world, steps=steps, warmup=warmup,
                    ms_per_step=dev_ms / steps, wall_ms_per_step=wall_ms / steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None,
                    dtype="f32",
                    data="synthetic", config=cfg, clocks=clk, e2e=e2e, gpu_launches=launches, roofline=roof, cpu_baseline=cpu,
                    accepted_fraction=acc_frac, proposals_per_s=value * len(w["blocks"]))

    res = measure(primary, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline)
    if args.workload is None and not args.no_also:
        # the 1M-chain MvNormal configuration (BASELINE configs[3]; 131072 chains per GPU), short run, same rules
        a = measure("c4", max(3, args.steps // 2), 3, False, False)
        res["also"] = {"c4": {k: a[k] for k in ("value", "unit", "ms_per_step", "config", "roofline", "gpu_launches", "accepted_fraction")}}
    if world > 1:
        barrier()
        tdist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
