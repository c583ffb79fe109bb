"""Host-side mirror of the DEMC.jl interface for the generation loop (same names, argument meaning
and defaults), driving libdemcz_cuda.so through its C ABI.

Reference interface mirrored (file:line under the upstream repository):
  DEMCopt / demcopt                     src/DEMC.jl:24-43
  MC                                    src/DEMC.jl:10-15
  demcz_sample (opts + positional)      src/demcz.jl:1-3, 9-63
  demcz_anneal (opts + positional)      src/demcz_anneal.jl:14-16, 18-65
  tempbaseline                          src/demcz_anneal.jl:1-3

Differences, all forced by the GPU boundary and documented in DESIGN.md / INTEGRATION.md:
  * `logobj` is a registered likelihood descriptor (MvNormalLogpdf, QuadraticLogobj, LinRegLoglik),
    not an arbitrary closure; a plain callable raises TypeError — there is NO CPU fallback.
  * `blockindex` holds 0-based column indices here (Python); the Julia shim converts from 1-based.
  * generation-synchronous archive appends and the intended init / temperature
    schedule (Q1, Q3); `init_mode="ref_serial_zeros"` reproduces Q1.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Sequence

import numpy as np

from . import _lib

__all__ = ["DEMCopt", "demcopt", "MC", "MvNormalLogpdf", "QuadraticLogobj", "LinRegLoglik", "LogisticLoglik", "demcz_sample",
           "demcz_anneal", "demcz_sample_par", "demcz_anneal_par", "tempbaseline", "Sampler"]


# ------------------------------------------------------------------------------------------------
# options / result records
# ------------------------------------------------------------------------------------------------
@dataclass
class DEMCopt:
    """mutable struct DEMCopt — src/DEMC.jl:24-39 (field names verbatim; `gamma` aliases `γ`)."""
    N: int
    K: int
    Ngeneration: int
    Nblocks: int
    blockindex: list
    eps_scale: np.ndarray
    γ: float
    verbose: bool
    print_step: int
    T0: float
    TN: float
    autostop: str
    autostop_every: int
    autostop_Rhat: float

    @property
    def gamma(self):
        return self.γ

    @gamma.setter
    def gamma(self, v):
        self.γ = v


def demcopt(Npar, *, N=4, K=10, Ngeneration=5000, Nblocks=1, blockindex=None, eps_scale=None, γ=2.38, gamma=None,
            verbose=True, print_step=100, T0=3, TN=1e-3, autostop="Rhat", autostop_every=1000, autostop_Rhat=1.05):
    """demcopt(Npar; ...) — src/DEMC.jl:41-43, same defaults."""
    return DEMCopt(N, K, Ngeneration, Nblocks, [list(range(Npar))] if blockindex is None else blockindex,
                   1e-4 * np.ones(Npar) if eps_scale is None else np.asarray(eps_scale, dtype=np.float64),
                   γ if gamma is None else gamma, verbose, print_step, T0, TN, autostop, autostop_every, autostop_Rhat)


@dataclass
class MC:
    """struct MC — src/DEMC.jl:10-15.  Arrays are Fortran-ordered Float64 like Julia's."""
    chain: np.ndarray            # N x d x G
    log_obj: np.ndarray          # N x G
    Xcurrent: np.ndarray         # N x d
    log_objcurrent: np.ndarray   # N
    # bookkeeping for prevrun continuation (not in the reference struct)
    generations_total: int = 0
    info: dict = field(default_factory=dict)


# ------------------------------------------------------------------------------------------------
# likelihood descriptors (replace the `logobj` closure, src/demcz.jl:189)
# ------------------------------------------------------------------------------------------------
class MvNormalLogpdf:
    """logpdf(MvNormal(μ, Σ), x) — the objective of test/example_normpdf.jl:13-16."""

    def __init__(self, μ, Σ, drop_constant=False):
        self.mu = np.ascontiguousarray(μ, dtype=np.float64)
        self.Sigma = np.asfortranarray(Σ, dtype=np.float64)
        self.drop_constant = bool(drop_constant)
        self.d = self.mu.size
        if self.Sigma.shape != (self.d, self.d):
            raise ValueError("Σ must be d x d")

    def _register(self, lib, ctx):
        return lib.demcz_set_likelihood_mvnormal(ctx, self.d, _dp(self.mu), _dp(self.Sigma), int(self.drop_constant))


class QuadraticLogobj(MvNormalLogpdf):
    """-sum((x .- μ).^2) — test/test_anneal.jl:10: MvNormal(μ, I/2) without the constant."""

    def __init__(self, μ):
        μ = np.asarray(μ, dtype=np.float64)
        super().__init__(μ, 0.5 * np.eye(μ.size), drop_constant=True)


class LinRegLoglik:
    """-0.5 * sum((y .- X*b).^2) — test/example_linreg.jl:32."""

    def __init__(self, X, y):
        self.X = np.asfortranarray(X, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.n, self.d = self.X.shape
        if self.y.shape != (self.n,):
            raise ValueError("y must have n entries")

    def _register(self, lib, ctx):
        return lib.demcz_set_likelihood_linreg(ctx, self.n, self.d, _dp(self.X), _dp(self.y))


class LogisticLoglik(LinRegLoglik):
    """sum(y .* eta .- log1p.(exp.(eta))), eta = X*b — Bernoulli-logit GLM on the same tensor-core contraction as
    LinRegLoglik with a softplus epilogue (an extension: the reference takes any closure, src/demcz.jl:189)."""

    def _register(self, lib, ctx):
        return lib.demcz_set_likelihood_logistic(ctx, self.n, self.d, _dp(self.X), _dp(self.y))


def tempbaseline(ig, Ng, T0, TN):
    """src/demcz_anneal.jl:1-3"""
    return T0 * (TN / T0) ** (ig / Ng)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _check_likelihood(logobj):
    if not hasattr(logobj, "_register"):
        raise TypeError(
            "logobj must be a registered GPU likelihood descriptor (MvNormalLogpdf, QuadraticLogobj, LinRegLoglik, LogisticLoglik); "
            "arbitrary closures are not supported and there is no CPU fallback")


# ------------------------------------------------------------------------------------------------
# context wrapper
# ------------------------------------------------------------------------------------------------
class Sampler:
    """One libdemcz_cuda context (one GPU).  Thin: every method is one ABI call."""

    def __init__(self, logobj, d, *, N, K, Ngeneration, blockindex, eps_scale, γ, anneal=False, T0=3.0, TN=1e-3,
                 autostop=False, autostop_every=1000, autostop_Rhat=1.05, adapt=False, min_gamma=0.1, max_gamma=4.0,
                 adapt_every=500, seed=31953150, gen_offset=0, store_chain=True, init_mode="last_N_of_Z",
                 archive_capacity=0, trace=False, profile=False, device=0, rank=0, world=1, N_global=None,
                 chain_offset=0, sync_mode="strict", exchange_layout="auto", push_engine="auto", thin=1, moments=False,
                 persistent="auto", snooker_prob=0.0, snooker_gamma=0.0):
        _check_likelihood(logobj)
        self.lib = _lib.load()
        self.d, self.N = int(d), int(N)
        blocks = [np.asarray(list(b), dtype=np.int32) for b in blockindex]
        self._bp = np.zeros(len(blocks) + 1, dtype=np.int32)
        self._bp[1:] = np.cumsum([b.size for b in blocks])
        self._bi = np.concatenate(blocks).astype(np.int32)
        self._eps = np.ascontiguousarray(eps_scale, dtype=np.float64)
        if self._eps.size != self.d:
            raise ValueError("eps_scale must have d entries")
        o = _lib.Opts()
        o.struct_size = C.sizeof(_lib.Opts)
        o.N, o.K, o.Ngeneration, o.Nblocks, o.d = self.N, int(K), int(Ngeneration), len(blocks), self.d
        o.block_ptr = self._bp.ctypes.data_as(C.POINTER(C.c_int32))
        o.block_idx = self._bi.ctypes.data_as(C.POINTER(C.c_int32))
        o.eps_scale = _dp(self._eps)
        o.gamma, o.T0, o.TN = float(γ), float(T0), float(TN)
        o.anneal, o.autostop, o.autostop_every, o.autostop_Rhat = int(anneal), int(autostop), int(autostop_every), float(autostop_Rhat)
        o.adapt, o.min_gamma, o.max_gamma, o.adapt_every = int(adapt), float(min_gamma), float(max_gamma), int(adapt_every)
        o.seed, o.gen_offset = int(seed) & (2**64 - 1), int(gen_offset)
        o.store_chain = 2 if store_chain == "stream" else int(bool(store_chain))
        o.init_mode = {"last_N_of_Z": 0, "ref_serial_zeros": 1}[init_mode]
        o.archive_capacity = int(archive_capacity)
        o.trace, o.profile = int(trace), int(profile)
        o.rank, o.world = int(rank), int(world)
        o.N_global = self.N if N_global is None else int(N_global)
        o.chain_offset = int(chain_offset)
        o.sync_mode = {"strict": 0, "overlap": 1}[sync_mode]
        if exchange_layout not in ("auto", "in_place", "dense_slots"):
            raise ValueError(f"exchange_layout must be auto, in_place or dense_slots (got {exchange_layout!r})")
        o.exchange_layout = {"auto": 0, "in_place": 1, "dense_slots": 3}[exchange_layout]
        o.push_engine = {"auto": 0, "kernel": 1, "copy_engine": 2}[push_engine]
        o.thin, o.moments = int(thin), int(bool(moments))
        o.persistent = {"auto": -1, True: 1, False: 0, 1: 1, 0: 0, -1: -1}[persistent]
        o.snooker_prob, o.snooker_gamma = float(snooker_prob), float(snooker_gamma)
        self.thin = max(int(thin), 1)
        self.opts = o
        self.Nblocks = len(blocks)
        self.ctx = C.c_void_p()
        rc = self.lib.demcz_create(C.byref(self.ctx), C.byref(o), int(device))
        if rc != _lib.OK:
            raise _lib.DemczError(rc, self.lib.demcz_last_error(None).decode())
        self._check(logobj._register(self.lib, self.ctx))
        self.generations_total = 0

    def _check(self, rc):
        if rc != _lib.OK:
            raise _lib.DemczError(rc, self.lib.demcz_last_error(self.ctx).decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.demcz_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- inputs ------------------------------------------------------------------------------
    def set_archive(self, Z):
        Z = np.asfortranarray(Z, dtype=np.float64)
        if Z.ndim != 2 or Z.shape[1] != self.d:
            raise ValueError("Zmat must be M x d")
        self._check(self.lib.demcz_set_archive(self.ctx, _dp(Z), Z.shape[0]))

    def set_state(self, X, logp=None):
        X = np.asfortranarray(X, dtype=np.float64)
        if X.shape != (self.N, self.d):
            raise ValueError("X must be N x d")
        lp = None if logp is None else np.ascontiguousarray(logp, dtype=np.float64)
        self._check(self.lib.demcz_set_state(self.ctx, _dp(X), _dp(lp)))

    def set_temperature_table(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64)
        self._check(self.lib.demcz_set_temperature_table(self.ctx, _dp(T), T.size))

    def comm_init(self, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.demcz_comm_init(self.ctx, buf))

    def peer_info(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._check(self.lib.demcz_comm_peer_info(self.ctx, buf))
        return bytes(buf)

    def open_peers(self, infos: bytes):
        buf = (C.c_uint8 * len(infos)).from_buffer_copy(infos)
        self._check(self.lib.demcz_comm_open_peers(self.ctx, buf))

    # -- run ---------------------------------------------------------------------------------
    def run(self, max_generations=0):
        info = _lib.RunInfo()
        self._check(self.lib.demcz_run(self.ctx, int(max_generations), C.byref(info)))
        self.generations_total = info.generations_total
        return {k: getattr(info, k) for k, _ in _lib.RunInfo._fields_}

    def kernel_info(self):
        """Which kernels this context runs (demcz_get_kernel_info) as a dict."""
        buf = C.create_string_buffer(1024)
        self.lib.demcz_get_kernel_info(self.ctx, buf, 1024)
        out = {}
        for kv in buf.value.decode().split(";"):
            if "=" in kv:
                k, v = kv.split("=", 1)
                out[k] = int(v) if v.lstrip("-").isdigit() else v
        return out

    def set_profile(self, on=True):
        self._check(self.lib.demcz_set_profile(self.ctx, int(bool(on))))

    # -- outputs -----------------------------------------------------------------------------
    def get_chain(self):
        G = self.generations_total // self.thin            # stored generations (opts.thin)
        chain = np.zeros((self.N, self.d, G), order="F")
        log_obj = np.zeros((self.N, G), order="F")
        if G:
            self._check(self.lib.demcz_get_chain(self.ctx, _dp(chain), _dp(log_obj)))
        return chain, log_obj

    def get_state(self):
        X = np.zeros((self.N, self.d), order="F")
        lp = np.zeros(self.N)
        self._check(self.lib.demcz_get_state(self.ctx, _dp(X), _dp(lp)))
        return X, lp

    def get_archive(self):
        M = C.c_int64()
        self._check(self.lib.demcz_get_archive(self.ctx, None, C.byref(M)))
        Z = np.zeros((M.value, self.d), order="F")
        self._check(self.lib.demcz_get_archive(self.ctx, _dp(Z), C.byref(M)))
        return Z

    def get_moments(self):
        mean = np.zeros(self.d)
        cov = np.zeros((self.d, self.d))
        self._check(self.lib.demcz_get_moments(self.ctx, _dp(mean), _dp(cov)))
        return mean, cov

    def get_archive_moments(self, first_row, n_rows):
        """mean / covariance of archive rows [first_row, first_row + n_rows): mean_cov_chain of the K-thinned history."""
        mean = np.zeros(self.d)
        cov = np.zeros((self.d, self.d))
        self._check(self.lib.demcz_get_archive_moments(self.ctx, int(first_row), int(n_rows), _dp(mean), _dp(cov)))
        return mean, cov

    def get_history_moments(self, reset=False):
        """(mean, cov, n) over every generation of every chain so far — mean_cov_chain (src/utils.jl:96-106) of the full
        history from the device accumulators (needs moments=True)."""
        mean = np.zeros(self.d)
        cov = np.zeros((self.d, self.d), order="F")
        n = C.c_int64()
        self._check(self.lib.demcz_get_history_moments(self.ctx, _dp(mean), _dp(cov), C.byref(n), int(bool(reset))))
        return mean, cov, n.value

    def get_archive_checksum(self, first_row=0, n_rows=None):
        """64-bit checksum of archive rows (equal on every rank of a sharded run iff the replicas are identical)."""
        if n_rows is None:
            M = C.c_int64()
            self._check(self.lib.demcz_get_archive(self.ctx, None, C.byref(M)))
            n_rows = M.value - first_row
        h = C.c_uint64()
        self._check(self.lib.demcz_get_archive_checksum(self.ctx, int(first_row), int(n_rows), C.byref(h)))
        return h.value

    def get_rhat(self):
        r = np.zeros(self.d)
        self._check(self.lib.demcz_get_rhat(self.ctx, _dp(r)))
        return r

    def get_accept_counts(self):
        c = np.zeros(self.N, dtype=np.int64)
        self._check(self.lib.demcz_get_accept_counts(self.ctx, c.ctypes.data_as(C.POINTER(C.c_int64))))
        return c

    def eval_logdensity(self, X):
        X = np.asfortranarray(np.atleast_2d(X), dtype=np.float64)
        out = np.zeros(X.shape[0])
        self._check(self.lib.demcz_eval_logdensity(self.ctx, _dp(X), X.shape[0], _dp(out)))
        return out

    def get_trace(self):
        G, nb, N, d = self.generations_total, self.Nblocks, self.N, self.d
        tr = dict(r1=np.zeros((G, nb, N), np.uint32), r2=np.zeros((G, nb, N), np.uint32),
                  u=np.zeros((G, nb, N), np.float32), xprop=np.zeros((G, nb, d, N), np.float32),
                  lp_prop=np.zeros((G, nb, N), np.float32), accepted=np.zeros((G, nb, N), np.uint8))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self._check(self.lib.demcz_get_trace(self.ctx, p(tr["r1"]), p(tr["r2"]), p(tr["u"]), p(tr["xprop"]),
                                             p(tr["lp_prop"]), p(tr["accepted"])))
        return tr


# ------------------------------------------------------------------------------------------------
# demcz_sample / demcz_anneal
# ------------------------------------------------------------------------------------------------
def _unpack(args, Zmat, defaults):
    """Positional tail (N, K, Ngeneration, Nblocks, blockindex, eps_scale, γ) — src/demcz.jl:9."""
    d = Zmat.shape[1]
    names = ["N", "K", "Ngeneration", "Nblocks", "blockindex", "eps_scale", "γ"]
    vals = dict(N=4, K=10, Ngeneration=5000, Nblocks=1, blockindex=[list(range(d))], eps_scale=1e-4 * np.ones(d), γ=2.38)
    vals.update(defaults)
    if len(args) > len(names):
        raise TypeError("too many positional arguments")
    for n, v in zip(names, args):
        vals[n] = v
    return vals


def _concat_prev(prevrun: Optional[MC], mc: MC) -> MC:
    if prevrun is None:
        return mc
    # MC(cat(prevrun.chain, mc.chain, dims=3), cat(prevrun.log_obj, mc.log_obj, dims=2), ...) src/demcz.jl:59
    return MC(np.asfortranarray(np.concatenate([prevrun.chain, mc.chain], axis=2)),
              np.asfortranarray(np.concatenate([prevrun.log_obj, mc.log_obj], axis=1)),
              mc.Xcurrent, mc.log_objcurrent, prevrun.generations_total + mc.generations_total, mc.info)


def _run(logobj, Zmat, v, *, anneal, prevrun, verbose, print_step, autostop, autostop_Rhat, autostop_every,
         temperaturefun, T0, TN, adaptγ, seed, device, store_chain, init_mode, archive_capacity, trace, profile,
         thin=1, moments=False, snooker_prob=0.0, snooker_gamma=0.0, persistent="auto", sharded=False, sync_mode="strict"):
    _check_likelihood(logobj)
    Zmat = np.asfortranarray(Zmat, dtype=np.float64)
    d = Zmat.shape[1]
    if len(v["blockindex"]) != v["Nblocks"]:
        raise ValueError("Nblocks != length(blockindex)")
    gen_offset = prevrun.generations_total if prevrun is not None else 0
    adaptγ = adaptγ or {}
    kw = dict(K=v["K"], Ngeneration=v["Ngeneration"], blockindex=v["blockindex"],
              eps_scale=v["eps_scale"], γ=v["γ"], anneal=anneal, T0=T0, TN=TN,
              autostop=(str(autostop).lstrip(":") == "Rhat") and not anneal, autostop_every=autostop_every,
              autostop_Rhat=autostop_Rhat, adapt=bool(adaptγ.get("adapt", False)) and anneal,
              min_gamma=adaptγ.get("minγ", adaptγ.get("min_gamma", 0.1)),
              max_gamma=adaptγ.get("maxγ", adaptγ.get("max_gamma", 4.0)), adapt_every=adaptγ.get("adapt_every", 500),
              seed=seed, gen_offset=gen_offset, store_chain=store_chain, init_mode=init_mode,
              archive_capacity=archive_capacity, trace=trace, profile=profile, thin=thin, moments=moments,
              snooker_prob=snooker_prob, snooker_gamma=snooker_gamma, persistent=persistent, sync_mode=sync_mode)
    from . import dist as _dist
    rank, world, local_rank = _dist.env_rank_world() if sharded else (0, 1, 0)
    if world > 1:
        # the multi-GPU runner (one process per GPU under torchrun): chains sharded contiguously, archive replicated,
        # appended rows exchanged every K generations — the sync mode of demcz_sample_par (src/demcz.jl:129-137)
        _dist.ensure_process_group(local_rank)
        N_global = int(v["N"])
        n_loc, off = _dist.shard(N_global, rank, world)
        cap = archive_capacity
        if cap == 0:                                      # segments must not straddle the ring end: pad M0 up to a multiple of N
            Zmat, cap = _dist.pad_archive_for_sharding(Zmat, N_global, v["Ngeneration"], v["K"])
            kw["archive_capacity"] = cap
        s = _dist.make_sharded_sampler(logobj, d, N_global=N_global, rank=rank, world=world, device=local_rank, **kw)
    else:
        n_loc, off = int(v["N"]), 0
        s = Sampler(logobj, d, N=v["N"], device=device, **kw)
    try:
        s.set_archive(Zmat)
        if world > 1:
            _dist.open_peer_archives(s, rank=rank, world=world, device=local_rank)
        if prevrun is not None:
            # X = prevrun.chain[:,:,end]; log_objcurrent = prevrun.log_objcurrent   (src/demcz.jl:20-21)
            s.set_state(prevrun.Xcurrent[off:off + n_loc], prevrun.log_objcurrent[off:off + n_loc])
        if anneal and temperaturefun is not tempbaseline:
            G = v["Ngeneration"]
            s.set_temperature_table([temperaturefun(ig, G, T0, TN) for ig in range(1, G + 1)])
        if verbose:
            info = {}
            while s.generations_total < v["Ngeneration"]:
                info = s.run(print_step)
                X, lp = s.get_state()
                if rank == 0:
                    print("-----------------------")
                    print(f"iteration {s.generations_total}")
                    if anneal:
                        print(f"bestval = {lp.max()}")
                    else:
                        print(f"average par = {X.mean(axis=0)}")
                        print(f"average val = {lp.mean()}")
                    print("-----------------------")
                if info["stopped_by_rhat"]:
                    break
        else:
            info = s.run(0)
        X, lp = s.get_state()
        if store_chain:
            chain, log_obj = s.get_chain()
        else:
            chain, log_obj = np.zeros((n_loc, d, 0), order="F"), np.zeros((n_loc, 0), order="F")
        info["rhat"] = s.get_rhat() if info.get("rhat_valid") else None
        if moments:
            info["history_moments"] = s.get_history_moments()
        if trace:
            info["trace"] = s.get_trace()
        Z = s.get_archive()
        if world > 1:                                     # every rank returns the whole population, like the reference's master
            info["accepted"] = int(_dist.allreduce_sum(info["accepted"], local_rank))
            info["proposals"] = int(_dist.allreduce_sum(info["proposals"], local_rank))
            X, lp = _dist.allgather_rows(X, local_rank), _dist.allgather_rows(lp, local_rank)
            chain, log_obj = _dist.allgather_rows(chain, local_rank), _dist.allgather_rows(log_obj, local_rank)
    finally:
        if world > 1:
            _dist.barrier()                               # no rank frees its archive while a peer still has it mapped
        s.close()
    mc = MC(np.asfortranarray(chain), np.asfortranarray(log_obj), X, lp, s.generations_total, info)
    return _concat_prev(prevrun, mc), Z


def demcz_sample(logobj, Zmat, *args, prevrun=None, verbose=None, print_step=None, autostop=None, autostop_Rhat=None,
                 autostop_every=None, seed=31953150, device=0, store_chain=True, init_mode="last_N_of_Z",
                 archive_capacity=0, trace=False, profile=False, **extra):
    """demcz_sample(logobj, Zmat, opts::DEMCopt; prevrun) — src/demcz.jl:1-3, or the positional form
    demcz_sample(logobj, Zmat, N, K, Ngeneration, Nblocks, blockindex, eps_scale, γ; ...) — src/demcz.jl:9.
    Returns (mc::MC, Z) with Z = Z[1:M, :]."""
    if len(args) == 1 and isinstance(args[0], DEMCopt):
        o = args[0]
        v = dict(N=o.N, K=o.K, Ngeneration=o.Ngeneration, Nblocks=o.Nblocks, blockindex=o.blockindex,
                 eps_scale=o.eps_scale, γ=o.γ)
        kw = dict(verbose=o.verbose, print_step=o.print_step, autostop=o.autostop, autostop_Rhat=o.autostop_Rhat,
                  autostop_every=o.autostop_every)
    else:
        v = _unpack(args, np.asarray(Zmat), {})
        kw = dict(verbose=True, print_step=100, autostop="no", autostop_Rhat=1.01, autostop_every=1000)  # src/demcz.jl:9
    for name, val in (("verbose", verbose), ("print_step", print_step), ("autostop", autostop),
                      ("autostop_Rhat", autostop_Rhat), ("autostop_every", autostop_every)):
        if val is not None:
            kw[name] = val
    return _run(logobj, Zmat, v, anneal=False, prevrun=prevrun, temperaturefun=tempbaseline, T0=3.0, TN=1e-3,
                adaptγ=None, seed=seed, device=device, store_chain=store_chain, init_mode=init_mode,
                archive_capacity=archive_capacity, trace=trace, profile=profile, **kw, **_extra(extra))


_EXTRA = ("thin", "moments", "snooker_prob", "snooker_gamma", "persistent", "sharded", "sync_mode")


def _extra(extra):
    """Keyword extensions that have no counterpart in the reference (history thinning, streaming moments, the snooker
    update, small-N persistent launches, multi-GPU sharding / sync mode)."""
    bad = set(extra) - set(_EXTRA)
    if bad:
        raise TypeError(f"unexpected keyword argument(s) {sorted(bad)}")
    return extra


_ADAPT_DEFAULT = {"adapt": True, "minγ": 0.1, "maxγ": 4.0, "adapt_every": 500}   # src/demcz_anneal.jl:14


def demcz_anneal(logobj, Zmat, *args, prevrun=None, verbose=None, print_step=None,
                 temperaturefun: Callable = tempbaseline, T0=None, TN=None, adaptγ=None, seed=31953150, device=0,
                 store_chain=True, init_mode="last_N_of_Z", archive_capacity=0, trace=False, profile=False, **extra):
    """demcz_anneal(logobj, Zmat, opts; prevrun, temperaturefun, adaptγ) — src/demcz_anneal.jl:14-16, or the
    positional form (:18).  The schedule is T0*(TN/T0)^(ig/Ngeneration) as intended."""
    if len(args) == 1 and isinstance(args[0], DEMCopt):
        o = args[0]
        v = dict(N=o.N, K=o.K, Ngeneration=o.Ngeneration, Nblocks=o.Nblocks, blockindex=o.blockindex,
                 eps_scale=o.eps_scale, γ=o.γ)
        kw = dict(verbose=o.verbose, print_step=o.print_step)
        T0 = o.T0 if T0 is None else T0
        TN = o.TN if TN is None else TN
    else:
        v = _unpack(args, np.asarray(Zmat), {})
        kw = dict(verbose=True, print_step=100)
        T0 = 3.0 if T0 is None else T0             # src/demcz_anneal.jl:18
        TN = 0.0 if TN is None else TN
    if verbose is not None:
        kw["verbose"] = verbose
    if print_step is not None:
        kw["print_step"] = print_step
    return _run(logobj, Zmat, v, anneal=True, prevrun=prevrun, autostop="no", autostop_Rhat=1.01, autostop_every=1000,
                temperaturefun=temperaturefun, T0=float(T0), TN=float(TN),
                adaptγ=dict(_ADAPT_DEFAULT) if adaptγ is None else adaptγ, seed=seed, device=device,
                store_chain=store_chain, init_mode=init_mode, archive_capacity=archive_capacity, trace=trace,
                profile=profile, **kw, **_extra(extra))


# ------------------------------------------------------------------------------------------------
# the reference's parallel entry points (src/demcz.jl:101-165, src/demcz_anneal.jl:82-140)
# ------------------------------------------------------------------------------------------------
def demcz_sample_par(logobj, Zmat, *args, sync_every=1000, prevrun=None, workerpool=None, autostop=None, autostop_Rhat=None,
                     verbose=None, **kw):
    """demcz_sample_par(logobj, Zmat, opts | N, K, ...; sync_every, prevrun, autostop, autostop_Rhat, workerpool)
    — src/demcz.jl:101-165.

    In the reference one Distributed.jl worker runs each chain for `sync_every` generations between barriers, all
    workers append to one shared Z, and the master checks R-hat on every sync block (:129-155).  Here the workers are
    the GPUs of the box: launched under torchrun (WORLD_SIZE > 1) the call runs the sharded sampler — chains
    partitioned over the ranks, archive replicated, every rank's appended rows exchanged over NVLink every K
    generations — and every rank returns the whole (mc, Z) like the reference's master.  `sync_every` is the R-hat
    check period (autostop_every); rows become visible to the other ranks every K <= sync_every generations, i.e. at
    least as often as a sync block ends.  `workerpool` is accepted and ignored.  With one process it is the same
    generation loop on one GPU.  Positional defaults follow src/demcz.jl:106 (autostop=:Rhat, Rhat < 1.1)."""
    if len(args) == 1 and isinstance(args[0], DEMCopt):
        o = args[0]
        kw2 = dict(autostop=o.autostop, autostop_Rhat=o.autostop_Rhat, verbose=o.verbose)
    else:
        kw2 = dict(autostop="Rhat", autostop_Rhat=1.1, verbose=True)
    for name, val in (("autostop", autostop), ("autostop_Rhat", autostop_Rhat), ("verbose", verbose)):
        if val is not None:
            kw2[name] = val
    kw.setdefault("sharded", True)
    return demcz_sample(logobj, Zmat, *args, prevrun=prevrun, autostop_every=int(sync_every), **kw2, **kw)


def demcz_anneal_par(logobj, Zmat, *args, sync_every=1000, prevrun=None, temperaturefun: Callable = tempbaseline, T0=None,
                     TN=None, adaptγ=None, verbose=None, workerpool=None, **kw):
    """demcz_anneal_par (src/demcz_anneal.jl:82-140): γ is adapted once per sync block there, i.e. adapt_every = sync_every;
    the positional form defaults to TN = 1e-3 (src/demcz_anneal.jl:86), unlike the serial one.  Under torchrun the
    chains are sharded over the GPUs like demcz_sample_par (the accept counter of the γ rule is all-reduced)."""
    ad = dict(_ADAPT_DEFAULT) if adaptγ is None else dict(adaptγ)
    ad["adapt_every"] = int(sync_every)
    if not (len(args) == 1 and isinstance(args[0], DEMCopt)) and TN is None:
        TN = 1e-3
    kw.setdefault("sharded", True)
    return demcz_anneal(logobj, Zmat, *args, prevrun=prevrun, temperaturefun=temperaturefun, T0=T0, TN=TN, adaptγ=ad,
                        verbose=verbose, **kw)
