"""numpy restatement of the DEMCz hot path of chrished/DEMC.jl + ctypes binding of the C one.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module, and only as the checker / timed CPU baseline.

PARITY UNPINNED w.r.t. the Julia reference (no Julia toolchain here, the reference's tests hold no
golden vectors, and its MersenneTwister stream is replaced by Philox4x32-10).  Pinned instead by:
Random123 Philox KATs, scipy's multivariate-normal logpdf, the FP64 Gram form of the regression
SSR, and agreement between this numpy restatement and the independent C one
(oracle/demcz_oracle.c) — see tests/test_oracle_*.py and tests/golden/.

Citations are file:line into /root/reference.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# ------------------------------------------------------------------------------------------------
# Philox4x32-10 (Random123).  Replaces Julia's global RNG used at src/demcz.jl:177,179,184,186,198.
# ------------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (2,) ints -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    c0, c1, c2, c3 = (ctr[..., i].copy() for i in range(4))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def uniform(w):
    """((w >> 9) + 0.5) * 2^-23 — in (0,1), exact in FP32 and FP64. Replaces rand() (src/demcz.jl:198)."""
    return ((np.asarray(w, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) + 0.5) * (1.0 / 8388608.0)


def draw_pair(w_r1, w_r2, M):
    """i1 = rand(1:M); deleteat!(set,i1); i2 = rand(set) — src/demcz.jl:176-179 (0-based)."""
    a = (np.asarray(w_r1, dtype=np.uint64) * np.uint64(M)) >> np.uint64(32)
    j = (np.asarray(w_r2, dtype=np.uint64) * np.uint64(M - 1)) >> np.uint64(32)
    return a.astype(np.int64), (j + (j >= a)).astype(np.int64)


def box_muller16(w):
    """One Philox word -> two N(0,1) draws (replaces randn of src/demcz.jl:184,186): 16-bit uniforms from the
    high / low half-word, u = (h + 0.5) 2^-16 (exact in FP32, in (0,1)); z0 = r cos(2 pi u2), z1 = r sin(2 pi u2)."""
    w = np.asarray(w, dtype=np.uint32)
    u1 = ((w >> np.uint32(16)).astype(np.float64) + 0.5) / 65536.0
    u2 = ((w & np.uint32(0xFFFF)).astype(np.float64) + 0.5) / 65536.0
    r = np.sqrt(-2.0 * np.log(u1))
    th = 2.0 * math.pi * u2
    return r * np.cos(th), r * np.sin(th)


def counters(chain_ids, gen_abs, block, call):
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    ctr = np.empty(chain_ids.shape + (4,), dtype=np.uint64)
    ctr[..., 0] = chain_ids & _MASK
    ctr[..., 1] = np.uint64(gen_abs & 0xFFFFFFFF)
    ctr[..., 2] = np.uint64(((block << 16) | call) & 0xFFFFFFFF)
    ctr[..., 3] = np.uint64((gen_abs >> 32) & 0xFFFFFFFF)
    return ctr


# ------------------------------------------------------------------------------------------------
# Likelihood descriptors (the reference takes a closure: src/demcz.jl:189)
# ------------------------------------------------------------------------------------------------
class MvNormalLik:
    """logpdf(MvNormal(mu, Sigma), x), test/example_normpdf.jl:9-16; quadratic objective
    -sum((x-mu)^2) of test/test_anneal.jl:10 via ``MvNormalLik.quadratic(mu)``."""

    kind = 0

    def __init__(self, mu, Sigma, drop_constant=False):
        self.mu = np.ascontiguousarray(mu, dtype=np.float64)
        self.Sigma = np.asarray(Sigma, dtype=np.float64)
        self.d = self.mu.size
        self.L = np.linalg.cholesky(self.Sigma)
        logdet = 2.0 * np.sum(np.log(np.diag(self.L)))
        self.c0 = 0.0 if drop_constant else -0.5 * (self.d * math.log(2 * math.pi) + logdet)

    @classmethod
    def quadratic(cls, mu):
        mu = np.asarray(mu, dtype=np.float64)
        return cls(mu, 0.5 * np.eye(mu.size), drop_constant=True)

    def logdensity(self, X):
        """X: (N, d) -> (N,)"""
        import scipy.linalg as sl
        y = sl.solve_triangular(self.L, (np.atleast_2d(X) - self.mu).T, lower=True)
        return self.c0 - 0.5 * np.sum(y * y, axis=0)


class LinRegLik:
    """-0.5 * sum((y .- X*b).^2), test/example_linreg.jl:32."""

    kind = 1

    def __init__(self, X, y):
        self.X = np.asfortranarray(X, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.n, self.d = self.X.shape

    def logdensity(self, B):
        R = self.y[:, None] - self.X @ np.atleast_2d(B).T
        return -0.5 * np.sum(R * R, axis=0)


class LogisticLik(LinRegLik):
    """Bernoulli-logit GLM: sum(y*eta - log(1 + exp(eta))), eta = X b.  Not in the reference (which takes any closure,
    src/demcz.jl:189); textbook formula, evaluated with the overflow-safe logaddexp."""

    kind = 2

    def logdensity(self, B):
        eta = self.X @ np.atleast_2d(B).T
        return self.y @ eta - np.sum(np.logaddexp(0.0, eta), axis=0)


# ------------------------------------------------------------------------------------------------
# One generation, all chains, generation-synchronous archive (update_blocks ->
# update_demcz_chain_block -> accept: src/demcz.jl:167-203; tempered: src/demcz_anneal.jl:142-178)
# ------------------------------------------------------------------------------------------------
def block_update(X, lp, Z, M, lik, block, ib, eps_scale, gamma, seed, gen_abs, T=None, chain_offset=0, chain_ids=None,
                 zstart=0, zcap=None):
    """update_demcz_chain_block for every chain (src/demcz.jl:174-195; tempered src/demcz_anneal.jl:149-170).

    X (N,d), lp (N,) are the states ENTERING block ib; returns a dict with r1, r2, u, xprop, lp_prop,
    accepted and margin = (lp_prop - lp) - T*log(u)  (accept <=> margin > 0; T=1 when T is None).
    chain_ids: the global chain id of every row of X (default chain_offset + 0..N-1) — lets a checker replay a sample
    of a large population.  zstart/zcap: logical row r of the drawable view lives at Z[(zstart + r) % zcap] (ring)."""
    X = np.asarray(X, dtype=np.float64)
    lp = np.asarray(lp, dtype=np.float64)
    N, d = X.shape
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    cid = (np.arange(N, dtype=np.uint64) if chain_ids is None else np.asarray(chain_ids, dtype=np.uint64)) + np.uint64(chain_offset)
    block = np.asarray(block, dtype=np.int64)
    w = philox4x32_10(counters(cid, gen_abs, ib, 0), key)
    r1, r2 = draw_pair(w[:, 0], w[:, 1], M)                                # :176-179
    p1, p2 = (r1, r2) if zcap is None else ((r1 + zstart) % zcap, (r2 + zstart) % zcap)
    blocklen = block.size
    c = gamma if blocklen == 1 else gamma / math.sqrt(2.0 * blocklen)      # :183-187
    zn = {}
    Xp = X.copy()                                                          # de_diffvec = zeros :180
    for k in block:
        q = int(k) >> 3
        if q not in zn:                                                    # randn(blocklen) :186
            wn = philox4x32_10(counters(cid, gen_abs, ib, 1 + q), key)
            zn[q] = [z for j in range(4) for z in box_muller16(wn[:, j])]   # columns 8q .. 8q+7
        diff = c * (Z[p1, k] - Z[p2, k]) + eps_scale[k] * zn[q][int(k) & 7]
        Xp[:, k] = X[:, k] + diff                                          # :188
    lpp = lik.logdensity(Xp)                                               # :189
    u = uniform(w[:, 2])
    with np.errstate(divide="ignore", invalid="ignore"):
        if T is None:
            acc = np.log(u) < (lpp - lp)                                   # :198
            margin = (lpp - lp) - np.log(u)
        else:
            acc = np.log(u) < (lpp - lp) / T                               # anneal :173
            margin = (lpp - lp) - T * np.log(u)
    return dict(r1=r1, r2=r2, u=u, xprop=Xp, lp_prop=lpp, accepted=acc, margin=margin)


SNOOKER_CALL = 0xFFFF0000


def is_snooker_gen(seed, gen_abs, prob):
    """Generation-level choice of the snooker update: ctr = (0xFFFFFFFF, gen lo, 0xFFFF0000, gen hi), word 0 -> u < prob."""
    if not prob > 0:
        return False
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    ctr = np.array([0xFFFFFFFF, gen_abs & 0xFFFFFFFF, SNOOKER_CALL, (gen_abs >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    return bool(uniform(philox4x32_10(ctr, key)[0]) < prob)


def snooker_update(X, lp, Z, M, lik, gamma_s, seed, gen_abs, T=None, chain_offset=0, chain_ids=None, zstart=0, zcap=None):
    """Snooker update of ter Braak & Vrugt (2008) for every chain — the update the reference's README cites
    (README.md:57-59) but does not implement; restated from the paper:
        z, z1, z2: three distinct archive rows; e = x - z;
        x' = x + gamma_s * e * ((z1 - z2) . e) / (e . e)          (the difference of the projections of z1, z2 on the line x - z)
        accept with probability min(1, pi(x')/pi(x) * (|x' - z| / |x - z|)^(d-1)),   |x' - z| / |x - z| = |1 + gamma_s proj|
    gamma_s = 0 -> uniform on [1.2, 2.2] (the paper's choice).  Tempered: log u < Delta/T + (d-1) log(...)."""
    X = np.asarray(X, dtype=np.float64)
    lp = np.asarray(lp, dtype=np.float64)
    N, d = X.shape
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    cid = (np.arange(N, dtype=np.uint64) if chain_ids is None else np.asarray(chain_ids, dtype=np.uint64)) + np.uint64(chain_offset)
    ctr = counters(cid, gen_abs, 0, 0)
    ctr[..., 2] = np.uint64(SNOOKER_CALL)
    w = philox4x32_10(ctr, key)
    z = ((w[:, 0].astype(np.uint64) * np.uint64(M)) >> np.uint64(32)).astype(np.int64)
    j1 = ((w[:, 1].astype(np.uint64) * np.uint64(M - 1)) >> np.uint64(32)).astype(np.int64)
    z1 = j1 + (j1 >= z)
    z2 = ((w[:, 2].astype(np.uint64) * np.uint64(M - 2)) >> np.uint64(32)).astype(np.int64)
    lo, hi = np.minimum(z, z1), np.maximum(z, z1)
    z2 = z2 + (z2 >= lo)
    z2 = z2 + (z2 >= hi)
    if gamma_s == 0:
        ctr[..., 2] = np.uint64(SNOOKER_CALL + 1)
        # the device forms 1.2f + u in FP32 (u is exact in FP32)
        gs = (np.float32(1.2) + uniform(philox4x32_10(ctr, key)[:, 0]).astype(np.float32)).astype(np.float64)
    else:
        gs = np.full(N, float(np.float32(gamma_s)))
    ph = (lambda r: r) if zcap is None else (lambda r: (r + zstart) % zcap)
    e = X - Z[ph(z)]
    ee = np.sum(e * e, axis=1)
    num = np.sum((Z[ph(z1)] - Z[ph(z2)]) * e, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        proj = np.where(ee > 0, gs * num / ee, 0.0)
        Xp = X + proj[:, None] * e
        lpp = lik.logdensity(Xp)
        geo = (d - 1) * np.log(np.abs(1.0 + proj))
        u = uniform(w[:, 3])
        Tm = 1.0 if T is None else T
        margin = (lpp - lp) + Tm * geo - Tm * np.log(u)
        acc = (ee > 0) & (margin > 0)
    return dict(z=z, r1=z1, r2=z2, u=u, xprop=Xp, lp_prop=lpp, accepted=acc, margin=margin, geo=geo)


def generation(X, lp, Z, M, lik, blocks, eps_scale, gamma, seed, gen_abs, T=None, chain_offset=0):
    """update_blocks for every chain, generation-synchronous archive (src/demcz.jl:167-172).
    X (N,d), lp (N,) are NOT modified; returns (Xnew, lpnew, trace[block])."""
    X = np.array(X, dtype=np.float64)
    lp = np.array(lp, dtype=np.float64)
    tr = {k: [] for k in ("r1", "r2", "u", "xprop", "lp_prop", "accepted")}
    for ib, block in enumerate(blocks):                                    # :168
        b = block_update(X, lp, Z, M, lik, block, ib, eps_scale, gamma, seed, gen_abs, T, chain_offset)
        for name in tr:
            tr[name].append(np.array(b[name]))
        acc = b["accepted"]
        X[acc] = b["xprop"][acc]                                           # :190-194
        lp[acc] = b["lp_prop"][acc]
    return X, lp, {k: np.stack(v) for k, v in tr.items()}


def tempbaseline(ig, Ng, T0, TN):
    """src/demcz_anneal.jl:1-3"""
    return T0 * (TN / T0) ** (ig / Ng)


def rhat_gelman(chain):
    """Rhat_gelman, src/utils.jl:2-20. chain: (N, d, G)."""
    N, d, G = chain.shape
    n = G // 2                                                             # :4
    cs = np.concatenate([chain[:, :, :n], chain[:, :, n:2 * n]], axis=0)   # :6-8
    m = 2 * N
    avg_par = cs.mean(axis=(0, 2), keepdims=True)                          # :10
    avg_chains = cs.mean(axis=2, keepdims=True)                            # :11
    B = n / (m - 1) * np.sum((avg_chains - avg_par) ** 2, axis=0)          # :13
    sj = 1.0 / (n - 1) * np.sum((cs - avg_chains) ** 2, axis=2, keepdims=True)   # :14
    W = 1.0 / m * np.sum(sj, axis=0)                                       # :15
    varhat = (n - 1) / n * W + 1.0 / n * B                                 # :16
    return np.sqrt(varhat / W).reshape(-1)                                 # :18


def flatten_chain(chain):
    """src/utils.jl:22-32 -> (d, N*G), generation-major, chain fastest."""
    N, d, G = chain.shape
    return chain.transpose(1, 2, 0).reshape(d, G * N)


def mean_cov_chain(chain):
    """src/utils.jl:96-106 (population covariance, 1/(N*G))."""
    f = flatten_chain(chain)
    b = f.mean(axis=1)
    c = (f - b[:, None]) @ (f - b[:, None]).T / f.shape[1]
    return b, c


def accept_ratio(log_obj):
    """src/utils.jl:61. log_obj: (N, G) -> (N,)"""
    return np.sum(np.diff(log_obj, axis=1) != 0.0, axis=1) / (log_obj.shape[1] - 1)


# ------------------------------------------------------------------------------------------------
# ctypes binding of the C restatement (oracle/demcz_oracle.c)
# ------------------------------------------------------------------------------------------------
class _Lik(C.Structure):
    _fields_ = [("kind", C.c_int), ("d", C.c_int), ("mu", C.c_void_p), ("L", C.c_void_p), ("c0", C.c_double),
                ("n", C.c_int64), ("X", C.c_void_p), ("y", C.c_void_p)]


class _Opts(C.Structure):
    _fields_ = [("N", C.c_int), ("K", C.c_int), ("Nblocks", C.c_int), ("d", C.c_int),
                ("block_ptr", C.c_void_p), ("block_idx", C.c_void_p), ("eps_scale", C.c_void_p),
                ("gamma", C.c_double), ("seed", C.c_uint64), ("gen_offset", C.c_int64),
                ("chain_offset", C.c_int64), ("append_mode", C.c_int), ("index_mode", C.c_int),
                ("nthreads", C.c_int), ("z_rowmajor", C.c_int)]


class _Trace(C.Structure):
    _fields_ = [("r1", C.c_void_p), ("r2", C.c_void_p), ("u", C.c_void_p), ("xprop", C.c_void_p),
                ("lp_prop", C.c_void_p), ("accepted", C.c_void_p)]


class _RunOpts(C.Structure):
    _fields_ = [("Ngeneration", C.c_int), ("anneal", C.c_int), ("Ttable", C.c_void_p), ("autostop", C.c_int),
                ("autostop_every", C.c_int), ("autostop_Rhat", C.c_double), ("adapt", C.c_int),
                ("min_gamma", C.c_double), ("max_gamma", C.c_double), ("adapt_every", C.c_int),
                ("init_zeros", C.c_int), ("store_chain", C.c_int)]


class _RunInfo(C.Structure):
    _fields_ = [("generations_done", C.c_int), ("stopped_by_rhat", C.c_int), ("final_gamma", C.c_double),
                ("M", C.c_int64)]


def build_c_oracle():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return os.path.join(_HERE, "libdemcz_oracle.so")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class COracle:
    """Thin binding of libdemcz_oracle.so."""

    def __init__(self, build=True):
        path = os.path.join(_HERE, "libdemcz_oracle.so")
        if build or not os.path.exists(path):
            build_c_oracle()
        self.lib = C.CDLL(path)
        self.lib.orc_logdensity.restype = C.c_double
        self.lib.orc_tempbaseline.restype = C.c_double
        self.lib.orc_tempbaseline.argtypes = [C.c_double] * 4
        self.lib.orc_uniform.restype = C.c_double

    # -- small pieces ------------------------------------------------------------------------
    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*[int(x) for x in ctr])
        k = (C.c_uint32 * 2)(*[int(x) for x in key])
        o = (C.c_uint32 * 4)()
        self.lib.orc_philox4x32_10(c, k, o)
        return [int(x) for x in o]

    def make_lik(self, lik):
        keep = []
        s = _Lik()
        s.kind, s.d = lik.kind, lik.d
        if lik.kind == 0:
            mu = np.ascontiguousarray(lik.mu, dtype=np.float64)
            L = np.asfortranarray(lik.L, dtype=np.float64)
            keep += [mu, L]
            s.mu, s.L, s.c0 = _ptr(mu), _ptr(L), lik.c0
        else:                                   # kind 1 (linreg) and 2 (logistic) carry X, y
            keep += [lik.X, lik.y]
            s.n, s.X, s.y = lik.n, _ptr(lik.X), _ptr(lik.y)
        s._keep = keep
        return s

    def make_opts(self, N, K, d, blocks, eps_scale, gamma, seed, gen_offset=0, chain_offset=0,
                  append_mode=0, index_mode=0, nthreads=1, z_rowmajor=0):
        bp = np.zeros(len(blocks) + 1, dtype=np.int32)
        bp[1:] = np.cumsum([len(b) for b in blocks])
        bi = np.concatenate([np.asarray(b, dtype=np.int32) for b in blocks]).astype(np.int32)
        eps = np.ascontiguousarray(eps_scale, dtype=np.float64)
        o = _Opts(N, K, len(blocks), d, _ptr(bp), _ptr(bi), _ptr(eps), gamma, seed, gen_offset, chain_offset,
                  append_mode, index_mode, nthreads, z_rowmajor)
        o._keep = [bp, bi, eps]
        return o

    def logdensity(self, lik, X):
        s = self.make_lik(lik)
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        out = np.empty(X.shape[0])
        for i in range(X.shape[0]):
            xi = np.ascontiguousarray(X[i])
            out[i] = self.lib.orc_logdensity(C.byref(s), _ptr(xi))
        return out

    def rhat(self, chain):
        N, d, G = chain.shape
        ch = np.asfortranarray(chain, dtype=np.float64)
        out = np.empty(d)
        self.lib.orc_rhat_gelman(_ptr(ch), N, G, d, _ptr(out))
        return out

    def tempbaseline(self, ig, Ng, T0, TN):
        return self.lib.orc_tempbaseline(ig, Ng, T0, TN)

    def is_snooker_gen(self, seed, gen_abs, prob):
        self.lib.orc_is_snooker_gen.argtypes = [C.c_uint64, C.c_int64, C.c_double]
        return bool(self.lib.orc_is_snooker_gen(seed, gen_abs, prob))

    def snooker_generation(self, X, lp, Z, M, lik, gamma_s, seed, gen_abs, T=None, chain_offset=0):
        """orc_snooker_generation: returns (Xnew, lpnew, trace) like generation()."""
        N, d = X.shape
        Xf = np.asfortranarray(X, dtype=np.float64).copy(order="F")
        lpf = np.array(lp, dtype=np.float64)
        Zf = np.asfortranarray(Z, dtype=np.float64)
        tr = dict(r1=np.zeros(N, np.uint32), r2=np.zeros(N, np.uint32), u=np.zeros(N), xprop=np.zeros((d, N)),
                  lp_prop=np.zeros(N), accepted=np.zeros(N, np.uint8))
        t = _Trace(*[_ptr(tr[k]) for k in ("r1", "r2", "u", "xprop", "lp_prop", "accepted")])
        o = self.make_opts(N, 1, d, [list(range(d))], np.zeros(d), 0.0, seed, 0, chain_offset, 0, 0, 1)
        s = self.make_lik(lik)
        self.lib.orc_snooker_generation(C.byref(o), C.byref(s), _ptr(Xf), _ptr(lpf), _ptr(Zf), C.c_int64(Zf.shape[0]), C.c_int64(M),
                                        C.c_int64(gen_abs), C.c_double(gamma_s), 0 if T is None else 1, C.c_double(1.0 if T is None else T),
                                        C.byref(t))
        tr["xprop"] = tr["xprop"].T
        tr["accepted"] = tr["accepted"].astype(bool)
        tr["r1"], tr["r2"] = tr["r1"].astype(np.int64), tr["r2"].astype(np.int64)
        return Xf, lpf, tr

    # -- one generation (teacher-forced parity) -------------------------------------------------
    def generation(self, X, lp, Z, M, lik, blocks, eps_scale, gamma, seed, gen_abs, T=None, chain_offset=0,
                   append_mode=0, append_now=False, nthreads=1):
        N, d = X.shape
        Xf = np.asfortranarray(X, dtype=np.float64).copy(order="F")
        lpf = np.array(lp, dtype=np.float64)
        Zf = np.asfortranarray(Z, dtype=np.float64)
        if append_mode == 1:
            Zf = Zf.copy(order="F")
        nb = len(blocks)
        tr = dict(r1=np.zeros((nb, N), np.uint32), r2=np.zeros((nb, N), np.uint32), u=np.zeros((nb, N)),
                  xprop=np.zeros((nb, d, N)), lp_prop=np.zeros((nb, N)), accepted=np.zeros((nb, N), np.uint8))
        t = _Trace(*[_ptr(tr[k]) for k in ("r1", "r2", "u", "xprop", "lp_prop", "accepted")])
        o = self.make_opts(N, 1, d, blocks, eps_scale, gamma, seed, 0, chain_offset, append_mode, 0, nthreads)
        s = self.make_lik(lik)
        Mv = C.c_int64(M)
        self.lib.orc_generation(C.byref(o), C.byref(s), _ptr(Xf), _ptr(lpf), _ptr(Zf), C.c_int64(Zf.shape[0]),
                                C.byref(Mv), C.c_int64(gen_abs), 0 if T is None else 1,
                                C.c_double(1.0 if T is None else T), 1 if append_now else 0, C.byref(t))
        tr["xprop"] = tr["xprop"].transpose(0, 2, 1)
        tr["accepted"] = tr["accepted"].astype(bool)
        tr["r1"] = tr["r1"].astype(np.int64)
        tr["r2"] = tr["r2"].astype(np.int64)
        return Xf, lpf, tr, Zf, Mv.value

    # -- full driver ----------------------------------------------------------------------------
    def run(self, lik, Z0, N, K, G, blocks, eps_scale, gamma, seed, *, anneal=False, Ttable=None, autostop=False,
            autostop_every=1000, autostop_Rhat=1.05, adapt=False, min_gamma=0.1, max_gamma=4.0, adapt_every=500,
            init_zeros=False, store_chain=True, X0=None, lp0=None, gen_offset=0, append_mode=0, index_mode=0,
            nthreads=1, z_rowmajor=0):
        Z0 = np.asarray(Z0, dtype=np.float64)
        M0, d = Z0.shape
        cap = M0 + -(-N * G // K)
        if z_rowmajor:
            Z = np.zeros((cap, d), order="C")
        else:
            Z = np.zeros((cap, d), order="F")
        Z[:M0] = Z0
        o = self.make_opts(N, K, d, blocks, eps_scale, gamma, seed, gen_offset, 0, append_mode, index_mode,
                           nthreads, z_rowmajor)
        s = self.make_lik(lik)
        Tt = np.ascontiguousarray(Ttable, dtype=np.float64) if Ttable is not None else None
        r = _RunOpts(G, 1 if anneal else 0, _ptr(Tt), 1 if autostop else 0, autostop_every, autostop_Rhat,
                     1 if adapt else 0, min_gamma, max_gamma, adapt_every, 1 if init_zeros else 0,
                     1 if store_chain else 0)
        chain = np.zeros((N, d, G), order="F") if store_chain else None
        log_obj = np.zeros((N, G), order="F") if store_chain else None
        Xout = np.zeros((N, d), order="F")
        lpout = np.zeros(N)
        rhat = np.zeros(d)
        info = _RunInfo()
        X0f = np.asfortranarray(X0, dtype=np.float64) if X0 is not None else None
        lp0f = np.ascontiguousarray(lp0, dtype=np.float64) if lp0 is not None else None
        import time
        t0 = time.perf_counter()
        rc = self.lib.orc_run(C.byref(o), C.byref(r), C.byref(s), _ptr(Z), C.c_int64(cap), C.c_int64(M0), _ptr(X0f),
                              _ptr(lp0f), _ptr(chain), _ptr(log_obj), _ptr(Xout), _ptr(lpout), _ptr(rhat),
                              C.byref(info))
        seconds = time.perf_counter() - t0
        assert rc == 0
        g = info.generations_done
        return dict(chain=chain[:, :, :g] if store_chain else None, log_obj=log_obj[:, :g] if store_chain else None,
                    X=Xout, lp=lpout, Z=np.array(Z[:info.M]), M=info.M, generations=g,
                    stopped_by_rhat=bool(info.stopped_by_rhat), gamma=info.final_gamma, rhat=rhat, seconds=seconds)
