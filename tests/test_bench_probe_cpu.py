"""CPU check of bench.py's parity-probe replay (the code that runs on the GPU box after the timed region): a stand-in
sampler that produces the trace from the numpy oracle itself must replay clean, and a corrupted trace must be caught.
Keeps Python-level bugs of the probe off the GPU budget."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import demcz_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_probe_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class FakeSampler:
    """The Sampler methods the probe uses, computed by the oracle (FP64) and stored in the GPU trace's layouts/dtypes."""

    def __init__(self, w, lik, N, K, seed, Tt=None, corrupt=None):
        self.w, self.lik, self.N, self.K, self.seed, self.Tt, self.corrupt = w, lik, N, K, seed, Tt, corrupt

    def set_archive(self, Z):
        self.Z = np.asarray(Z)
        self.X = self.Z[-self.N:].copy()
        self.lp = self.lik.logdensity(self.X)

    def get_state(self):
        return self.X.copy(), self.lp.copy()

    def run(self, n=0):
        w, N, K = self.w, self.N, self.K
        nb, d = len(w["blocks"]), w["d"]
        tr = dict(r1=np.zeros((K, nb, N), np.uint32), r2=np.zeros((K, nb, N), np.uint32), u=np.zeros((K, nb, N), np.float32),
                  xprop=np.zeros((K, nb, d, N), np.float32), lp_prop=np.zeros((K, nb, N), np.float32), accepted=np.zeros((K, nb, N), np.uint8))
        X, lp = self.X.copy(), self.lp.copy()
        for g in range(1, K + 1):
            T = None if self.Tt is None else float(np.float32(self.Tt[g - 1]))
            X, lp, t = orc.generation(X, lp, self.Z, self.Z.shape[0], self.lik, w["blocks"], 1e-5 * np.ones(d), w["gamma"], self.seed, g, T=T)
            for k in tr:
                v = t[k]
                tr[k][g - 1] = v.transpose(0, 2, 1) if k == "xprop" else v
        if self.corrupt == "index":
            tr["r1"][K // 2, 0, :] ^= 1
        if self.corrupt == "decision":
            tr["accepted"][:, 0, :] ^= 1
        self.tr = tr

    def get_trace(self):
        return self.tr

    def close(self):
        pass


@pytest.mark.parametrize("name", ["c1", "c3"])
def test_probe_replay_is_clean_on_an_oracle_made_trace_and_catches_corruption(name):
    b = _bench()
    w = b.workload(name, chains_per_gpu=96, archive_rows=400)
    lik = orc.MvNormalLik(w["mu"], w["Sigma"], drop_constant=(w["kind"] == "quad"))
    Z0 = b.make_archive(w, 400)
    K = w["K"]
    Tt = np.array([2.0 * (1e-4 / 2.0) ** (ig / K) for ig in range(1, K + 1)]) if w["anneal"] else None
    for corrupt, field in ((None, None), ("index", "index_or_uniform_mismatch"), ("decision", "decision_mismatch_outside_band")):
        st = b.replay_traced_kblock(w, lambda **kw: FakeSampler(w, lik, 96, K, kw["seed"], Tt, corrupt), Z0, 96, K, 4242, 64, anneal_T=Tt)
        assert st["proposals"] == 64 * K * len(w["blocks"])
        if corrupt is None:
            # FP32 storage of x', lp' in the trace: within the probe's tolerances
            assert st["ok"], st
        else:
            assert not st["ok"] and st[field] > 0, st


def test_probe_replay_linreg_chunked_logdensity():
    b = _bench()
    rng = np.random.default_rng(0)
    n, p = 5000, 6
    X = np.ones((n, p), order="F"); X[:, 1:] = rng.normal(size=(n, p - 1))
    y = X @ (1 + rng.random(p)) + rng.normal(size=n)
    w = dict(kind="linreg", X=X, y=y, d=p, n=n, K=4, blocks=[list(range(p))], gamma=2.0, anneal=False)
    B = rng.normal(size=(7, p))
    np.testing.assert_allclose(b._linreg_logdensity_chunked(w, B, rows_per_chunk=999), orc.LinRegLik(X, y).logdensity(B), rtol=1e-12)
    w["bols"] = np.linalg.solve(X.T @ X, X.T @ y); w["post_sd"] = np.sqrt(np.diag(np.linalg.inv(X.T @ X)))
    Z0 = b.make_archive(w, 300)
    lik = orc.LinRegLik(X, y)
    st = b.replay_traced_kblock(w, lambda **kw: FakeSampler(w, lik, 64, 4, kw["seed"]), Z0, 64, 4, 7, 32)
    assert st["ok"], st
