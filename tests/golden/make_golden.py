"""Generates tests/golden/oracle_golden.npz: outputs of the CPU oracle (C restatement, oracle/demcz_oracle.c) on three
small seeded configurations.  The reference itself (Julia) cannot run in this image and holds no golden vectors, so
these fixtures pin the ORACLE (and through it the documented Philox stream layout, the proposal / accept arithmetic and
the R-hat formula) against accidental change between rounds; tests/test_oracle_golden.py replays them with both the C
and the numpy restatement.  Re-run only when the stream layout is changed on purpose:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import demcz_oracle as orc  # noqa: E402


def configs():
    rng = np.random.default_rng(20261018)
    out = {}
    # (a) test/example_normpdf.jl shape: d=5 MvNormal, one block
    d = 5
    mu = rng.random(d); A = rng.random((d, d)); S = A.T @ A + 2 * np.eye(d); S = S / S.max() / 100
    out["normpdf"] = dict(kind="mvn", mu=mu, Sigma=S, Z0=rng.normal(size=(30, d)), N=6, K=4, G=17, blocks=[list(range(d))],
                          eps=1e-5 * np.ones(d), gamma=2.38, seed=31953150, anneal=False)
    # (b) test/test_anneal.jl shape: quadratic objective, two blocks, tempered, blocklen-1 block
    d = 4
    mu = rng.random(d)
    out["anneal"] = dict(kind="quad", mu=mu, Sigma=0.5 * np.eye(d), Z0=rng.normal(size=(20, d)), N=5, K=3, G=12,
                         blocks=[[0, 1, 2], [3]], eps=1e-3 * np.ones(d), gamma=2.0, seed=7, anneal=True, T0=2.0, TN=1e-2)
    # (c) test/example_linreg.jl shape
    n, p = 200, 3
    X = np.ones((n, p)); X[:, 1:] = rng.normal(size=(n, p - 1)); beta = 1 + 3 * rng.random(p); y = X @ beta + rng.normal(size=n)
    out["linreg"] = dict(kind="linreg", X=X, y=y, Z0=beta + 0.1 * rng.normal(size=(16, p)), N=4, K=5, G=10, blocks=[list(range(p))],
                         eps=1e-5 * np.ones(p), gamma=2.0, seed=11, anneal=False)
    return out


def run(cfg, co):
    if cfg["kind"] == "linreg":
        lik = orc.LinRegLik(cfg["X"], cfg["y"])
    else:
        lik = orc.MvNormalLik(cfg["mu"], cfg["Sigma"], drop_constant=cfg["kind"] == "quad")
    kw = {}
    if cfg["anneal"]:
        kw = dict(anneal=True, Ttable=np.array([orc.tempbaseline(ig, cfg["G"], cfg["T0"], cfg["TN"]) for ig in range(1, cfg["G"] + 1)]))
    r = co.run(lik, cfg["Z0"], cfg["N"], cfg["K"], cfg["G"], cfg["blocks"], cfg["eps"], cfg["gamma"], cfg["seed"], **kw)
    return dict(chain=r["chain"], log_obj=r["log_obj"], Z=r["Z"], rhat=co.rhat(np.asfortranarray(r["chain"])))


if __name__ == "__main__":
    co = orc.COracle(build=True)
    blob = {}
    for name, cfg in configs().items():
        res = run(cfg, co)
        for k, v in res.items():
            blob[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **blob)
    print({k: v.shape for k, v in blob.items()})
