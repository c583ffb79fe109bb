"""Generates tests/golden/extensions_golden.npz: the oracle's output for the round-2 extensions — one snooker generation
(ter Braak & Vrugt 2008; indices, uniform, proposal, log-density, decision) on a MvNormal and on a logistic target, and
logistic log-likelihood values.  Like oracle_golden.npz these fixtures pin the ORACLE (the Philox stream layout of the
snooker draws in particular) against accidental change; tests/test_oracle_golden.py replays them with the C and the numpy
restatement.  Re-run only when the stream layout is changed on purpose:  python tests/golden/make_golden_ext.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import demcz_oracle as orc  # noqa: E402


def cases():
    rng = np.random.default_rng(20261019)
    d, N, M = 4, 12, 40
    mv = orc.MvNormalLik(rng.random(d), np.eye(d) * 0.5 + 0.2)
    n, p = 150, 3
    X = np.asfortranarray(np.column_stack([np.ones(n), rng.normal(size=(n, p - 1))]))
    y = (rng.random(n) < 0.45).astype(float)
    lg = orc.LogisticLik(X, y)
    out = {}
    for name, lik, dd in (("mvn", mv, d), ("logistic", lg, p)):
        Z = rng.normal(size=(M, dd)) * 0.4
        Xs = rng.normal(size=(N, dd)) * 0.4
        out[name] = dict(lik=lik, Z=Z, X=Xs, M=M, seed=31953150, gen=77, gamma_s=0.0 if name == "mvn" else 1.7, T=None if name == "mvn" else 0.7,
                         chain_offset=5)
    out["logistic"]["B"] = rng.normal(size=(5, p)) * np.array([1.0, 4.0, 0.2])
    return out


def run_numpy(c):
    lp = c["lik"].logdensity(c["X"])
    o = orc.snooker_update(c["X"], lp, c["Z"], c["M"], c["lik"], c["gamma_s"], c["seed"], c["gen"], T=c["T"], chain_offset=c["chain_offset"])
    return {k: np.asarray(o[k]) for k in ("r1", "r2", "u", "xprop", "lp_prop", "accepted")}


if __name__ == "__main__":
    blob = {}
    for name, c in cases().items():
        for k, v in run_numpy(c).items():
            blob[f"{name}.{k}"] = v
        if "B" in c:
            blob[f"{name}.logdensity"] = c["lik"].logdensity(c["B"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "extensions_golden.npz"), **blob)
    print({k: v.shape for k, v in blob.items()})
