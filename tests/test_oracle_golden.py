"""The oracle against its committed golden vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py):
the C restatement must reproduce them exactly, and the numpy restatement — an independent implementation of the same
reference lines — must reproduce the same trajectories generation by generation."""
import os

import numpy as np
import pytest

from oracle import demcz_oracle as orc
from tests.golden.make_golden import configs, run

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))


@pytest.mark.parametrize("name", ["normpdf", "anneal", "linreg"])
def test_c_oracle_reproduces_golden(coracle, name):
    res = run(configs()[name], coracle)
    for k, v in res.items():
        np.testing.assert_allclose(v, GOLD[f"{name}.{k}"], rtol=1e-12, atol=1e-14, err_msg=f"{name}.{k}")


@pytest.mark.parametrize("name", ["normpdf", "anneal", "linreg"])
def test_numpy_oracle_follows_golden(name):
    cfg = configs()[name]
    lik = orc.LinRegLik(cfg["X"], cfg["y"]) if cfg["kind"] == "linreg" else orc.MvNormalLik(cfg["mu"], cfg["Sigma"], drop_constant=cfg["kind"] == "quad")
    chain, log_obj, Zg = GOLD[f"{name}.chain"], GOLD[f"{name}.log_obj"], GOLD[f"{name}.Z"]
    N, K, G, M0 = cfg["N"], cfg["K"], cfg["G"], cfg["Z0"].shape[0]
    X = cfg["Z0"][M0 - N:].copy()
    lp = lik.logdensity(X)
    for g in range(1, G + 1):
        M = M0 + N * ((g - 1) // K)
        T = orc.tempbaseline(g, G, cfg["T0"], cfg["TN"]) if cfg["anneal"] else None
        X, lp, _ = orc.generation(X, lp, Zg, M, lik, cfg["blocks"], cfg["eps"], cfg["gamma"], cfg["seed"], g, T=T)
        np.testing.assert_allclose(X, chain[:, :, g - 1], rtol=1e-10, atol=1e-12, err_msg=f"generation {g}")
        np.testing.assert_allclose(lp, log_obj[:, g - 1], rtol=1e-9)
    np.testing.assert_allclose(orc.rhat_gelman(chain), GOLD[f"{name}.rhat"], rtol=1e-10)
