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


# ---- round-2 extensions: snooker update, logistic likelihood (tests/golden/extensions_golden.npz) ---------------------
from tests.golden.make_golden_ext import cases, run_numpy  # noqa: E402

GOLD_EXT = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extensions_golden.npz"))


@pytest.mark.parametrize("name", ["mvn", "logistic"])
def test_snooker_golden_both_restatements(coracle, name):
    c = cases()[name]
    got = run_numpy(c)
    for k, v in got.items():
        if v.dtype.kind in "iub":
            assert np.array_equal(v, GOLD_EXT[f"{name}.{k}"]), k
        else:
            np.testing.assert_allclose(v, GOLD_EXT[f"{name}.{k}"], rtol=1e-12, atol=1e-14, err_msg=k)
    lp = c["lik"].logdensity(c["X"])
    _, _, tr = coracle.snooker_generation(c["X"], lp, c["Z"], c["M"], c["lik"], c["gamma_s"], c["seed"], c["gen"], T=c["T"],
                                          chain_offset=c["chain_offset"])
    for k in ("r1", "r2", "accepted"):
        assert np.array_equal(tr[k], GOLD_EXT[f"{name}.{k}"]), k
    for k in ("u", "xprop", "lp_prop"):
        np.testing.assert_allclose(tr[k], GOLD_EXT[f"{name}.{k}"], rtol=1e-11, atol=1e-13, err_msg=k)
    if "B" in c:
        np.testing.assert_allclose(coracle.logdensity(c["lik"], c["B"]), GOLD_EXT[f"{name}.logdensity"], rtol=1e-12)
