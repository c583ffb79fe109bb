"""Pins the oracle's arithmetic: MvNormal logpdf vs scipy, linreg vs the FP64 Gram form,
split-Rhat (C) vs the numpy transcription of src/utils.jl:2-20, and the C generation step vs
the numpy one (two independent restatements of src/demcz.jl:167-203)."""
import numpy as np
import pytest
import scipy.stats

from oracle import demcz_oracle as orc


def _mvn(d, seed=0):
    rng = np.random.default_rng(seed)
    mu = rng.random(d)
    A = rng.random((d, d))
    S = A.T @ A + 2 * np.eye(d)                      # test/example_normpdf.jl:9-12
    S = S / S.max() / 100
    return mu, S


@pytest.mark.parametrize("d", [1, 5, 10, 20])
def test_mvnormal_vs_scipy(coracle, d):
    mu, S = _mvn(d)
    lik = orc.MvNormalLik(mu, S)
    X = np.random.default_rng(2).normal(size=(50, d)) * 0.2 + mu
    ref = scipy.stats.multivariate_normal(mu, S).logpdf(X)
    np.testing.assert_allclose(lik.logdensity(X), ref, rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(coracle.logdensity(lik, X), ref, rtol=1e-11, atol=1e-9)


def test_quadratic_is_mvnormal_half_identity(coracle):
    mu = np.random.default_rng(3).random(10)
    lik = orc.MvNormalLik.quadratic(mu)
    X = np.random.default_rng(4).normal(size=(20, 10))
    ref = -np.sum((X - mu) ** 2, axis=1)             # test/test_anneal.jl:10
    np.testing.assert_allclose(lik.logdensity(X), ref, rtol=1e-12)
    np.testing.assert_allclose(coracle.logdensity(lik, X), ref, rtol=1e-12)


def test_linreg_vs_gram_form(coracle):
    rng = np.random.default_rng(5)
    n, p = 2000, 8
    X = rng.normal(size=(n, p)); X[:, 0] = 1
    beta = 1 + 3 * rng.random(p)
    y = X @ beta + rng.normal(size=n)
    lik = orc.LinRegLik(X, y)
    B = beta + 0.05 * rng.normal(size=(10, p))
    gram = -0.5 * (y @ y - 2 * B @ (X.T @ y) + np.einsum("ci,ij,cj->c", B, X.T @ X, B))
    np.testing.assert_allclose(lik.logdensity(B), gram, rtol=1e-10)
    np.testing.assert_allclose(coracle.logdensity(lik, B), gram, rtol=1e-10)


def test_rhat_c_vs_numpy(coracle):
    rng = np.random.default_rng(6)
    for (N, d, G) in [(5, 5, 100), (4, 3, 51), (100, 2, 20)]:
        chain = rng.normal(size=(N, d, G)) + rng.normal(size=(N, d, 1)) * 0.3
        np.testing.assert_allclose(coracle.rhat(chain), orc.rhat_gelman(chain), rtol=1e-12)
    # iid chains -> Rhat ~ 1
    assert np.all(np.abs(orc.rhat_gelman(rng.normal(size=(50, 3, 2000))) - 1) < 0.01)


def test_tempbaseline(coracle):
    for ig in (1, 500, 5000):
        assert coracle.tempbaseline(ig, 5000, 2.0, 1e-4) == pytest.approx(2.0 * (1e-4 / 2.0) ** (ig / 5000), rel=1e-14)
    assert coracle.tempbaseline(10, 100, 5.0, 0.0) == 0.0      # Q4: TN = 0 -> T = 0


@pytest.mark.parametrize("blocks,T", [([list(range(5))], None), ([[0, 1, 2], [3], [4]], None),
                                      ([list(range(5))], 0.37), ([[4, 2], [0, 1, 3]], 0.0)])
def test_generation_c_vs_numpy(coracle, blocks, T):
    d, N, M = 5, 64, 50
    mu, S = _mvn(d, 1)
    lik = orc.MvNormalLik(mu, S)
    rng = np.random.default_rng(8)
    Z = rng.normal(size=(M, d))
    X = mu + 0.1 * rng.normal(size=(N, d))
    lp = lik.logdensity(X)
    eps = 1e-3 * (1 + np.arange(d))
    for gen in (1, 2, 2**33 + 5):
        Xn, lpn, tr = orc.generation(X, lp, Z, M, lik, blocks, eps, 2.38, 31953150, gen, T=T, chain_offset=7)
        Xc, lpc, trc, _, _ = coracle.generation(X, lp, Z, M, lik, blocks, eps, 2.38, 31953150, gen, T=T, chain_offset=7)
        assert np.array_equal(tr["r1"], trc["r1"]) and np.array_equal(tr["r2"], trc["r2"])
        assert np.array_equal(tr["u"], trc["u"])
        np.testing.assert_allclose(tr["xprop"], trc["xprop"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(tr["lp_prop"], trc["lp_prop"], rtol=1e-10)
        assert np.array_equal(tr["accepted"], trc["accepted"])
        np.testing.assert_allclose(Xn, Xc, rtol=1e-13)
        np.testing.assert_allclose(lpn, lpc, rtol=1e-10)
        assert np.all(tr["r1"] != tr["r2"])


def test_faithful_index_mode_same_draws(coracle):
    d, N = 3, 8
    lik = orc.MvNormalLik(np.zeros(d), np.eye(d))
    Z0 = np.random.default_rng(9).normal(size=(30, d))
    a = coracle.run(lik, Z0, N, 5, 40, [list(range(d))], 1e-4 * np.ones(d), 2.38, 11, index_mode=0)
    b = coracle.run(lik, Z0, N, 5, 40, [list(range(d))], 1e-4 * np.ones(d), 2.38, 11, index_mode=1)
    assert np.array_equal(a["chain"], b["chain"]) and np.array_equal(a["Z"], b["Z"])
    assert a["M"] == 30 + N * 40 // 5


def test_run_threads_bit_identical(coracle):
    d, N = 4, 33
    mu, S = _mvn(d, 2)
    lik = orc.MvNormalLik(mu, S)
    Z0 = np.random.default_rng(10).normal(size=(60, d))
    a = coracle.run(lik, Z0, N, 10, 50, [list(range(d))], 1e-5 * np.ones(d), 2.38, 5, nthreads=1)
    b = coracle.run(lik, Z0, N, 10, 50, [list(range(d))], 1e-5 * np.ones(d), 2.38, 5, nthreads=4)
    c = coracle.run(lik, Z0, N, 10, 50, [list(range(d))], 1e-5 * np.ones(d), 2.38, 5, z_rowmajor=1)
    assert np.array_equal(a["chain"], b["chain"]) and np.array_equal(a["chain"], c["chain"])
    assert np.array_equal(a["Z"], c["Z"])
