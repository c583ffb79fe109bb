"""Statistical gates (-m gpu).  First the reference's own assertions on its own test configurations
(test/example_normpdf.jl:49-51, test/test_anneal.jl:31 with the evident intent), then the
north star's: posterior mean / covariance equal to the analytic target within Monte-Carlo error
and Rhat < 1.01."""
import numpy as np
import pytest

import demc_jl_b200 as DEMC
from tests.parity import example_normpdf_target

pytestmark = pytest.mark.gpu


def test_example_normpdf_reference_assertions():
    # test/example_normpdf.jl:6-51 (N=5 chains, K=10, Z0 = randn(10d, d), 5000 + 5000 generations via prevrun)
    ndim = 5
    mu, S = example_normpdf_target(ndim, seed=31953150)
    logobj = DEMC.MvNormalLogpdf(mu, S)
    blockindex = [list(range(ndim))]
    eps_scale = 1e-5 * np.ones(ndim)
    N, K, Ngen = 5, 10, 5000
    Z = np.random.default_rng(31953150).normal(size=(10 * ndim, ndim))
    mc, Z = DEMC.demcz_sample(logobj, Z, N, K, 5000, 1, blockindex, eps_scale, 2.38, verbose=False)
    mc, Z = DEMC.demcz_sample(logobj, Z[-10 * ndim - 1:], N, K, Ngen, 1, blockindex, eps_scale, 2.38, prevrun=mc,
                              verbose=False)
    Ntot = mc.chain.shape[2]
    assert Ntot == 10000
    keep = slice(Ntot - Ngen // 2, Ntot)
    chain_burned = mc.chain[:, :, keep]
    logobj_burned = mc.log_obj[:, keep]
    accept_ratio, Rhat = DEMC.convergence_check(chain_burned, logobj_burned, verbose=False)
    assert np.all(Rhat < 1.1)
    assert np.all(accept_ratio > 0.1)
    assert np.all(accept_ratio < 0.45)
    b, Sb = DEMC.mean_cov_chain(chain_burned)
    sd = np.sqrt(np.diag(S))
    assert np.all(np.abs(b - mu) < 0.25 * sd)            # printed, not asserted, by the reference (:42)


def test_test_anneal_reference_configuration():
    # test/test_anneal.jl:7-31: quadratic objective, d=10, N=5, 5000 generations, T0=2 -> TN=1e-4
    ndim = 10
    mu = np.random.default_rng(31953150).random(ndim)
    Z = np.random.default_rng(1).normal(size=(10 * ndim, ndim))
    mc, Z = DEMC.demcz_anneal(DEMC.QuadraticLogobj(mu), Z, 5, 10, 5000, 1, [list(range(ndim))], 1e-5 * np.ones(ndim),
                              2.38, verbose=False, TN=1e-4, T0=2)
    bestval = mc.log_obj.max()
    assert bestval > -1e-1                                 # evident intent of `abs(bestval) > -1e-1` (Q8)
    i, g = np.unravel_index(np.argmax(mc.log_obj), mc.log_obj.shape)
    assert np.sum((mc.chain[i, :, g] - mu) ** 2) < 1e-1


def test_posterior_moments_and_rhat_many_chains():
    # north star: mean/cov match the analytic target within Monte-Carlo error, Rhat < 1.01
    d, N, G, W = 5, 1024, 6000, 4000
    mu, S = example_normpdf_target(d, seed=3)
    Z0 = np.random.default_rng(5).normal(size=(2 * N, d))
    mc, Z = DEMC.demcz_sample(DEMC.MvNormalLogpdf(mu, S), Z0, N, 10, G, 1, [list(range(d))], 1e-5 * np.ones(d), 2.38,
                              verbose=False)
    burned = mc.chain[:, :, G - W:]         # split-Rhat < 1.01 needs ~100 autocorrelation times per half-window
    Rhat = DEMC.Rhat_gelman(burned)
    assert np.all(Rhat < 1.01), Rhat
    b, C = DEMC.mean_cov_chain(burned)
    sd = np.sqrt(np.diag(S))
    assert np.all(np.abs(b - mu) < 0.02 * sd), (b - mu) / sd
    assert np.linalg.norm(C - S) / np.linalg.norm(S) < 0.03
    ar = np.mean(np.diff(mc.log_obj[:, G - W:], axis=1) != 0)
    assert 0.1 < ar < 0.45


def test_correlated_d20_autostop_rhat():
    # BASELINE configs[3] shape at reduced N: 20-dim equicorrelated (rho = 0.95) target, Rhat autostop at 1.01
    # (split-Rhat over a window needs ~100 autocorrelation times per half-chain to fall below 1.01: calibrated
    # with the CPU oracle, windows of 1000 / 2500 / 5000 generations plateau at 1.066 / 1.028 / 1.013)
    d, N = 20, 4096
    S = 0.95 * np.ones((d, d)) + 0.05 * np.eye(d)
    mu = np.linspace(-1, 1, d)
    Z0 = 3.0 * np.random.default_rng(9).normal(size=(2 * N, d))
    mc, Z = DEMC.demcz_sample(DEMC.MvNormalLogpdf(mu, S), Z0, N, 10, 60000, 1, [list(range(d))], 1e-5 * np.ones(d), 2.38,
                              verbose=False, autostop="Rhat", autostop_every=10000, autostop_Rhat=1.01, store_chain=False)
    assert mc.info["stopped_by_rhat"], mc.info
    assert mc.info["rhat_max"] < 1.01
    X = mc.Xcurrent
    assert np.all(np.abs(X.mean(axis=0) - mu) < 0.08)
    C = np.cov(X.T, bias=True)
    # one time slice of N = 4096 draws: the top eigenvalue (19.05) alone has a relative sampling s.d. of sqrt(2/N) = 2.2 %
    assert np.linalg.norm(C - S) / np.linalg.norm(S) < 0.08
