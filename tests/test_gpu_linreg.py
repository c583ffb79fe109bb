"""GPU parity tests of the linear-regression likelihood path (-m gpu): the tcgen05 3xTF32 residual GEMM fused
with the sum of squares, against the FP64 oracle (test/example_linreg.jl:32 upstream: -0.5*sum((y .- X*b).^2))."""
import numpy as np
import pytest

import demc_jl_b200 as DEMC
from oracle import demcz_oracle as orc
from tests.parity import check_run_against_oracle

pytestmark = pytest.mark.gpu


def example_linreg_data(nobs, npar, seed=319531501):
    """test/example_linreg.jl:9-29: intercept + npar equicorrelated (0.25) regressors, beta = 1 + 3 U, unit noise."""
    rng = np.random.default_rng(seed)
    z0 = rng.normal(size=(nobs, 1))
    X = np.ones((nobs, npar + 1))
    X[:, 1:] = np.sqrt(0.25) * z0 + np.sqrt(0.75) * rng.normal(size=(nobs, npar))
    beta = 1 + 3 * rng.random(npar + 1)
    y = X @ beta + rng.normal(size=nobs)
    return X, y, beta


def _sampler(lo, N=8, G=1, **kw):
    d = lo.d
    return DEMC.Sampler(lo, d, N=N, K=10, Ngeneration=G, blockindex=[list(range(d))], eps_scale=1e-5 * np.ones(d), γ=2.0, **kw)


@pytest.mark.parametrize("nobs,npar", [(1000, 25), (777, 4), (5000, 63), (100000, 31), (256, 7), (300, 39)])
def test_eval_logdensity_vs_fp64(nobs, npar):
    X, y, beta = example_linreg_data(nobs, npar)
    lo = DEMC.LinRegLoglik(X, y)
    ref = orc.LinRegLik(X, y)
    rng = np.random.default_rng(1)
    bols = np.linalg.solve(X.T @ X, X.T @ y)
    B = np.concatenate([bols + 0.03 * rng.normal(size=(150, npar + 1)),      # posterior scale
                        bols[None, :],
                        rng.normal(size=(149, npar + 1))])                   # far away (the reference's Z0 = randn)
    with _sampler(lo) as s:
        got = s.eval_logdensity(B)
    want = ref.logdensity(B)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() < 1e-5, rel.max()                   # north-star tolerance
    assert rel.max() < 2e-6, rel.max()                   # what 3xTF32 + the centred form actually deliver
    # differences of log-objectives near the mode (what accept decisions see) must be accurate in absolute terms
    near = slice(0, 151)
    assert np.max(np.abs((got[near] - got[150]) - (want[near] - want[150]))) < 2e-3 * max(1.0, nobs / 1e4)


def test_more_points_than_chains_and_single_row_tail():
    X, y, _ = example_linreg_data(513, 9)                # 513 rows: a 1-row tail tile
    lo = DEMC.LinRegLoglik(X, y)
    B = np.random.default_rng(2).normal(size=(1000, 10))  # 1000 points through a context of 8 chains
    with _sampler(lo) as s:
        got = s.eval_logdensity(B)
    np.testing.assert_allclose(got, orc.LinRegLik(X, y).logdensity(B), rtol=2e-6)


@pytest.mark.parametrize("nobs,npar,N,blocks", [
    (1000, 25, 64, None),                                 # test/example_linreg.jl shape
    (2000, 9, 100, [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9]]),   # two blocks, ragged N
    (600, 39, 130, None),                                 # p = 40 > 32: two 128-byte k-blocks per row
])
def test_linreg_stepwise(nobs, npar, N, blocks):
    X, y, beta = example_linreg_data(nobs, npar)
    d = npar + 1
    bols = np.linalg.solve(X.T @ X, X.T @ y)
    Z0 = bols + 0.05 * np.random.default_rng(3).normal(size=(3 * N, d))
    st = check_run_against_oracle(DEMC.LinRegLoglik(X, y), Z0, N=N, K=5, G=15, blocks=blocks or [list(range(d))],
                                  eps=1e-5 * np.ones(d), gamma=2.0, seed=11)
    assert st["accepted"] > 0 and st["max_lp_relerr"] < 2e-6


def test_linreg_anneal_stepwise():
    X, y, _ = example_linreg_data(800, 5)
    d = 6
    Z0 = np.random.default_rng(4).normal(size=(200, d))
    check_run_against_oracle(DEMC.LinRegLoglik(X, y), Z0, N=48, K=4, G=12, blocks=[list(range(d))],
                             eps=1e-5 * np.ones(d), gamma=2.0, seed=5, anneal=True, T0=2.0, TN=1e-3)


def test_linreg_runs_are_reproducible():
    X, y, _ = example_linreg_data(3000, 15)
    d = 16
    lo = DEMC.LinRegLoglik(X, y)
    Z0 = np.random.default_rng(6).normal(size=(100, d))
    args = (32, 10, 30, 1, [list(range(d))], 1e-5 * np.ones(d), 2.0)
    a, Za = DEMC.demcz_sample(lo, Z0, *args, verbose=False)
    b, Zb = DEMC.demcz_sample(lo, Z0, *args, verbose=False)
    np.testing.assert_array_equal(a.chain, b.chain)       # fixed tile schedule + ordered partial sums: bit-reproducible
    np.testing.assert_array_equal(a.log_obj, b.log_obj)
    np.testing.assert_array_equal(Za, Zb)


def test_example_linreg_posterior_matches_ols():
    # test/example_linreg.jl:36-76 — the reference prints b .- bols and asserts nothing; the analytic posterior is
    # N(bols, (X'X)^-1) since sigma^2 = 1
    nobs, npar = 1000, 25
    X, y, _ = example_linreg_data(nobs, npar)
    ndim = npar + 1
    opts = DEMC.demcopt(ndim)
    opts.blockindex = [list(range(ndim))]
    opts.Nblocks = 1
    opts.eps_scale = 1e-5 * np.ones(ndim)
    opts.γ = 2.0
    opts.N = 128
    opts.K = 10
    opts.Ngeneration = 12000
    opts.autostop = "no"
    opts.verbose = False
    Z = np.random.default_rng(7).normal(size=(2 * opts.N, ndim))
    mc, Z = DEMC.demcz_sample(DEMC.LinRegLoglik(X, y), Z, opts)
    W = 6000          # d = 26 at ~10 % acceptance mixes slowly: split-Rhat over a short window sits above 1.1
    keep = mc.chain[:, :, -W:]
    b, Sb = DEMC.mean_cov_chain(keep)
    bols = np.linalg.solve(X.T @ X, X.T @ y)
    S = np.linalg.inv(X.T @ X)
    sd = np.sqrt(np.diag(S))
    assert np.all(np.abs(b - bols) < 0.08 * sd), (b - bols) / sd
    assert np.linalg.norm(Sb - S) / np.linalg.norm(S) < 0.15
    accept_ratio, Rhat = DEMC.convergence_check(keep, mc.log_obj[:, -W:], verbose=False)
    assert np.all(Rhat < 1.1)                             # the reference's autostop threshold in this example
    assert 0.05 < accept_ratio.mean() < 0.5


def test_linreg_streaming_rhat_and_par_entry_point():
    # autostop R-hat from the streaming accumulators of the linreg accept kernel vs Rhat_gelman (src/utils.jl:2-20) on the
    # stored window; driven through the reference's parallel entry point demcz_sample_par (sync_every = R-hat window)
    X, y, _ = example_linreg_data(1500, 7)
    d, N, every = 8, 96, 60
    bols = np.linalg.solve(X.T @ X, X.T @ y)
    Z0 = bols + 0.05 * np.random.default_rng(8).normal(size=(4 * N, d))
    mc, Z = DEMC.demcz_sample_par(DEMC.LinRegLoglik(X, y), Z0, N, 10, 2 * every, 1, [list(range(d))], 1e-5 * np.ones(d), 2.0,
                                  sync_every=every, autostop="Rhat", autostop_Rhat=0.0, verbose=False)
    assert mc.chain.shape == (N, d, 2 * every) and mc.info["rhat_valid"]
    ref = orc.rhat_gelman(mc.chain[:, :, every:])
    np.testing.assert_allclose(mc.info["rhat"], ref, rtol=5e-4)


# ---- the 2-way FP16 split on tcgen05 kind::f16 (default) and the 3xTF32 kernel it falls back to ---------------------
def _ill_scaled(nobs, p, rng):
    """Regressor scales spread over 6 decades, heavy-tailed rows (10^4 x outliers), intercept."""
    X = rng.normal(size=(nobs, p))
    X[:, 0] = 1.0
    X *= 10.0 ** np.linspace(-3, 3, p)[None, :]
    out = rng.choice(nobs, size=max(nobs // 200, 3), replace=False)
    X[out, 1:] *= 1e4
    beta = (1 + rng.random(p)) / 10.0 ** np.linspace(-3, 3, p)
    y = X @ beta + rng.normal(size=nobs)
    return X, y


@pytest.mark.parametrize("kind", ["f16", "tf32"])
@pytest.mark.parametrize("nobs,p", [(4000, 12), (3000, 40), (2500, 64)])
def test_split_kinds_on_ill_scaled_regressors(kind, nobs, p, monkeypatch):
    """North-star accuracy on data chosen to hurt an FP16 / TF32 operand split: log-objective within 1e-5 relative of
    FP64 for points at posterior scale, at the mode and far away; differences near the mode accurate in absolute terms;
    a traced run replays with exact decisions outside the tolerance band."""
    monkeypatch.setenv("DEMCZ_LINREG_KIND", kind)
    rng = np.random.default_rng(100 + p)
    X, y = _ill_scaled(nobs, p, rng)
    lo, ref = DEMC.LinRegLoglik(X, y), orc.LinRegLik(X, y)
    bols = np.linalg.lstsq(X, y, rcond=None)[0]
    S = np.linalg.inv(X.T @ X)
    sd = np.sqrt(np.diag(S))
    Lp = np.linalg.cholesky(S)
    B = np.concatenate([bols + rng.normal(size=(200, p)) @ Lp.T,             # posterior draws (correlated: intercept vs near-constant column)
                        bols[None, :],
                        bols + sd * rng.normal(size=(55, p)),                 # independent moves: far outside the posterior
                        bols * (1 + 0.5 * rng.normal(size=(44, p)))])         # very far
    # the state is FP32 on the device: compare at the points the GPU actually evaluates (an ill-conditioned posterior moves
    # by more than the tolerance under a 6e-8 relative perturbation of b)
    B = B.astype(np.float32).astype(np.float64)
    with _sampler(lo) as s:
        assert s.kernel_info()["linreg_kind"] == kind
        got = s.eval_logdensity(B)
    want = ref.logdensity(B)
    rel = np.abs(got - want) / np.abs(want)
    worst = int(np.argmax(rel))
    assert rel.max() < 1e-5, (rel.max(), worst, got[worst], want[worst], [float(rel[a:b].max()) for a, b in ((0, 200), (200, 201), (201, 256), (256, 300))])
    # Differences near the mode.  Yardstick: the same centred form in plain float32 numpy (sgemm).  A 3-term operand split
    # (2 x FP16 or 3 x TF32) carries 22 significant bits per operand against float32's 24 and drops the lo . lo term, so
    # its products are ~2^-22 accurate where sgemm's are exact: allow 8x the sgemm error (or 5e-3 absolute).
    near = slice(0, 201)
    dgot, dwant = got[near] - got[200], want[near] - want[200]
    X32, D32 = X.astype(np.float32), (B[near] - bols).astype(np.float32)
    q32 = np.sum((X32 @ D32.T).astype(np.float32) ** 2, axis=0, dtype=np.float32).astype(np.float64)
    q64 = np.sum((X @ (B[near] - bols).T) ** 2, axis=0)
    yard = np.max(np.abs(0.5 * (q32 - q64)))
    assert np.max(np.abs(dgot - dwant)) < max(5e-3, 8.0 * yard), (np.max(np.abs(dgot - dwant)), yard)
    N = 96
    Z0 = bols + 2.0 * rng.normal(size=(3 * N, p)) @ Lp.T           # over-dispersed along the posterior's own geometry
    st = check_run_against_oracle(lo, Z0, N=N, K=4, G=8, blocks=[list(range(p))], eps=1e-5 * sd, gamma=2.0, seed=3)
    assert st["max_lp_relerr"] < 1e-5


def test_fp16_split_equals_tf32_split_decisions():
    """Both kernels run the same chain: with decisions exact outside the band and log-objectives equal to ~1e-6 relative,
    a short free-running chain on well-separated proposals takes the same accept decisions under either MMA kind."""
    import os
    X, y, _ = example_linreg_data(2000, 15)
    d = 16
    bols = np.linalg.solve(X.T @ X, X.T @ y)
    Z0 = bols + 0.05 * np.random.default_rng(9).normal(size=(200, d))
    res = {}
    for kind in ("f16", "tf32"):
        os.environ["DEMCZ_LINREG_KIND"] = kind
        try:
            mc, _ = DEMC.demcz_sample(DEMC.LinRegLoglik(X, y), Z0, 64, 10, 20, 1, [list(range(d))], 1e-5 * np.ones(d), 2.0, verbose=False)
        finally:
            os.environ.pop("DEMCZ_LINREG_KIND", None)
        res[kind] = mc
    np.testing.assert_allclose(res["f16"].log_obj[:, 0], res["tf32"].log_obj[:, 0], rtol=2e-6)
    agree = np.mean(np.all(res["f16"].chain[:, :, -1] == res["tf32"].chain[:, :, -1], axis=1))
    assert agree > 0.9                                      # a near-tie may flip a chain; most chains are bit-identical


# ---- Bernoulli-logit GLM on the same tcgen05 contraction (softplus epilogue) ------------------------------------------
def _logit_data(n, p, rng):
    X = np.ones((n, p)); X[:, 1:] = rng.normal(size=(n, p - 1))
    beta = rng.normal(size=p) / np.sqrt(p)
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-X @ beta))).astype(np.float64)
    return X, y, beta


@pytest.mark.parametrize("n,p", [(1000, 8), (777, 3), (5000, 40), (20000, 64), (193, 5)])
def test_logistic_eval_vs_fp64(n, p):
    rng = np.random.default_rng(n + p)
    X, y, beta = _logit_data(n, p, rng)
    ref = orc.LogisticLik(X, y)
    B = np.concatenate([beta + 0.1 * rng.normal(size=(100, p)), np.zeros((1, p)), 3.0 * rng.normal(size=(60, p))])   # near, zero, far (saturated)
    B = B.astype(np.float32).astype(np.float64)
    with _sampler(DEMC.LogisticLoglik(X, y)) as s:
        assert s.kernel_info()["linreg_link"] == "logit"
        got = s.eval_logdensity(B)
    want = ref.logdensity(B)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() < 1e-5, (rel.max(), int(np.argmax(rel)))
    assert rel[:101].max() < 2e-6, rel[:101].max()        # MUFU ex2 / lg2: ~2e-7 relative per row, unbiased over the rows


@pytest.mark.parametrize("n,p,N,blocks", [(1200, 8, 96, None), (900, 10, 70, [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9]])])
def test_logistic_stepwise(n, p, N, blocks):
    rng = np.random.default_rng(7)
    X, y, beta = _logit_data(n, p, rng)
    Z0 = beta + 0.2 * rng.normal(size=(3 * N, p))
    st = check_run_against_oracle(DEMC.LogisticLoglik(X, y), Z0, N=N, K=5, G=15, blocks=blocks or [list(range(p))],
                                  eps=1e-5 * np.ones(p), gamma=2.0, seed=13, snooker_prob=0.2)
    assert st["accepted"] > 0 and st["max_lp_relerr"] < 2e-6


def test_logistic_posterior_is_centred_on_the_mle():
    rng = np.random.default_rng(17)
    n, p, N = 4000, 6, 128
    X, y, beta = _logit_data(n, p, rng)
    b = np.zeros(p)                                        # Newton iterations for the MLE and the observed information
    for _ in range(30):
        mu_ = 1.0 / (1.0 + np.exp(-X @ b))
        H = X.T @ (X * (mu_ * (1 - mu_))[:, None])
        b = b + np.linalg.solve(H, X.T @ (y - mu_))
    Sinv = np.linalg.inv(H)
    Z0 = b + 2.0 * np.sqrt(np.diag(Sinv)) * rng.normal(size=(3 * N, p))
    mc, _ = DEMC.demcz_sample(DEMC.LogisticLoglik(X, y), Z0, N, 10, 3000, 1, [list(range(p))], 1e-5 * np.ones(p), 2.38 , verbose=False)
    m, C = DEMC.mean_cov_chain(mc.chain[:, :, 1500:])
    sd = np.sqrt(np.diag(Sinv))
    assert np.all(np.abs(m - b) < 0.15 * sd), (m - b) / sd            # Laplace approximation: mean ~ MLE, cov ~ H^-1
    assert np.linalg.norm(C - Sinv) / np.linalg.norm(Sinv) < 0.15
