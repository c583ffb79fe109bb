"""CPU evidence for the one semantic deviation of the GPU path from the reference: the reference appends
chain ic's row INSIDE the chain loop (src/demcz.jl:31-33, 88-91), so later chains of the same generation can already
draw it; the chain-parallel path appends after the generation (generation-synchronous).  Both are restated by the oracle
(append_mode); this test shows they sample the same posterior: identical first K-1 generations, and matching moments,
acceptance band and split-Rhat afterwards (the reference's own statistical assertions, test/example_normpdf.jl:49-51)."""
import numpy as np

from oracle import demcz_oracle as orc


def _target(d=4, seed=3):
    rng = np.random.default_rng(seed)
    mu = rng.random(d)
    A = rng.random((d, d))
    S = A.T @ A + 2 * np.eye(d)
    return mu, S / S.max() / 100


def test_inloop_and_gen_sync_append_sample_the_same_posterior(coracle):
    d, N, K, G = 4, 24, 10, 3000
    mu, S = _target(d)
    lik = orc.MvNormalLik(mu, S)
    Z0 = np.random.default_rng(0).normal(size=(10 * d + N, d))
    kw = dict(store_chain=True)
    a = coracle.run(lik, Z0, N, K, G, [list(range(d))], 1e-5 * np.ones(d), 2.38, 31953150, append_mode=0, **kw)   # gen-sync
    b = coracle.run(lik, Z0, N, K, G, [list(range(d))], 1e-5 * np.ones(d), 2.38, 31953150, append_mode=1, **kw)   # reference in-loop
    # until the first append both are the same process with the same Philox stream
    np.testing.assert_array_equal(a["chain"][:, :, :K - 1], b["chain"][:, :, :K - 1])   # generation K itself already differs
    assert a["M"] == b["M"] == Z0.shape[0] + N * (G // K)
    sd = np.sqrt(np.diag(S))
    for r in (a, b):
        burned = r["chain"][:, :, G // 2:]
        m, C = orc.mean_cov_chain(burned)
        assert np.all(np.abs(m - mu) < 0.1 * sd)
        assert np.linalg.norm(C - S) / np.linalg.norm(S) < 0.15
        assert np.all(orc.rhat_gelman(burned) < 1.1)
        ar = orc.accept_ratio(r["log_obj"][:, G // 2:])
        assert np.all(ar > 0.1) and np.all(ar < 0.45)
    ma, Ca = orc.mean_cov_chain(a["chain"][:, :, G // 2:])
    mb, Cb = orc.mean_cov_chain(b["chain"][:, :, G // 2:])
    assert np.all(np.abs(ma - mb) < 0.1 * sd)
    assert np.linalg.norm(Ca - Cb) / np.linalg.norm(S) < 0.15
