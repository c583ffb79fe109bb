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


def test_snooker_update_draws_distinct_rows_and_keeps_the_target_invariant():
    """Restatement of the snooker update of ter Braak & Vrugt (2008) (cited at README.md:57-59 upstream, no reference code):
    z, z1, z2 are distinct; with the (d-1) log|x'-z|/|x-z| Jacobian term the update leaves N(0, I) invariant."""
    import numpy as np
    from oracle import demcz_oracle as orc
    rng = np.random.default_rng(0)
    d, M, N = 3, 1500, 400
    lik = orc.MvNormalLik(np.zeros(d), np.eye(d))
    Z = rng.normal(size=(M, d))
    X = rng.normal(size=(N, d))
    lp = lik.logdensity(X)
    xs, acc = [], 0.0
    for g in range(1, 301):
        o = orc.snooker_update(X, lp, Z, M, lik, 0.0, 11, g)
        assert np.all(o["z"] != o["r1"]) and np.all(o["z"] != o["r2"]) and np.all(o["r1"] != o["r2"])
        assert o["r1"].max() < M and o["r2"].max() < M and o["z"].max() < M
        a = o["accepted"]
        X[a], lp[a] = o["xprop"][a], o["lp_prop"][a]
        acc += a.mean()
        if g > 100:
            xs.append(X.copy())
    xs = np.concatenate(xs)
    assert 0.2 < acc / 300 < 0.7
    assert np.all(np.abs(xs.mean(0)) < 0.05) and np.all(np.abs(xs.var(0) - 1.0) < 0.05)
    # exhaustive check of the index mapping on a tiny archive: the ordered triple is uniform over distinct triples
    cnt = {}
    for g in range(1, 400):
        o = orc.snooker_update(X[:200], lp[:200], Z, 4, lik, 1.7, 5, g)
        for t in zip(o["z"], o["r1"], o["r2"]):
            cnt[t] = cnt.get(t, 0) + 1
    assert len(cnt) == 24 and min(cnt.values()) > 0.8 * 399 * 200 / 24 and max(cnt.values()) < 1.2 * 399 * 200 / 24


def test_snooker_generation_choice_has_the_asked_frequency():
    from oracle import demcz_oracle as orc
    hits = sum(orc.is_snooker_gen(31953150, g, 0.1) for g in range(1, 20001))
    assert 1800 < hits < 2200
    assert not any(orc.is_snooker_gen(1, g, 0.0) for g in range(1, 100))


def test_logistic_lik_matches_a_direct_evaluation():
    import numpy as np
    from oracle import demcz_oracle as orc
    rng = np.random.default_rng(1)
    X = rng.normal(size=(50, 3)); y = (rng.random(50) < 0.5).astype(float)
    B = rng.normal(size=(4, 3))
    want = [np.sum(y * (X @ b) - np.log1p(np.exp(X @ b))) for b in B]
    np.testing.assert_allclose(orc.LogisticLik(X, y).logdensity(B), want, rtol=1e-12)


def test_c_and_numpy_restatements_agree_on_snooker_and_logistic(coracle):
    """The two independent restatements (C, numpy) of the round-2 extensions against each other: indices and uniforms
    bit-exact, proposals / log-densities to FP64 rounding, decisions identical; generation-level choice identical."""
    import numpy as np
    from oracle import demcz_oracle as orc
    rng = np.random.default_rng(2)
    # logistic log-likelihood
    X = np.asfortranarray(np.column_stack([np.ones(300), rng.normal(size=(300, 4))])); y = (rng.random(300) < 0.4).astype(float)
    lik = orc.LogisticLik(X, y)
    B = rng.normal(size=(6, 5)) * np.array([1, 1, 5, 0.1, 30])[None, :]          # incl. saturated eta
    np.testing.assert_allclose(coracle.logdensity(lik, B), lik.logdensity(B), rtol=1e-12)
    # snooker generation, MvNormal and logistic targets, fixed and random jump size, plain and tempered
    d, N, M = 5, 64, 200
    mv = orc.MvNormalLik(rng.normal(size=d), np.eye(d) + 0.3)
    for lk, dd in ((mv, d), (lik, 5)):
        Z = rng.normal(size=(M, dd)) * 0.3
        Xs = rng.normal(size=(N, dd)) * 0.3
        lp = lk.logdensity(Xs)
        for gs, T in ((0.0, None), (1.7, None), (0.0, 0.5)):
            a = orc.snooker_update(Xs, lp, Z, M, lk, gs, 99, 12345, T=T, chain_offset=7)
            Xc, lpc, tr = coracle.snooker_generation(Xs, lp, Z, M, lk, gs, 99, 12345, T=T, chain_offset=7)
            assert np.array_equal(tr["r1"], a["r1"]) and np.array_equal(tr["r2"], a["r2"]) and np.array_equal(tr["u"], a["u"])
            np.testing.assert_allclose(tr["xprop"], a["xprop"], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(tr["lp_prop"], a["lp_prop"], rtol=1e-11)
            assert np.array_equal(tr["accepted"], a["accepted"])
            Xn = Xs.copy(); Xn[a["accepted"]] = a["xprop"][a["accepted"]]
            np.testing.assert_allclose(Xc, Xn, rtol=1e-12, atol=1e-14)
    for g in range(1, 200):
        assert coracle.is_snooker_gen(31953150, g, 0.3) == orc.is_snooker_gen(31953150, g, 0.3)
