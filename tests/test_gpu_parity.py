"""GPU parity tests proper (-m gpu): the CUDA generation loop, called through the C ABI, against
the CPU oracle on the same seeded inputs — see tests/parity.py for the comparison rules."""
import numpy as np
import pytest
import scipy.stats

import demc_jl_b200 as DEMC
from demc_jl_b200 import _lib
from oracle import demcz_oracle as orc
from tests.parity import check_run_against_oracle, example_normpdf_target

pytestmark = pytest.mark.gpu


def _z0(rows, d, seed=1, scale=1.0):
    return scale * np.random.default_rng(seed).normal(size=(rows, d))


def test_config1_normpdf_stepwise():
    # BASELINE configs[0]: test/example_normpdf.jl — d=5 MvNormal, N=100, K=10
    d, N = 5, 100
    mu, S = example_normpdf_target(d)
    st = check_run_against_oracle(DEMC.MvNormalLogpdf(mu, S), _z0(10 * d + N, d), N=N, K=10, G=60,
                                  blocks=[list(range(d))], eps=1e-5 * np.ones(d), gamma=2.38, seed=31953150)
    assert st["proposals"] == 100 * 60 and st["band_disagree"] <= st["in_band"]


@pytest.mark.parametrize("d,N,K,blocks", [
    (5, 77, 7, [[0, 1, 2], [3], [4]]),            # ragged N, blocklen-1 blocks
    (5, 33, 3, [[4, 2], [0, 1, 3]]),              # non-contiguous blocks
    (1, 40, 5, [[0]]),                            # scalar parameter
    (10, 256, 10, [list(range(5)), list(range(5, 10))]),
    (20, 1000, 10, [list(range(20))]),
    (32, 64, 4, [list(range(16)), list(range(16, 32))]),   # DEMCZ_MAX_DIM
    (7, 31, 1, [[0, 1, 2, 3, 4, 5, 6]]),          # K=1: append every generation
])
def test_shapes_and_blocks_stepwise(d, N, K, blocks):
    mu, S = example_normpdf_target(d, seed=d)
    eps = 1e-3 * (1 + np.arange(d))
    st = check_run_against_oracle(DEMC.MvNormalLogpdf(mu, S), _z0(3 * N + 11, d, seed=N, scale=0.3), N=N, K=K, G=24,
                                  blocks=blocks, eps=eps, gamma=2.38, seed=12345 + d)
    assert st["accepted"] > 0


def test_correlated_d20_stepwise():
    # BASELINE configs[3] target: 20-dim highly correlated MvNormal
    d, N = 20, 512
    S = 0.95 * np.ones((d, d)) + 0.05 * np.eye(d)
    mu = np.linspace(-1, 1, d)
    Z0 = np.random.default_rng(3).multivariate_normal(mu, S, size=4 * N)
    st = check_run_against_oracle(DEMC.MvNormalLogpdf(mu, S), Z0, N=N, K=10, G=30, blocks=[list(range(d))],
                                  eps=1e-5 * np.ones(d), gamma=2.38, seed=7)
    assert 0.05 < st["accept_frac"] < 0.6


@pytest.mark.parametrize("TN", [1e-4, 0.0])
def test_anneal_config3_stepwise(TN):
    # BASELINE configs[2]: test/test_anneal.jl — quadratic objective, d=10, Nblocks=2; TN=0 is SURVEY Q4
    d, N = 10, 200
    mu = np.random.default_rng(2).random(d)
    st = check_run_against_oracle(DEMC.QuadraticLogobj(mu), _z0(10 * d + N, d), N=N, K=10, G=40,
                                  blocks=[list(range(5)), list(range(5, 10))], eps=1e-5 * np.ones(d), gamma=2.38,
                                  seed=99, anneal=True, T0=2.0, TN=TN)
    assert st["accepted"] > 0


def test_custom_temperature_table_and_quirk_Q3():
    # serial reference quirk Q3: T = 1e-3^(ig/1000) whatever opts say — reproducible through the table
    d, N, G = 4, 64, 20
    mu = np.zeros(d)
    T = np.array([orc.tempbaseline(ig, 1000, 1.0, 1e-3) for ig in range(1, G + 1)])
    check_run_against_oracle(DEMC.QuadraticLogobj(mu), _z0(2 * N, d), N=N, K=5, G=G, blocks=[list(range(d))],
                             eps=1e-4 * np.ones(d), gamma=2.38, seed=5, anneal=True, Ttable=T)


def test_init_modes_and_gen_offset_and_chunked_run():
    d, N = 5, 48
    mu, S = example_normpdf_target(d)
    lo = DEMC.MvNormalLogpdf(mu, S)
    kw = dict(N=N, K=10, G=30, blocks=[list(range(d))], eps=1e-5 * np.ones(d), gamma=2.38, seed=3)
    check_run_against_oracle(lo, _z0(60, d), init_mode="ref_serial_zeros", **kw)
    check_run_against_oracle(lo, _z0(60, d), gen_offset=2**33 + 17, **kw)               # 64-bit generation counter
    check_run_against_oracle(lo, _z0(60, d), chunks=[7, 1, 12], **kw)                   # resume inside a context
    X0 = mu + 0.01 * np.random.default_rng(5).normal(size=(N, d))
    check_run_against_oracle(lo, _z0(60, d), X0=X0, **kw)                               # prevrun-style state


def test_overlap_sync_mode_visibility_stepwise():
    # DEMCZ_SYNC_OVERLAP: rows appended after K-block j become drawable from K-block j+2 (also on one GPU)
    d, N = 5, 64
    mu, S = example_normpdf_target(d)
    check_run_against_oracle(DEMC.MvNormalLogpdf(mu, S), _z0(100, d), N=N, K=4, G=30, blocks=[list(range(d))],
                             eps=1e-5 * np.ones(d), gamma=2.38, seed=3, sync_mode="overlap", chunks=[5, 9])


def test_prevrun_continuation_is_bit_identical():
    # mc, Z = demcz_sample(...); mc, Z = demcz_sample(..., Z; prevrun=mc)   (test/example_normpdf.jl:30-32)
    d, N, K = 5, 20, 10
    mu, S = example_normpdf_target(d)
    lo = DEMC.MvNormalLogpdf(mu, S)
    args = (N, K, 40, 1, [list(range(d))], 1e-5 * np.ones(d), 2.38)
    Z0 = _z0(50, d)
    mc_full, Z_full = DEMC.demcz_sample(lo, Z0, *args, verbose=False)
    args2 = (N, K, 20, 1, [list(range(d))], 1e-5 * np.ones(d), 2.38)
    mc1, Z1 = DEMC.demcz_sample(lo, Z0, *args2, verbose=False)
    mc2, Z2 = DEMC.demcz_sample(lo, Z1, *args2, prevrun=mc1, verbose=False)
    assert mc2.chain.shape == (N, d, 40)
    np.testing.assert_array_equal(mc2.chain, mc_full.chain)
    np.testing.assert_array_equal(mc2.log_obj, mc_full.log_obj)
    np.testing.assert_array_equal(Z2, Z_full)


def test_eval_logdensity_vs_scipy():
    for d in (1, 5, 20, 32):
        mu, S = example_normpdf_target(d, seed=d)
        X = mu + 0.2 * np.random.default_rng(d).normal(size=(1000, d))
        with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=4, K=10, Ngeneration=1, blockindex=[list(range(d))],
                          eps_scale=np.ones(d), γ=2.38) as s:
            got = s.eval_logdensity(X)
        dist = scipy.stats.multivariate_normal(mu, S)
        ref = dist.logpdf(X)
        # lp = c0 - q/2 can cancel to ~0: the 1e-5 is relative to the two terms, not to their difference
        scale = np.maximum(np.abs(ref), abs(dist.logpdf(mu)))
        assert np.all(np.abs(got - ref) <= 1e-5 * scale)


def test_error_paths():
    d = 3
    with pytest.raises(_lib.DemczError) as ei:
        DEMC.Sampler(DEMC.MvNormalLogpdf(np.zeros(d), -np.eye(d)), d, N=4, K=10, Ngeneration=1,
                     blockindex=[list(range(d))], eps_scale=np.ones(d), γ=2.38)
    assert ei.value.code == -5                              # DEMCZ_ERR_NOT_PD
    with DEMC.Sampler(DEMC.MvNormalLogpdf(np.zeros(d), np.eye(d)), d, N=4, K=10, Ngeneration=5,
                      blockindex=[list(range(d))], eps_scale=np.ones(d), γ=2.38) as s:
        with pytest.raises(_lib.DemczError):
            s.run()                                         # no archive yet
        with pytest.raises(_lib.DemczError):
            s.set_archive(np.zeros((1, d)))                 # < 2 rows: cannot draw r1 != r2
        s.set_archive(np.zeros((3, d)))
        with pytest.raises(_lib.DemczError):
            s.run()                                         # archive smaller than N for last_N_of_Z init


def test_streaming_rhat_matches_Rhat_gelman():
    # the autostop Rhat comes from streaming accumulators in the step kernel; compare it with
    # Rhat_gelman (src/utils.jl:2-20) evaluated by the oracle on the stored window
    d, N, every = 5, 64, 50
    mu, S = example_normpdf_target(d)
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=10, Ngeneration=2 * every, blockindex=[list(range(d))],
                      eps_scale=1e-5 * np.ones(d), γ=2.38, autostop=True, autostop_every=every, autostop_Rhat=0.0,
                      seed=11) as s:
        s.set_archive(_z0(200, d))
        for w in range(2):
            info = s.run(every)
            assert info["rhat_valid"] and not info["stopped_by_rhat"]
            chain, _ = s.get_chain()
            ref = orc.rhat_gelman(chain[:, :, w * every:(w + 1) * every])
            got = s.get_rhat()
            np.testing.assert_allclose(got, ref, rtol=2e-4)
            assert info["rhat_max"] == pytest.approx(got.max())


def test_rhat_autostop_stops_and_odd_window():
    d, N, every = 3, 128, 101            # odd window: n = floor(every/2), last generation of the window unused
    mu, S = np.zeros(d), np.eye(d)
    Z0 = np.random.default_rng(1).normal(size=(4 * N, d))
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=10, Ngeneration=20 * every, blockindex=[list(range(d))],
                      eps_scale=1e-5 * np.ones(d), γ=2.38, autostop=True, autostop_every=every, autostop_Rhat=1.2,
                      seed=13) as s:
        s.set_archive(Z0)
        info = s.run()
        assert info["stopped_by_rhat"] and info["generations_total"] % every == 0
        assert info["generations_total"] < 20 * every
        chain, _ = s.get_chain()
        g = info["generations_total"]
        ref = orc.rhat_gelman(chain[:, :, g - every:g])
        np.testing.assert_allclose(s.get_rhat(), ref, rtol=2e-4)
        assert ref.max() < 1.2


def test_adapt_gamma_follows_reference_rule():
    # src/demcz_anneal.jl:48-57 replayed on the GPU's own log_obj history
    d, N, G, every = 6, 96, 90, 15
    mu = np.random.default_rng(4).random(d)
    with DEMC.Sampler(DEMC.QuadraticLogobj(mu), d, N=N, K=10, Ngeneration=G, blockindex=[list(range(d))],
                      eps_scale=1e-5 * np.ones(d), γ=2.38, anneal=True, T0=2.0, TN=1e-4, adapt=True, adapt_every=every,
                      min_gamma=0.1, max_gamma=4.0, seed=21) as s:
        s.set_archive(_z0(300, d))
        info = s.run()
        _, log_obj = s.get_chain()
    gamma = 2.38
    for w in range(G // every):
        win = log_obj[:, w * every:(w + 1) * every]
        ar = np.count_nonzero(np.diff(win, axis=1) != 0) / (N * every)
        if ar < 0.1:
            gamma = max(0.1, gamma / 2)
        elif ar > 0.5:
            gamma = min(4.0, 1.5 * gamma)
    assert info["final_gamma"] == gamma


def test_ring_buffer_archive_wraps():
    d, N, K, G, cap = 4, 32, 2, 40, 200
    mu, S = example_normpdf_target(d)
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=K, Ngeneration=G, blockindex=[list(range(d))],
                      eps_scale=1e-5 * np.ones(d), γ=2.38, archive_capacity=cap, trace=True, seed=2) as s:
        s.set_archive(_z0(64, d))
        s.run()
        chain, _ = s.get_chain()
        Z = s.get_archive()
        tr = s.get_trace()
    assert Z.shape == (cap, d)
    total = 64 + N * (G // K)
    assert total > cap
    # logical order = oldest first; the newest N rows are the last append (generation G)
    np.testing.assert_array_equal(Z[-N:], chain[:, :, G - 1])
    np.testing.assert_array_equal(Z[-2 * N:-N], chain[:, :, G - 1 - K])
    assert tr["r1"].max() < cap and tr["r2"].max() < cap and np.all(tr["r1"] != tr["r2"])


def test_moments_api():
    d, N = 5, 500
    mu, S = example_normpdf_target(d)
    X0 = np.random.default_rng(6).multivariate_normal(mu, S, size=N)
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=10, Ngeneration=1, blockindex=[list(range(d))],
                      eps_scale=np.ones(d), γ=2.38) as s:
        s.set_archive(_z0(N, d))
        s.set_state(X0)
        m, c = s.get_moments()
    X32 = X0.astype(np.float32).astype(np.float64)
    np.testing.assert_allclose(m, X32.mean(axis=0), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(c, np.cov(X32.T, bias=True), rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("d,N,K,G,M0", [
    (3, 1, 2, 9, 2),        # a single chain, the smallest legal archive (2 rows: r1 != r2 is forced)
    (2, 5, 50, 12, 7),      # K > G: the archive never grows
    (4, 33, 1, 6, 40),      # append after every generation, ragged N
])
def test_edge_sizes_stepwise(d, N, K, G, M0):
    mu, S = example_normpdf_target(d, seed=d + N)
    st = check_run_against_oracle(DEMC.MvNormalLogpdf(mu, S), _z0(M0, d, seed=M0, scale=0.2), N=N, K=K, G=G,
                                  blocks=[list(range(d))], eps=1e-4 * np.ones(d), gamma=2.38, seed=99,
                                  X0=(mu + 0.05 * np.random.default_rng(N).normal(size=(N, d))) if M0 < N else None)
    assert st["proposals"] == N * G


def test_zero_generations_and_state_only():
    d, N = 3, 8
    mu, S = example_normpdf_target(d)
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=10, Ngeneration=0, blockindex=[list(range(d))],
                      eps_scale=np.ones(d), γ=2.38) as s:
        s.set_archive(_z0(20, d))
        info = s.run()
        assert info["generations_total"] == 0 and info["M"] == 20
        X, lp = s.get_state()
        np.testing.assert_array_equal(X, _z0(20, d)[-N:].astype(np.float32).astype(np.float64))
        assert s.get_archive().shape == (20, d)


def test_archive_moments_are_mean_cov_of_thinned_history():
    # rows appended during the run = population every K-th generation -> mean_cov_chain (src/utils.jl:96-106) of the
    # K-thinned chain, available without MC.chain; checked against the stored history and through the ring buffer
    d, N, K, G = 5, 64, 5, 60
    mu, S = example_normpdf_target(d)
    for cap in (0, 500):
        with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=K, Ngeneration=G, blockindex=[list(range(d))],
                          eps_scale=1e-5 * np.ones(d), γ=2.38, archive_capacity=cap, seed=4) as s:
            s.set_archive(_z0(100, d))
            s.run()
            chain, _ = s.get_chain()
            Z = s.get_archive()
            n_new = N * 6                                  # the last 6 appended segments
            m, C = s.get_archive_moments(Z.shape[0] - n_new, n_new)
            with pytest.raises(_lib.DemczError):
                s.get_archive_moments(Z.shape[0] - 10, 11)
        thinned = chain[:, :, K - 1::K][:, :, -6:]
        mref, Cref = DEMC.mean_cov_chain(thinned)
        np.testing.assert_allclose(m, mref, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(C, Cref, rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize("kind", ["mvn", "linreg"])
def test_streamed_history_equals_resident_history(kind):
    # store_chain="stream" (demcz_opts.store_chain = 2): each launch's history slice goes to pinned host memory under
    # the following launches; MC.chain / MC.log_obj must be identical to the device-resident history
    rng = np.random.default_rng(12)
    if kind == "mvn":
        d = 5
        mu, S = example_normpdf_target(d)
        lo = DEMC.MvNormalLogpdf(mu, S)
        Z0 = _z0(80, d)
    else:
        n, d = 900, 6
        X = np.ones((n, d)); X[:, 1:] = rng.normal(size=(n, d - 1)); y = X @ (1 + rng.random(d)) + rng.normal(size=n)
        lo = DEMC.LinRegLoglik(X, y)
        Z0 = np.linalg.solve(X.T @ X, X.T @ y) + 0.05 * rng.normal(size=(80, d))
    args = (40, 7, 33, 1, [list(range(d))], 1e-5 * np.ones(d), 2.2)
    a, Za = DEMC.demcz_sample(lo, Z0, *args, verbose=False, store_chain=True)
    b, Zb = DEMC.demcz_sample(lo, Z0, *args, verbose=True, print_step=5, store_chain="stream")     # verbose: chunked demcz_run calls
    np.testing.assert_array_equal(a.chain, b.chain)
    np.testing.assert_array_equal(a.log_obj, b.log_obj)
    np.testing.assert_array_equal(Za, Zb)


@pytest.mark.parametrize("d,blocks", [(3, None), (10, None), (10, "two"), (13, None), (20, None), (20, "three"), (22, None), (27, "two")])
def test_dense_slot_archive_is_bit_identical_to_padded_slots(d, blocks):
    """exchange_layout="dense_slots" (archive rows back to back; the 8-GPU layout) changes where rows live, nothing
    else: state, log-objective, history, archive and accept counts equal the padded-slot run bit for bit (Rhat to 1e-6) — ragged gather
    tails (64 % rows-per-instruction != 0), block masks, Rhat accumulators, ring wrap and a ragged last warp included.
    The padded-slot path itself is checked step-wise against the oracle above."""
    N, K, G = 1000, 4, 40                                   # 1000 chains: the last warp is ragged
    rng = np.random.default_rng(100 + d)
    A = rng.normal(size=(d, d))
    S = A @ A.T / d + 0.5 * np.eye(d)
    mu = rng.normal(size=d)
    Z0 = rng.normal(size=(2 * N + 7, d))
    if blocks == "two":
        bi = [list(range(0, d // 2)), list(range(d // 2, d))]
    elif blocks == "three":
        bi = [[0, 5, 7], list(range(8, d)), [1, 2, 3, 4, 6]]
    else:
        bi = [list(range(d))]
    out = {}
    for lay in ("in_place", "dense_slots"):
        with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=K, Ngeneration=G, blockindex=bi, eps_scale=1e-5 * np.ones(d),
                          γ=2.38, seed=5, store_chain=True, autostop=True, autostop_every=20, autostop_Rhat=0.0,
                          archive_capacity=5 * N, exchange_layout=lay) as s:
            s.set_archive(Z0)
            s.run(17)
            info = s.run()
            out[lay] = (s.get_state(), s.get_archive(), s.get_chain(), s.get_rhat(), info["accepted"], s.get_accept_counts())
    a, b = out["in_place"], out["dense_slots"]
    np.testing.assert_array_equal(a[0][0], b[0][0])
    np.testing.assert_array_equal(a[0][1], b[0][1])
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2][0], b[2][0])
    np.testing.assert_array_equal(a[2][1], b[2][1])
    np.testing.assert_allclose(a[3], b[3], rtol=1e-6)      # FP64 atomics across CTAs: summation order varies run to run
    assert a[4] == b[4]
    np.testing.assert_array_equal(a[5], b[5])


def _small_linreg(rng, n=900, d=6, rows=80):
    X = np.ones((n, d)); X[:, 1:] = rng.normal(size=(n, d - 1)); y = X @ (1 + rng.random(d)) + rng.normal(size=n)
    return DEMC.LinRegLoglik(X, y), np.linalg.solve(X.T @ X, X.T @ y) + 0.05 * rng.normal(size=(rows, d))


@pytest.mark.parametrize("kind", ["mvn", "linreg"])
@pytest.mark.parametrize("store", [True, "stream"])
def test_thinned_history_is_every_thin_th_generation(kind, store):
    # opts.thin: MC.chain keeps generation g iff g % thin == 0 -> N x d x floor(G/thin), identical to the
    # corresponding slices of the full history; state, archive and counters unchanged
    rng = np.random.default_rng(21)
    if kind == "mvn":
        d = 5
        mu, S = example_normpdf_target(d)
        lo, Z0 = DEMC.MvNormalLogpdf(mu, S), _z0(80, d)
    else:
        lo, Z0 = _small_linreg(rng)
        d = 6
    args = (40, 7, 38, 1, [list(range(d))], 1e-5 * np.ones(d), 2.2)
    full, Zf = DEMC.demcz_sample(lo, Z0, *args, verbose=False, store_chain=True)
    for thin in (3, 7, 50):
        th, Zt = DEMC.demcz_sample(lo, Z0, *args, verbose=(thin == 3), print_step=5, store_chain=store, thin=thin)
        assert th.chain.shape == (40, d, 38 // thin) and th.log_obj.shape == (40, 38 // thin)
        np.testing.assert_array_equal(th.chain, full.chain[:, :, thin - 1::thin][:, :, :38 // thin])
        np.testing.assert_array_equal(th.log_obj, full.log_obj[:, thin - 1::thin][:, :38 // thin])
        np.testing.assert_array_equal(th.Xcurrent, full.Xcurrent)
        np.testing.assert_array_equal(Zt, Zf)


@pytest.mark.parametrize("kind", ["mvn", "linreg", "mvn_blocks"])
def test_history_moments_equal_mean_cov_chain_of_the_stored_history(kind):
    # opts.moments: flatten_chain + mean_cov_chain (src/utils.jl:22-32, 96-106) of EVERY generation from
    # device accumulators, without MC.chain; checked against utils.mean_cov_chain on the stored history of the same run
    rng = np.random.default_rng(22)
    if kind == "linreg":
        lo, Z0 = _small_linreg(rng, rows=300)
        d, blocks = 6, [list(range(6))]
    else:
        d = 7
        mu, S = example_normpdf_target(d)
        lo, Z0 = DEMC.MvNormalLogpdf(mu, S), _z0(300, d, scale=0.3)
        blocks = [list(range(d))] if kind == "mvn" else [[0, 2, 4], [1, 3, 5, 6]]
    N, K, G = 200, 5, 60
    with DEMC.Sampler(lo, d, N=N, K=K, Ngeneration=G, blockindex=blocks, eps_scale=1e-5 * np.ones(d), γ=2.0, seed=8,
                      store_chain=True, moments=True) as s:
        s.set_archive(Z0)
        s.run(23)
        s.run()
        chain, _ = s.get_chain()
        m, c, n = s.get_history_moments()
    assert n == N * G
    mref, cref = DEMC.mean_cov_chain(chain)
    np.testing.assert_allclose(m, mref, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(c, cref, rtol=2e-5, atol=1e-9 * np.abs(cref).max())
    # and the run itself is the run without the accumulators
    with DEMC.Sampler(lo, d, N=N, K=K, Ngeneration=G, blockindex=blocks, eps_scale=1e-5 * np.ones(d), γ=2.0, seed=8,
                      store_chain=True) as s:
        s.set_archive(Z0)
        s.run()
        chain0, _ = s.get_chain()
        with pytest.raises(_lib.DemczError):
            s.get_history_moments()
    np.testing.assert_array_equal(chain, chain0)


def test_archive_checksum_tracks_the_rows():
    d, N = 5, 64
    mu, S = example_normpdf_target(d)
    Z0 = _z0(128, d)
    hs = []
    for Z in (Z0, Z0.copy(), np.roll(Z0, 1, axis=0)):
        with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=4, Ngeneration=16, blockindex=[list(range(d))],
                          eps_scale=1e-5 * np.ones(d), γ=2.38, seed=3, store_chain=False, archive_capacity=256) as s:
            s.set_archive(Z)
            h0 = s.get_archive_checksum()
            s.run()                                         # 4 appends of 64 rows into a 256-row ring: wraps
            hs.append((h0, s.get_archive_checksum(), s.get_archive_checksum(0, 64)))
    assert hs[0] == hs[1]                                   # same rows, same order -> same checksum (before, after, partial)
    assert hs[0][0] != hs[2][0]                             # position-sensitive
    assert hs[0][0] != hs[0][1] and hs[0][1] != hs[0][2]


# ---- snooker update (ter Braak & Vrugt 2008, cited by the reference's README.md:57-59) -----------------
@pytest.mark.parametrize("d,blocks,anneal,gamma_s", [(5, None, False, 0.0), (10, "two", False, 1.7), (20, None, True, 0.0), (3, None, False, 0.0)])
def test_snooker_generations_stepwise(d, blocks, anneal, gamma_s):
    mu, S = example_normpdf_target(d, seed=d)
    bi = [list(range(d))] if blocks is None else [list(range(0, d // 2)), list(range(d // 2, d))]
    st = check_run_against_oracle(DEMC.MvNormalLogpdf(mu, S), _z0(300, d, seed=d, scale=0.3), N=130, K=4, G=40, blocks=bi,
                                  eps=1e-5 * np.ones(d), gamma=2.38, seed=21 + d, anneal=anneal, T0=2.0, TN=1e-2,
                                  snooker_prob=0.35, snooker_gamma=gamma_s, chunks=[7])
    assert st["snooker_gens"] >= 5 and st["accepted"] > 0
    assert st["proposals"] == 130 * (st["snooker_gens"] + (40 - st["snooker_gens"]) * len(bi))


def test_snooker_linreg_stepwise():
    rng = np.random.default_rng(31)
    n, p, N = 1500, 9, 96
    X = np.ones((n, p)); X[:, 1:] = rng.normal(size=(n, p - 1)); y = X @ (1 + rng.random(p)) + rng.normal(size=n)
    Z0 = np.linalg.solve(X.T @ X, X.T @ y) + 0.05 * rng.normal(size=(3 * N, p))
    st = check_run_against_oracle(DEMC.LinRegLoglik(X, y), Z0, N=N, K=5, G=20, blocks=[list(range(p))], eps=1e-5 * np.ones(p),
                                  gamma=2.0, seed=5, snooker_prob=0.3)
    assert st["snooker_gens"] >= 3


def test_snooker_mixture_samples_the_target():
    # 10 % snooker generations (the paper's mixture): posterior moments, R-hat and accept band as for the parallel update
    d, N = 5, 256
    mu, S = example_normpdf_target(d)
    mc, Z = DEMC.demcz_sample(DEMC.MvNormalLogpdf(mu, S), _z0(10 * d + N, d), N, 10, 6000, 1, [list(range(d))], 1e-5 * np.ones(d), 2.38,
                              verbose=False, snooker_prob=0.1, seed=77)
    keep = mc.chain[:, :, 2000:]
    b, C = DEMC.mean_cov_chain(keep)
    sd = np.sqrt(np.diag(S))
    assert np.all(np.abs(b - mu) < 0.05 * sd)
    assert np.linalg.norm(C - S) / np.linalg.norm(S) < 0.05
    ar, Rhat = DEMC.convergence_check(keep, mc.log_obj[:, 2000:], verbose=False)
    assert np.all(Rhat < 1.01) and 0.1 < ar.mean() < 0.5


# ---- small-N persistent launches (opts.persistent): every K-block between two host decisions in ONE launch ----------
@pytest.mark.parametrize("case", ["plain", "blocks", "anneal_adapt", "rhat", "overlap", "chunks", "tiny_ring", "snooker"])
def test_persistent_small_n_launches_are_bit_identical_to_one_launch_per_kblock(case):
    d, N, K, G = 5, 100, 10, 97
    mu, S = example_normpdf_target(d)
    kw = dict(K=K, Ngeneration=G, blockindex=[list(range(d))], eps_scale=1e-5 * np.ones(d), γ=2.38, seed=5, store_chain=True, trace=True)
    chunks = []
    if case == "blocks":
        kw["blockindex"] = [[0, 2, 4], [1, 3]]
    elif case == "anneal_adapt":
        kw.update(anneal=True, T0=2.0, TN=1e-3, adapt=True, adapt_every=25)
    elif case == "rhat":
        kw.update(autostop=True, autostop_every=30, autostop_Rhat=0.0)
    elif case == "overlap":
        kw.update(sync_mode="overlap")
    elif case == "chunks":
        chunks = [7, 33, 1]
    elif case == "tiny_ring":
        kw.update(archive_capacity=6 * N)             # the ring wraps during the run: persistent launches fall back near the wrap
    elif case == "snooker":
        kw.update(snooker_prob=0.2)
    out = {}
    for pers in (False, True):
        with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, persistent=pers, **kw) as s:
            s.set_archive(_z0(3 * N, d, scale=0.3))
            for c in chunks:
                s.run(c)
            info = s.run()
            assert s.kernel_info()["step_persistent"] == int(pers)
            out[pers] = (s.get_state(), s.get_chain(), s.get_archive(), s.get_trace(), s.get_accept_counts(), info)
    a, b = out[False], out[True]
    np.testing.assert_array_equal(a[0][0], b[0][0]); np.testing.assert_array_equal(a[0][1], b[0][1])
    np.testing.assert_array_equal(a[1][0], b[1][0]); np.testing.assert_array_equal(a[1][1], b[1][1])
    np.testing.assert_array_equal(a[2], b[2])
    for k in a[3]:
        np.testing.assert_array_equal(a[3][k], b[3][k], err_msg=k)
    np.testing.assert_array_equal(a[4], b[4])
    assert a[5]["accepted"] == b[5]["accepted"] and a[5]["final_gamma"] == b[5]["final_gamma"] and a[5]["M"] == b[5]["M"]
    if case not in ("tiny_ring",):
        assert b[5]["step_launches"] < a[5]["step_launches"]          # fewer, longer launches


def test_persistent_path_stepwise_against_oracle():
    d, N = 5, 100                                            # BASELINE config 1 (test/example_normpdf.jl shape)
    mu, S = example_normpdf_target(d)
    st = check_run_against_oracle(DEMC.MvNormalLogpdf(mu, S), _z0(10 * d + N, d), N=N, K=10, G=60, blocks=[list(range(d))],
                                  eps=1e-5 * np.ones(d), gamma=2.38, seed=31953150)
    assert st["accepted"] > 0


def test_anneal_par_entry_point_adapts_gamma_per_sync_block():
    # demcz_anneal_par (src/demcz_anneal.jl:82-140): γ is adapted once per sync block from the block's accept ratio
    # (:123-132 — the rule of :48-57 with adapt_every = sync_every); positional default TN = 1e-3 (:86)
    d, N, every, G = 4, 64, 15, 60
    mu = np.linspace(-1, 1, d)
    mc, Z = DEMC.demcz_anneal_par(DEMC.QuadraticLogobj(mu), _z0(3 * N, d), N, 5, G, 1, [list(range(d))], 1e-5 * np.ones(d), 2.38,
                                  sync_every=every, T0=2.0, verbose=False)
    assert mc.chain.shape == (N, d, G) and Z.shape[0] == 3 * N + N * (G // 5)
    gamma = 2.38
    for w in range(G // every):
        win = mc.log_obj[:, w * every:(w + 1) * every]
        ar = np.count_nonzero(np.diff(win, axis=1) != 0) / (N * every)
        gamma = max(0.1, gamma / 2) if ar < 0.1 else (min(4.0, 1.5 * gamma) if ar > 0.5 else gamma)
    assert mc.info["final_gamma"] == gamma


def test_rhat_and_moments_with_snooker_generations_match_the_stored_history():
    # the snooker kernel feeds the same streaming accumulators (split-Rhat sums, history moments, accept counts) as the fused
    # step kernel: Rhat_gelman (src/utils.jl:2-20) and mean_cov_chain (:96-106) on the stored chain must agree
    d, N, every, G = 6, 300, 40, 80
    mu, S = example_normpdf_target(d)
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=5, Ngeneration=G, blockindex=[list(range(d))], eps_scale=1e-5 * np.ones(d),
                      γ=2.38, seed=12, store_chain=True, autostop=True, autostop_every=every, autostop_Rhat=0.0, moments=True,
                      snooker_prob=0.25) as s:
        s.set_archive(_z0(2 * N, d, scale=0.3))
        s.run()
        chain, log_obj = s.get_chain()
        rhat = s.get_rhat()
        m, c, n = s.get_history_moments()
        counts = s.get_accept_counts()
    np.testing.assert_allclose(rhat, orc.rhat_gelman(chain[:, :, every:]), rtol=5e-4)
    mref, cref = DEMC.mean_cov_chain(chain)
    assert n == N * G
    np.testing.assert_allclose(m, mref, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(c, cref, rtol=2e-5, atol=1e-9 * np.abs(cref).max())
    np.testing.assert_array_equal(counts, np.sum(np.diff(log_obj, axis=1) != 0.0, axis=1))
