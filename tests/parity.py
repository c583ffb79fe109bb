"""Step-wise (teacher-forced) parity of the CUDA generation loop against the CPU oracle.

FP32 libm on the GPU and FP64 libm on the CPU differ by ulps, so a free-running comparison
diverges after the first near-tie of an accept decision.  Instead, for every generation and block
the oracle is fed the state the GPU entered that proposal with, the same archive and the same
Philox counters, and the two are compared proposal by proposal:

  * archive indices r1, r2 and the accept uniform u ........ bit-exact (integer / exact-FP32 work)
  * proposal x' .............................................. |dx| <= XPROP_ATOL + XPROP_RTOL*|x'|
  * log-density of the proposal ............................. LP_RTOL relative (north star: 1e-5 in FP32)
  * accept/reject decision .................................. exact wherever the decision margin
        |(lp' - lp) - T log u| exceeds LP_RTOL * max(|lp'|, |lp|, 1); inside the band either answer is legal
  * history, state and archive appends ...................... bit-exact consequences of the GPU's own decisions

Test infrastructure: imports oracle/ as the checker only.
"""
from __future__ import annotations

import numpy as np

import demc_jl_b200 as DEMC
from oracle import demcz_oracle as orc

LP_RTOL = 1e-5
XPROP_RTOL = 2e-6
XPROP_ATOL = 1e-6


def oracle_lik(logobj):
    if isinstance(logobj, DEMC.LogisticLoglik):
        return orc.LogisticLik(logobj.X, logobj.y)
    if isinstance(logobj, DEMC.LinRegLoglik):
        return orc.LinRegLik(logobj.X, logobj.y)
    return orc.MvNormalLik(logobj.mu, logobj.Sigma, drop_constant=logobj.drop_constant)


def check_run_against_oracle(logobj, Z0, *, N, K, G, blocks, eps, gamma, seed, anneal=False, T0=3.0, TN=1e-3,
                             Ttable=None, gen_offset=0, init_mode="last_N_of_Z", X0=None, device=0, chunks=None,
                             lp_rtol=LP_RTOL, sync_mode="strict", snooker_prob=0.0, snooker_gamma=0.0):
    """Runs G generations on the GPU with trace + history and replays every proposal through the
    numpy oracle.  `chunks`: optional list of max_generations per demcz_run call (exercises resume
    inside one context).  Returns a dict of statistics; raises AssertionError on any mismatch."""
    Z0 = np.asarray(Z0, dtype=np.float64)
    M0, d = Z0.shape
    lik = oracle_lik(logobj)
    s = DEMC.Sampler(logobj, d, N=N, K=K, Ngeneration=G, blockindex=blocks, eps_scale=eps, γ=gamma, anneal=anneal,
                     T0=T0, TN=TN, seed=seed, gen_offset=gen_offset, store_chain=True, init_mode=init_mode,
                     trace=True, device=device, sync_mode=sync_mode, snooker_prob=snooker_prob, snooker_gamma=snooker_gamma)
    try:
        s.set_archive(Z0)
        if X0 is not None:
            s.set_state(X0)
        if Ttable is not None:
            s.set_temperature_table(Ttable)
        X_init, lp_init = s.get_state()
        if chunks:
            for c in chunks:
                s.run(c)
            info = s.run(0)
        else:
            info = s.run(0)
        assert info["generations_total"] == G
        chain, log_obj = s.get_chain()
        Xfin, lpfin = s.get_state()
        Z = s.get_archive()
        tr = s.get_trace()
        acc_counts = s.get_accept_counts()
    finally:
        s.close()

    # ---- initial state and its log-density (src/demcz.jl:15-17) ---------------------------------
    if X0 is None:
        if init_mode == "last_N_of_Z":
            np.testing.assert_array_equal(X_init, Z0[M0 - N:].astype(np.float32).astype(np.float64))
        else:
            assert np.all(X_init == 0)
    ref_lp0 = lik.logdensity(X_init)
    np.testing.assert_allclose(lp_init, ref_lp0, rtol=lp_rtol, atol=lp_rtol)

    if anneal:
        Tt = np.asarray(Ttable, dtype=np.float64) if Ttable is not None else \
            np.array([orc.tempbaseline(ig, G, T0, TN) for ig in range(1, G + 1)])
        Tt = Tt.astype(np.float32).astype(np.float64)       # the device table is FP32

    # ---- archive: appended rows are the generation-K states in chain order (src/demcz.jl:88-91) --
    n_app = G // K
    assert Z.shape[0] == M0 + N * n_app, (Z.shape, M0, N, n_app)
    np.testing.assert_array_equal(Z[:M0], Z0.astype(np.float32).astype(np.float64))
    for j in range(n_app):
        np.testing.assert_array_equal(Z[M0 + N * j: M0 + N * (j + 1)], chain[:, :, K * (j + 1) - 1])

    stats = dict(proposals=0, accepted=0, in_band=0, band_disagree=0, max_lp_relerr=0.0, max_xprop_err=0.0, snooker_gens=0)
    X = X_init.copy()
    lp = lp_init.copy()
    nb = len(blocks)
    for g in range(1, G + 1):
        n_seg = (g - 1) // K                                  # generation-synchronous archive
        if sync_mode == "overlap":                            # the newest segment is still in flight: not drawable yet
            n_seg = max(n_seg - 1, 0)
        M = M0 + N * n_seg
        T = Tt[g - 1] if anneal else None
        if orc.is_snooker_gen(seed, gen_offset + g, snooker_prob):
            # snooker generation (ter Braak & Vrugt 2008): one proposal per chain, all coordinates; trace in block slot 0
            stats["snooker_gens"] += 1
            o = orc.snooker_update(X, lp, Z, M, lik, snooker_gamma, seed, gen_offset + g, T=T)
            assert np.array_equal(tr["r1"][g - 1, 0], o["r1"]) and np.array_equal(tr["r2"][g - 1, 0], o["r2"]), f"snooker z1/z2 mismatch gen {g}"
            assert np.array_equal(tr["u"][g - 1, 0].astype(np.float64), o["u"]), f"snooker u mismatch gen {g}"
            g_xp = tr["xprop"][g - 1, 0].T.astype(np.float64)
            err = np.abs(g_xp - o["xprop"])
            scale = np.abs(X) + np.abs(Z[o["z"]]) + np.abs(Z[o["r1"]]) + np.abs(Z[o["r2"]])
            amp = 1.0 + np.abs(o["xprop"] - X).max(axis=1, keepdims=True) / np.maximum(np.abs(X - Z[o["z"]]).max(axis=1, keepdims=True), 1e-30)
            tol = XPROP_ATOL + 4e-6 * amp * (scale + np.abs(o["xprop"]))          # FP32 dot products: relative error ~d eps of proj
            assert np.all(err <= tol), f"snooker xprop mismatch gen {g}: max err {err.max():.3e}"
            ref_lpp = lik.logdensity(g_xp)
            g_lpp = tr["lp_prop"][g - 1, 0].astype(np.float64)
            rel = np.abs(g_lpp - ref_lpp) / np.maximum(np.abs(ref_lpp), 1.0)
            assert np.all(rel <= lp_rtol), f"snooker lp_prop mismatch gen {g}: max rel {rel.max():.3e}"
            stats["max_lp_relerr"] = max(stats["max_lp_relerr"], float(rel.max()))
            g_acc = tr["accepted"][g - 1, 0].astype(bool)
            Tm = 1.0 if T is None else T
            # geometry factor from the GPU's own proposal: |x' - z| / |x - z|
            with np.errstate(divide="ignore", invalid="ignore"):
                rr = np.sqrt(np.sum((g_xp - Z[o["z"]]) ** 2, axis=1) / np.sum((X - Z[o["z"]]) ** 2, axis=1))
                geo = (d - 1) * np.log(rr)
                margin = (ref_lpp - lp) + Tm * geo - Tm * np.log(o["u"])
            band = lp_rtol * np.maximum(np.maximum(np.abs(ref_lpp), np.abs(lp)), 1.0) * 2.0 + 1e-4 * (d - 1) * Tm
            ref_acc = margin > 0
            outside = np.abs(margin) > band
            assert np.array_equal(g_acc[outside], ref_acc[outside]), f"snooker accept decision mismatch outside the band, gen {g}"
            stats["in_band"] += int((~outside).sum())
            stats["band_disagree"] += int((g_acc != ref_acc).sum())
            stats["proposals"] += N
            stats["accepted"] += int(g_acc.sum())
            X[g_acc] = g_xp[g_acc]
            lp[g_acc] = g_lpp[g_acc]
            for ib in range(1, len(blocks)):
                assert not tr["accepted"][g - 1, ib].any()
            np.testing.assert_array_equal(chain[:, :, g - 1], X, err_msg=f"history mismatch gen {g} (snooker)")
            if isinstance(logobj, DEMC.LinRegLoglik):
                np.testing.assert_allclose(log_obj[:, g - 1], lp, rtol=2e-7, err_msg=f"log_obj mismatch gen {g}")
                lp = log_obj[:, g - 1].copy()
            else:
                np.testing.assert_array_equal(log_obj[:, g - 1], lp, err_msg=f"log_obj mismatch gen {g}")
            continue
        for ib, block in enumerate(blocks):
            o = orc.block_update(X, lp, Z, M, lik, block, ib, np.asarray(eps, dtype=np.float32).astype(np.float64),
                                 _f32(gamma, len(block)), seed, gen_offset + g, T=T)
            g_r1, g_r2 = tr["r1"][g - 1, ib], tr["r2"][g - 1, ib]
            assert np.array_equal(g_r1, o["r1"]), f"r1 mismatch gen {g} block {ib}"
            assert np.array_equal(g_r2, o["r2"]), f"r2 mismatch gen {g} block {ib}"
            assert np.array_equal(tr["u"][g - 1, ib].astype(np.float64), o["u"]), f"u mismatch gen {g} block {ib}"
            g_xp = tr["xprop"][g - 1, ib].T.astype(np.float64)            # (N, d)
            err = np.abs(g_xp - o["xprop"])
            tol = XPROP_ATOL + XPROP_RTOL * np.abs(o["xprop"])
            # the eps jitter adds eps*|z| * (relative error of the fast FP32 Box-Muller)
            tol = tol + 1e-4 * np.abs(np.asarray(eps))[None, :]
            # FP32 rounding of c*(Z[r1]-Z[r2]) scales with the archive rows, not with x'
            tol = tol + 4e-7 * abs(gamma) * (np.abs(Z[o["r1"]]) + np.abs(Z[o["r2"]]) + np.abs(X))
            assert np.all(err <= tol), f"xprop mismatch gen {g} block {ib}: max err {err.max():.3e}"
            stats["max_xprop_err"] = max(stats["max_xprop_err"], float(err.max()))
            untouched = np.setdiff1d(np.arange(d), np.asarray(block))
            if untouched.size:                                              # block mask: other columns unchanged
                assert np.array_equal(g_xp[:, untouched], X[:, untouched])
            # log-density of the GPU's own proposal, in FP64
            ref_lpp = lik.logdensity(g_xp)
            g_lpp = tr["lp_prop"][g - 1, ib].astype(np.float64)
            rel = np.abs(g_lpp - ref_lpp) / np.maximum(np.abs(ref_lpp), 1.0)
            assert np.all(rel <= lp_rtol), f"lp_prop mismatch gen {g} block {ib}: max rel {rel.max():.3e}"
            stats["max_lp_relerr"] = max(stats["max_lp_relerr"], float(rel.max()))
            # decision
            g_acc = tr["accepted"][g - 1, ib].astype(bool)
            logu = np.log(o["u"])
            with np.errstate(invalid="ignore"):
                margin = (ref_lpp - lp) - (logu if T is None else T * logu)
            band = lp_rtol * np.maximum(np.maximum(np.abs(ref_lpp), np.abs(lp)), 1.0) * 2.0
            ref_acc = margin > 0
            outside = np.abs(margin) > band
            assert np.array_equal(g_acc[outside], ref_acc[outside]), \
                f"accept decision mismatch outside the tolerance band, gen {g} block {ib}"
            stats["in_band"] += int((~outside).sum())
            stats["band_disagree"] += int((g_acc != ref_acc).sum())
            stats["proposals"] += N
            stats["accepted"] += int(g_acc.sum())
            # carry the GPU's decision forward (teacher forcing)
            X[g_acc] = g_xp[g_acc]
            lp[g_acc] = g_lpp[g_acc]
        np.testing.assert_array_equal(chain[:, :, g - 1], X, err_msg=f"history mismatch gen {g}")
        if isinstance(logobj, DEMC.LinRegLoglik):
            # the linreg path keeps log-objectives in FP64; the trace reports them rounded to FP32
            np.testing.assert_allclose(log_obj[:, g - 1], lp, rtol=2e-7, err_msg=f"log_obj mismatch gen {g}")
            lp = log_obj[:, g - 1].copy()
        else:
            np.testing.assert_array_equal(log_obj[:, g - 1], lp, err_msg=f"log_obj mismatch gen {g}")
    np.testing.assert_array_equal(Xfin, X)
    np.testing.assert_array_equal(lpfin, lp)            # (linreg: lp was re-read from the FP64 history above)
    assert info["accepted"] == stats["accepted"]
    assert info["proposals"] == stats["proposals"]
    # accept_ratio numerator of convergence_check (src/utils.jl:61)
    if G > 1:
        np.testing.assert_array_equal(acc_counts, np.sum(np.diff(log_obj, axis=1) != 0.0, axis=1))
    stats["accept_frac"] = stats["accepted"] / max(stats["proposals"], 1)
    stats["final_gamma"] = info["final_gamma"]
    return stats


def _f32(gamma, blocklen):
    """The device applies c = fl32(γ/sqrt(2 blocklen)); give the oracle a γ that reproduces it."""
    c = gamma if blocklen == 1 else gamma / np.sqrt(2.0 * blocklen)
    c32 = float(np.float32(c))
    return c32 if blocklen == 1 else c32 * np.sqrt(2.0 * blocklen)


def example_normpdf_target(d, seed=0):
    """μ, Σ of test/example_normpdf.jl:9-12."""
    rng = np.random.default_rng(seed)
    mu = rng.random(d)
    A = rng.random((d, d))
    S = A.T @ A + 2 * np.eye(d)
    S = S / S.max() / 100
    return mu, S
