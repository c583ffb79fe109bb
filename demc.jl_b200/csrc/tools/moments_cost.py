"""Cost of opts.moments on the C4 shape (one GPU): ms per K-block with and without the in-kernel accumulators, and the
accumulated moments against the target.  python tools/moments_cost.py"""
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demc_jl_b200 as DEMC  # noqa: E402

d, N = 20, 131072
S = 0.95 * np.ones((d, d)) + 0.05 * np.eye(d)
mu = np.linspace(-1, 1, d)
Z0 = (mu + np.random.default_rng(0).standard_normal((1 << 21, d)) @ np.linalg.cholesky(S).T)
for mom in (False, True):
    with DEMC.Sampler(DEMC.MvNormalLogpdf(mu, S), d, N=N, K=10, Ngeneration=2000, blockindex=[list(range(d))],
                      eps_scale=1e-5 * np.ones(d), γ=2.38, store_chain=False, moments=mom, archive_capacity=(1 << 21) + N * 40) as s:
        s.set_archive(Z0)
        s.run(200)
        info = s.run(1000)
        print("moments", mom, "us per K-block (10 generations)", 1e3 * info["device_ms"] / 100)
        if mom:
            m, c, n = s.get_history_moments()
            print("points", n, "max |mean - mu|", np.abs(m - mu).max(), "max |cov - Sigma|", np.abs(c - S).max())
