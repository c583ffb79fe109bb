"""demc.jl_b200 — B200-native DEMCz sampler behind the DEMC.jl interface (generation loop only).

Usage mirrors the reference (README.md:19-33 upstream):

    import demc_jl_b200 as DEMC
    logobj = DEMC.MvNormalLogpdf(mu, Sigma)
    opts = DEMC.demcopt(ndim)
    mc, Z = DEMC.demcz_sample(logobj, Zinit, opts)

Importing the package does not load CUDA; the first sampler call loads libdemcz_cuda.so and fails
loudly if it (or a GPU) is missing — there is no CPU fallback.
"""
from .api import (DEMCopt, MC, LinRegLoglik, LogisticLoglik, MvNormalLogpdf, QuadraticLogobj, Sampler, demcopt, demcz_anneal,
                  demcz_anneal_par, demcz_sample, demcz_sample_par, tempbaseline)
from .utils import Rhat_gelman, convergence_check, extract_best, flatten_chain, mean_cov_chain
from . import _lib

__all__ = ["DEMCopt", "MC", "LinRegLoglik", "LogisticLoglik", "MvNormalLogpdf", "QuadraticLogobj", "Sampler", "demcopt", "demcz_anneal",
           "demcz_sample", "demcz_sample_par", "demcz_anneal_par", "tempbaseline", "Rhat_gelman", "convergence_check", "flatten_chain", "mean_cov_chain", "extract_best"]
