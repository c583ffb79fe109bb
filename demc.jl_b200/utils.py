"""Post-processing diagnostics of DEMC.jl on host arrays (the MC record returned by the sampler).
Mirrors src/utils.jl: Rhat_gelman (:2-20), flatten_chain (:22-32), convergence_check (:34-94),
mean_cov_chain (:96-111).  These run on the arrays the caller already holds in host memory; the
in-loop Rhat used by autostop is computed on the GPU from streaming accumulators (csrc)."""
from __future__ import annotations

import numpy as np

__all__ = ["Rhat_gelman", "flatten_chain", "mean_cov_chain", "convergence_check", "extract_best"]


def Rhat_gelman(chain, Npop=None, Ngeneration=None, Npar=None):
    """src/utils.jl:2-20. chain: Npop x Npar x Ngeneration."""
    chain = np.asarray(chain, dtype=np.float64)
    N, d, G = chain.shape
    Npop = N if Npop is None else Npop
    Ngeneration = G if Ngeneration is None else Ngeneration
    n = Ngeneration // 2
    m = Npop * 2
    cs = np.concatenate([chain[:Npop, :, :n], chain[:Npop, :, n:2 * n]], axis=0)
    avg_par = cs.mean(axis=(0, 2), keepdims=True)
    avg_chains = cs.mean(axis=2, keepdims=True)
    B = n / (m - 1) * np.sum((avg_chains - avg_par) ** 2, axis=0)
    sj = 1.0 / (n - 1) * np.sum((cs - avg_chains) ** 2, axis=2, keepdims=True)
    W = 1.0 / m * np.sum(sj, axis=0)
    varhat = (n - 1) / n * W + 1.0 / n * B
    return np.sqrt(varhat / W).reshape(-1)


def flatten_chain(chain, Npop=None, Ngeneration=None, Npar=None):
    """src/utils.jl:22-32 -> Npar x (Npop*Ngeneration), generation-major with the chain index fastest."""
    chain = np.asarray(chain)
    N, d, G = chain.shape
    return np.ascontiguousarray(chain.transpose(1, 2, 0).reshape(d, G * N))


def mean_cov_chain(chain, Npop=None, Ngeneration=None, Npar=None):
    """src/utils.jl:96-111 (1/(Ngeneration*Npop) normalisation)."""
    f = flatten_chain(chain)
    b = f.mean(axis=1)
    c = (f - b[:, None]) @ (f - b[:, None]).T / f.shape[1]
    return b, c


def convergence_check(chain, log_obj, figure_path=None, *, verbose=True, parnames=()):
    """src/utils.jl:34-94: (accept_ratio per chain, Rhat per parameter).  Plotting is disabled in
    the reference (commented out) and is not reproduced."""
    chain = np.asarray(chain)
    log_obj = np.asarray(log_obj)
    Npop, Npar, Ngeneration = chain.shape
    assert log_obj.shape[0] == Npop, "number of rows of log_obj is not equal to number of chains"
    assert log_obj.shape[1] == Ngeneration, "number of columns of log_obj is not equal to number of generations simulated"
    accept_ratio = np.sum(np.diff(log_obj, axis=1) != 0.0, axis=1) / (Ngeneration - 1)     # :61
    Rhat = Rhat_gelman(chain, Npop, Ngeneration, Npar)                                     # :82
    if verbose:
        print("Summary Checks\n\nAcceptance Ratio of each chain:")
        print(accept_ratio)
        print(f"\nRhat Gelman: {Rhat}\n")
    return accept_ratio, Rhat


def extract_best(mc):
    """(best log-objective, its parameter vector) over the whole history — what src/utils.jl:120-125 intends (the
    upstream body uses `findn`, gone since Julia 1.0); first maximum in column-major order like Julia's `findn`."""
    lo = np.asarray(mc.log_obj)
    if lo.size == 0:
        raise ValueError("extract_best: the MC record holds no history (store_chain was off)")
    flat = int(np.argmax(lo.ravel(order="F")))            # NaN-free by construction: a NaN proposal is never accepted
    row, col = flat % lo.shape[0], flat // lo.shape[0]
    return float(lo[row, col]), np.array(mc.chain[row, :, col], dtype=np.float64)
