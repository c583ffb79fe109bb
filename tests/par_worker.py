"""Worker of tests/test_gpu_multi.py::test_par_entry_points_under_torchrun: run under `torchrun --nproc-per-node 2`.
Calls the reference's parallel entry points (demcz_sample_par / demcz_anneal_par: src/demcz.jl:101-165,
src/demcz_anneal.jl:82-140), which shard the chains over the ranks, and writes what rank 0 got to an .npz."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demc_jl_b200 as DEMC  # noqa: E402


def problem():
    d, N = 6, 256
    rng = np.random.default_rng(3)
    A = rng.normal(size=(d, d))
    S = A @ A.T / d + 0.5 * np.eye(d)
    mu = rng.normal(size=d)
    Z0 = rng.normal(size=(2 * N, d))
    return d, N, mu, S, Z0


def run(out_path):
    d, N, mu, S, Z0 = problem()
    blocks = [list(range(d))]
    eps = 1e-5 * np.ones(d)
    mc, Z = DEMC.demcz_sample_par(DEMC.MvNormalLogpdf(mu, S), Z0, N, 5, 40, 1, blocks, eps, 2.38, sync_every=20, autostop="Rhat",
                                  autostop_Rhat=0.0, verbose=False, seed=9)
    ma, Za = DEMC.demcz_anneal_par(DEMC.QuadraticLogobj(mu), Z0, N, 5, 40, 1, blocks, eps, 2.38, sync_every=10, T0=2.0, TN=1e-3,
                                   verbose=False, seed=9)
    if int(os.environ.get("RANK", 0)) == 0:
        np.savez(out_path, chain=mc.chain, log_obj=mc.log_obj, X=mc.Xcurrent, lp=mc.log_objcurrent, Z=Z, rhat=mc.info["rhat"],
                 accepted=mc.info["accepted"], a_chain=ma.chain, a_X=ma.Xcurrent, a_lp=ma.log_objcurrent, a_Z=Za,
                 a_gamma=ma.info["final_gamma"], a_accepted=ma.info["accepted"])


if __name__ == "__main__":
    run(sys.argv[1])
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
