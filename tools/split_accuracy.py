"""CPU study (numpy; no GPU involved) for DESIGN.md section 9, item 1: operand-split accuracy of the linreg residual
GEMM  D[chain,row] = sum_k delta[chain,k] X[row,k]  and of  ||X delta||^2:

  3xTF32  : the shipped kernel's split — hi = tf32(v), lo = tf32(v - hi), products hi*hi + lo*hi + hi*lo
  2xFP16  : hi = fp16(v'), lo = fp16(v' - hi) on operands pre-scaled by powers of two (per regressor column for X, the
            inverse scale and one global scale for delta), the same three products — three kind::f16 MMAs at twice the
            TF32 issue rate
  1xTF32 / 1xFP16 : no split, for scale

Products and sums are taken in FP64 here, so the numbers isolate the operand error (the tensor core adds FP32
accumulation on top of either).  Run:  python tools/split_accuracy.py > profiles/r1_split_accuracy_cpu.log
"""
import numpy as np


def tf32(v):
    """cvt.rna.tf32.f32: 10 explicit mantissa bits, round to nearest, ties away from zero."""
    b = np.asarray(v, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000
    return b.astype(np.uint32).view(np.float32)


def split_tf32(v):
    hi = tf32(v)
    lo = tf32(np.asarray(v, np.float32) - hi)
    return hi.astype(np.float64), lo.astype(np.float64)


def split_fp16(v):
    hi = np.asarray(v, np.float32).astype(np.float16)
    lo = (np.asarray(v, np.float32) - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def pow2_scale(maxabs, target_exp):
    """power of two s with maxabs * s in [2^(target_exp-1), 2^target_exp)"""
    return 2.0 ** (target_exp - 1 - np.floor(np.log2(maxabs)))


def gemm3(dh, dl, xh, xl):
    return dh @ xh.T + dl @ xh.T + dh @ xl.T


def study(name, X, delta):
    X = X.astype(np.float32).astype(np.float64)           # both schemes start from FP32 data
    delta = delta.astype(np.float32).astype(np.float64)
    D = delta @ X.T
    q = (D * D).sum(axis=1)
    out = {}
    # --- TF32
    dh, dl = split_tf32(delta); xh, xl = split_tf32(X)
    out["1xTF32"] = dh @ xh.T
    out["3xTF32"] = gemm3(dh, dl, xh, xl)
    # --- FP16 with power-of-two scaling: X columns to [2^13, 2^14), delta/sx globally to [2^13, 2^14)
    sx = pow2_scale(np.abs(X).max(axis=0), 14)
    Xs = X * sx
    ds = delta / sx
    sd = pow2_scale(np.abs(ds).max(), 14)
    ds = ds * sd
    dh, dl = split_fp16(ds); xh, xl = split_fp16(Xs)
    out["1xFP16"] = (dh @ xh.T) / sd
    out["2xFP16"] = gemm3(dh, dl, xh, xl) / sd
    print(f"--- {name}: n={X.shape[0]} p={X.shape[1]} chains={delta.shape[0]}  |delta| ~ {np.abs(delta).mean():.2e}  "
          f"||X delta||^2 ~ {q.mean():.3e}")
    for k in ("1xTF32", "1xFP16", "3xTF32", "2xFP16"):
        Dk = out[k]
        qk = (Dk * Dk).sum(axis=1)
        eD = np.abs(Dk - D).max() / np.abs(D).max()
        eq = np.abs(qk - q) / q
        print(f"  {k:7s}: max |dD|/max|D| = {eD:9.2e}   rel err of ||X delta||^2: max {eq.max():9.2e}  mean {eq.mean():9.2e}   "
              f"abs err max {np.abs(qk - q).max():9.2e}")


def main():
    rng = np.random.default_rng(31953150)
    n, p, N = 100_000, 32, 64
    z0 = rng.normal(size=(n, 1))
    X = np.ones((n, p))
    X[:, 1:] = np.sqrt(0.25) * z0 + np.sqrt(0.75) * rng.normal(size=(n, p - 1))     # example_linreg.jl generator (C2)
    post = rng.normal(size=(N, p)) * 1e-3                                           # posterior scale: sd ~ 1/sqrt(n)
    study("C2 shape, posterior-scale delta", X, post)
    study("C2 shape, start-up delta ~ 1", X, rng.normal(size=(N, p)))
    scales = 10.0 ** rng.uniform(-3, 3, size=p)
    study("ill-scaled regressors (column scales 1e-3..1e3), delta ~ 1e-3/scale", X * scales, post / scales)
    d2 = post.copy(); d2[:, ::2] *= 1e-6                                            # half of the coefficients 1e6 times smaller
    study("C2 shape, delta with components spread over 6 decades", X, d2)
    Xh = X.copy(); Xh[rng.integers(0, n, 50), rng.integers(1, p, 50)] *= 1e4        # outliers in X: the per-column scale follows them
    study("C2 shape with 50 outliers of 1e4 in X", Xh, post)


if __name__ == "__main__":
    main()
