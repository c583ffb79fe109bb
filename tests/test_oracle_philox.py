"""Pins the oracle's RNG layer: Philox4x32-10 against the Random123 known-answer vectors,
and the C restatement against the numpy one."""
import numpy as np

from oracle import demcz_oracle as orc

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_kat_numpy():
    for ctr, key, out in KAT:
        got = orc.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == out


def test_philox_kat_c(coracle):
    for ctr, key, out in KAT:
        assert tuple(coracle.philox(ctr, key)) == out


def test_philox_c_vs_numpy_random(coracle):
    rng = np.random.default_rng(1)
    ctrs = rng.integers(0, 2**32, size=(64, 4), dtype=np.uint64).astype(np.uint32)
    key = (0x12345678, 0x9ABCDEF0)
    got = orc.philox4x32_10(ctrs, key)
    for i in range(64):
        assert coracle.philox(ctrs[i], key) == [int(x) for x in got[i]]


def test_uniform_open_interval_and_fp32_exact():
    w = np.array([0, 1, 511, 512, 2**32 - 1], dtype=np.uint32)
    u = orc.uniform(w)
    assert np.all(u > 0) and np.all(u < 1)
    assert np.all(u.astype(np.float32).astype(np.float64) == u)   # exactly representable in FP32
    assert u[0] == 2.0**-24 and u[-1] == 1 - 2.0**-24


def test_draw_pair_exhaustive_small_M():
    # r1 != r2 always; the ordered pair is uniform over the M*(M-1) possibilities (Q9)
    for M in (2, 3, 5, 7):
        # words chosen so that mulhi hits every bucket exactly once
        w1 = ((np.arange(M, dtype=np.uint64) << np.uint64(32)) // np.uint64(M) + np.uint64(1)).astype(np.uint32)
        w2 = ((np.arange(M - 1, dtype=np.uint64) << np.uint64(32)) // np.uint64(M - 1) + np.uint64(1)).astype(np.uint32)
        A, B = np.meshgrid(w1, w2, indexing="ij")
        r1, r2 = orc.draw_pair(A.ravel(), B.ravel(), M)
        assert np.all(r1 != r2)
        assert np.all((r1 >= 0) & (r1 < M) & (r2 >= 0) & (r2 < M))
        pairs = set(zip(r1.tolist(), r2.tolist()))
        assert len(pairs) == M * (M - 1)


def test_box_muller_moments():
    rng = np.random.default_rng(7)
    w = rng.integers(0, 2**32, size=200000, dtype=np.uint64).astype(np.uint32)
    z0, z1 = orc.box_muller16(w)
    z = np.concatenate([z0, z1])
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.mean(z**4) - 3) < 0.1
