"""Model check (pure Python, no GPU) of the packed exchange's staging-ring protocol in overlap sync mode
(demc.jl_b200/csrc/demcz_comm.cu, demcz_internal.h: kStageDepth).

Per rank the library enforces only these orderings:
  * push of segment k starts after the rank's own K-block k ended and after its previous push finished (one side stream);
  * K-block k+1 (k+1 >= 2) starts after K-block k ended, after the pushes of segment k-1 of EVERY rank have landed
    (own event + the peers' flag words), and after the rank has then unpacked segment k-1 out of staging buffer
    (k-1) mod depth.
Nobody acknowledges an unpack.  The claim in the header is that depth 4 makes that safe: a sender refills buffer
s mod 4 with segment s+4 only after every receiver has unpacked segment s — whatever the speeds of the ranks.
The model draws adversarial durations and checks the claim; it also shows that the argument is tight (depth 3 can be
overrun), which is why the ring is not shallower.
"""
import random


def simulate(world, n_seg, rng, slow_rank=None):
    """Event times of a run of n_seg K-blocks.  Returns (push_start[r][k], unpack_end[r][k])."""
    INF = float("inf")
    S_end = [[0.0] * n_seg for _ in range(world)]
    P_start = [[INF] * n_seg for _ in range(world)]
    P_end = [[INF] * n_seg for _ in range(world)]
    U_end = [[INF] * n_seg for _ in range(world)]

    def dur(r, lo, hi):
        x = rng.uniform(lo, hi)
        return x * (rng.choice([1, 1, 20]) if r == slow_rank else 1)        # one rank stalls now and then

    for k in range(n_seg):                      # K-block k of every rank; pushes of segment k follow it
        for r in range(world):
            start = S_end[r][k - 1] if k else 0.0
            if k >= 2:                           # segment k-2 becomes drawable: landed from everyone, then unpacked here
                start = max([start] + [P_end[q][k - 2] for q in range(world)])
                start += dur(r, 0.0, 0.6)        # unpack
                U_end[r][k - 2] = start
            S_end[r][k] = start + dur(r, 0.5, 1.5)
        for r in range(world):
            P_start[r][k] = max(S_end[r][k], P_end[r][k - 1] if k else 0.0)
            P_end[r][k] = P_start[r][k] + dur(r, 0.1, 2.5)
    return P_start, U_end


def overruns(world, n_seg, depth, P_start, U_end):
    bad = 0
    for s in range(n_seg - depth):
        for a in range(world):
            for b in range(world):
                if a != b and not (P_start[a][s + depth] >= U_end[b][s]):
                    bad += 1                     # a writes segment s+depth into b's buffer before b has emptied it
    return bad


def test_depth_4_is_never_overrun():
    rng = random.Random(1234)
    for trial in range(300):
        world = rng.choice([2, 3, 8])
        P, U = simulate(world, 24, rng, slow_rank=rng.choice([None, 0, world - 1]))
        assert overruns(world, 24, 4, P, U) == 0


def test_depth_3_can_be_overrun_so_the_ring_is_not_shallower():
    rng = random.Random(99)
    worst = 0
    for trial in range(300):
        P, U = simulate(2, 24, rng, slow_rank=1)
        worst = max(worst, overruns(2, 24, 3, P, U))
    assert worst > 0
