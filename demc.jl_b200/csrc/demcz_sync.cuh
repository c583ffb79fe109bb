// demcz_sync.cuh — device-side waits shared by the kernels: the segment flag words of the multi-GPU archive
// exchange and a wall-clock bound for every spin (a protocol bug or a dead peer must end as a failed launch,
// never as a hung GPU; a healthy launch that is merely preempted or time-sliced must not trap).
#pragma once
#include <stdint.h>

namespace demcz {

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s of wall time

// Segment flags: peer r publishes `seq` into word [r] of THIS rank's flag block (st.release.sys after a system-scope
// fence over its row stores) once its rows of segments 0 .. seq-1 have landed here.  Wait until word `f` >= need.
// Normally zero spins: the segment a launch needs was pushed one K-block earlier (overlap sync mode).  Bounded by wall
// time through the sleep: 2^26 polls of >= ~100 ns each are at least several seconds (a dead peer ends as a failed
// launch, never as a hung GPU; a healthy but slow peer is waited for).  Kept tiny and inline: it runs in the prologue
// of the step kernel, which sits exactly at its 128-register budget.
__device__ __forceinline__ void flag_wait(const unsigned long long* f, unsigned long long need)
{
    unsigned long long v;
    unsigned int spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= need) return;
        if (++spins == (1u << 26)) __trap();     // a peer never delivered
        __nanosleep(200);
    }
}

// Prologue of every kernel that gathers archive rows on a sharded run: threads 0 .. world-1 each watch one peer's
// word; the caller's next __syncthreads orders the rest of the CTA behind the acquire (causality order is cumulative).
__device__ __forceinline__ void wait_peer_segments(const unsigned long long* flags, int world, int my_rank,
                                                   unsigned long long need)
{
    if (need != 0ull && (int)threadIdx.x < world && (int)threadIdx.x != my_rank) flag_wait(flags + threadIdx.x, need);
}

}  // namespace demcz
