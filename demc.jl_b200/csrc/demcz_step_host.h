// demcz_step_host.h — host-side glue of the fused step kernel shared by demcz_lib.cu (generation loop, batched
// evaluation) and demcz_step_launch.cu (the kernel instantiations, one translation unit per width group).
#pragma once
#include <cstring>
#include "demcz_step.cuh"
#include "demcz_internal.h"

// closed-form likelihood constants of the context in the kernel's parameter layout (see MvnParams)
template <int LD>
inline demcz::MvnParams<LD> demcz_make_mvn_params(const demcz_ctx* c, bool jmajor)
{
    using namespace demcz;
    using MP = MvnParams<LD>;
    MvnParams<LD> P;
    memset(&P, 0, sizeof P);
    const int d = c->o.d;
    auto W = [&](int i, int j) { return (i < d && j < d && j <= i) ? c->mvn_W[(size_t)i * d + j] : 0.f; };
    for (int k = 0; k < LD / 2; ++k)
        P.mu2[k] = make_float2(2 * k < d ? c->mvn_mu[2 * k] : 0.f, 2 * k + 1 < d ? c->mvn_mu[2 * k + 1] : 0.f);
    if (jmajor) {
        for (int j = 0; j < LD; ++j)
            for (int i2 = j / 2; i2 < LD / 2; ++i2) P.W[MP::wj_off(j) + i2 - j / 2] = make_float2(W(2 * i2, j), W(2 * i2 + 1, j));
    } else {
        for (int i2 = 0; i2 < LD / 2; ++i2)
            for (int j = 0; j < 2 * i2 + 2; ++j) P.W[i2 * (i2 + 1) + j] = make_float2(W(2 * i2, j), W(2 * i2 + 1, j));
    }
    P.c0 = c->mvn_c0;
    return P;
}


// one launcher per register width (LD = 4*ceil(d/4)), each in its own translation unit
cudaError_t demcz_launch_step_ld4(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full);
cudaError_t demcz_launch_step_ld8(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full);
cudaError_t demcz_launch_step_ld12(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full);
cudaError_t demcz_launch_step_ld16(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full);
cudaError_t demcz_launch_step_ld20(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full);
cudaError_t demcz_launch_step_ld24(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full);
cudaError_t demcz_launch_step_ld28(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full);
cudaError_t demcz_launch_step_ld32(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full);

inline cudaError_t demcz_launch_step(demcz_ctx* c, const demcz::StepArgs& a, bool acc, bool full)
{
    switch (c->LD) {
        case 4:  return demcz_launch_step_ld4(c, a, acc, full);
        case 8:  return demcz_launch_step_ld8(c, a, acc, full);
        case 12: return demcz_launch_step_ld12(c, a, acc, full);
        case 16: return demcz_launch_step_ld16(c, a, acc, full);
        case 20: return demcz_launch_step_ld20(c, a, acc, full);
        case 24: return demcz_launch_step_ld24(c, a, acc, full);
        case 28: return demcz_launch_step_ld28(c, a, acc, full);
        case 32: return demcz_launch_step_ld32(c, a, acc, full);
    }
    return cudaErrorInvalidValue;
}
