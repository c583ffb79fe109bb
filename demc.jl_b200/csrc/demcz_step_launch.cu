// demcz_step_launch.cu — instantiations and launch of the fused step kernel (demcz_step.cuh).  Compiled once per
// register-width group (-DDEMCZ_LD_GROUP=0..3) so that the build parallelises: every (LD, ACC, FULL, DENSE)
// combination is its own kernel.
#include "demcz_step_host.h"
#include <cstdlib>

using namespace demcz;

namespace {

constexpr int kWarps = 4;

// production structure of the step kernel (see tools/step_tune.cu and profiles/ for how it was chosen)
constexpr int kStepOpt = DEMCZ_STEP_OPT;

// resident blocks per SM the register allocation is held to (4 warps per block)
template <int LD> constexpr int min_blocks() { return LD <= 20 ? 4 : LD <= 28 ? 3 : 2; }

template <int LD, bool ACC, bool FULL, bool DENSE>
cudaError_t launch_step_t(demcz_ctx* c, const StepArgs& a)
{
    using SL = StageLayout<LD>;
    auto kern = demcz_step_mvn_kernel<LD, ACC, FULL, kWarps, min_blocks<LD>(), kStepOpt, DENSE>;
    const size_t smem = (size_t)kWarps * (2 * SL::FLOATS) * sizeof(float);
    static bool configured[64] = {false};       // per instantiation and per device (function attributes are per device)
    const int dev = c->device & 63;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured[dev] = true;
    }
    const int threads = kWarps * 32;
    const int grid = (a.N + threads - 1) / threads;
    const MvnParams<LD> P = demcz_make_mvn_params<LD>(c, kStepOpt & OPT_JMAJOR);
    // Programmatic dependent launch: consecutive step launches of a stream overlap the next launch's scheduling and
    // prologue with this one's tail; the kernel waits (griddepcontrol.wait) before it touches anything a predecessor wrote.
    static const bool pdl = !(getenv("DEMCZ_PDL") && atoi(getenv("DEMCZ_PDL")) == 0);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a, P);
}

template <int LD, bool DENSE>
cudaError_t launch_step_lay(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    if (full) return acc ? launch_step_t<LD, true, true, DENSE>(c, a) : launch_step_t<LD, false, true, DENSE>(c, a);
    return acc ? launch_step_t<LD, true, false, DENSE>(c, a) : launch_step_t<LD, false, false, DENSE>(c, a);
}

template <int LD>
cudaError_t launch_step_ld(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    if constexpr (zslot(LD) != LD) {            // dense slots (DEMCZ_XCHG_DENSE_SLOTS) only differ where the slot has padding
        if (c->ZS == LD) return launch_step_lay<LD, true>(c, a, acc, full);
    }
    return launch_step_lay<LD, false>(c, a, acc, full);
}

}  // namespace

#ifndef DEMCZ_LD_GROUP
#error "compile with -DDEMCZ_LD_GROUP=0..3"
#endif

#if DEMCZ_LD_GROUP == 0
cudaError_t demcz_launch_step_g0(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    switch (c->LD) {
        case 4:  return launch_step_ld<4>(c, a, acc, full);
        case 8:  return launch_step_ld<8>(c, a, acc, full);
        case 12: return launch_step_ld<12>(c, a, acc, full);
    }
    return cudaErrorInvalidValue;
}
#elif DEMCZ_LD_GROUP == 1
cudaError_t demcz_launch_step_g1(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    switch (c->LD) {
        case 16: return launch_step_ld<16>(c, a, acc, full);
        case 20: return launch_step_ld<20>(c, a, acc, full);
    }
    return cudaErrorInvalidValue;
}
#elif DEMCZ_LD_GROUP == 2
cudaError_t demcz_launch_step_g2(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    switch (c->LD) {
        case 24: return launch_step_ld<24>(c, a, acc, full);
        case 28: return launch_step_ld<28>(c, a, acc, full);
    }
    return cudaErrorInvalidValue;
}
#else
cudaError_t demcz_launch_step_g3(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    return c->LD == 32 ? launch_step_ld<32>(c, a, acc, full) : cudaErrorInvalidValue;
}
#endif
