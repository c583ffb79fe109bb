// demcz_step_launch.cu — instantiations and launch of the fused step kernel (demcz_step.cuh).  Compiled once per
// register width (-DDEMCZ_LD=4,8,...,32) so that the build parallelises: every (LD, ACC, FULL, DENSE, MOM)
// combination is its own kernel.
#include "demcz_step_host.h"
#include <cstdlib>

using namespace demcz;

namespace {

constexpr int kWarps = 4;

// production structure of the step kernel (see tools/step_tune.cu and profiles/ for how it was chosen)
constexpr int kStepOpt = DEMCZ_STEP_OPT;

// resident blocks per SM the register allocation is held to (4 warps per block); the moments variant keeps its
// cross-product accumulators in registers and runs one block fewer
template <int LD, bool MOM> constexpr int min_blocks() { return (LD <= 20 ? 4 : LD <= 28 ? 3 : 2) - (MOM ? 1 : 0); }

template <int LD, bool ACC, bool FULL, bool DENSE, bool MOM, bool PERS = false>
cudaError_t launch_step_t(demcz_ctx* c, const StepArgs& a)
{
    using SL = StageLayout<LD>;
    auto kern = demcz_step_mvn_kernel<LD, ACC, FULL, kWarps, min_blocks<LD, MOM>(), kStepOpt, DENSE, MOM, PERS>;
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
    // Not on the push-kernel transport: there the launch attribute costs the overlap of the push with the step kernel
    // (8 GPUs, C4: 49.9 G chain-updates/s with it, 70.6 G without; profiles/r2_c4_8gpu_pdl_*.json); with the copy engines
    // it is worth +2 % on 2 GPUs and nothing on 4.  DEMCZ_PDL=0/1 forces it.
    static const int pdl_env = getenv("DEMCZ_PDL") ? atoi(getenv("DEMCZ_PDL")) : -1;
    const bool pdl = pdl_env >= 0 ? pdl_env != 0 : (c->o.world == 1 || c->inkernel_poll);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a, P);
}

template <int LD, bool DENSE, bool MOM>
cudaError_t launch_step_lay(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    if (full) return acc ? launch_step_t<LD, true, true, DENSE, MOM>(c, a) : launch_step_t<LD, false, true, DENSE, MOM>(c, a);
    return acc ? launch_step_t<LD, true, false, DENSE, MOM>(c, a) : launch_step_t<LD, false, false, DENSE, MOM>(c, a);
}

template <int LD>
cudaError_t launch_step_ld(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    const bool mom = a.mom != nullptr;
    if (a.pers_K > 0) {                         // small-N persistent launch (one CTA; padded slots, no moments: demcz_run arranges that)
        if (full) return acc ? launch_step_t<LD, true, true, false, false, true>(c, a) : launch_step_t<LD, false, true, false, false, true>(c, a);
        return acc ? launch_step_t<LD, true, false, false, false, true>(c, a) : launch_step_t<LD, false, false, false, false, true>(c, a);
    }
    if constexpr (zslot(LD) != LD) {            // dense slots (DEMCZ_XCHG_DENSE_SLOTS) only differ where the slot has padding
        if (c->ZS == LD) return mom ? launch_step_lay<LD, true, true>(c, a, acc, full) : launch_step_lay<LD, true, false>(c, a, acc, full);
    }
    return mom ? launch_step_lay<LD, false, true>(c, a, acc, full) : launch_step_lay<LD, false, false>(c, a, acc, full);
}

}  // namespace

#ifndef DEMCZ_LD
#error "compile with -DDEMCZ_LD=4|8|...|32"
#endif

#define DEMCZ_CAT2(a, b) a##b
#define DEMCZ_CAT(a, b) DEMCZ_CAT2(a, b)
cudaError_t DEMCZ_CAT(demcz_launch_step_ld, DEMCZ_LD)(demcz_ctx* c, const StepArgs& a, bool acc, bool full)
{
    return launch_step_ld<DEMCZ_LD>(c, a, acc, full);
}
