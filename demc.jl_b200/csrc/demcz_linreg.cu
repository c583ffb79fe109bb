// demcz_linreg.cu — linear-regression Gaussian log-likelihood (test/example_linreg.jl:32) path.
// Placeholder translation unit: the tcgen05 implementation lands in a later commit.
#include "demcz_internal.h"

extern "C" int demcz_set_likelihood_linreg(demcz_ctx* c, int64_t, int32_t, const double*, const double*)
{
    if (!c) return DEMCZ_ERR_INVALID;
    return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "linreg likelihood: not built yet");
}
int demcz_linreg_eval(demcz_ctx* c, const float*, int, int, float*) { return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "linreg: not built yet"); }
int demcz_linreg_generations(demcz_ctx* c, int64_t, int, bool) { return demcz_fail(c, DEMCZ_ERR_UNSUPPORTED, "linreg: not built yet"); }
void demcz_linreg_free(demcz_ctx*) {}
