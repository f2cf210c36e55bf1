// pd_synth.cu -- on-device synthetic cohort generator. Placeholder until the generator kernels land.
#include "pd_context.h"

extern "C" int pd_contig_synthesize(pd_ctx * c, uint64_t, uint64_t, double, uint32_t, const uint32_t *, const uint32_t *,
                                    const uint8_t *)
{
    if (!c) return PD_ERR_ARG;
    return pd_fail(c, PD_ERR_ARG, "pd_contig_synthesize: not available in this build");
}
