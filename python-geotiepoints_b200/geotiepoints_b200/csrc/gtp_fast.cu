// gtp_fast.cu - float64 anchor-relative fast path (placeholder: not enabled yet).
#include "gtp_kernels.h"

bool gtp_cviirs_fast_f64_supported(const CviirsCfg&) { return false; }

cudaError_t gtp_launch_cviirs_fast_f64(const double*, const double*, const double*, double*, double*,
                                       long long, const CviirsCfg&, cudaStream_t)
{
    return cudaErrorNotSupported;
}
