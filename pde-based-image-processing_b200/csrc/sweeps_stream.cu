// sweeps_stream.cu -- kernel generation 1 (streaming relaxation kernels). Placeholder until the
// kernels land: reports "unsupported" so that pdegpu_dev_relax falls back to generation 0.
#include "stencil_math.cuh"

int relax_stream(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, int solver)
{
    (void)ctx; (void)sys; (void)iter; (void)omega; (void)solver;
    return PDEGPU_ERR_UNSUPPORTED;
}
