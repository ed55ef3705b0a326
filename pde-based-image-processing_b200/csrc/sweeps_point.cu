// sweeps_point.cu -- kernel generation 1, solver 1 (placeholder: not implemented yet).
#include "stencil_math.cuh"

int relax_stream_point(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega)
{
    (void)ctx; (void)sys; (void)iter; (void)omega;
    return PDEGPU_ERR_UNSUPPORTED;
}
