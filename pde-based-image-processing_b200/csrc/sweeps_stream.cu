// sweeps_stream.cu -- dispatch of kernel generation 1. Returns PDEGPU_ERR_UNSUPPORTED for cases it has
// no kernel for (pdegpu_dev_relax then runs generation 0).
#include "stencil_math.cuh"

int relax_stream_line(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega);
int relax_stream_point(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega);

int relax_stream(pdegpu_ctx *ctx, const pdegpu_system *sys, int iter, float omega, int solver)
{
    if (solver == 2) return relax_stream_line(ctx, sys, iter, omega);
    return relax_stream_point(ctx, sys, iter, omega);
}
