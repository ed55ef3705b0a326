// sweeps_window2_d.cu -- kernel generation 2b (sweeps_window2_impl.cuh), families: PDE4.
#include "sweeps_window2_impl.cuh"

#ifndef W2_PROBE
PDEGPU_W2_FAMILY(4)
#endif
