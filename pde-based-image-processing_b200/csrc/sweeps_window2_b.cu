// sweeps_window2_b.cu -- kernel generation 2b (sweeps_window2_impl.cuh), families: early-linearisation flow and disparity.
#include "sweeps_window2_impl.cuh"

#ifndef W2_PROBE
PDEGPU_W2_FAMILY(0)
PDEGPU_W2_FAMILY(3)
#endif
