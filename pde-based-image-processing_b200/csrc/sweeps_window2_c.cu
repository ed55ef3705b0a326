// sweeps_window2_c.cu -- kernel generation 2b (sweeps_window2_impl.cuh), families: 8-neighbour late-linearisation flow.
#include "sweeps_window2_impl.cuh"

#ifndef W2_PROBE
PDEGPU_W2_FAMILY(2)
#endif
