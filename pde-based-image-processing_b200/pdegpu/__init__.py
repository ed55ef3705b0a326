"""pdegpu -- Python face of libpdegpu (B200-native variational PDE solver core).

Two layers, both thin:

* ``pdegpu.mex``  : the 13 MEX functions of the reference (``Oflow_sor_elin4_2d`` ...), same names,
  same argument order, ``single`` everywhere; they call the C gateways in gateways/pdegpu_mex.so,
  which call the C ABI (include/pdegpu.h). This is the drop-in surface.
* ``pdegpu.lib``  : ctypes binding of the C ABI itself (contexts, device-pointer entry points) for
  device-resident pipelines, benchmarks and multi-GPU sharding.

There is no CPU fallback: importing works anywhere, calling needs the built library and a B200.
"""
from . import mex_harness  # noqa: F401
from .mex_harness import MexError  # noqa: F401

__all__ = ["mex", "lib", "synth", "MexError"]
