"""B200-native implementation of GeneEvolve's per-generation reproduction hot path.

The product is libgeneevolve_b200.so (hand-written sm_100a CUDA behind the C-ABI of
include/geneevolve_b200.h); this package holds its sources (csrc/), the ctypes binding (capi.py), the host
mirror of the reference's generation loop (simulation.py) and the multi-GPU plumbing (dist.py).
"""
from .capi import (Engine, Draws, GeneEvolveError, gen_params, load_library, GE_REP_BITS, GE_REP_SEGMENTS,  # noqa: F401
                   GE_RNG_PHILOX, GE_RNG_REPLAY)

__version__ = "0.1.0"
