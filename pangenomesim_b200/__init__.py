"""pangenomesim_b200 -- B200-native sequence-evolution engine for pangenomesim.

The product is the C-ABI shared library ``lib/libpgs.so`` (hand-written sm_100a CUDA kernels,
``include/pgs.h``) plus the C++ host program ``bin/pangenomesim-b200``.  This Python package is
only a thin ctypes binding used by the tests and by ``bench.py``; it has no compute path of its
own and raises if the CUDA library is missing.
"""
from .capi import Engine, PgsError, load_library, lib_path  # noqa: F401

__all__ = ["Engine", "PgsError", "load_library", "lib_path"]
