"""kangaroo_b200 -- B200-native (sm_100a) census / semi-global-matching engine behind the
operator surface of arpg/Kangaroo's `roo::` namespace.

  kangaroo_b200.capi   ctypes binding of include/roo_b200.h (the drop-in C ABI)
  kangaroo_b200.roo    Python mirror of the roo:: operators + the fused StereoEngine (needs torch + a GPU)
  kangaroo_b200.synth  deterministic synthetic stereo pairs for tests and bench.py
  kangaroo_b200.pxm    the applications' on-disk outputs (SavePXM P5/P7, .pdm), host side

The compute path is the CUDA library kangaroo_b200/lib/libroo_b200.so (sources in csrc/); there is
no CPU fallback: loading fails loudly if the library is missing.
"""
__version__ = "0.1.0"
