"""esloco_b200 -- B200-native engine for esloco's Monte-Carlo read-placement / target-tally path.

Host side (Python, mirrors the reference's module and function names for the hot path) over a C-ABI
shared library of hand-written sm_100a CUDA kernels (include/esloco_b200.h, esloco_b200/csrc/).
There is no CPU fallback: using the engine without the built library or without a B200 raises.
"""
__version__ = "0.1.0"
