"""graphphysics_b200 -- B200-native (sm_100a) message-passing hot path of graph-physics.

Host-side mirror of the reference's `graphphysics/models` API over hand-written CUDA kernels
(libgp_b200.so, C ABI in include/gp_b200.h).  No CPU fallback.
"""
__version__ = "0.1.0"
