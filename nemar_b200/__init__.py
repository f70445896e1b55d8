"""nemar_b200 — B200-native engine for NeMAR's training hot path (NEMARModel.optimize_parameters).

csrc/     hand-written sm_100a CUDA behind the C ABI in include/nemar_b200.h
engine/   ctypes binding, autograd glue, flat Adam, one-process-per-GPU data parallelism
models/, options/, data/, util/   host-side mirror of the reference's plugin surface
"""
__version__ = "0.1.0"
