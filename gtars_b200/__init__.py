"""gtars_b200 — B200-native (sm_100a) implementation of gtars' interval-overlap hot path.

`ffi` binds the C ABI of include/gtars_gpu.h (libgtars_gpu.so); the classes mirroring the reference's
public API (Tokenizer, MultiChromOverlapper, Igd, …) sit on top of it.  There is no CPU fallback.
"""
__version__ = "0.1.0"
