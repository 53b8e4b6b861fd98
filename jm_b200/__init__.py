"""jm_b200 -- B200-native motion estimation + transform/quantisation for the JM 19.0 H.264 encoder.

The product is ``jm_b200/lib/libjmb200.so`` (hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/jmb200.h``) plus the C shim ``jm_b200/shim/`` that links it behind JM's own call sites.
This Python package is only the harness-side mirror of that C ABI (ctypes) used by tests and bench.py.
"""
