"""savont_b200 -- B200 (sm_100a) implementation of savont's read-vs-ASV assignment path.

The product is savont_b200/lib/libsavont_b200.so (C ABI in include/savont_b200.h); the Python
modules here are the ctypes doorway (_lib), the host layer mirroring the reference's function
names (api), the multi-GPU plumbing (dist) and the synthetic workload harness (synth)."""
__version__ = "0.1.0"
