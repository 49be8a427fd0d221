"""shark_b200 - B200-native implementation of Shark's k-mer Bloom-filter hot path.

The product is libshark_b200.so (hand-written sm_100a CUDA kernels behind the C ABI of
include/shark_b200.h) plus the `shark-b200` command-line drop-in; this package is the thin
Python host layer used by the tests and bench.py.  Build with `python shark_b200/build.py`.
"""
__version__ = "0.1.0"
