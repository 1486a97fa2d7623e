"""
tredparse_b200 — B200-native re-implementation of tredparse's genotyping hot path.

Host code in Python (same BamParser / IntegratedCaller / tred.py API and JSON layout as the
reference), hand-written sm_100a CUDA behind a C-ABI shared library (``libtredsw.so``, declared in
``include/tredsw.h``) that replaces the reference's ctypes binding to ``src/ssw.c``.
"""
__version__ = "0.1.0"
