"""The drop-in boundary without a GPU: libtredsw.so loads, exports every entry point that
include/tredsw.h declares (parsed from the header itself), the ctypes mirrors have the C layouts, the
pure-host legacy helpers behave like src/ssw.c:878-904, and without a CUDA device every compute entry
point fails loudly (no CPU fallback).  No compute is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from tredparse_b200 import _lib, cohort

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tredsw.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)   # struct bodies hold no prototypes
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_declares_what_the_binding_lists():
    decl = _declared_functions()
    assert set(decl) == set(_lib.EXPORTS), (sorted(set(decl) ^ set(_lib.EXPORTS)))
    # the six libssw symbols the reference binds (src/ssw_wrap.py:69-83,274-280)
    for n in ("ssw_init", "init_destroy", "ssw_align", "align_destroy", "cigar_int_to_op", "cigar_int_to_len"):
        assert n in decl


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared_functions():
        assert hasattr(lib, name), "libtredsw.so lacks " + name
    assert lib.tredsw_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_struct_layouts_match_the_header():
    class SAlign(ctypes.Structure):            # s_align, src/ssw.h:42-52 == CAlignRes, src/ssw_wrap.py:43-51
        _fields_ = [("score1", ctypes.c_uint16), ("score2", ctypes.c_uint16), ("ref_begin1", ctypes.c_int32),
                    ("ref_end1", ctypes.c_int32), ("read_begin1", ctypes.c_int32), ("read_end1", ctypes.c_int32),
                    ("ref_end2", ctypes.c_int32), ("cigar", ctypes.POINTER(ctypes.c_uint32)),
                    ("cigarLen", ctypes.c_int32)]
    assert ctypes.sizeof(SAlign) == 40 and SAlign.cigar.offset == 24
    assert ctypes.sizeof(_lib.Family) == 128 == _lib.FAMILY_DTYPE.itemsize
    assert ctypes.sizeof(_lib.GridProblem) == 16 * 4 + 5 * 8 + 10 * 8 == _lib.GRID_PROBLEM_DTYPE.itemsize
    assert ctypes.sizeof(_lib.GridResult) == 48 == _lib.GRID_RESULT_DTYPE.itemsize
    assert _lib.POSTERIOR_DTYPE.itemsize == 24
    assert cohort.PROBLEM_DTYPE.itemsize == 40 and cohort.LOCUS_DTYPE.itemsize == 32
    assert cohort.CALL_DTYPE.itemsize == 64
    assert ctypes.sizeof(cohort.Cohort) % 8 == 0
    # field offsets the C side relies on (include/tredsw.h: tredsw_cohort)
    assert cohort.Cohort.families.offset == 64 and cohort.Cohort.stutter_w.offset == 88


def test_cigar_helpers_are_host_functions():
    lib = _lib.load()
    lib.cigar_int_to_len.restype = ctypes.c_uint32
    lib.cigar_int_to_len.argtypes = [ctypes.c_uint32]
    lib.cigar_int_to_op.restype = ctypes.c_char
    lib.cigar_int_to_op.argtypes = [ctypes.c_uint32]
    for length, op, ch in ((37, 0, b"M"), (2, 1, b"I"), (5, 2, b"D")):       # src/ssw.h:132-170
        w = (length << 4) | op
        assert lib.cigar_int_to_len(w) == length and lib.cigar_int_to_op(w) == ch


def test_no_cpu_fallback_without_a_device():
    lib = _lib.load()
    if lib.tredsw_device_count() > 0:
        pytest.skip("a CUDA device is visible: covered by the -m gpu tests")
    h = lib.tredsw_create(0, None)
    assert not h
    assert "no CUDA device" in _lib.last_error()
    with pytest.raises(_lib.TredswError):
        _lib.Context(0)
    # the product API must raise, not fall back to the oracle or any CPU path
    from tredparse_b200 import ssw
    with pytest.raises(_lib.TredswError):
        ssw.align_pairs(["ACGT" * 10], ["ACGT" * 10], np.zeros(1, np.int32), np.zeros(1, np.int32))


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "tredparse_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_python_mirrors_equal_the_compiled_header(tmp_path):
    """sizeof / offsetof as gcc sees include/tredsw.h == the ctypes / numpy mirrors the host side marshals with."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "tredsw.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(s_align), sizeof(tredsw_family), sizeof(tredsw_grid_problem),
           sizeof(tredsw_grid_result), sizeof(tredsw_posterior), sizeof(tredsw_problem), sizeof(tredsw_locus),
           sizeof(tredsw_cohort), sizeof(tredsw_call));
    printf("%zu %zu %zu %zu %zu %zu\\n", offsetof(tredsw_cohort, families), offsetof(tredsw_cohort, stutter_w),
           offsetof(tredsw_cohort, input_flags), offsetof(tredsw_cohort, norepeatpairs), offsetof(tredsw_cohort, n_bases),
           offsetof(tredsw_cohort, read_name));
    printf("%zu %zu\\n", offsetof(tredsw_grid_result, sum_uniq), offsetof(tredsw_posterior, p));
    return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    lines = subprocess.check_output([str(exe)]).decode().split("\n")
    sizes = [int(x) for x in lines[0].split()]
    C = cohort.Cohort
    assert sizes == [40, ctypes.sizeof(_lib.Family), ctypes.sizeof(_lib.GridProblem), ctypes.sizeof(_lib.GridResult),
                     _lib.POSTERIOR_DTYPE.itemsize, cohort.PROBLEM_DTYPE.itemsize, cohort.LOCUS_DTYPE.itemsize,
                     ctypes.sizeof(C), cohort.CALL_DTYPE.itemsize]
    assert [int(x) for x in lines[1].split()] == [C.families.offset, C.stutter_w.offset, C.input_flags.offset,
                                                  C.norepeatpairs.offset, C.n_bases.offset, C.read_name.offset]
    assert [int(x) for x in lines[2].split()] == [_lib.GRID_RESULT_DTYPE.fields["sum_uniq"][1],
                                                  _lib.POSTERIOR_DTYPE.fields["p"][1]]
