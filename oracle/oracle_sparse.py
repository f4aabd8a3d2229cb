"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/oracle_sparse.c (the sparse CPU restatement).

Used by tests/ (mid-size parity) and by bench.py's cpu_baseline / --impl reference legs.  The reference itself is pure
Python over dense matrices and cannot hold the benchmark shapes; see the C file's header."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle_sparse.so")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class SparseOracle:
    """CSC + CSR copies (int64 pointers, int32 indices, uint16 codes) built with numpy on the host."""

    def __init__(self, n, m, cell, snv, code, L0, L1, varpos):
        cell = np.asarray(cell, np.int64)
        snv = np.asarray(snv, np.int64)
        code = np.asarray(code, np.uint16)
        self.n, self.m = int(n), int(m)
        o = np.lexsort((cell, snv))
        self.colptr = np.zeros(m + 1, np.int64)
        np.cumsum(np.bincount(snv, minlength=m), out=self.colptr[1:])
        self.row = np.ascontiguousarray(cell[o], np.int32)
        self.ccode = np.ascontiguousarray(code[o])
        o = np.lexsort((snv, cell))
        self.rowptr = np.zeros(n + 1, np.int64)
        np.cumsum(np.bincount(cell, minlength=n), out=self.rowptr[1:])
        self.col = np.ascontiguousarray(snv[o], np.int32)
        self.rcode = np.ascontiguousarray(code[o])
        self.L0 = np.ascontiguousarray(L0, np.float64)
        self.L1 = np.ascontiguousarray(L1, np.float64)
        self.varpos = np.ascontiguousarray(varpos, np.uint8)

    @classmethod
    def from_csc(cls, n, m, colptr, row, code, L0, L1, varpos):
        self = cls.__new__(cls)
        self.n, self.m = int(n), int(m)
        self.colptr = np.ascontiguousarray(colptr, np.int64)
        self.row = np.ascontiguousarray(row, np.int32)
        self.ccode = np.ascontiguousarray(code, np.uint16)
        self.rowptr = self.col = self.rcode = None
        self.L0 = np.ascontiguousarray(L0, np.float64)
        self.L1 = np.ascontiguousarray(L1, np.float64)
        self.varpos = np.ascontiguousarray(varpos, np.uint8)
        return self

    def snv_table(self, cell_label, Cn):
        m = self.m
        S0 = np.empty((m, Cn)); S1 = np.empty((m, Cn))
        No = np.empty((m, Cn), np.int32); Nv = np.empty((m, Cn), np.int32)
        lab = np.ascontiguousarray(cell_label, np.uint8)
        lib().orc_snv_table(C.c_int32(m), _p(self.colptr), _p(self.row), _p(self.ccode), _p(self.L0), _p(self.L1),
                            _p(self.varpos), _p(lab), C.c_int32(Cn), _p(S0), _p(S1), _p(No), _p(Nv))
        return S0, S1, No, Nv

    @classmethod
    def from_csc_csr(cls, n, m, colptr, row, ccode, rowptr, col, rcode, L0, L1, varpos):
        """Both orientations handed in ready-made (the full-size tests sort on the GPU)."""
        self = cls.from_csc(n, m, colptr, row, ccode, L0, L1, varpos)
        self.rowptr = np.ascontiguousarray(rowptr, np.int64)
        self.col = np.ascontiguousarray(col, np.int32)
        self.rcode = np.ascontiguousarray(rcode, np.uint16)
        return self

    def cell_table(self, snv_label, Q):
        n = self.n
        A0 = np.empty((n, Q)); A1 = np.empty((n, Q))
        No = np.empty((n, Q), np.int32); Nv = np.empty((n, Q), np.int32)
        lab = np.ascontiguousarray(snv_label, np.uint8)
        lib().orc_cell_table(C.c_int32(n), _p(self.rowptr), _p(self.col), _p(self.rcode), _p(self.L0), _p(self.L1),
                             _p(self.varpos), _p(lab), C.c_int32(Q), _p(A0), _p(A1), _p(No), _p(Nv))
        return A0, A1, No, Nv

    def assign_linear(self, cell_label, lenA):
        lab = np.ascontiguousarray(cell_label, np.uint8)
        out = np.empty(self.m, np.uint8); nobs = np.empty(self.m, np.int32)
        lib().orc_assign_linear(C.c_int32(self.m), _p(self.colptr), _p(self.row), _p(self.ccode), _p(self.L0), _p(self.L1),
                                _p(lab), C.c_double(lenA), _p(out), _p(nobs))
        return out, nobs

    def assign_branching(self, cell_label):
        lab = np.ascontiguousarray(cell_label, np.uint8)
        out = np.empty(self.m, np.uint8); nobs = np.empty((self.m, 2), np.int32)
        lib().orc_assign_branching(C.c_int32(self.m), _p(self.colptr), _p(self.row), _p(self.ccode), _p(self.L0),
                                   _p(self.L1), _p(lab), _p(out), _p(nobs))
        return out, nobs

    def tree_loglik(self, cell_node, snv_node, present):
        K = present.shape[0]
        pc = np.empty(self.n)
        cn = np.ascontiguousarray(cell_node, np.uint8); sn = np.ascontiguousarray(snv_node, np.uint8)
        pr = np.ascontiguousarray(present, np.uint8)
        lib().orc_tree_loglik(C.c_int32(self.n), _p(self.rowptr), _p(self.col), _p(self.rcode), _p(self.L0), _p(self.L1),
                              _p(cn), _p(sn), _p(pr), C.c_int32(K), _p(pc))
        return pc


def threads():
    return int(lib().orc_threads())


def use_all_cores():
    """All cores this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for its workers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_threads(int(n))
    return threads()
