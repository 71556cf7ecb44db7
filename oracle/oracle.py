"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libceed_oracle.so (ceed_oracle.c), the CPU restatement of the
reference's operator-apply path.  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libceed_oracle.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_restriction_offset.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, ip, C.c_int, dp, dp]
        _lib.oracle_restriction_offset.restype = None
        _lib.oracle_restriction_strided.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int, dp, dp]
        _lib.oracle_restriction_strided.restype = None
        _lib.oracle_tensor_contract.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, dp, dp]
        _lib.oracle_tensor_contract.restype = None
        _lib.oracle_bp_qdata.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, ip, dp, dp]
        _lib.oracle_bp_apply.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, ip, C.c_int, dp, dp, dp, C.c_int]
    return _lib


def _d(a):
    return a.ctypes.data_as(dp)


def gauss(Q):
    x, w = np.zeros(Q), np.zeros(Q)
    lib().oracle_gauss_quadrature(Q, _d(x), _d(w))
    return x, w


def lobatto(Q):
    x, w = np.zeros(Q), np.zeros(Q)
    lib().oracle_lobatto_quadrature(Q, _d(x), _d(w))
    return x, w


def lagrange_1d(P, Q, qmode):
    interp, grad, qref, qw = np.zeros(P * Q), np.zeros(P * Q), np.zeros(Q), np.zeros(Q)
    assert lib().oracle_lagrange_1d(P, Q, qmode, _d(interp), _d(grad), _d(qref), _d(qw)) == 0
    return interp, grad, qref, qw


def collocated_grad(P, Q, interp, grad):
    out = np.zeros(Q * Q)
    assert lib().oracle_collocated_grad(P, Q, _d(np.ascontiguousarray(interp)), _d(np.ascontiguousarray(grad)), _d(out)) == 0
    return out


def restriction_offset(num_elem, elem_size, num_comp, comp_stride, offsets, transpose, u, out_len):
    """returns E-vector [elem][comp][node] (gather) or the L-vector (scatter-add into zeros)."""
    off = np.ascontiguousarray(offsets, dtype=np.int32)
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.zeros(out_len)
    lib().oracle_restriction_offset(num_elem, elem_size, num_comp, int(comp_stride), off.ctypes.data_as(ip), int(transpose), _d(u), _d(v))
    return v


def restriction_strided(num_elem, elem_size, num_comp, strides, transpose, u, out_len):
    s = (C.c_int64 * 3)(*[int(x) for x in strides])
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.zeros(out_len)
    lib().oracle_restriction_strided(num_elem, elem_size, num_comp, s, int(transpose), _d(u), _d(v))
    return v


def bp_sizes(bp, gallery, p):
    P, Q, nc, ncq = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert lib().oracle_bp_sizes(bp, int(gallery), p, C.byref(P), C.byref(Q), C.byref(nc), C.byref(ncq)) == 0
    return P.value, Q.value, nc.value, ncq.value


def bp_qdata(bp, p, offsets, coords, gallery=False):
    """qdata [comp][elem][qpt] of the BP setup operator."""
    P, Q, nc, ncq = bp_sizes(bp, gallery, p)
    off = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1)
    num_elem = off.size // P ** 3
    coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1)
    num_nodes = coords.size // 3
    qd = np.zeros(num_elem * Q ** 3 * ncq)
    assert lib().oracle_bp_qdata(bp, int(gallery), p, num_elem, num_nodes, off.ctypes.data_as(ip), _d(coords), _d(qd)) == 0
    return qd


def bp_apply(bp, p, offsets, num_nodes, qdata, u, gallery=False, interlaced=False, v0=None):
    P, Q, nc, ncq = bp_sizes(bp, gallery, p)
    off = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1)
    num_elem = off.size // P ** 3
    u = np.ascontiguousarray(u, dtype=np.float64)
    qdata = np.ascontiguousarray(qdata, dtype=np.float64)
    v = np.zeros(nc * num_nodes) if v0 is None else np.array(v0, dtype=np.float64)
    assert lib().oracle_bp_apply(bp, int(gallery), p, num_elem, int(num_nodes), off.ctypes.data_as(ip), int(interlaced), _d(qdata), _d(u), _d(v),
                                 0 if v0 is None else 1) == 0
    return v
