"""TEST INFRASTRUCTURE ONLY -- ctypes driver of the UNMODIFIED reference library built into oracle/_ref/ by oracle/Makefile.

Used by tests/golden/make_golden.py (fixture generation), by the live-reference parity tests, and by bench.py's
cpu_baseline / `--impl reference` legs.  Never imported by the product package (libceed_b200/).
Only the reference's public C API (include/ceed/ceed.h) is called.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
QF_DIR = os.path.join(os.path.dirname(HERE), "libceed_b200", "qfunctions")

MEM_HOST, MEM_DEVICE = 0, 1
COPY_VALUES, USE_POINTER = 0, 1
EVAL_NONE, EVAL_INTERP, EVAL_GRAD, EVAL_WEIGHT = 0, 1, 2, 16
GAUSS, GAUSS_LOBATTO = 0, 1
NOTRANSPOSE, TRANSPOSE = 0, 1


def available(cuda=False):
    return os.path.exists(os.path.join(REF_DIR, "lib-cuda" if cuda else "lib", "libceed.so"))


class RefCeed:
    """A Ceed of the reference library: RefCeed("/cpu/self/ref/serial")."""
    _libs = {}

    def __init__(self, resource="/cpu/self/ref/serial", cuda=False):
        key = "lib-cuda" if cuda else "lib"
        if key not in RefCeed._libs:
            RefCeed._libs[key] = C.CDLL(os.path.join(REF_DIR, key, "libceed.so"), mode=C.RTLD_GLOBAL)
        self.lib = lib = RefCeed._libs[key]
        self.bpqf = None
        self.ceed = C.c_void_p()
        self._chk(lib.CeedInit(resource.encode(), C.byref(self.ceed)))
        # abort-free error handling: return codes + stored message
        lib.CeedSetErrorHandler.argtypes = [C.c_void_p, C.c_void_p]
        lib.CeedSetErrorHandler(self.ceed, C.cast(lib.CeedErrorStore, C.c_void_p))
        self.VECTOR_ACTIVE = C.c_void_p.in_dll(lib, "CEED_VECTOR_ACTIVE")
        self.VECTOR_NONE = C.c_void_p.in_dll(lib, "CEED_VECTOR_NONE")
        self.BASIS_NONE = C.c_void_p.in_dll(lib, "CEED_BASIS_NONE")
        self.RSTR_NONE = C.c_void_p.in_dll(lib, "CEED_ELEMRESTRICTION_NONE")
        self.STRIDES_BACKEND = (C.c_int32 * 3).in_dll(lib, "CEED_STRIDES_BACKEND")
        self.REQUEST_IMMEDIATE = C.c_void_p.in_dll(lib, "CEED_REQUEST_IMMEDIATE")
        self._keep = []

    def _chk(self, code):
        if code:
            msg = C.c_char_p()
            try:
                self.lib.CeedGetErrorMessage(self.ceed, C.byref(msg))
            except Exception:
                pass
            raise RuntimeError(f"reference libCEED error {code}: {msg.value.decode() if msg.value else ''}")

    # ---- vectors
    def vector(self, n, array=None):
        v = C.c_void_p()
        self._chk(self.lib.CeedVectorCreate(self.ceed, C.c_ssize_t(int(n)), C.byref(v)))
        if array is not None:
            self.set_array(v, array)
        return v

    def set_array(self, v, array):
        a = np.ascontiguousarray(array, dtype=np.float64)
        self._chk(self.lib.CeedVectorSetArray(v, MEM_HOST, COPY_VALUES, a.ctypes.data_as(C.c_void_p)))

    def set_value(self, v, value):
        self.lib.CeedVectorSetValue.argtypes = [C.c_void_p, C.c_double]
        self._chk(self.lib.CeedVectorSetValue(v, float(value)))

    def get_array(self, v, n):
        p = C.POINTER(C.c_double)()
        self._chk(self.lib.CeedVectorGetArrayRead(v, MEM_HOST, C.byref(p)))
        out = np.ctypeslib.as_array(p, shape=(int(n),)).copy() if n else np.zeros(0)
        self._chk(self.lib.CeedVectorRestoreArrayRead(v, C.byref(p)))
        return out

    # ---- restrictions / bases
    def restriction(self, nelem, elemsize, ncomp, compstride, lsize, offsets):
        r = C.c_void_p()
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        self._chk(self.lib.CeedElemRestrictionCreate(self.ceed, nelem, elemsize, ncomp, int(compstride), C.c_ssize_t(int(lsize)), MEM_HOST,
                                                     COPY_VALUES, off.ctypes.data_as(C.c_void_p), C.byref(r)))
        return r

    def restriction_strided(self, nelem, elemsize, ncomp, lsize, strides=None):
        r = C.c_void_p()
        s = self.STRIDES_BACKEND if strides is None else (C.c_int32 * 3)(*[int(x) for x in strides])
        self._chk(self.lib.CeedElemRestrictionCreateStrided(self.ceed, nelem, elemsize, ncomp, C.c_ssize_t(int(lsize)), s, C.byref(r)))
        return r

    def restriction_apply(self, r, tmode, u, v):
        self._chk(self.lib.CeedElemRestrictionApply(r, tmode, u, v, self.REQUEST_IMMEDIATE))

    def restriction_e_layout(self, r):
        lay = (C.c_int32 * 3)()
        self._chk(self.lib.CeedElemRestrictionGetELayout(r, lay))
        return tuple(lay)

    def basis_lagrange(self, dim, ncomp, P, Q, qmode):
        b = C.c_void_p()
        self._chk(self.lib.CeedBasisCreateTensorH1Lagrange(self.ceed, dim, ncomp, P, Q, qmode, C.byref(b)))
        return b

    def basis_matrices(self, b, P, Q):
        out = {}
        for name, fn, n in (("interp_1d", "CeedBasisGetInterp1D", P * Q), ("grad_1d", "CeedBasisGetGrad1D", P * Q),
                            ("q_ref_1d", "CeedBasisGetQRef", Q), ("q_weight_1d", "CeedBasisGetQWeights", Q)):
            p = C.POINTER(C.c_double)()
            self._chk(getattr(self.lib, fn)(b, C.byref(p)))
            out[name] = np.ctypeslib.as_array(p, shape=(n,)).copy()
        if Q >= P:
            cg = np.zeros(Q * Q)
            self._chk(self.lib.CeedBasisGetCollocatedGrad(b, cg.ctypes.data_as(C.c_void_p)))
            out["collo_grad_1d"] = cg
        return out

    def basis_apply(self, b, nelem, tmode, emode, u, v):
        self._chk(self.lib.CeedBasisApply(b, nelem, tmode, emode, u if u is not None else self.VECTOR_NONE, v))

    # ---- qfunctions / operators
    def qfunction_by_name(self, name):
        qf = C.c_void_p()
        self._chk(self.lib.CeedQFunctionCreateInteriorByName(self.ceed, name.encode(), C.byref(qf)))
        return qf

    def qfunction_bp(self, name, header):
        """User QFunction from libceed_b200/qfunctions/<header> (CPU function pointer from oracle/_ref/lib/libbpqf.so)."""
        if self.bpqf is None:
            self.bpqf = C.CDLL(os.path.join(REF_DIR, "lib", "libbpqf.so"))
            self.bpqf.bpqf_get.restype = C.c_void_p
            self.bpqf.bpqf_get.argtypes = [C.c_char_p]
        f = self.bpqf.bpqf_get(name.encode())
        assert f, name
        qf = C.c_void_p()
        src = f"{os.path.join(QF_DIR, header)}:{name}"
        self.lib.CeedQFunctionCreateInterior.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_char_p, C.c_void_p]
        self._chk(self.lib.CeedQFunctionCreateInterior(self.ceed, 1, C.c_void_p(f), src.encode(), C.byref(qf)))
        return qf

    def qf_add_input(self, qf, name, size, emode):
        self._chk(self.lib.CeedQFunctionAddInput(qf, name.encode(), size, emode))

    def qf_add_output(self, qf, name, size, emode):
        self._chk(self.lib.CeedQFunctionAddOutput(qf, name.encode(), size, emode))

    def operator(self, qf):
        op = C.c_void_p()
        self._chk(self.lib.CeedOperatorCreate(self.ceed, qf, None, None, C.byref(op)))
        return op

    def op_set_field(self, op, name, rstr, basis, vec):
        self._chk(self.lib.CeedOperatorSetField(op, name.encode(), rstr if rstr is not None else self.RSTR_NONE,
                                                basis if basis is not None else self.BASIS_NONE, vec))

    def op_apply(self, op, u, v):
        self._chk(self.lib.CeedOperatorApply(op, u, v, self.REQUEST_IMMEDIATE))

    def op_apply_add(self, op, u, v):
        self._chk(self.lib.CeedOperatorApplyAdd(op, u, v, self.REQUEST_IMMEDIATE))


BP_TABLE = {  # bp: (ncomp, is_diff, q_extra, qmode, setup, apply, ncq) -- same table as libceed_b200/bp.py
    1: (1, False, 2, GAUSS, "BPSetupMassGeo", "BPMass", 1),
    2: (3, False, 2, GAUSS, "BPSetupMassGeo", "BPMass3", 1),
    3: (1, True, 2, GAUSS, "BPSetupDiffGeo", "BPDiff", 7),
    4: (3, True, 2, GAUSS, "BPSetupDiffGeo", "BPDiff3", 7),
    5: (1, True, 1, GAUSS_LOBATTO, "BPSetupDiffGeo", "BPDiff", 7),
    6: (3, True, 1, GAUSS_LOBATTO, "BPSetupDiffGeo", "BPDiff3", 7),
}


class RefBP:
    """The same BP operator wiring as libceed_b200.bp.BPProblem, on the reference library."""

    def __init__(self, rc, bp, p, num_elem, num_nodes, offsets, coords, gallery=False, interlaced=False):
        ncomp, is_diff, q_extra, qmode, setup_name, apply_name, ncq = BP_TABLE[bp]
        if gallery:
            assert ncomp == 1
            ncq = 6 if is_diff else 1
        P, Q = p + 1, p + q_extra
        self.rc, self.ncomp, self.ncq, self.P, self.Q, self.num_elem, self.num_nodes = rc, ncomp, ncq, P, Q, num_elem, num_nodes
        nn = num_nodes
        offsets = np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1)
        self.rstr_x = rc.restriction(num_elem, P ** 3, 3, nn, 3 * nn, offsets)
        if interlaced and ncomp > 1:
            self.rstr_u = rc.restriction(num_elem, P ** 3, ncomp, 1, ncomp * nn, offsets * ncomp)
        else:
            self.rstr_u = rc.restriction(num_elem, P ** 3, ncomp, nn, ncomp * nn, offsets)
        self.qd_len = num_elem * Q ** 3 * ncq
        self.rstr_qd = rc.restriction_strided(num_elem, Q ** 3, ncq, self.qd_len, None)
        self.basis_x = rc.basis_lagrange(3, 3, P, Q, qmode)
        self.basis_u = rc.basis_lagrange(3, ncomp, P, Q, qmode)
        self.x = rc.vector(3 * nn, coords.reshape(-1))
        self.qdata = rc.vector(self.qd_len)
        if gallery:
            qs = rc.qfunction_by_name("Poisson3DBuild" if is_diff else "Mass3DBuild")
        else:
            qs = rc.qfunction_bp(setup_name, "bp_geo.h")
            rc.qf_add_input(qs, "x", 3, EVAL_INTERP)
            rc.qf_add_input(qs, "dx", 9, EVAL_GRAD)
            rc.qf_add_input(qs, "weight", 1, EVAL_WEIGHT)
            rc.qf_add_output(qs, "qdata", ncq, EVAL_NONE)
        self.op_setup = rc.operator(qs)
        if not gallery:
            rc.op_set_field(self.op_setup, "x", self.rstr_x, self.basis_x, rc.VECTOR_ACTIVE)
        rc.op_set_field(self.op_setup, "dx", self.rstr_x, self.basis_x, rc.VECTOR_ACTIVE)
        rc.op_set_field(self.op_setup, "weights" if gallery else "weight", None, self.basis_x, rc.VECTOR_NONE)
        rc.op_set_field(self.op_setup, "qdata", self.rstr_qd, None, rc.VECTOR_ACTIVE)
        rc.op_apply(self.op_setup, self.x, self.qdata)
        if gallery:
            qa = rc.qfunction_by_name("Poisson3DApply" if is_diff else "MassApply")
            names = ("du", "qdata", "dv") if is_diff else ("u", "qdata", "v")
        else:
            qa = rc.qfunction_bp(apply_name, "bp_apply.h")
            names = ("u", "qdata", "v")
            emode, size = (EVAL_GRAD, 3 * ncomp) if is_diff else (EVAL_INTERP, ncomp)
            rc.qf_add_input(qa, "u", size, emode)
            rc.qf_add_input(qa, "qdata", ncq, EVAL_NONE)
            rc.qf_add_output(qa, "v", size, emode)
        self.op = rc.operator(qa)
        rc.op_set_field(self.op, names[0], self.rstr_u, self.basis_u, rc.VECTOR_ACTIVE)
        rc.op_set_field(self.op, names[1], self.rstr_qd, None, self.qdata)
        rc.op_set_field(self.op, names[2], self.rstr_u, self.basis_u, rc.VECTOR_ACTIVE)
        self.u = rc.vector(ncomp * nn)
        self.v = rc.vector(ncomp * nn)

    def apply(self, u):
        self.rc.set_array(self.u, u)
        self.rc.op_apply(self.op, self.u, self.v)
        return self.rc.get_array(self.v, self.ncomp * self.num_nodes)

    def qdata_array(self):
        """qdata as [comp][elem][qpt] whatever the backend's own strided layout is."""
        rc = self.rc
        ev = rc.vector(self.qd_len)
        rc.restriction_apply(self.rstr_qd, NOTRANSPOSE, self.qdata, ev)
        lay = rc.restriction_e_layout(self.rstr_qd)  # strides for (node, comp, elem)
        e = rc.get_array(ev, self.qd_len)
        nq = self.Q ** 3
        n, c, el = np.meshgrid(np.arange(nq), np.arange(self.ncq), np.arange(self.num_elem), indexing="ij")
        out = np.zeros((self.ncq, self.num_elem, nq))
        out[c, el, n] = e[n * lay[0] + c * lay[1] + el * lay[2]]
        return out.reshape(-1)
