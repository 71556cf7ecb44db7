"""Host-side mirror of the reference's object interface for the operator-apply path, over the C ABI.

Names and argument meaning follow the reference's Python binding (python/ceed.py, ceed_vector.py,
ceed_elemrestriction.py, ceed_basis.py, ceed_qfunction.py, ceed_operator.py in /root/reference): a `Ceed` context
creates Vectors, ElemRestrictions, Bases, QFunctions and Operators; `Operator.apply(u, v)` is CeedOperatorApply.
Errors raise `CeedError` carrying the backend's message (the reference's error handler prints the same text).
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from ._lib import (COPY_VALUES, EVAL_GRAD, EVAL_INTERP, EVAL_NONE, EVAL_WEIGHT, GAUSS, GAUSS_LOBATTO, MEM_DEVICE,  # noqa: F401
                   MEM_HOST, NORM_1, NORM_2, NORM_MAX, NOTRANSPOSE, OWN_POINTER, SCATTER_ATOMIC, SCATTER_DETERMINISTIC,
                   SCATTER_EVECTOR, TRANSPOSE, USE_POINTER)

QFUNCTION_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "qfunctions")

VECTOR_ACTIVE = C.c_void_p(1)
VECTOR_NONE = C.c_void_p(0)
ELEMRESTRICTION_NONE = None
BASIS_NONE = None


class CeedError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"ceed-b200 error {code}: {message}")
        self.code = code


class _Object:
    _destroy = None

    def __init__(self, ceed, ptr):
        self._ceed = ceed
        self._ptr = ptr

    def _chk(self, code):
        self._ceed._chk(code)

    def __del__(self):
        try:
            if self._ptr and self._destroy and L._lib is not None:
                getattr(L.lib(), self._destroy)(self._ptr)
                self._ptr = None
        except Exception:
            pass


class Ceed:
    """CeedInit("/gpu/cuda/b200:device_id=N") equivalent (interface/ceed.c:1175)."""

    def __init__(self, resource="/gpu/cuda/b200"):
        self._lib = L.lib()
        device_id = 0
        if ":device_id=" in resource:
            device_id = int(resource.split(":device_id=")[1])
        root = resource.split(":")[0]
        if not "/gpu/cuda/b200".startswith(root.rstrip("/")) and root != "/gpu/cuda/b200":
            raise CeedError(-3, f"No suitable backend: {resource}")
        self._ptr = C.c_void_p()
        code = self._lib.ceedb200_init(device_id, C.byref(self._ptr))
        if code:
            raise CeedError(code, self._lib.ceedb200_last_error(None).decode())
        self._lib.ceedb200_add_jit_source_root(self._ptr, QFUNCTION_DIR.encode())

    def _chk(self, code):
        if code:
            raise CeedError(code, self._lib.ceedb200_last_error(self._ptr).decode())

    def __del__(self):
        try:
            if self._ptr:
                self._lib.ceedb200_destroy(self._ptr)
                self._ptr = None
        except Exception:
            pass

    # -- context configuration
    def add_jit_source_root(self, path):
        self._chk(self._lib.ceedb200_add_jit_source_root(self._ptr, str(path).encode()))

    def add_jit_define(self, define):
        self._chk(self._lib.ceedb200_add_jit_define(self._ptr, str(define).encode()))

    def set_stream(self, stream_ptr):
        self._chk(self._lib.ceedb200_set_stream(self._ptr, C.c_void_p(stream_ptr)))

    def set_scatter_mode(self, mode):
        self._chk(self._lib.ceedb200_set_scatter_mode(self._ptr, mode))

    def set_autotune(self, level=1):
        """0 off, 1 tune fused operators that have no tuning-table entry on their first apply, 2 always."""
        self._chk(self._lib.ceedb200_set_autotune(self._ptr, level))

    def synchronize(self):
        self._chk(self._lib.ceedb200_synchronize(self._ptr))

    def launch_count(self):
        return int(self._lib.ceedb200_launch_count(self._ptr))

    # -- object factories
    def Vector(self, size):
        return Vector(self, size)

    def ElemRestriction(self, nelem, elemsize, ncomp, compstride, lsize, offsets):
        return ElemRestriction(self, nelem, elemsize, ncomp, compstride, lsize, offsets)

    def StridedElemRestriction(self, nelem, elemsize, ncomp, lsize, strides=None):
        return ElemRestriction(self, nelem, elemsize, ncomp, 0, lsize, None, strides=strides, strided=True)

    def BasisTensorH1Lagrange(self, dim, ncomp, P, Q, qmode):
        return Basis(self, dim, ncomp, P, Q, qmode=qmode)

    def BasisTensorH1(self, dim, ncomp, P1d, Q1d, interp1d, grad1d, qref1d, qweight1d):
        return Basis(self, dim, ncomp, P1d, Q1d, matrices=(interp1d, grad1d, qref1d, qweight1d))

    def QFunction(self, source, name):
        return QFunction(self, source, name)

    def QFunctionContext(self):
        return QFunctionContext(self)

    def Operator(self, qf):
        return Operator(self, qf)


class Vector(_Object):
    _destroy = "ceedb200_vector_destroy"

    def __init__(self, ceed, size):
        ptr = C.c_void_p()
        ceed._chk(ceed._lib.ceedb200_vector_create(ceed._ptr, int(size), C.byref(ptr)))
        super().__init__(ceed, ptr)
        self.size = int(size)
        self._keep = None

    def __len__(self):
        return self.size

    def set_array(self, array, memtype=MEM_HOST, cmode=COPY_VALUES):
        """numpy array (host) or raw device pointer / object with data_ptr() (device)."""
        lib = self._ceed._lib
        if memtype == MEM_HOST:
            arr = np.ascontiguousarray(array, dtype=np.float64)
            assert arr.size == self.size
            if cmode != COPY_VALUES:
                self._keep = arr
                cmode = USE_POINTER
            self._chk(lib.ceedb200_vector_set_array(self._ptr, MEM_HOST, cmode, arr.ctypes.data_as(C.c_void_p)))
        else:
            dptr = array.data_ptr() if hasattr(array, "data_ptr") else int(array)
            if cmode != COPY_VALUES:
                self._keep = array
                cmode = USE_POINTER
            self._chk(lib.ceedb200_vector_set_array(self._ptr, MEM_DEVICE, cmode, C.c_void_p(dptr)))

    def take_array(self, memtype=MEM_HOST):
        out = C.c_void_p()
        self._chk(self._ceed._lib.ceedb200_vector_take_array(self._ptr, memtype, C.byref(out)))
        keep, self._keep = self._keep, None
        return keep

    def set_value(self, value):
        self._chk(self._ceed._lib.ceedb200_vector_set_value(self._ptr, float(value)))

    def sync_array(self, memtype=MEM_HOST):
        self._chk(self._ceed._lib.ceedb200_vector_sync_array(self._ptr, memtype))

    def get_array_read(self, memtype=MEM_HOST):
        """Host: numpy copy of the data.  Device: raw device pointer (int)."""
        out = C.c_void_p()
        self._chk(self._ceed._lib.ceedb200_vector_get_array_read(self._ptr, memtype, C.byref(out)))
        if memtype == MEM_DEVICE:
            return out.value
        if self.size == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(self.size,)).copy()

    to_numpy = get_array_read

    def device_pointer(self, write=False):
        out = C.c_void_p()
        fn = self._ceed._lib.ceedb200_vector_get_array_write if write else self._ceed._lib.ceedb200_vector_get_array
        self._chk(fn(self._ptr, MEM_DEVICE, C.byref(out)))
        return out.value

    def norm(self, normtype=NORM_2):
        out = C.c_double()
        self._chk(self._ceed._lib.ceedb200_vector_norm(self._ptr, normtype, C.byref(out)))
        return out.value

    def scale(self, alpha):
        self._chk(self._ceed._lib.ceedb200_vector_scale(self._ptr, float(alpha)))

    def reciprocal(self):
        self._chk(self._ceed._lib.ceedb200_vector_reciprocal(self._ptr))

    def axpy(self, alpha, x):
        self._chk(self._ceed._lib.ceedb200_vector_axpy(self._ptr, float(alpha), x._ptr))

    def axpby(self, alpha, beta, x):
        self._chk(self._ceed._lib.ceedb200_vector_axpby(self._ptr, float(alpha), float(beta), x._ptr))

    def pointwise_mult(self, x, y):
        self._chk(self._ceed._lib.ceedb200_vector_pointwise_mult(self._ptr, x._ptr, y._ptr))


class ElemRestriction(_Object):
    _destroy = "ceedb200_restriction_destroy"

    def __init__(self, ceed, nelem, elemsize, ncomp, compstride, lsize, offsets, strides=None, strided=False):
        ptr = C.c_void_p()
        lib = ceed._lib
        if strided:
            s = None if strides is None else (C.c_int32 * 3)(*[int(x) for x in strides])
            ceed._chk(lib.ceedb200_restriction_create_strided(ceed._ptr, nelem, elemsize, ncomp, int(lsize), s, C.byref(ptr)))
        else:
            off = np.ascontiguousarray(offsets, dtype=np.int32)
            assert off.size == nelem * elemsize
            ceed._chk(lib.ceedb200_restriction_create(ceed._ptr, nelem, elemsize, ncomp, int(compstride), int(lsize), MEM_HOST,
                                                      COPY_VALUES, off.ctypes.data_as(C.c_void_p), C.byref(ptr)))
        super().__init__(ceed, ptr)
        self.nelem, self.elemsize, self.ncomp, self.lsize = nelem, elemsize, ncomp, int(lsize)
        self.esize = nelem * elemsize * ncomp

    def apply(self, u, v, tmode=NOTRANSPOSE):
        self._chk(self._ceed._lib.ceedb200_restriction_apply(self._ptr, tmode, u._ptr, v._ptr))

    def set_split(self, split_elem):
        self._chk(self._ceed._lib.ceedb200_restriction_set_split(self._ptr, int(split_elem)))

    def T_apply(self, u, v):
        self.apply(u, v, TRANSPOSE)

    def get_e_layout(self):
        lay = (C.c_int32 * 3)()
        self._chk(self._ceed._lib.ceedb200_restriction_get_e_layout(self._ptr, lay))
        return tuple(lay)

    def create_evector(self):
        return Vector(self._ceed, self.esize)

    def create_lvector(self):
        return Vector(self._ceed, self.lsize)


class Basis(_Object):
    _destroy = "ceedb200_basis_destroy"

    def __init__(self, ceed, dim, ncomp, P, Q, qmode=None, matrices=None):
        ptr = C.c_void_p()
        lib = ceed._lib
        if matrices is None:
            ceed._chk(lib.ceedb200_basis_create_tensor_h1_lagrange(ceed._ptr, dim, ncomp, P, Q, qmode, C.byref(ptr)))
        else:
            m = [np.ascontiguousarray(x, dtype=np.float64) for x in matrices]
            ceed._chk(lib.ceedb200_basis_create_tensor_h1(ceed._ptr, dim, ncomp, P, Q, *[x.ctypes.data_as(L.c_scalar_p) for x in m],
                                                          C.byref(ptr)))
        super().__init__(ceed, ptr)
        self.dim, self.ncomp, self.P, self.Q = dim, ncomp, P, Q

    def _matrix(self, which, n):
        out = np.zeros(n)
        self._chk(self._ceed._lib.ceedb200_basis_get_matrix(self._ptr, which, out.ctypes.data_as(L.c_scalar_p)))
        return out

    interp_1d = property(lambda s: s._matrix(0, s.P * s.Q).reshape(s.Q, s.P))
    grad_1d = property(lambda s: s._matrix(1, s.P * s.Q).reshape(s.Q, s.P))
    q_ref_1d = property(lambda s: s._matrix(2, s.Q))
    q_weight_1d = property(lambda s: s._matrix(3, s.Q))
    collo_grad_1d = property(lambda s: s._matrix(4, s.Q * s.Q).reshape(s.Q, s.Q))

    def apply(self, nelem, emode, u, v, tmode=NOTRANSPOSE):
        uptr = u._ptr if u is not None else VECTOR_NONE
        self._chk(self._ceed._lib.ceedb200_basis_apply(self._ptr, nelem, tmode, emode, uptr, v._ptr))

    def apply_add(self, nelem, emode, u, v, tmode=NOTRANSPOSE):
        self._chk(self._ceed._lib.ceedb200_basis_apply_add(self._ptr, nelem, tmode, emode, u._ptr, v._ptr))


class QFunctionContext(_Object):
    _destroy = "ceedb200_qfcontext_destroy"

    def __init__(self, ceed):
        ptr = C.c_void_p()
        ceed._chk(ceed._lib.ceedb200_qfcontext_create(ceed._ptr, C.byref(ptr)))
        super().__init__(ceed, ptr)

    def set_data(self, data, memtype=MEM_HOST, cmode=COPY_VALUES):
        arr = np.ascontiguousarray(data)
        self._keep = arr
        self._chk(self._ceed._lib.ceedb200_qfcontext_set_data(self._ptr, memtype, COPY_VALUES, arr.nbytes, arr.ctypes.data_as(C.c_void_p)))


class QFunction(_Object):
    _destroy = "ceedb200_qfunction_destroy"

    def __init__(self, ceed, source, name):
        ptr = C.c_void_p()
        ceed._chk(ceed._lib.ceedb200_qfunction_create(ceed._ptr, str(source).encode(), name.encode(), C.byref(ptr)))
        super().__init__(ceed, ptr)
        self.inputs, self.outputs = [], []
        self._ctx = None

    def add_input(self, name, size, emode):
        self._chk(self._ceed._lib.ceedb200_qfunction_add_input(self._ptr, name.encode(), size, emode))
        self.inputs.append((name, size, emode))

    def add_output(self, name, size, emode):
        self._chk(self._ceed._lib.ceedb200_qfunction_add_output(self._ptr, name.encode(), size, emode))
        self.outputs.append((name, size, emode))

    def set_context(self, ctx):
        self._ctx = ctx
        self._chk(self._ceed._lib.ceedb200_qfunction_set_context(self._ptr, ctx._ptr))

    def apply(self, q, inputs, outputs):
        U = (C.c_void_p * max(1, len(inputs)))(*[v._ptr for v in inputs])
        V = (C.c_void_p * max(1, len(outputs)))(*[v._ptr for v in outputs])
        self._chk(self._ceed._lib.ceedb200_qfunction_apply(self._ptr, q, U, V))


class Operator(_Object):
    _destroy = "ceedb200_operator_destroy"

    def __init__(self, ceed, qf):
        ptr = C.c_void_p()
        ceed._chk(ceed._lib.ceedb200_operator_create(ceed._ptr, qf._ptr, C.byref(ptr)))
        super().__init__(ceed, ptr)
        self._refs = [qf]

    def set_field(self, fieldname, rstr, basis, vector):
        rp = rstr._ptr if rstr is not None else None
        bp = basis._ptr if basis is not None else None
        vp = vector._ptr if isinstance(vector, Vector) else vector
        self._refs += [rstr, basis, vector]
        self._chk(self._ceed._lib.ceedb200_operator_set_field(self._ptr, fieldname.encode(), rp, bp, vp))

    def apply(self, u, v):
        self._chk(self._ceed._lib.ceedb200_operator_apply(self._ptr, u._ptr if u is not None else None, v._ptr))

    def apply_add(self, u, v):
        self._chk(self._ceed._lib.ceedb200_operator_apply_add(self._ptr, u._ptr if u is not None else None, v._ptr))

    def apply_part(self, u, v, part):
        """part 1: the interface-touching elements [0, split) of a partitioned mesh, part 2: the interior elements
        (ElemRestriction.set_split); 1 followed by 2 equals apply()."""
        self._chk(self._ceed._lib.ceedb200_operator_apply_part(self._ptr, u._ptr, v._ptr, part))

    def apply_streamed(self, u, v, num_chunks=0):
        """v = A u with u valid on the (pinned) host only and v wanted on the host: chunked H2D / apply / D2H pipeline
        (ceedb200_operator_apply_streamed).  Returns True when the streamed path ran, False when it fell back to apply()."""
        used = C.c_int(0)
        self._chk(self._ceed._lib.ceedb200_operator_apply_streamed(self._ptr, u._ptr, v._ptr, num_chunks, C.byref(used)))
        return bool(used.value)

    def set_tuning(self, elems_per_block=0, blocks_per_sm=0):
        self._chk(self._ceed._lib.ceedb200_operator_set_tuning(self._ptr, elems_per_block, blocks_per_sm))

    SHAPE_KEYS = ("elems_per_group", "group_warps", "cta_warps", "min_blocks_per_sm", "qf_mode", "qf_unroll", "stage_mask")

    def set_kernel_shape(self, **shape):
        vals = [shape.get(k, -1 if k in ("qf_mode", "stage_mask") else 0) for k in self.SHAPE_KEYS]
        self._chk(self._ceed._lib.ceedb200_operator_set_kernel_shape(self._ptr, (C.c_int * 7)(*vals)))

    def get_kernel_shape(self):
        vals, sig = (C.c_int * 7)(), C.create_string_buffer(512)
        self._chk(self._ceed._lib.ceedb200_operator_get_kernel_shape(self._ptr, vals, sig, 512))
        return dict(zip(self.SHAPE_KEYS, list(vals)), signature=sig.value.decode())

    def set_timing(self, enabled=True):
        self._chk(self._ceed._lib.ceedb200_operator_set_timing(self._ptr, 1 if enabled else 0))

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        self._chk(self._ceed._lib.ceedb200_operator_last_kernel_ms(self._ptr, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def is_fused(self):
        out = C.c_int()
        self._chk(self._ceed._lib.ceedb200_operator_is_fused(self._ptr, C.byref(out)))
        return bool(out.value)

    def kernel_source(self):
        return self._ceed._lib.ceedb200_operator_kernel_source(self._ptr).decode()

    def kernel_info(self):
        vals = [C.c_int() for _ in range(6)]
        self._chk(self._ceed._lib.ceedb200_operator_kernel_info(self._ptr, *[C.byref(v) for v in vals]))
        keys = ["regs", "smem_bytes", "threads", "elems_per_block", "grid", "local_bytes"]
        return dict(zip(keys, [v.value for v in vals]))
