"""ctypes binding of the C ABI declared in include/ceed_b200.h.

The shared library (libceed_b200/lib/libceed_b200.so) is the product: hand-written sm_100a CUDA kernels plus the
NVRTC host layer.  There is NO fallback: if the library is missing the import fails, and every compute entry point
fails loudly when no B200 / CUDA driver is present.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libceed_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ceed_b200.h")

MEM_HOST, MEM_DEVICE = 0, 1
COPY_VALUES, USE_POINTER, OWN_POINTER = 0, 1, 2
NORM_1, NORM_2, NORM_MAX = 0, 1, 2
NOTRANSPOSE, TRANSPOSE = 0, 1
EVAL_NONE, EVAL_INTERP, EVAL_GRAD, EVAL_WEIGHT = 0, 1, 2, 16
GAUSS, GAUSS_LOBATTO = 0, 1
SCATTER_DETERMINISTIC, SCATTER_ATOMIC, SCATTER_EVECTOR = 0, 1, 2

c_scalar_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)
handle = C.c_void_p


def declared_symbols(header=HEADER_PATH):
    """Names of all CEEDB200_EXPORT functions declared in the public header."""
    text = open(header).read()
    return sorted(set(re.findall(r"CEEDB200_EXPORT\s+[\w\s\*]+?\b(ceedb200_\w+)\s*\(", text)))


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C libceed_b200/csrc). ceed-b200 has no CPU fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    P = C.POINTER
    sigs = {
        "ceedb200_init": [C.c_int, P(handle)],
        "ceedb200_destroy": [handle],
        "ceedb200_add_jit_source_root": [handle, C.c_char_p],
        "ceedb200_add_jit_define": [handle, C.c_char_p],
        "ceedb200_set_stream": [handle, C.c_void_p],
        "ceedb200_get_stream": [handle, P(C.c_void_p)],
        "ceedb200_synchronize": [handle],
        "ceedb200_set_scatter_mode": [handle, C.c_int],
        "ceedb200_vector_create": [handle, C.c_int64, P(handle)],
        "ceedb200_vector_destroy": [handle],
        "ceedb200_vector_length": [handle, P(C.c_int64)],
        "ceedb200_vector_has_valid_array": [handle, P(C.c_int)],
        "ceedb200_vector_has_borrowed_array_of_type": [handle, C.c_int, P(C.c_int)],
        "ceedb200_vector_set_array": [handle, C.c_int, C.c_int, C.c_void_p],
        "ceedb200_vector_take_array": [handle, C.c_int, P(C.c_void_p)],
        "ceedb200_vector_set_value": [handle, C.c_double],
        "ceedb200_vector_set_value_strided": [handle, C.c_int64, C.c_int64, C.c_int64, C.c_double],
        "ceedb200_vector_sync_array": [handle, C.c_int],
        "ceedb200_vector_get_array": [handle, C.c_int, P(C.c_void_p)],
        "ceedb200_vector_get_array_read": [handle, C.c_int, P(C.c_void_p)],
        "ceedb200_vector_get_array_write": [handle, C.c_int, P(C.c_void_p)],
        "ceedb200_vector_copy_strided": [handle, C.c_int64, C.c_int64, C.c_int64, handle],
        "ceedb200_vector_norm": [handle, C.c_int, P(C.c_double)],
        "ceedb200_vector_scale": [handle, C.c_double],
        "ceedb200_vector_reciprocal": [handle],
        "ceedb200_vector_filter": [handle, C.c_double],
        "ceedb200_vector_axpy": [handle, C.c_double, handle],
        "ceedb200_vector_axpby": [handle, C.c_double, C.c_double, handle],
        "ceedb200_vector_pointwise_mult": [handle, handle, handle],
        "ceedb200_restriction_create": [handle, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int, C.c_int,
                                        C.c_void_p, P(handle)],
        "ceedb200_restriction_create_strided": [handle, C.c_int32, C.c_int32, C.c_int32, C.c_int64, c_int_p, P(handle)],
        "ceedb200_restriction_destroy": [handle],
        "ceedb200_restriction_apply": [handle, C.c_int, handle, handle],
        "ceedb200_restriction_get_offsets": [handle, C.c_int, P(C.c_void_p)],
        "ceedb200_restriction_get_e_layout": [handle, c_int_p],
        "ceedb200_restriction_get_info": [handle, c_int_p, c_int_p, c_int_p, P(C.c_int64), P(C.c_int64)],
        "ceedb200_basis_create_tensor_h1": [handle, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_scalar_p, c_scalar_p,
                                            c_scalar_p, c_scalar_p, P(handle)],
        "ceedb200_basis_create_tensor_h1_lagrange": [handle, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int,
                                                     P(handle)],
        "ceedb200_basis_destroy": [handle],
        "ceedb200_basis_apply": [handle, C.c_int32, C.c_int, C.c_int, handle, handle],
        "ceedb200_basis_apply_add": [handle, C.c_int32, C.c_int, C.c_int, handle, handle],
        "ceedb200_basis_get_matrix": [handle, C.c_int, c_scalar_p],
        "ceedb200_host_gauss_quadrature": [C.c_int32, c_scalar_p, c_scalar_p],
        "ceedb200_host_lobatto_quadrature": [C.c_int32, c_scalar_p, c_scalar_p],
        "ceedb200_host_lagrange_1d": [C.c_int32, C.c_int32, C.c_int, c_scalar_p, c_scalar_p, c_scalar_p, c_scalar_p],
        "ceedb200_host_collocated_grad_1d": [C.c_int32, C.c_int32, c_scalar_p, c_scalar_p, c_scalar_p],
        "ceedb200_qfcontext_create": [handle, P(handle)],
        "ceedb200_qfcontext_destroy": [handle],
        "ceedb200_qfcontext_set_data": [handle, C.c_int, C.c_int, C.c_size_t, C.c_void_p],
        "ceedb200_qfcontext_take_data": [handle, C.c_int, P(C.c_void_p)],
        "ceedb200_qfcontext_get_data": [handle, C.c_int, P(C.c_void_p)],
        "ceedb200_qfcontext_get_data_read": [handle, C.c_int, P(C.c_void_p)],
        "ceedb200_qfcontext_has_valid_data": [handle, P(C.c_int)],
        "ceedb200_qfcontext_has_borrowed_data_of_type": [handle, C.c_int, P(C.c_int)],
        "ceedb200_qfunction_create": [handle, C.c_char_p, C.c_char_p, P(handle)],
        "ceedb200_qfunction_destroy": [handle],
        "ceedb200_qfunction_add_input": [handle, C.c_char_p, C.c_int32, C.c_int],
        "ceedb200_qfunction_add_output": [handle, C.c_char_p, C.c_int32, C.c_int],
        "ceedb200_qfunction_set_context": [handle, handle],
        "ceedb200_qfunction_apply": [handle, C.c_int32, P(handle), P(handle)],
        "ceedb200_operator_create": [handle, handle, P(handle)],
        "ceedb200_operator_destroy": [handle],
        "ceedb200_operator_set_field": [handle, C.c_char_p, handle, handle, handle],
        "ceedb200_operator_apply": [handle, handle, handle],
        "ceedb200_operator_apply_add": [handle, handle, handle],
        "ceedb200_operator_is_fused": [handle, P(C.c_int)],
        "ceedb200_operator_kernel_info": [handle, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int)],
        "ceedb200_operator_set_timing": [handle, C.c_int],
        "ceedb200_operator_last_kernel_ms": [handle, P(C.c_float), P(C.c_float)],
        "ceedb200_operator_set_tuning": [handle, C.c_int, C.c_int],
        "ceedb200_operator_set_kernel_shape": [handle, P(C.c_int)],
        "ceedb200_operator_get_kernel_shape": [handle, P(C.c_int), C.c_char_p, C.c_int],
        "ceedb200_set_autotune": [handle, C.c_int],
        "ceedb200_cg_dot": [handle, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p],
        "ceedb200_cg_update": [handle, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p],
        "ceedb200_cg_direction": [handle, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p],
        "ceedb200_cg_constrain": [handle, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong],
        "ceedb200_restriction_debug_scatter_tables": [handle, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64],
        "ceedb200_iface_pack": [handle, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p],
        "ceedb200_iface_unpack_sum": [handle, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        "ceedb200_iface_put": [handle, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p],
        "ceedb200_iface_wait_unpack_sum": [handle, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p,
                                           C.c_int, C.c_void_p],
        "ceedb200_ipc_alloc": [handle, C.c_size_t, P(C.c_void_p), C.c_char_p],
        "ceedb200_ipc_open": [handle, C.c_char_p, P(C.c_void_p)],
        "ceedb200_ipc_close": [handle, C.c_void_p],
        "ceedb200_ipc_free": [handle, C.c_void_p],
        "ceedb200_vector_valid_sides": [handle, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "ceedb200_operator_apply_part": [handle, handle, handle, C.c_int],
        "ceedb200_operator_apply_streamed": [handle, handle, handle, C.c_int, C.POINTER(C.c_int)],
        "ceedb200_operator_debug_launch": [handle, handle, handle, C.c_int, C.c_int, C.c_void_p],
        "ceedb200_operator_debug_stream_plan": [handle, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)],
        "ceedb200_restriction_set_split": [handle, C.c_int32],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.ceedb200_last_error.argtypes = [handle]
    lib.ceedb200_last_error.restype = C.c_char_p
    lib.ceedb200_version.argtypes = []
    lib.ceedb200_version.restype = C.c_char_p
    lib.ceedb200_launch_count.argtypes = [handle]
    lib.ceedb200_launch_count.restype = C.c_int64
    lib.ceedb200_operator_kernel_source.argtypes = [handle]
    lib.ceedb200_operator_kernel_source.restype = C.c_char_p
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib
