"""TEST INFRASTRUCTURE: CPU emulation of the generated fused-operator kernels (no GPU needed).

In compile-only mode (CEED_B200_COMPILE_ONLY=1) the core's "device" memory is host memory and every launch fails loudly.  The debug
entry point ceedb200_operator_debug_launch describes the launch the fused apply WOULD perform: the kernel's argument block, launch shape,
element range, the generated source of the variant and the tables of the finalize pass.  This module compiles that very source for the
host with g++ (tests/emu/b200-jit.h maps the CUDA constructs: one OS thread per CUDA thread, barriers for __syncwarp / __syncthreads),
runs it against the argument block and performs the finalize pass in numpy -- so the code generator's lane maps, table use, tail
handling and scatter are checked against the oracle on a machine without a GPU.  cp.async is emulated as an immediate copy, cp.async.bulk +
mbarrier by tests/emu/b200-tma.h (copy at once, byte-counted completion, phase parity), named barriers as pthread barriers, the E-vector mode's
transpose restriction in numpy; the in-kernel ordered completion / in-kernel finalize (acquire / release flags between CTAs) are not emulated.  Nothing of this is reachable from the product."""
import atexit
import ctypes as C
import hashlib
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
_cache = {}


class DebugLaunch(C.Structure):
    """B200DebugLaunch of include/ceed_b200.h"""
    _fields_ = [("args", C.c_ubyte * 1024),
                ("args_size", C.c_int), ("grid", C.c_int), ("threads", C.c_int), ("smem_bytes", C.c_int), ("run_mode", C.c_int),
                ("kernel_add", C.c_int), ("zero_first", C.c_int), ("fin_slot", C.c_int), ("num_comp", C.c_int),
                ("e_begin", C.c_longlong), ("e_end", C.c_longlong), ("comp_stride", C.c_longlong), ("num_shared", C.c_longlong),
                ("num_halo", C.c_longlong),
                ("halo_node", C.POINTER(C.c_int)), ("halo_ptr", C.POINTER(C.c_int)), ("halo", C.POINTER(C.c_double)),
                ("v", C.POINTER(C.c_double)), ("source", C.c_char_p),
                ("scatter_mode", C.c_int), ("e_entries", C.c_longlong), ("offsets", C.POINTER(C.c_int)), ("evec", C.POINTER(C.c_double))]


def build(source):
    """g++-compile one generated kernel source (+ the thread driver) into a shared library; cached by content."""
    key = hashlib.sha1(source.encode()).hexdigest()
    if key in _cache:
        return _cache[key]
    # cp.async (LDGSTS) is emulated as an immediate copy -- a legal execution: the sources are read-only for the kernel and the
    # destination is not read before the matching wait; named barriers of multi-warp element groups become pthread barriers
    source = re.sub(r'asm volatile\("cp\.async\.c[ag]\.shared\.global \[%0\], \[%1\], (\d+);"[^;]*;', r"memcpy(dst, src, \1);", source)
    source = re.sub(r'asm volatile\("cp\.async\.(commit_group|wait_group \d+);"[^;]*;', "", source)
    source = re.sub(r'asm volatile\("bar\.sync %0, (\d+);" ::"r"\((.*?)\) : "memory"\);', r"b200_emu_named_barrier(\2, \1);", source)
    source = re.sub(r'asm volatile\("prefetch\.global\.L2 \[%0\];"[^;]*;', "", source)
    if "asm volatile" in source:
        raise NotImplementedError("kernel uses inline PTX that is not emulated (acquire / release flags of the in-kernel ordered completion)")
    name = re.search(r"__global__ void\s+(?:__launch_bounds__\([^)]*\)\s*)?(b200_operator_\w+)\s*\(", source).group(1)
    d = tempfile.mkdtemp(prefix="b200emu_")
    atexit.register(shutil.rmtree, d, ignore_errors=True)
    cu = os.path.join(d, "kernel.cpp")
    with open(cu, "w") as f:
        f.write(source + "\n#include \"emu_driver.inc\"\n")
    so = os.path.join(d, "kernel.so")
    cmd = ["g++", "-O0", "-std=c++17", "-march=x86-64-v3", "-fPIC", "-shared", "-pthread", "-w", "-DEMU_KERNEL=" + name, "-I" + EMU,
           "-I" + os.path.join(ROOT, "libceed_b200", "csrc", "jit"), "-o", so, cu]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("host compile of the generated kernel failed:\n" + r.stderr[-4000:])
    lib = C.CDLL(so)
    lib.b200_emu_launch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_int]
    lib.b200_emu_launch.restype = C.c_int
    _cache[key] = lib
    return lib


def emulated_apply(op, u, v, add=False, part=0, grid=3):
    """v = A u (or v += A u) of a fused operator of the Python mirror, executed on CPU threads.  Returns the launch description."""
    lib = op._ceed._lib
    desc = DebugLaunch()
    op._chk(lib.ceedb200_operator_debug_launch(op._ptr, u._ptr, v._ptr, int(bool(add)), int(part), C.byref(desc)))
    n = len(v)
    out = np.ctypeslib.as_array(desc.v, shape=(n,)) if desc.v else None  # (NULL: no offset-restricted output, e.g. a setup operator)
    if desc.zero_first and part <= 1:
        assert out is not None
        out[:] = 0.0
    if desc.e_end > desc.e_begin:
        kernel = build(desc.source.decode())
        rc = kernel.b200_emu_launch(C.cast(desc.args, C.c_void_p), desc.args_size, max(1, min(grid, desc.grid) if desc.grid > 0 else grid), desc.threads,
                                    desc.smem_bytes, desc.e_begin, desc.e_end, desc.run_mode)
        assert rc == 0, rc
    if desc.fin_slot >= 0 and desc.num_shared > 0:
        # k_halo_finalize (b200_restriction.cu): owner value + halo slots in ascending order
        ns, nh = desc.num_shared, desc.num_halo
        node = np.ctypeslib.as_array(desc.halo_node, shape=(ns,)).astype(np.int64)
        ptr = np.ctypeslib.as_array(desc.halo_ptr, shape=(ns + 1,)).astype(np.int64)
        halo = np.ctypeslib.as_array(desc.halo, shape=(nh * desc.num_comp,))
        cnt = ptr[1:] - ptr[:-1]
        for c in range(desc.num_comp):
            for k in range(int(cnt.max())):
                m = cnt > k
                out[node[m] + c * desc.comp_stride] += halo[ptr[:-1][m] + k + c * nh]
    if desc.scatter_mode == 2 and desc.e_entries > 0:
        # E-vector mode: the transpose restriction kernel that follows (k_offset_transpose, b200_restriction.cu) adds the E-vector into v in
        # ascending E-index per node -- np.add.at accumulates in index order
        ne = desc.e_entries
        off = np.ctypeslib.as_array(desc.offsets, shape=(ne,)).astype(np.int64)
        ev = np.ctypeslib.as_array(desc.evec, shape=(ne * desc.num_comp,))
        for c in range(desc.num_comp):
            np.add.at(out, off + c * desc.comp_stride, ev[c * ne:(c + 1) * ne])
    return desc
