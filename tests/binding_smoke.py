"""Smoke program of the reference's PYTHON binding (the unmodified `libceed` package, built into oracle/_ref/python by
oracle/build_ref_python.py) on a resource given on the command line: a 3-D mass operator (BP1 shape) and a 3-D Poisson operator (BP3 shape)
from the gallery QFunctions, through libceed.Ceed / Vector / ElemRestriction / BasisTensorH1Lagrange / QFunctionByName / Operator --
the call path of python/tests/test-5-operator.py.  Prints one JSON line with the results' norms and a checksum vector; the test compares
resources.  usage: python tests/binding_smoke.py <resource> [plugin.so]"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
resource = sys.argv[1]
# the reference library first (the binding links it by rpath; the same object is reused), then -- for /gpu/cuda/b200 -- the backend plugin,
# whose constructor registers the resource (LD_PRELOAD does the same for a C program)
# (BINDING_SMOKE_LIB=lib selects the CPU-only build of the reference on a machine without a CUDA driver; the caller then also puts
# oracle/_ref/lib first on LD_LIBRARY_PATH, which takes precedence over the extension's RUNPATH)
ctypes.CDLL(os.path.join(REF, os.environ.get("BINDING_SMOKE_LIB", "lib-cuda"), "libceed.so"), mode=ctypes.RTLD_GLOBAL)
if len(sys.argv) > 2:
    ctypes.CDLL(sys.argv[2], mode=ctypes.RTLD_GLOBAL)
sys.path.insert(0, os.path.join(REF, "python"))
sys.path.insert(0, ROOT)
import libceed  # noqa: E402  (the reference's package)
from libceed_b200 import mesh as M  # noqa: E402  (host-side mesh builder only)

ceed = libceed.Ceed(resource)
p, nel = 3, (4, 3, 3)
P, Q = p + 1, p + 2
off = M.hex_offsets(*nel, p).astype(np.int32)
coords = M.hex_coords(*nel, p)
ne, nn = off.shape[0], coords.shape[1]
out = {"resource": ceed.get_resource() if hasattr(ceed, "get_resource") else resource, "num_elem": int(ne), "num_nodes": int(nn)}

x = ceed.Vector(3 * nn)
x.set_array(np.ascontiguousarray(coords.reshape(-1)), cmode=libceed.COPY_VALUES)
rx = ceed.ElemRestriction(ne, P ** 3, 3, nn, 3 * nn, off.reshape(-1), cmode=libceed.COPY_VALUES)
ru = ceed.ElemRestriction(ne, P ** 3, 1, 1, nn, off.reshape(-1), cmode=libceed.COPY_VALUES)
bx = ceed.BasisTensorH1Lagrange(3, 3, P, Q, libceed.GAUSS)
bu = ceed.BasisTensorH1Lagrange(3, 1, P, Q, libceed.GAUSS)
rng = np.random.Generator(np.random.PCG64(11))
u_host = rng.uniform(-1.0, 1.0, nn)

for name, build, apply, nq, emode in (("mass", "Mass3DBuild", "MassApply", 1, "interp"), ("poisson", "Poisson3DBuild", "Poisson3DApply", 6, "grad")):
    strides = np.array([1, Q ** 3, Q ** 3 * nq], dtype="int32")
    rq = ceed.StridedElemRestriction(ne, Q ** 3, nq, nq * ne * Q ** 3, strides)
    qdata = ceed.Vector(nq * ne * Q ** 3)
    qf_setup = ceed.QFunctionByName(build)
    op_setup = ceed.Operator(qf_setup)
    op_setup.set_field("dx", rx, bx, libceed.VECTOR_ACTIVE)
    op_setup.set_field("weights", libceed.ELEMRESTRICTION_NONE, bx, libceed.VECTOR_NONE)
    op_setup.set_field("qdata", rq, libceed.BASIS_NONE, libceed.VECTOR_ACTIVE)
    op_setup.apply(x, qdata)
    qf = ceed.QFunctionByName(apply)
    op = ceed.Operator(qf)
    op.set_field("du" if emode == "grad" else "u", ru, bu, libceed.VECTOR_ACTIVE)
    op.set_field("qdata", rq, libceed.BASIS_NONE, qdata)
    op.set_field("dv" if emode == "grad" else "v", ru, bu, libceed.VECTOR_ACTIVE)
    u, v = ceed.Vector(nn), ceed.Vector(nn)
    u.set_array(u_host, cmode=libceed.COPY_VALUES)
    v.set_value(0.0)
    op.apply(u, v)
    with v.array_read() as a:
        res = np.array(a, dtype=np.float64)
    out[name] = {"norm2": float(np.linalg.norm(res)), "sum": float(res.sum()), "v": res[:: max(1, nn // 64)].tolist()}
    # ApplyAdd on the same objects
    op.apply_add(u, v)
    with v.array_read() as a:
        out[name]["norm2_after_add"] = float(np.linalg.norm(np.array(a)))
print("RESULT" + json.dumps(out))
