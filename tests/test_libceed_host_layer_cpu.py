"""The libCEED host layer (libceed_b200/backend/*.c, the plugin) on a machine WITHOUT a GPU: loaded into the reference's CPU-only library in
the core's compile-only mode.  What can be checked there: the plugin registers the resource, CeedInit resolves it, objects are created
through libCEED's public API, host-side data round-trips, the fused kernel of an operator built through libCEED is generated and
NVRTC-compiled -- and an apply FAILS LOUDLY with libCEED's backend error code and a message that says why (no CPU fallback anywhere),
leaving the user's vectors unlocked (restore-on-error).  The compute itself is covered on the GPU (tests/test_gpu_libceed_backend.py)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "lib", "libceed.so")
PLUGIN = os.path.join(ROOT, "libceed_b200", "lib", "libceed_b200_backend.so")

BODY = r"""
import ctypes as C, json, os, sys
sys.path.insert(0, %r)
import numpy as np
from oracle import refceed as R
from libceed_b200 import mesh as M
out = {}
R.RefCeed._libs["lib"] = C.CDLL(os.path.join(R.REF_DIR, "lib", "libceed.so"), mode=C.RTLD_GLOBAL)
C.CDLL(%r, mode=C.RTLD_GLOBAL)            # constructor: CeedRegister("/gpu/cuda/b200", ...)
rc = R.RefCeed("/gpu/cuda/b200")
res = C.c_char_p(); rc.lib.CeedGetResource(rc.ceed, C.byref(res)); out["resource"] = res.value.decode()
det = C.c_bool(); rc.lib.CeedIsDeterministic(rc.ceed, C.byref(det)); out["deterministic"] = bool(det.value)
pref = C.c_int(); rc.lib.CeedGetPreferredMemType(rc.ceed, C.byref(pref)); out["preferred_mem_type"] = pref.value
# host data round trip through the plugin's vector
a = np.linspace(-1.0, 1.0, 17)
v = rc.vector(17, a)
out["vector_roundtrip"] = bool(np.array_equal(rc.get_array(v, 17), a))
# basis matrices of the plugin's basis object against the CPU reference backend's
cpu = R.RefCeed("/cpu/self/ref/serial")
mb, mc = rc.basis_matrices(rc.basis_lagrange(3, 1, 4, 6, 0), 4, 6), cpu.basis_matrices(cpu.basis_lagrange(3, 1, 4, 6, 0), 4, 6)
out["basis_equal"] = all(np.array_equal(mb[k], mc[k]) for k in mc)
# an operator built through libCEED: the apply must fail loudly (backend error code, message), never fall back to a CPU path
p, nel = 3, (3, 2, 2)
off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
try:
    R.RefBP(rc, 1, p, off.shape[0], coords.shape[1], off, coords)
    out["apply_error"] = None
except RuntimeError as e:
    out["apply_error"] = str(e)
# restore-on-error: a failed CeedOperatorApply leaves the user's vectors accessible
nn = coords.shape[1]
ru = rc.restriction(off.shape[0], 64, 1, nn, nn, off.reshape(-1))
rq = rc.restriction_strided(off.shape[0], 125, 1, off.shape[0] * 125, None)
bu = rc.basis_lagrange(3, 1, 4, 5, 0)
qf = rc.qfunction_by_name("MassApply")
op = rc.operator(qf)
qd = rc.vector(off.shape[0] * 125, np.ones(off.shape[0] * 125))
rc.op_set_field(op, "u", ru, bu, rc.VECTOR_ACTIVE); rc.op_set_field(op, "qdata", rq, None, qd); rc.op_set_field(op, "v", ru, bu, rc.VECTOR_ACTIVE)
u, w = rc.vector(nn, np.ones(nn)), rc.vector(nn, np.zeros(nn))
code = rc.lib.CeedOperatorApply(op, u, w, rc.REQUEST_IMMEDIATE)
out["apply_code"] = code
try:
    out["u_after_error"] = bool(np.array_equal(rc.get_array(u, nn), np.ones(nn)))
    rc.set_array(w, np.full(nn, 2.0))
    out["w_after_error"] = bool(np.array_equal(rc.get_array(w, nn), np.full(nn, 2.0)))
except RuntimeError as e:
    out["u_after_error"] = out["w_after_error"] = str(e)
print("RESULT" + json.dumps(out))
""" % (ROOT, PLUGIN)


@pytest.mark.skipif(not (os.path.exists(REF_LIB) and os.path.exists(PLUGIN)), reason="oracle/_ref or the backend plugin not built")
def test_host_layer_registers_creates_objects_and_fails_loudly_without_a_gpu():
    env = dict(os.environ, CEED_B200_COMPILE_ONLY="1")
    r = subprocess.run([sys.executable, "-c", BODY], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    assert res["resource"] == "/gpu/cuda/b200" and res["deterministic"] is True and res["preferred_mem_type"] == 1  # CEED_MEM_DEVICE
    assert res["vector_roundtrip"] and res["basis_equal"]
    # CEED_ERROR_BACKEND (-2), raised by the host layer with the core's message: the kernels were generated, nothing ran, nothing was faked
    assert res["apply_error"] and "error -2" in res["apply_error"] and "COMPILE_ONLY" in res["apply_error"], res["apply_error"]
    assert res["apply_code"] == -2
    assert res["u_after_error"] is True and res["w_after_error"] is True, res
