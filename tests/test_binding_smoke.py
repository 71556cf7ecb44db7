"""Language-binding smoke test: the reference's unmodified PYTHON binding -- package `libceed`, cffi extension built
by the reference's own python/build_ceed_cffi.py into oracle/_ref/python (oracle/build_ref_python.py) -- drives a 3-D mass and a 3-D Poisson
operator from the gallery QFunctions.  On the CPU the reference's backends must agree with each other (validates the build of the binding);
on the GPU `/gpu/cuda/b200` (plugin loaded into the Python process) must match `/cpu/self/ref/serial` to 1e-12 through the same calls."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PLUGIN = os.path.join(ROOT, "libceed_b200", "lib", "libceed_b200_backend.so")
SMOKE = os.path.join(ROOT, "tests", "binding_smoke.py")
have_binding = os.path.exists(os.path.join(REF, "python", "libceed", "__init__.py"))


def smoke(resource, cuda, plugin=None):
    env = dict(os.environ)
    if not cuda:
        env["BINDING_SMOKE_LIB"] = "lib"
        env["LD_LIBRARY_PATH"] = os.path.join(REF, "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([sys.executable, SMOKE, resource] + ([plugin] if plugin else []), capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])


def close(a, b, tol):
    for name in ("mass", "poisson"):
        va, vb = np.array(a[name]["v"]), np.array(b[name]["v"])
        scale = np.abs(va).max()
        assert np.abs(va - vb).max() <= tol * scale, name
        assert abs(a[name]["norm2"] - b[name]["norm2"]) <= tol * a[name]["norm2"], name
        assert abs(a[name]["norm2_after_add"] - 2 * a[name]["norm2"]) <= 1e-12 * a[name]["norm2"], name


@pytest.mark.skipif(not have_binding, reason="oracle/_ref/python not built (needs /root/reference)")
def test_reference_python_binding_runs_on_the_cpu_reference():
    a = smoke("/cpu/self/ref/serial", cuda=False)
    b = smoke("/cpu/self/opt/blocked", cuda=False)
    assert a["num_elem"] == 36 and a["mass"]["norm2"] > 0 and a["poisson"]["norm2"] > 0
    close(a, b, 1e-12)


@pytest.mark.gpu
@pytest.mark.skipif(not have_binding, reason="oracle/_ref/python not built (needs /root/reference)")
def test_reference_python_binding_on_gpu_cuda_b200():
    """A Python caller of the reference switches resources and nothing else."""
    if not os.path.exists(PLUGIN):
        pytest.skip("backend plugin not built")
    ref = smoke("/cpu/self/ref/serial", cuda=True)
    got = smoke("/gpu/cuda/b200", cuda=True, plugin=PLUGIN)
    assert "b200" in got["resource"], got["resource"]
    close(ref, got, 1e-12)
