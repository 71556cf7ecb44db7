"""Run the reference's own C test suite (binaries prebuilt into oracle/_ref/tests by `make -C oracle ref-tests`) against a
resource, with the b200 backend plugin preloaded.  Pass/skip/fail rules follow the reference's runner (tests/junit.py:100-183
in /root/reference): exit code 0 and stdout empty or equal to tests/output/<test>.out is a pass; stderr containing
"Backend does not implement" is a skip; a handful of tests are REQUIRED to fail with a specific message.
    python tests/run_ref_suite.py /gpu/cuda/b200 [filter]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TDIR = os.path.join(ROOT, "oracle", "_ref", "tests")
PLUGIN = os.path.join(ROOT, "libceed_b200", "lib", "libceed_b200_backend.so")

REQUIRED_FAILURE = {
    "t006": "No suitable backend:", "t007": "No suitable backend:", "t008": "Available backend resources:",
    "t110": "Cannot grant CeedVector array access", "t111": "Cannot grant CeedVector array access", "t112": "Cannot grant CeedVector array access",
    "t113": "Cannot grant CeedVector array access", "t114": "Cannot grant CeedVector array access",
    "t115": "Cannot grant CeedVector read-only array access, the access lock is already in use",
    "t116": "Cannot destroy CeedVector, the writable access lock is in use", "t117": "Cannot restore CeedVector array access, access was not granted",
    "t118": "Cannot sync CeedVector, the access lock is already in use",
    "t215": "Cannot destroy CeedElemRestriction, a process has read access to the offset data",
    "t303": "Input/output vectors too short for basis and evaluation mode",
    "t408": "CeedQFunctionContextGetData(): Cannot grant CeedQFunctionContext data access, a process has read access",
}
SKIP_STRINGS = ["Backend does not implement", "Can only provide HOST memory for this backend", "Can only set HOST memory for this backend"]


def load_manifest():
    specs = {}
    path = os.path.join(TDIR, "manifest.txt")
    for line in open(path):
        m = re.match(r".*/([\w\-]+)\.c://TESTARGS(\(.*?\))?\s*(.*)$", line.strip())
        if not m:
            continue
        name, kv, args = m.group(1), m.group(2) or "", m.group(3)
        specs.setdefault(name, []).append((kv, args))
    return specs


def run_suite(resource, pattern=None, timeout=120, verbose=False):
    specs = load_manifest()
    tests = sorted(os.listdir(os.path.join(TDIR, "bin")))
    results = []
    env = dict(os.environ, LD_PRELOAD=PLUGIN, CEED_ERROR_HANDLER="exit")
    for t in tests:
        if pattern and not re.search(pattern, t):
            continue
        for kv, args in specs.get(t, [("", "{ceed_resource}")]):
            label = t + (" " + kv if kv else "")
            if 'only="cpu"' in kv and "gpu" in resource:
                results.append((label, "skip", "CPU only test with GPU backend"))
                continue
            argv = [a.replace("{ceed_resource}", resource) for a in args.split()]
            try:
                r = subprocess.run([os.path.join(TDIR, "bin", t)] + argv, capture_output=True, text=True, timeout=timeout, env=env,
                                   cwd=os.path.join(TDIR, "tests"))
                out, err, rc = r.stdout, r.stderr, r.returncode
            except subprocess.TimeoutExpired:
                results.append((label, "fail", "timeout"))
                continue
            tid = t[:4]
            if any(s in err for s in SKIP_STRINGS):
                results.append((label, "skip", [s for s in SKIP_STRINGS if s in err][0]))
            elif tid in REQUIRED_FAILURE:
                ok = REQUIRED_FAILURE[tid] in err
                results.append((label, "pass" if ok else "fail", "required failure message " + ("seen" if ok else "MISSING: " + err[-300:])))
            else:
                exp = os.path.join(TDIR, "tests", "output", t + ".out")
                expected = open(exp).read() if os.path.exists(exp) else ""
                good = rc == 0 and err.strip() == "" and (out == expected or (not expected and out.strip() == "") or t.startswith("ex") or tid == "t003")
                results.append((label, "pass" if good else "fail", "" if good else f"rc={rc} stderr={err[-400:]!r} stdout={out[-200:]!r}"))
            if verbose:
                print(results[-1][0], results[-1][1], results[-1][2][:200], flush=True)
    return results


if __name__ == "__main__":
    res = run_suite(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, verbose=True)
    counts = {k: sum(1 for r in res if r[1] == k) for k in ("pass", "skip", "fail")}
    print("SUMMARY", sys.argv[1], counts)
    for r in res:
        if r[1] == "fail":
            print("FAIL", r[0], r[2][:600])
    sys.exit(1 if counts["fail"] else 0)
