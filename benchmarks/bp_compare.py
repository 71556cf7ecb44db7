"""CeedOperatorApply through libCEED's public C API on several resources of ONE libceed.so (the unmodified reference built
into oracle/_ref/lib-cuda, with the b200 backend plugin registered): /gpu/cuda/b200 against the reference's own GPU
backends (/gpu/cuda/gen is the baseline to beat) -- same mesh, same seeded input, same protocol.
Reports GDoF/s (median of the timed applies, device-resident vectors, cudaDeviceSynchronize on both sides) and the
difference of every result to the first resource.
    python benchmarks/bp_compare.py [--bp 3] [--p 1..8] [--dofs 10e6] [--resources /gpu/cuda/b200,/gpu/cuda/gen]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libceed_b200 import mesh as M  # noqa: E402
from libceed_b200.bp import BP_TABLE, algorithmic_bytes, seeded_uniform  # noqa: E402
from oracle import refceed as R  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bp", type=int, nargs="+", default=[3])
ap.add_argument("--p", type=int, nargs="+", default=[6])
ap.add_argument("--dofs", type=float, default=10e6)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--resources", default="/gpu/cuda/b200,/gpu/cuda/gen")
ap.add_argument("--json", default=None)
args = ap.parse_args()

lib = C.CDLL(os.path.join(R.REF_DIR, "lib-cuda", "libceed.so"), mode=C.RTLD_GLOBAL)
R.RefCeed._libs["lib-cuda"] = R.RefCeed._libs["lib"] = lib
C.CDLL(os.path.join(ROOT, "libceed_b200", "lib", "libceed_b200_backend.so"), mode=C.RTLD_GLOBAL)
cudart = C.CDLL("libcudart.so.12")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6550.1

rows = []
for bp in args.bp:
    for p in args.p:
        ncomp = BP_TABLE[bp][0]
        nel = M.choose_elements(args.dofs, p, ncomp)
        off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
        nn = coords.shape[1]
        u = seeded_uniform(ncomp * nn)
        ref_v = None
        for res in args.resources.split(","):
            try:
                rc = R.RefCeed(res, cuda=True)
                t0 = time.time()
                prob = R.RefBP(rc, bp, p, off.shape[0], nn, off, coords)
                rc.set_array(prob.u, u)
                # move data to the device / JIT-compile: warm-up applies
                for _ in range(3):
                    rc.op_apply(prob.op, prob.u, prob.v)
                cudart.cudaDeviceSynchronize()
                setup = time.time() - t0
                times = []
                for _ in range(args.reps):
                    cudart.cudaDeviceSynchronize()
                    t = time.perf_counter()
                    rc.op_apply(prob.op, prob.u, prob.v)
                    cudart.cudaDeviceSynchronize()
                    times.append(time.perf_counter() - t)
                v = rc.get_array(prob.v, ncomp * nn)
                if ref_v is None:
                    ref_v = v
                err = float(np.abs(v - ref_v).max() / np.abs(ref_v).max())
                ms = float(np.median(times) * 1e3)
                gb = algorithmic_bytes(bp, p, off.shape[0], nn) / 1e9
                row = dict(bp=bp, p=p, dofs=ncomp * nn, resource=res, ms=ms, gdofs=ncomp * nn / ms / 1e6, frac_hbm=gb / ms * 1e3 / peak,
                           rel_diff_to_first=err, setup_s=setup)
            except Exception as exc:
                row = dict(bp=bp, p=p, resource=res, error=str(exc)[:300])
            rows.append(row)
            print(json.dumps(row), flush=True)
if args.json:
    json.dump(rows, open(args.json, "w"), indent=1)
