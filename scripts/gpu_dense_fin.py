"""Dense finalize pass (CEED_B200_DENSE_FINALIZE) against the list form: bitwise equality and time of the second scatter pass at 10 M DoFs.
usage: python scripts/gpu_dense_fin.py"""
import os, re, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libceed_b200 import Ceed
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.mesh import choose_elements
out = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02c_dense_finalize.txt"), "w")
def say(s):
    print(s, flush=True); out.write(s + "\n"); out.flush()
bad = 0
for bp, p, dofs in ((1, 3, 10e6), (3, 6, 10e6), (6, 4, 10e6), (2, 3, 3e6)):
    ceed = Ceed()
    prob = BPProblem(ceed, bp, p, choose_elements(dofs, p, BP_TABLE[bp][0]), interlaced=(bp == 6))
    prob.u.set_array(seeded_uniform(prob.num_dofs))
    prob.op.set_timing(True)
    res = {}
    for dense in ("0", "1", "0", "1"):
        os.environ["CEED_B200_DENSE_FINALIZE"] = dense
        t = []
        for i in range(9):
            prob.op.apply(prob.u, prob.v)
            if i >= 3: t.append(prob.op.last_kernel_ms())
        f, a = float(np.median([x[0] for x in t])), float(np.median([x[1] for x in t]))
        v = prob.v.get_array_read()
        res.setdefault(dense, []).append((f, a, v.copy() if dense not in res or len(res[dense]) == 0 else None))
        say(f"bp{bp} p={p} {prob.num_dofs/1e6:.2f} M DoFs dense={dense}: fused {f:.4f} + finalize {a:.4f} = {f+a:.4f} ms  {prob.num_dofs/(f+a)/1e6:.2f} GDoF/s  frac {prob.bytes_per_apply()/(f+a)/1e6/6550.1:.3f}")
    same = bool(np.array_equal(res["0"][0][2], res["1"][0][2]))
    bad += not same
    say(f"  bitwise equal: {same}")
    del prob, ceed
say("all bitwise equal" if not bad else f"{bad} MISMATCHES")
sys.exit(1 if bad else 0)
