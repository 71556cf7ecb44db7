"""Fused kernel vs the unfused twins (CEED_B200_NO_FUSE=1: restriction, basis, QFunction kernels with E- and Q-vectors in HBM, the path of
operators the generator rejects) at 10 M DoFs.  Wall clock around synchronised batches of applies.  usage: python scripts/gpu_unfused.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libceed_b200 import Ceed
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.mesh import choose_elements
out = open(os.path.join(ROOT, "gpurun_out", "r02c_unfused.txt"), "w")
def say(s):
    print(s, flush=True); out.write(s + "\n"); out.flush()
for bp, p in ((3, 6),):
    res = {}
    for nofuse in (0, 1):
        if nofuse: os.environ["CEED_B200_NO_FUSE"] = "1"
        else: os.environ.pop("CEED_B200_NO_FUSE", None)
        ceed = Ceed()
        prob = BPProblem(ceed, bp, p, choose_elements(10e6, p, BP_TABLE[bp][0]))
        prob.u.set_array(seeded_uniform(prob.num_dofs))
        for _ in range(3): prob.op.apply(prob.u, prob.v)
        ceed.synchronize()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n): prob.op.apply(prob.u, prob.v)
        ceed.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        res[nofuse] = (ms, prob.v.get_array_read().copy())
        say(f"bp{bp} p={p} {prob.num_dofs/1e6:.2f} M DoFs {'unfused twins' if nofuse else 'fused kernel '} fused={prob.op.is_fused}: {ms:.3f} ms per apply, {prob.num_dofs/ms/1e6:.2f} GDoF/s, "
            f"{prob.bytes_per_apply()/ms/1e6/6550.1:.3f} of the roofline")
        del prob, ceed
    say(f"  max rel diff fused vs unfused: {np.abs(res[0][1] - res[1][1]).max() / np.abs(res[0][1]).max():.2e}; unfused / fused time: {res[1][0] / res[0][0]:.1f}x")
