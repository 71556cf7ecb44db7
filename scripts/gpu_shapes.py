"""Run every kernel shape of tests/test_gpu_parity.py::SHAPES on one BP case and report the error vs the default shape.
usage: python scripts/gpu_shapes.py bp p nx ny nz"""
import os, sys
os.environ["CEED_B200_NO_TUNE_TABLE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libceed_b200 import Ceed
from libceed_b200.bp import BPProblem, seeded_uniform
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
bp, p, nel = int(sys.argv[1]), int(sys.argv[2]), tuple(int(x) for x in sys.argv[3:6])
SHAPES = [dict(group_warps=2, cta_warps=2), dict(group_warps=2, cta_warps=8), dict(group_warps=4, cta_warps=4), dict(qf_mode=1, qf_unroll=2),
          dict(qf_mode=2, qf_unroll=2), dict(qf_mode=2, group_warps=2, cta_warps=4, elems_per_group=3), dict(stage_mask=17), dict(stage_mask=9),
          dict(stage_mask=0, cta_warps=1), dict(stage_mask=19, group_warps=2, cta_warps=4), dict(qf_mode=3), dict(qf_mode=0)]
ceed = Ceed(); prob = BPProblem(ceed, bp, p, nel)
prob.u.set_array(seeded_uniform(prob.num_dofs, 23))
prob.op.apply(prob.u, prob.v); v0 = prob.v.get_array_read().copy()
print("default", prob.op.get_kernel_shape(), flush=True)
for sh in SHAPES:
    prob.op.set_kernel_shape(**sh)
    prob.v.set_value(-7.0)
    try:
        prob.op.apply(prob.u, prob.v); ceed.synchronize()
        v = prob.v.get_array_read()
        print(sh, "err", np.abs(v - v0).max() / np.abs(v0).max(), prob.op.kernel_info(), flush=True)
    except Exception as e:
        print(sh, "FAILED", str(e)[:200], flush=True); break
