"""Run one workload a few times (for ncu captures / quick timing).  usage: python scripts/gpu_one.py bp3p6 [dofs] [epb] [scatter]"""
import os, re, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libceed_b200 import Ceed
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.mesh import choose_elements
wl = sys.argv[1]; dofs = float(sys.argv[2]) if len(sys.argv) > 2 else 10e6
epb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
scatter = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m = re.fullmatch(r"bp(\d)p(\d)", wl); bp, p = int(m.group(1)), int(m.group(2))
ceed = Ceed(); ceed.set_scatter_mode(scatter)
prob = BPProblem(ceed, bp, p, choose_elements(dofs, p, BP_TABLE[bp][0]))
prob.u.set_array(seeded_uniform(prob.num_dofs))
if epb: prob.op.set_tuning(epb, 0)
prob.op.set_timing(True)
t = []
for i in range(8):
    prob.op.apply(prob.u, prob.v); t.append(prob.op.last_kernel_ms())
f, a = np.median([x[0] for x in t[3:]]), np.median([x[1] for x in t[3:]])
print(wl, prob.op.kernel_info(), f"fused {f:.3f} aux {a:.3f} ms {prob.num_dofs/(f+a)/1e6:.2f} GDoF/s {prob.bytes_per_apply()/(f+a)/1e6/6550.1*100:.1f}%")
