"""Small BP applies on the kernel shapes that use the round-2 mechanisms, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/gpu_sanitize.py     (bulk copies that start one double early for odd Q^3, ...)
    compute-sanitizer --tool racecheck python scripts/gpu_sanitize.py     (shared-memory hazards: stage barriers, mbarrier hand-over, swizzle)
Every result is also checked against the default shape."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CEED_B200_NO_TUNE_TABLE"] = "1"
from libceed_b200 import Ceed
from libceed_b200.bp import BPProblem, seeded_uniform

CASES = [(1, 3, (3, 2, 2), dict(stage_mask=33)), (3, 2, (3, 3, 2), dict(stage_mask=33, group_warps=2, cta_warps=4, elems_per_group=3)),
         (3, 6, (2, 1, 1), dict(stage_mask=257)), (5, 7, (1, 2, 1), dict(stage_mask=257, group_warps=2, cta_warps=4)),
         (3, 6, (1, 2, 1), dict(stage_mask=289, group_warps=2, cta_warps=2)), (3, 5, (2, 1, 1), dict(stage_mask=297, group_warps=1, cta_warps=2)),
         (3, 3, (3, 2, 2), dict(stage_mask=65)), (6, 2, (2, 2, 1), dict(stage_mask=257)),
         (3, 4, (3, 2, 2), dict(stage_mask=513)), (5, 5, (2, 2, 1), dict(stage_mask=513, group_warps=2, cta_warps=4, elems_per_group=2)), (4, 1, (3, 2, 2), dict(stage_mask=512)),
         # lean in-place-plane kernel (layout 4): direct loads with parked targets, bulk pipelines (odd Q^3: sources aligned down at run time),
         # partial tail batches, 3 components; in-kernel finalize with per-part release / acquire counters (needs >= 64 batches per part)
         (1, 3, (5, 3, 3), dict(qf_mode=4, group_warps=1, cta_warps=4, elems_per_group=6, stage_mask=0)),
         (1, 3, (5, 3, 3), dict(qf_mode=4, group_warps=1, cta_warps=2, elems_per_group=4, stage_mask=40)),
         (2, 2, (3, 3, 2), dict(qf_mode=4, group_warps=1, cta_warps=2, elems_per_group=3, stage_mask=8)),
         (1, 4, (3, 2, 3), dict(qf_mode=4, group_warps=1, cta_warps=4, elems_per_group=5, stage_mask=96)),
         (1, 2, (12, 12, 11), dict(qf_mode=4, group_warps=1, cta_warps=4, elems_per_group=2, stage_mask=128))]
os.environ["CEED_B200_PARTS"] = "3"
for bp, p, nel, shape in CASES:
    prob = BPProblem(Ceed(), bp, p, nel)
    u = seeded_uniform(prob.num_dofs, 3)
    prob.u.set_array(u)
    prob.op.apply(prob.u, prob.v)
    v0 = prob.v.get_array_read().copy()
    prob.op.set_kernel_shape(**shape)
    prob.v.set_value(9.0)
    prob.op.apply(prob.u, prob.v)
    prob.op.apply_add(prob.u, prob.v)
    err = np.abs(prob.v.get_array_read() - 2 * v0).max() / np.abs(v0).max()
    print(f"BP{bp} p={p} {nel} {shape}: rel diff to the default shape {err:.1e}", flush=True)
    assert err < 1e-13
print("sanitize cases done")
