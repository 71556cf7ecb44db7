"""DoFs/sec in CG -- the reference's own figure of merit (examples/petsc/bps.c:218-288) -- for the /gpu/cuda/b200 path.
    python benchmarks/bp_cg.py [--workload bp3p6] [--dofs 10e6] [--iters 50]
Unpreconditioned CG (BP runs with -pc_type none), b = A x_true for a seeded x_true; every scalar device-resident."""
import argparse, json, os, re, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from libceed_b200 import ceed as cm, mesh as M
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.cg import DeviceCG

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="bp3p6")
ap.add_argument("--dofs", type=float, default=10e6)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
m = re.fullmatch(r"bp(\d)p(\d)", args.workload); bp, p = int(m.group(1)), int(m.group(2))
dev = torch.device("cuda", 0)
ceed = cm.Ceed("/gpu/cuda/b200")
stream = torch.cuda.current_stream()
ceed.set_stream(stream.cuda_stream)
prob = BPProblem(ceed, bp, p, M.choose_elements(args.dofs, p, BP_TABLE[bp][0]))
n = prob.num_dofs
cg = DeviceCG(ceed, prob.op, prob.u, prob.v, n, dev)
x_true = torch.from_numpy(seeded_uniform(n, 7)).to(dev)
cg.p.copy_(x_true); cg.apply(); b = cg.Ap.clone()
cg.start(b); r0 = cg.residual_norm2()
cg.iterate(3); torch.cuda.synchronize()           # warm-up (JIT)
cg.start(b)
l0 = ceed.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); cg.iterate(args.iters); e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps(dict(metric="DoFs/sec in CG", workload=args.workload, dofs=n, iterations=args.iters, ms_per_iteration=ms / args.iters,
                      gdofs_per_s=n * args.iters / ms / 1e6, residual_reduction=cg.residual_norm2() / r0,
                      kernel_launches_per_iteration=(ceed.launch_count() - l0) / args.iters)))
