"""DoFs/sec in CG -- the reference's own figure of merit (examples/petsc/bps.c:218-288) -- for the /gpu/cuda/b200 path.
    python benchmarks/bp_cg.py [--workload bp3p6] [--dofs 10e6] [--iters 50]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/bp_cg.py ...      (one rank per GPU)
Unpreconditioned CG (BP runs with -pc_type none), b = A x_true for a seeded x_true; every scalar device-resident.  With N ranks the
operator is the overlapped multi-GPU step (parallel.DistributedOperator) and the two dot products of an iteration are all-reduced
over NCCL; the global problem has N x dofs DoFs (weak scaling) and the figure is global DoFs x iterations / max-over-ranks time."""
import argparse, json, os, re, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from libceed_b200 import ceed as cm, mesh as M
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.cg import DeviceCG

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="bp3p6")
ap.add_argument("--dofs", type=float, default=10e6, help="DoFs per GPU")
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
m = re.fullmatch(r"bp(\d)p(\d)", args.workload); bp, p = int(m.group(1)), int(m.group(2))
rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ceed = cm.Ceed(f"/gpu/cuda/b200:device_id={local_rank}")
stream = torch.cuda.current_stream()
ceed.set_stream(stream.cuda_stream)
ncomp = BP_TABLE[bp][0]
nel = M.choose_elements(args.dofs, p, ncomp)
if world > 1:
    from libceed_b200.parallel import DistributedOperator
    grid = M.split3(world)
    part = M.Partition(tuple(nel[d] * grid[d] for d in range(3)), p, world, rank)
    dop = DistributedOperator(ceed, bp, p, part, dev, overlap=True)
    n = dop.n_local
    mask = np.tile(part.owned_mask(), ncomp)
    cg = DeviceCG(ceed, None, None, None, n, dev, owned_mask=mask, dist_op=dop)
    t = torch.tensor([int(mask.sum())], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    n_global = int(t.item())
    # a right-hand side that is consistent on interface copies: b = A x with x a function of the global node id
    gid = part.global_node_ids()
    x_np = np.concatenate([np.sin(0.001 * (gid + 1) * (c + 1)) for c in range(ncomp)])
    transport = dop.transport
else:
    prob = BPProblem(ceed, bp, p, nel)
    n = n_global = prob.num_dofs
    cg = DeviceCG(ceed, prob.op, prob.u, prob.v, n, dev)
    x_np = seeded_uniform(n, 7)
    transport = "none"
x_true = torch.from_numpy(x_np).to(dev)
cg.p.copy_(x_true); cg.apply(); b = cg.Ap.clone()
cg.start(b); r0 = cg.residual_norm2()
cg.iterate(3); torch.cuda.synchronize()           # warm-up (JIT)
cg.start(b)
l0 = ceed.launch_count()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); cg.iterate(args.iters); e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
res = cg.residual_norm2() / r0
if rank == 0:
    print(json.dumps(dict(metric="DoFs/sec in CG", workload=args.workload, n_gpus=world, dofs=n_global, iterations=args.iters, ms_per_iteration=ms / args.iters,
                          gdofs_per_s=n_global * args.iters / ms / 1e6, residual_reduction=res, interface_transport=transport,
                          kernel_launches_per_iteration=(ceed.launch_count() - l0) / args.iters)))
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)
