"""Worker of tests/test_gpu_multirank.py -- launched with `python -m torch.distributed.run --nproc-per-node N` (one rank per GPU, NCCL).

Every rank builds its element box of a global structured hex mesh (mesh.Partition), applies the BP operator on /gpu/cuda/b200
through the C ABI, sums the interface DoFs over NCCL (parallel.InterfaceExchange, the product's multi-GPU step, overlapped
variant included) and writes its local vector to <out_dir>/rank<r>.npz.  Rank 0 additionally applies the same operator on the
WHOLE mesh on its own GPU (single-rank result) for the comparison done by the test.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from libceed_b200 import ceed as cm
    from libceed_b200 import mesh as M
    from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
    from libceed_b200.parallel import DistributedOperator

    out_dir, bp, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    n_global = tuple(int(x) for x in sys.argv[4:7])
    cg_iters = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    ncomp = BP_TABLE[bp][0]
    ceed = cm.Ceed(f"/gpu/cuda/b200:device_id={local_rank}")
    ceed.set_stream(torch.cuda.current_stream().cuda_stream)
    part = M.Partition(n_global, p, world, rank)
    n_glob = int(np.prod([n * p + 1 for n in n_global]))
    u_glob = seeded_uniform(ncomp * n_glob)
    gid = part.global_node_ids()
    nloc = part.num_local_nodes
    u_loc = np.concatenate([u_glob[gid + c * n_glob] for c in range(ncomp)])
    res = {}
    for overlap in (False, True):
        dop = DistributedOperator(ceed, bp, p, part, dev, overlap=overlap, transport=os.environ.get("CEED_B200_TEST_TRANSPORT", "auto"))
        dop.u_t.copy_(torch.from_numpy(u_loc))
        dop.apply()
        torch.cuda.synchronize()
        v1 = dop.v_t.cpu().numpy().copy()
        dop.v_t.fill_(7.0)  # Apply overwrites; a second step must give the same bits (deterministic order)
        dop.apply()
        torch.cuda.synchronize()
        assert np.array_equal(v1, dop.v_t.cpu().numpy()), "multi-rank step is not bitwise reproducible"
        res["v_overlap" if overlap else "v_serial"] = v1
        res["transport"] = dop.transport
        if overlap and cg_iters:
            # distributed CG (dot products all-reduced over NCCL): solve A x = A x_true for the mass operator
            from libceed_b200.cg import DeviceCG
            cg = DeviceCG(ceed, None, None, None, dop.n_local, dev, owned_mask=np.tile(part.owned_mask(), ncomp), dist_op=dop)
            x_true = torch.from_numpy(u_loc).to(dev)
            cg.p.copy_(x_true)
            cg.apply()
            b = cg.Ap.clone()
            cg.start(b)
            r0 = cg.residual_norm2()
            cg.iterate(cg_iters)
            res["cg_r0"], res["cg_r"] = r0, cg.residual_norm2()
            res["cg_x"] = cg.x.cpu().numpy()
        del dop
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), gid=gid, owned=part.owned_mask(), nc=ncomp, u=u_loc, **res)
    if rank == 0:
        prob = BPProblem(ceed, bp, p, n_global)
        prob.u.set_array(u_glob)
        prob.op.apply(prob.u, prob.v)
        np.savez(os.path.join(out_dir, "single.npz"), v=prob.v.get_array_read(), n_glob=n_glob)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
