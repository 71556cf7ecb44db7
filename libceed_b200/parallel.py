"""Multi-GPU layer: element-partitioned hex mesh, one process per GPU, interface-DoF sum over NCCL (NVLink 5 / NVSwitch).

The reference library has no communication; its examples delegate the local-to-global sum to PETSc
VecScatter ADD_VALUES (examples/petsc/bpsraw.c:240-262).  Here every rank owns a local L-vector that includes a copy of the
interface nodes (mesh.Partition); after the local CeedOperatorApply the interface values are exchanged with grouped
point-to-point sends/receives and added in ascending neighbour-rank order (deterministic, same bits on both sides of an
interface because both sides add the same two-or-more partial sums in the same global order).

torch.distributed is plumbing only: `backend="nccl"` on GPUs, `backend="gloo"` for the CPU tests of this logic.
"""
import numpy as np
import torch
import torch.distributed as dist


class InterfaceExchange:
    def __init__(self, part, ncomp, comp_stride, device, group=None):
        """part: mesh.Partition; the L-vector entry of (node n, component c) is n + c * comp_stride."""
        self.part, self.device, self.group = part, device, group
        self.neighbors = []
        for rank, idx in part.neighbors:
            full = np.concatenate([idx + c * comp_stride for c in range(ncomp)])
            t = torch.from_numpy(full).to(device)
            self.neighbors.append((rank, t, torch.empty(full.size, dtype=torch.float64, device=device),
                                   torch.empty(full.size, dtype=torch.float64, device=device)))
        self.bytes_per_exchange = sum(2 * 8 * t.numel() for _, t, _, _ in self.neighbors)

    def sum_interfaces(self, v):
        """v: 1-D float64 torch tensor (the local L-vector after the local apply); updated in place."""
        if not self.neighbors:
            return v
        ops = []
        # pack the local partial sums BEFORE anything is added, so every rank sends its own contribution only
        for rank, idx, send, recv in self.neighbors:
            torch.index_select(v, 0, idx, out=send)
            ops.append(dist.P2POp(dist.isend, send, rank, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv, rank, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        # global order: contributions are added by ascending rank on every owner of the node.  A node shared by ranks
        # {a < b < c} gets partial_a + partial_b + partial_c on all three: each rank inserts its own partial at its position.
        me = self.part.rank
        lower = [(r, i, rc) for r, i, _, rc in self.neighbors if r < me]
        upper = [(r, i, rc) for r, i, _, rc in self.neighbors if r > me]
        if lower:
            # ((p_r0 + p_r1) + ... ) + p_me : start from the lowest rank's value
            acc = torch.zeros_like(v)
            touched = torch.zeros(v.numel(), dtype=torch.bool, device=v.device)
            for r, idx, rc in lower:
                acc.index_add_(0, idx, rc)  # indices within one neighbour are unique -> deterministic
                touched[idx] = True
            v[touched] = acc[touched] + v[touched]
        for r, idx, rc in upper:
            v.index_add_(0, idx, rc)
        return v
