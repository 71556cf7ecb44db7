"""Multi-GPU layer: element-partitioned hex mesh, one process per GPU, interface-DoF sum over NCCL (NVLink 5 / NVSwitch).

The reference library has no communication; its examples delegate the local-to-global sum to PETSc
VecScatter ADD_VALUES (examples/petsc/bpsraw.c:240-262).  Here every rank owns a local L-vector that includes a copy of the
interface nodes (mesh.Partition).  After the local CeedOperatorApply:

  1. `ceedb200_iface_pack` (CUDA) gathers this rank's partial sums of all interface entries into ONE send buffer
     (one contiguous segment per neighbour rank);
  2. the segments are exchanged with grouped NCCL send/recv (torch.distributed.batch_isend_irecv: plumbing only);
  3. `ceedb200_iface_unpack_sum` (CUDA) rebuilds every interface entry as the sum of all partial values in ASCENDING RANK
     ORDER with the own value at its place, so all copies of a node carry identical bits on every rank (deterministic).

The tables (what to send where, in which order to add) are built once on the host with numpy.  The two kernels are the
backend's own; `kernels=` lets the CPU tests of the table logic (gloo, tests/test_parallel_gloo.py) inject a host executor --
the product default requires the CUDA library and fails loudly without it.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class CudaInterfaceKernels:
    """pack / unpack through the C ABI (include/ceed_b200.h) on the Ceed's stream."""

    def __init__(self, ceed):
        if ceed is None:
            raise RuntimeError("InterfaceExchange needs a libceed_b200 Ceed (CUDA kernels); there is no host fallback")
        self.ceed = ceed

    def pack(self, v, idx, send):
        assert v.is_cuda and idx.is_cuda and send.is_cuda
        self.ceed._chk(self.ceed._lib.ceedb200_iface_pack(self.ceed._ptr, C.c_void_p(v.data_ptr()), C.c_void_p(idx.data_ptr()), idx.numel(),
                                                          C.c_void_p(send.data_ptr())))

    def unpack_sum(self, v, node, ptr, src, recv):
        self.ceed._chk(self.ceed._lib.ceedb200_iface_unpack_sum(self.ceed._ptr, C.c_void_p(v.data_ptr()), node.numel(), C.c_void_p(node.data_ptr()),
                                                                C.c_void_p(ptr.data_ptr()), C.c_void_p(src.data_ptr()), C.c_void_p(recv.data_ptr())))


def build_interface_tables(part, ncomp, comp_stride):
    """Host tables of the exchange.  Returns (ranks, seg_offsets, send_idx, node, ptr, src):
    ranks[k] / seg_offsets[k]..seg_offsets[k+1] = neighbour k and its segment of the send and receive buffers;
    send_idx[i] = L-index whose value goes to send position i;
    node / ptr / src = CSR over the unique interface entries: contributions in ascending rank order, src -1 = own value."""
    ranks, fulls = [], []
    for rank, idx in part.neighbors:  # sorted by rank; both sides list the shared nodes in the same (global lexicographic) order
        ranks.append(rank)
        fulls.append(np.concatenate([idx + c * comp_stride for c in range(ncomp)]).astype(np.int64))
    sizes = np.array([f.size for f in fulls], dtype=np.int64)
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    if not fulls:
        z = np.zeros(0, dtype=np.int64)
        return ranks, seg, z, z, np.zeros(1, dtype=np.int32), np.zeros(0, dtype=np.int32)
    send_idx = np.concatenate(fulls)
    # contributions: (L-index, rank, position in the receive buffer); the own value takes part with position -1
    l_all = np.concatenate(fulls)
    r_all = np.concatenate([np.full(f.size, r, dtype=np.int64) for r, f in zip(ranks, fulls)])
    p_all = np.arange(l_all.size, dtype=np.int64)
    own = np.unique(l_all)
    l_all = np.concatenate([l_all, own])
    r_all = np.concatenate([r_all, np.full(own.size, part.rank, dtype=np.int64)])
    p_all = np.concatenate([p_all, np.full(own.size, -1, dtype=np.int64)])
    order = np.lexsort((r_all, l_all))  # primary key L-index, secondary key rank
    l_s, p_s = l_all[order], p_all[order]
    node, counts = np.unique(l_s, return_counts=True)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    assert p_s.size < 2 ** 31
    return ranks, seg, send_idx, node.astype(np.int64), ptr, p_s.astype(np.int32)


class InterfaceExchange:
    def __init__(self, part, ncomp, comp_stride, device, ceed=None, group=None, kernels=None):
        """part: mesh.Partition; the L-vector entry of (node n, component c) is n + c * comp_stride."""
        self.part, self.device, self.group = part, device, group
        self.kernels = kernels if kernels is not None else CudaInterfaceKernels(ceed)
        ranks, seg, send_idx, node, ptr, src = build_interface_tables(part, ncomp, comp_stride)
        self.ranks, self.seg = ranks, [int(x) for x in seg]
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.send_idx, self.node, self.ptr, self.src = dev(send_idx), dev(node), dev(ptr), dev(src)
        self.send = torch.empty(send_idx.size, dtype=torch.float64, device=device)
        self.recv = torch.empty(send_idx.size, dtype=torch.float64, device=device)
        self.bytes_per_exchange = 2 * 8 * int(send_idx.size)

    def sum_interfaces(self, v):
        """v: 1-D float64 torch tensor (the local L-vector after the local apply); updated in place."""
        if not self.ranks:
            return v
        self.kernels.pack(v, self.send_idx, self.send)
        ops = []
        for k, rank in enumerate(self.ranks):
            a, b = self.seg[k], self.seg[k + 1]
            ops.append(dist.P2POp(dist.isend, self.send[a:b], rank, group=self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv[a:b], rank, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.kernels.unpack_sum(v, self.node, self.ptr, self.src, self.recv)
        return v
