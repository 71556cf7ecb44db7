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
    """Interface-DoF sum with NCCL send/recv as the transport (pack -> grouped P2P -> rank-ordered unpack)."""
    transport = "nccl"

    def __init__(self, part, ncomp, comp_stride, device, ceed=None, group=None, kernels=None):
        """part: mesh.Partition; the L-vector entry of (node n, component c) is n + c * comp_stride."""
        self.part, self.device, self.group = part, device, group
        self.ceed = ceed
        self.kernels = kernels if kernels is not None else CudaInterfaceKernels(ceed)
        ranks, seg, send_idx, node, ptr, src = build_interface_tables(part, ncomp, comp_stride)
        self.ranks, self.seg = ranks, [int(x) for x in seg]
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.send_idx, self.node, self.ptr, self.src = dev(send_idx), dev(node), dev(ptr), dev(src)
        self.send = torch.empty(send_idx.size, dtype=torch.float64, device=device)
        self.recv = torch.empty(send_idx.size, dtype=torch.float64, device=device)
        self.bytes_per_exchange = 2 * 8 * int(send_idx.size)

    def begin(self, v, stream=None):
        """Start the exchange of the interface values of v (complete partial sums).  `stream`: a torch CUDA stream to run the
        pack kernel and the NCCL operations on (overlap with work on the caller's stream); None = the current stream."""
        if not self.ranks:
            return
        ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
        main_ptr = self._main_stream_ptr()
        with ctx:
            if stream is not None and self.ceed is not None:
                self.ceed.set_stream(stream.cuda_stream)
            try:
                self.kernels.pack(v, self.send_idx, self.send)
            finally:
                if stream is not None and self.ceed is not None:
                    self.ceed.set_stream(main_ptr)
            ops = []
            for k, rank in enumerate(self.ranks):
                a, b = self.seg[k], self.seg[k + 1]
                ops.append(dist.P2POp(dist.isend, self.send[a:b], rank, group=self.group))
                ops.append(dist.P2POp(dist.irecv, self.recv[a:b], rank, group=self.group))
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def _main_stream_ptr(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0

    def end(self, v):
        """Finish the exchange: every interface entry of v becomes the rank-ordered sum of all partial values."""
        if not self.ranks:
            return v
        self.kernels.unpack_sum(v, self.node, self.ptr, self.src, self.recv)
        return v

    def sum_interfaces(self, v):
        """v: 1-D float64 torch tensor (the local L-vector after the local apply); updated in place."""
        self.begin(v)
        return self.end(v)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class PeerInterfaceExchange:
    """The same interface sum over NVLink peer memory, without NCCL on the data path.

    Every rank exposes one device allocation (CUDA IPC): double recv[2][total] + one arrival flag per neighbour.  `begin`
    launches ONE kernel (ceedb200_iface_put) that gathers the interface values and stores them straight into the neighbours'
    receive areas, then raises their flags; `end` launches ceedb200_iface_wait_unpack_sum, which waits for this rank's flags and
    forms the sums in ascending rank order (same bits as the NCCL transport).  The step counter lives on the device and the
    buffer halves alternate with its parity: nothing changes on the host from step to step, so the pair can be captured in a
    CUDA graph, and the put -- a few CTAs without shared memory -- runs next to the interior-element kernel.
    torch.distributed is used once, at construction, to swap the IPC handles and buffer offsets."""
    transport = "peer"

    def __init__(self, part, ncomp, comp_stride, device, ceed, group=None):
        if ceed is None:
            raise RuntimeError("PeerInterfaceExchange needs a libceed_b200 Ceed (CUDA kernels); there is no host fallback")
        self.part, self.device, self.group, self.ceed = part, device, group, ceed
        ranks, seg, send_idx, node, ptr, src = build_interface_tables(part, ncomp, comp_stride)
        self.ranks, self.seg = ranks, [int(x) for x in seg]
        total = int(send_idx.size)
        self.total = total
        lib, cp = ceed._lib, ceed._ptr
        # this rank's exposed allocation
        self._local = C.c_void_p()
        handle = C.create_string_buffer(64)
        nbytes = 2 * total * 8 + max(1, len(ranks)) * 8
        ceed._chk(lib.ceedb200_ipc_alloc(cp, nbytes, C.byref(self._local), handle))
        info = dict(rank=part.rank, handle=handle.raw, total=total, seg={r: self.seg[k] for k, r in enumerate(ranks)},
                    flag={r: k for k, r in enumerate(ranks)})
        world = dist.get_world_size(group)
        infos = [None] * world
        dist.all_gather_object(infos, info, group=group)
        self._remote = []
        peer_recv, peer_half, peer_flag = [], [], []
        for r in ranks:
            ri = infos[r]
            base = C.c_void_p()
            ceed._chk(lib.ceedb200_ipc_open(cp, ri["handle"], C.byref(base)))
            self._remote.append(base)
            peer_recv.append(base.value + 8 * ri["seg"][part.rank])          # my segment inside the neighbour's receive area
            peer_half.append(ri["total"])
            peer_flag.append(base.value + 16 * ri["total"] + 8 * ri["flag"][part.rank])
        dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=dt))).to(device)
        nb_of = np.zeros(total, dtype=np.int32)
        for k in range(len(ranks)):
            nb_of[self.seg[k]:self.seg[k + 1]] = k
        self.send_idx, self.nb_of, self.seg_d = dev(send_idx, np.int64), dev(nb_of, np.int32), dev(seg, np.int64)
        self.node, self.ptr, self.src = dev(node, np.int64), dev(ptr, np.int32), dev(src, np.int32)
        self.peer_recv, self.peer_half, self.peer_flag = dev(peer_recv, np.uint64), dev(peer_half, np.int64), dev(peer_flag, np.uint64)
        self.ctr = torch.zeros(2, dtype=torch.int64, device=device)
        self.bytes_per_exchange = 2 * 8 * total
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # every rank has mapped its neighbours before the first put

    def begin(self, v, stream=None):
        if not self.ranks:
            return
        p = lambda t: C.c_void_p(t.data_ptr())
        self.ceed._chk(self.ceed._lib.ceedb200_iface_put(self.ceed._ptr, C.c_void_p(stream.cuda_stream if stream is not None else 0), p(v), p(self.send_idx),
                                                         p(self.nb_of), p(self.seg_d), self.total, len(self.ranks), p(self.peer_recv), p(self.peer_half),
                                                         p(self.peer_flag), p(self.ctr)))

    def end(self, v):
        if not self.ranks:
            return v
        p = lambda t: C.c_void_p(t.data_ptr())
        self.ceed._chk(self.ceed._lib.ceedb200_iface_wait_unpack_sum(self.ceed._ptr, p(v), self.node.numel(), p(self.node), p(self.ptr), p(self.src),
                                                                     C.c_void_p(self._local.value), self.total, C.c_void_p(self._local.value + 16 * self.total),
                                                                     len(self.ranks), p(self.ctr)))
        return v

    def sum_interfaces(self, v):
        self.begin(v)
        return self.end(v)

    def close(self):
        lib, cp = self.ceed._lib, self.ceed._ptr
        for base in self._remote:
            lib.ceedb200_ipc_close(cp, base)
        self._remote = []
        if self._local:
            lib.ceedb200_ipc_free(cp, self._local)
            self._local = None


class DistributedOperator:
    """One BP operator on one rank's element box of a partitioned mesh + the interface-DoF sum: `apply()` is the multi-GPU step.

    overlap=True: the elements touching the rank interface are ordered first; the step applies them (operator part 1), starts
    the exchange of their -- now complete -- interface values on a second, high-priority stream, applies the interior elements
    (part 2) on the main stream meanwhile, and finishes with the rank-ordered sum.  overlap=False: apply, then exchange, serially.
    transport: "peer" (NVLink peer-memory stores from our own kernel), "nccl" (grouped send/recv), "auto" (peer, else nccl)."""

    def __init__(self, ceed, bp, p, part, device, overlap=True, transport="auto", group=None):
        from . import ceed as cm
        from .bp import BP_TABLE, BPProblem
        self.ceed, self.part, self.device, self.overlap = ceed, part, device, overlap
        ncomp = BP_TABLE[bp][0]
        perm, split = part.boundary_first_permutation() if overlap else (None, None)
        self.prob = BPProblem(ceed, bp, p, part.n_local, part=part, elem_perm=perm, split=split)
        self.n_local = self.prob.num_dofs
        self.u_t = torch.zeros(self.n_local, dtype=torch.float64, device=device)
        self.v_t = torch.zeros(self.n_local, dtype=torch.float64, device=device)
        self.prob.u.set_array(self.u_t, cm.MEM_DEVICE, cm.USE_POINTER)
        self.prob.v.set_array(self.v_t, cm.MEM_DEVICE, cm.USE_POINTER)
        self.exchange = None
        if part.size > 1:
            if transport in ("auto", "peer"):
                try:
                    self.exchange = PeerInterfaceExchange(part, ncomp, self.prob.num_nodes, device, ceed, group=group)
                except Exception as exc:  # noqa: BLE001 -- e.g. CUDA IPC not permitted in this container
                    if transport == "peer":
                        raise
                    print(f"[libceed_b200] peer-memory exchange unavailable ({exc}); using NCCL send/recv", flush=True)
                ok = torch.tensor([1 if self.exchange is not None else 0], device=device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # all ranks use the same transport
                if int(ok.item()) == 0:
                    self.exchange = None
            if self.exchange is None:
                self.exchange = InterfaceExchange(part, ncomp, self.prob.num_nodes, device, ceed=ceed, group=group)
        self.comm_stream = torch.cuda.Stream(device=device, priority=-1) if (overlap and self.exchange is not None) else None
        self.ev_boundary, self.ev_comm = torch.cuda.Event(), torch.cuda.Event()
        self.timeline = None  # set to a dict of timing events by record_timeline()

    @property
    def transport(self):
        return self.exchange.transport if self.exchange is not None else "none"

    def apply(self, u_vec=None, v_vec=None, v_t=None):
        """v = A u on this rank's box, interface DoFs summed over all ranks (every copy of a shared node ends with the same bits)."""
        prob = self.prob
        u_vec, v_vec = u_vec or prob.u, v_vec or prob.v
        v_t = self.v_t if v_t is None else v_t
        if self.exchange is None:
            prob.op.apply(u_vec, v_vec)
        elif not self.overlap:
            prob.op.apply(u_vec, v_vec)
            self.exchange.begin(v_t)
            self.exchange.end(v_t)
        else:
            main = torch.cuda.current_stream(self.device)
            tl = self.timeline
            if tl is not None:
                tl["start"].record(main)
            prob.op.apply_part(u_vec, v_vec, 1)        # boundary elements: interface values are complete after this
            self.ev_boundary.record(main)
            if tl is not None:
                tl["boundary_done"].record(main)
            self.comm_stream.wait_event(self.ev_boundary)
            self.exchange.begin(v_t, stream=self.comm_stream)
            self.ev_comm.record(self.comm_stream)
            if tl is not None:
                tl["send_done"].record(self.comm_stream)
            prob.op.apply_part(u_vec, v_vec, 2)        # interior elements, concurrently with the exchange
            if tl is not None:
                tl["interior_done"].record(main)
            main.wait_event(self.ev_comm)
            self.exchange.end(v_t)
            if tl is not None:
                tl["sum_done"].record(main)
        return v_t

    def record_timeline(self, steps=5):
        """Per-phase device timestamps (ms since the start of the step, median over `steps` steps) of the overlapped step on this
        rank: end of the boundary elements (+ their finalize), end of the interface send on the communication stream, end of the
        interior elements, end of the rank-ordered sum.  send_done < interior_done means the exchange was hidden."""
        if not (self.overlap and self.exchange is not None):
            return None
        names = ("start", "boundary_done", "send_done", "interior_done", "sum_done")
        rows = []
        for _ in range(steps):
            self.timeline = {n: torch.cuda.Event(enable_timing=True) for n in names}
            torch.cuda.synchronize(self.device)
            self.apply()
            torch.cuda.synchronize(self.device)
            rows.append([self.timeline["start"].elapsed_time(self.timeline[n]) for n in names[1:]])
        self.timeline = None
        med = np.median(np.array(rows), axis=0)
        return dict(zip(names[1:], [float(x) for x in med]))
