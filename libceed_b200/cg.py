"""Conjugate gradients around CeedOperatorApply, device-resident.

The reference's published figure of merit is "DoFs/sec in CG" (examples/petsc/bps.c:218-288: PETSc KSPCG around the libCEED
operator, VecDot / VecAXPY from PETSc).  Here one iteration is

    Ap = A p                      fused operator kernel (+ halo finalize) [+ interface sum over NCCL]
    pAp = <p, Ap>_w               ceedb200_cg_dot        [+ all-reduce]
    x += a p; r -= a Ap; rr' = <r, r>_w     ceedb200_cg_update     [+ all-reduce]      a = rr / pAp
    p = r + (rr'/rr) p            ceedb200_cg_direction

with every scalar on the device: no host synchronisation inside the loop, so `iterations` steps can be enqueued back to back
(or captured in a CUDA graph).  `w` is the ownership mask of the partition (1 = owned entry) so that the distributed dot
products count every DoF once.  Unpreconditioned, like the reference's BP runs with -pc_type none.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class DeviceCG:
    def __init__(self, ceed, op, u_vec, v_vec, n, device, exchange=None, owned_mask=None, group=None, dist_op=None, free_mask=None):
        """op: libceed_b200 Operator; u_vec / v_vec: its active input / output Vectors (length n) that wrap the torch tensors
        created here (USE_POINTER), so the operator reads p and writes Ap in place.
        dist_op: a parallel.DistributedOperator instead of (op, u_vec, v_vec, exchange): A p is then its (overlapped) multi-GPU step.
        free_mask: 1 = free DoF, 0 = essential (Dirichlet) DoF: constrained rows of A act as identity rows (ceedb200_cg_constrain), the
        right-hand side must be zero there (homogeneous conditions; lift inhomogeneous data into b beforehand)."""
        from . import ceed as cm
        self.dist_op = dist_op
        if dist_op is not None:
            op, u_vec, v_vec, exchange = dist_op.prob.op, dist_op.prob.u, dist_op.prob.v, dist_op.exchange
        self.ceed, self.op, self.n, self.exchange, self.group = ceed, op, n, exchange, group
        f64 = dict(dtype=torch.float64, device=device)
        self.x, self.r, self.p, self.Ap = (torch.zeros(n, **f64) for _ in range(4))
        self.w = None if owned_mask is None else torch.from_numpy(np.ascontiguousarray(owned_mask, dtype=np.float64)).to(device)
        self.free = None if free_mask is None else torch.from_numpy(np.ascontiguousarray(free_mask, dtype=np.float64)).to(device)
        self.scal = torch.zeros(4, **f64)  # rr, pAp, rr_new, spare
        self.u_vec, self.v_vec = u_vec, v_vec
        u_vec.set_array(self.p, cm.MEM_DEVICE, cm.USE_POINTER)
        v_vec.set_array(self.Ap, cm.MEM_DEVICE, cm.USE_POINTER)
        self.distributed = exchange is not None and dist.is_initialized() and dist.get_world_size(group) > 1

    def _ptr(self, t, offset=0):
        return C.c_void_p(t.data_ptr() + 8 * offset)

    def _allreduce(self, i):
        if self.distributed:
            dist.all_reduce(self.scal[i:i + 1], group=self.group)

    def apply(self):
        """Ap = A p (including the interface sum)."""
        if self.dist_op is not None:
            self.dist_op.apply(self.u_vec, self.v_vec, v_t=self.Ap)
        else:
            self.op.apply(self.u_vec, self.v_vec)
            if self.exchange is not None:
                self.exchange.sum_interfaces(self.Ap)
        if self.free is not None:
            self.ceed._chk(self.ceed._lib.ceedb200_cg_constrain(self.ceed._ptr, self._ptr(self.Ap), self._ptr(self.p), self._ptr(self.free), self.n))

    def start(self, b):
        """x = 0, r = p = b (b: torch tensor on the device, already consistent on interface copies); rr = <r, r>_w."""
        lib, cp = self.ceed._lib, self.ceed._ptr
        self.x.zero_()
        self.r.copy_(b)
        self.p.copy_(b)
        wp = self._ptr(self.w) if self.w is not None else None
        self.ceed._chk(lib.ceedb200_cg_dot(cp, self._ptr(self.r), self._ptr(self.r), wp, self.n, self._ptr(self.scal, 0)))
        self._allreduce(0)

    def iterate(self, iterations):
        lib, cp = self.ceed._lib, self.ceed._ptr
        wp = self._ptr(self.w) if self.w is not None else None
        rr, pAp, rrn = self._ptr(self.scal, 0), self._ptr(self.scal, 1), self._ptr(self.scal, 2)
        for _ in range(iterations):
            self.apply()
            self.ceed._chk(lib.ceedb200_cg_dot(cp, self._ptr(self.p), self._ptr(self.Ap), wp, self.n, pAp))
            self._allreduce(1)
            self.ceed._chk(lib.ceedb200_cg_update(cp, self._ptr(self.x), self._ptr(self.r), self._ptr(self.p), self._ptr(self.Ap), wp, self.n, rr, pAp, rrn))
            self._allreduce(2)
            self.ceed._chk(lib.ceedb200_cg_direction(cp, self._ptr(self.p), self._ptr(self.r), self.n, rrn, rr))
            self.scal[0:1].copy_(self.scal[2:3])  # rr = rr_new (device to device, stream ordered)

    def residual_norm2(self):
        """sqrt(<r, r>_w) (host value: synchronises)."""
        return float(self.scal[0].sqrt().item())
