"""CEED bake-off problems BP1-BP6 as libCEED operators (host-side wiring, no kernels here).

Problem definitions follow examples/petsc/include/bpsproblemdata.h:29-133 and the operator wiring of
examples/petsc/src/libceedsetup.c:73-140:
    BP1/BP2  mass,      Gauss quadrature  Q = P + q_extra(=1)... here q = p + 2 as in the bake-off
    BP3/BP4  diffusion, Gauss quadrature  q = p + 2
    BP5/BP6  diffusion, Gauss-Lobatto     q = p + 1 (collocated)
    even numbers: 3 components
setup operator:  x INTERP + dx GRAD + weight WEIGHT -> qdata NONE   (strided, CEED_STRIDES_BACKEND)
apply operator:  u INTERP|GRAD + qdata NONE        -> v INTERP|GRAD
"""
import os

import numpy as np

from . import ceed as cm
from . import mesh as M

BP_TABLE = {
    #      ncomp, kind,  q_extra, qmode,            setup QF,          apply QF,   qdata comps
    1: (1, "mass", 2, cm.GAUSS, "BPSetupMassGeo", "BPMass", 1),
    2: (3, "mass", 2, cm.GAUSS, "BPSetupMassGeo", "BPMass3", 1),
    3: (1, "diff", 2, cm.GAUSS, "BPSetupDiffGeo", "BPDiff", 7),
    4: (3, "diff", 2, cm.GAUSS, "BPSetupDiffGeo", "BPDiff3", 7),
    5: (1, "diff", 1, cm.GAUSS_LOBATTO, "BPSetupDiffGeo", "BPDiff", 7),
    6: (3, "diff", 1, cm.GAUSS_LOBATTO, "BPSetupDiffGeo", "BPDiff3", 7),
}

GEO_H = os.path.join(cm.QFUNCTION_DIR, "bp_geo.h")
APPLY_H = os.path.join(cm.QFUNCTION_DIR, "bp_apply.h")


def algorithmic_bytes(bp, p, num_elem, num_nodes):
    """Compulsory HBM traffic of one CeedOperatorApply:
    read u + write v + read qdata + read int32 offsets."""
    ncomp, _, q_extra, _, _, _, ncq = BP_TABLE[bp]
    P, Q = p + 1, p + q_extra
    return 16 * ncomp * num_nodes + (8 * ncq * Q ** 3 + 4 * P ** 3) * num_elem


def seeded_uniform(n, seed=0x5EED):
    """uniform(-1, 1) from a fixed seed; identical bytes for every backend."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-1.0, 1.0, size=n)


class BPProblem:
    """One BP operator on one structured hex (sub-)mesh, built on a `Ceed` context of the b200 backend."""

    def __init__(self, ceed, bp, p, nelem_xyz, part=None, interlaced=False, build_qdata=True, elem_perm=None, split=None):
        """elem_perm / split: element order of a partitioned mesh with the `split` interface-touching elements first
        (mesh.Partition.boundary_first_permutation) -- enables Operator.apply_part."""
        ncomp, kind, q_extra, qmode, setup_name, apply_name, ncq = BP_TABLE[bp]
        self.ceed, self.bp, self.p, self.ncomp, self.kind, self.ncq = ceed, bp, p, ncomp, kind, ncq
        P, Q = p + 1, p + q_extra
        self.P, self.Q = P, Q
        if part is None:
            nx, ny, nz = nelem_xyz
            coords = M.hex_coords(nx, ny, nz, p)
        else:
            nx, ny, nz = part.n_local
            coords = M.hex_coords(nx, ny, nz, p, n_global=part.n_global, e0=part.e0)
        self.nelem_xyz = (nx, ny, nz)
        self.num_elem = nx * ny * nz
        self.num_nodes = (nx * p + 1) * (ny * p + 1) * (nz * p + 1)
        self.num_dofs = self.num_nodes * ncomp
        offsets = M.hex_offsets(nx, ny, nz, p)
        if elem_perm is not None:
            offsets = np.ascontiguousarray(offsets[elem_perm])
        self.offsets = offsets
        nn = self.num_nodes
        # restrictions
        self.rstr_x = ceed.ElemRestriction(self.num_elem, P ** 3, 3, nn, 3 * nn, offsets)
        if interlaced and ncomp > 1:
            # PETSc-style interlaced components (examples/petsc/bpsraw.c:92-94): offset * ncomp, comp_stride 1
            self.rstr_u = ceed.ElemRestriction(self.num_elem, P ** 3, ncomp, 1, ncomp * nn, offsets * ncomp)
        else:
            self.rstr_u = ceed.ElemRestriction(self.num_elem, P ** 3, ncomp, nn, ncomp * nn, offsets)
        if split is not None:
            self.rstr_u.set_split(split)
        self.rstr_qd = ceed.StridedElemRestriction(self.num_elem, Q ** 3, ncq, self.num_elem * Q ** 3 * ncq, None)
        # bases
        self.basis_x = ceed.BasisTensorH1Lagrange(3, 3, P, Q, qmode)
        self.basis_u = ceed.BasisTensorH1Lagrange(3, ncomp, P, Q, qmode)
        # vectors
        self.coords = coords
        self.x = ceed.Vector(3 * nn)
        self.x.set_array(coords.reshape(-1))
        self.qdata = ceed.Vector(self.num_elem * Q ** 3 * ncq)
        # setup operator
        qf_setup = ceed.QFunction(GEO_H, setup_name)
        qf_setup.add_input("x", 3, cm.EVAL_INTERP)
        qf_setup.add_input("dx", 9, cm.EVAL_GRAD)
        qf_setup.add_input("weight", 1, cm.EVAL_WEIGHT)
        qf_setup.add_output("qdata", ncq, cm.EVAL_NONE)
        self.op_setup = ceed.Operator(qf_setup)
        self.op_setup.set_field("x", self.rstr_x, self.basis_x, cm.VECTOR_ACTIVE)
        self.op_setup.set_field("dx", self.rstr_x, self.basis_x, cm.VECTOR_ACTIVE)
        self.op_setup.set_field("weight", cm.ELEMRESTRICTION_NONE, self.basis_x, cm.VECTOR_NONE)
        self.op_setup.set_field("qdata", self.rstr_qd, cm.BASIS_NONE, cm.VECTOR_ACTIVE)
        if build_qdata:
            self.op_setup.apply(self.x, self.qdata)
        # apply operator
        qf = ceed.QFunction(APPLY_H, apply_name)
        if kind == "mass":
            qf.add_input("u", ncomp, cm.EVAL_INTERP)
            qf.add_input("qdata", ncq, cm.EVAL_NONE)
            qf.add_output("v", ncomp, cm.EVAL_INTERP)
        else:
            qf.add_input("u", 3 * ncomp, cm.EVAL_GRAD)
            qf.add_input("qdata", ncq, cm.EVAL_NONE)
            qf.add_output("v", 3 * ncomp, cm.EVAL_GRAD)
        self.qf = qf
        self.op = ceed.Operator(qf)
        self.op.set_field("u", self.rstr_u, self.basis_u, cm.VECTOR_ACTIVE)
        self.op.set_field("qdata", self.rstr_qd, cm.BASIS_NONE, self.qdata)
        self.op.set_field("v", self.rstr_u, self.basis_u, cm.VECTOR_ACTIVE)
        self.u = ceed.Vector(self.num_dofs)
        self.v = ceed.Vector(self.num_dofs)

    def bytes_per_apply(self):
        return algorithmic_bytes(self.bp, self.p, self.num_elem, self.num_nodes)
