"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (oracle/_ref/lib/libceed.so,
resource /cpu/self/ref/serial) through its public C API.  Run in the dev container (needs oracle/_ref, built by
`make -C oracle ref`); the fixtures are committed so that tests on the GPU box do not need /root/reference.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libceed_b200 import mesh as M  # noqa: E402  (host-side mesh builder only)
from libceed_b200.bp import seeded_uniform  # noqa: E402
from oracle import refceed as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

BASIS_CASES = [(P, Q, qm) for P in range(2, 10) for (Q, qm) in ((P + 1, 0), (P, 1), (P + 2, 1), (max(2, P - 1), 0))]
BP_CASES = [  # bp, p, (nx, ny, nz), gallery, interlaced
    (1, 1, (3, 2, 2), False, False), (1, 3, (2, 2, 2), False, False), (3, 1, (2, 3, 2), False, False), (3, 2, (2, 2, 2), False, False),
    (3, 3, (2, 2, 1), False, False), (3, 4, (2, 1, 1), False, False), (3, 6, (1, 1, 2), False, False), (3, 8, (1, 1, 1), False, False),
    (5, 4, (2, 1, 1), False, False), (5, 7, (1, 1, 1), False, False), (2, 2, (2, 2, 1), False, False), (4, 1, (2, 2, 2), False, False),
    (4, 2, (2, 1, 2), False, True), (6, 2, (2, 2, 1), False, False), (6, 3, (1, 2, 1), False, True), (1, 2, (2, 2, 2), True, False),
    (3, 2, (2, 2, 2), True, False), (3, 5, (2, 1, 1), False, False), (3, 7, (1, 1, 1), False, False), (5, 5, (1, 2, 1), False, False),
    (5, 6, (1, 1, 2), False, False), (6, 4, (2, 1, 1), False, False), (6, 6, (1, 1, 1), False, False), (4, 3, (2, 1, 1), False, False),
    (1, 2, (2, 2, 2), False, False), (1, 4, (2, 1, 1), False, False), (1, 5, (1, 1, 2), False, False), (2, 3, (2, 1, 1), False, False),
]


def main():
    rc = R.RefCeed("/cpu/self/ref/serial")
    out = {}
    # ---- basis matrices (interface/ceed-basis.c)
    for P, Q, qm in BASIS_CASES:
        m = rc.basis_matrices(rc.basis_lagrange(3, 1, P, Q, qm), P, Q)
        for k, v in m.items():
            out[f"basis_P{P}_Q{Q}_m{qm}_{k}"] = v
    # ---- restriction (backends/ref/ceed-ref-restriction.c), E-layout of the CPU reference reported alongside
    rng = np.random.default_rng(7)
    nelem, esize, ncomp, lsize = 5, 6, 2, 14
    offsets = rng.integers(0, lsize, size=nelem * esize).astype(np.int32)
    u = rng.uniform(-1, 1, ncomp * lsize)
    r = rc.restriction(nelem, esize, ncomp, lsize, ncomp * lsize, offsets)
    ev = rc.vector(nelem * esize * ncomp)
    rc.restriction_apply(r, R.NOTRANSPOSE, rc.vector(ncomp * lsize, u), ev)
    e = rc.get_array(ev, nelem * esize * ncomp)
    lv = rc.vector(ncomp * lsize)
    rc.set_value(lv, 0.0)
    w = rng.uniform(-1, 1, nelem * esize * ncomp)
    rc.restriction_apply(r, R.TRANSPOSE, rc.vector(w.size, w), lv)
    out.update(rstr_offsets=offsets, rstr_u=u, rstr_e=e, rstr_layout=np.array(rc.restriction_e_layout(r)), rstr_w=w,
               rstr_lt=rc.get_array(lv, ncomp * lsize), rstr_dims=np.array([nelem, esize, ncomp, lsize]))
    # ---- BP operators
    for bp, p, nel, gallery, interlaced in BP_CASES:
        nx, ny, nz = nel
        off = M.hex_offsets(nx, ny, nz, p)
        coords = M.hex_coords(nx, ny, nz, p)
        nn = coords.shape[1]
        ref = R.RefBP(rc, bp, p, off.shape[0], nn, off, coords, gallery=gallery, interlaced=interlaced)
        u = seeded_uniform(ref.ncomp * nn)
        key = f"bp{bp}_p{p}_n{nx}x{ny}x{nz}_g{int(gallery)}_i{int(interlaced)}"
        out[key + "_v"] = ref.apply(u)
        out[key + "_qdata"] = ref.qdata_array()
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
