"""Even-Q linear plane layout (stage bit 512) against the tuned table shape: same group shape with the bit added, plus a small sweep.
usage: python scripts/gpu_lin.py bp3p4 bp5p5 ..."""
import os, re, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libceed_b200 import Ceed, ceed as cm
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.mesh import choose_elements

for wl in sys.argv[1:]:
    m = re.fullmatch(r"bp(\d)p(\d)", wl); bp, p = int(m.group(1)), int(m.group(2))
    ceed = Ceed()
    base = BPProblem(ceed, bp, p, choose_elements(10e6, p, BP_TABLE[bp][0]))
    base.u.set_array(seeded_uniform(base.num_dofs))
    vref = [None]

    def run(tag, **shape):
        op = ceed.Operator(base.qf)
        op.set_field("u", base.rstr_u, base.basis_u, cm.VECTOR_ACTIVE)
        op.set_field("qdata", base.rstr_qd, cm.BASIS_NONE, base.qdata)
        op.set_field("v", base.rstr_u, base.basis_u, cm.VECTOR_ACTIVE)
        if shape: op.set_kernel_shape(**shape)
        op.set_timing(True)
        try:
            for _ in range(3): op.apply(base.u, base.v)
            t = []
            for _ in range(10):
                op.apply(base.u, base.v); t.append(op.last_kernel_ms())
        except Exception as e:
            print(f"{tag:44s} FAILED {str(e)[:140]}", flush=True); return None
        f, a = np.median([x[0] for x in t]), np.median([x[1] for x in t])
        i, s = op.kernel_info(), op.get_kernel_shape()
        v = base.v.get_array_read().copy()
        if vref[0] is None: vref[0] = v
        err = np.abs(v - vref[0]).max() / np.abs(vref[0]).max()
        print(f"{tag:44s} {f:.3f}+{a:.3f} ms {base.num_dofs/(f+a)/1e6:6.2f} GDoF/s {base.bytes_per_apply()/(f+a)/1e6/6550.1*100:5.1f}% regs={i['regs']} epw={i['elems_per_block']} "
              f"thr={i['threads']} grid={i['grid']} smem={i['smem_bytes']} stage={s['stage_mask']} err={err:.1e}", flush=True)
        return s

    print(f"{wl}: {base.num_dofs/1e6:.2f}M DoFs, {base.num_elem} elements")
    s0 = run("table")
    keys = ("elems_per_group", "group_warps", "cta_warps", "min_blocks_per_sm", "qf_mode", "qf_unroll")
    tab = {k: s0[k] for k in keys}
    st0 = s0["stage_mask"] & ~(256 | 512)
    run("table shape + lin (512)", **tab, stage_mask=st0 | 512)
    for E in sorted({max(1, tab["elems_per_group"] - 1), tab["elems_per_group"] + 1, tab["elems_per_group"] + 2}):
        run(f"lin E={E}", **{**tab, "elems_per_group": E}, stage_mask=st0 | 512)
    for gw, w in ((1, 4), (2, 4), (2, 8), (4, 4)):
        run(f"lin gw={gw} warps={w} (heuristic E)", group_warps=gw, cta_warps=w, qf_mode=0, stage_mask=513)
