"""Round 2c, second GPU session: wider lean-kernel sweep on the headline (BP1 p=3: batch widths 3..7, warps per CTA, occupancy targets,
with / without the interleaved columns), and lean vs the shipped general x-line entry for BP2 p=4, BP1 p=5 (where the new stage bits may
move the crossover).  Results: gpurun_out/r02c_lean_sweep2.txt.  usage: python scripts/gpu_r02c2.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libceed_b200 import Ceed, ceed as cm
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.mesh import choose_elements
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
log = open(os.path.join(OUT, "r02c_lean_sweep3.txt" if "high" in sys.argv else "r02c_lean_sweep2.txt"), "w")
T0 = time.time()
BUDGET = float(os.environ.get("R02C_SWEEP_BUDGET", "200"))


def say(s):
    print(s, flush=True); log.write(s + "\n"); log.flush()


def sweep(bp, p, cands):
    ceed = Ceed()
    base = BPProblem(ceed, bp, p, choose_elements(10e6, p, BP_TABLE[bp][0]))
    base.u.set_array(seeded_uniform(base.num_dofs))
    say(f"bp{bp} p={p}: {base.num_dofs / 1e6:.2f} M DoFs, {base.num_elem} elements")
    vref = [None]

    def run(tag, **shape):
        if time.time() - T0 > BUDGET: return
        op = ceed.Operator(base.qf)
        op.set_field("u", base.rstr_u, base.basis_u, cm.VECTOR_ACTIVE)
        op.set_field("qdata", base.rstr_qd, cm.BASIS_NONE, base.qdata)
        op.set_field("v", base.rstr_u, base.basis_u, cm.VECTOR_ACTIVE)
        if shape: op.set_kernel_shape(**shape)
        op.set_timing(True)
        try:
            for _ in range(3): op.apply(base.u, base.v)
            t = []
            for _ in range(9):
                op.apply(base.u, base.v); t.append(op.last_kernel_ms())
        except Exception as e:  # noqa
            say(f"  {tag:48s} FAILED {str(e)[:200]}"); return
        f, a = float(np.median([x[0] for x in t])), float(np.median([x[1] for x in t]))
        i, got = op.kernel_info(), op.get_kernel_shape()
        v = base.v.get_array_read()
        if vref[0] is None: vref[0] = v.copy()
        say(f"  {tag:48s} {f:.4f}+{a:.4f} = {f + a:.4f} ms {base.num_dofs / (f + a) / 1e6:6.2f} GDoF/s regs={i['regs']} smem={i['smem_bytes']} grid={i['grid']} loc={i['local_bytes']} "
            f"minb={got['min_blocks_per_sm']} bitwise={bool(np.array_equal(v, vref[0]))}")

    run("table entry (as shipped)")
    for tag, shape in cands:
        run(tag, **shape)


def lean(E, warps, stage, minb=0):
    mb = minb or max(1, 65536 // (warps * 32 * 96))
    return (f"lean E={E} warps={warps} stage={stage} minb={mb}", dict(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage, min_blocks_per_sm=mb, qf_unroll=4))


if "high" in sys.argv:
    # third session: lean kernel with one or two elements per warp against the shipped general x-line entries at p = 6..8
    for bp, p in ((1, 6), (2, 6), (1, 7), (2, 7), (1, 8)):
        sweep(bp, p, [lean(E, w, s) for E in (1, 2) for w in (4, 8) for s in (0, 1)])
else:
    c13 = [lean(6, 4, s, mb) for s in (0, 1) for mb in (4, 6)]
    c13 += [lean(E, w, s) for E in (4, 5, 7, 3) for w in (4, 2) for s in (0, 1)]
    c13 += [lean(4, 4, 3), lean(4, 8, 1), lean(5, 4, 3), lean(4, 4, 1, 4), lean(4, 4, 1, 6)]
    sweep(1, 3, c13)
    sweep(2, 4, [lean(E, w, s) for E in (2, 3, 4) for w in (4, 8) for s in (0, 1, 5)])
    sweep(1, 5, [lean(E, w, s) for E in (2, 3, 4) for w in (4, 8) for s in (0, 1, 5)])
    sweep(2, 5, [lean(E, w, s) for E in (1, 2, 3) for w in (4, 8) for s in (0, 1)])
say(f"done ({time.time() - T0:.0f} s)")
