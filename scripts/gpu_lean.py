"""Lean in-place-plane kernel (QFunction layout 4) on a B200: parity against the oracle over operators / batch widths / tails /
scatter modes, then a shape sweep on the headline workload.  usage: python scripts/gpu_lean.py [parity|sweep|all] [bpXpY] [dofs]"""
import os, re, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CEED_B200_NO_TUNE_TABLE", "1")
from libceed_b200 import Ceed, ceed as cm
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.mesh import choose_elements

what = sys.argv[1] if len(sys.argv) > 1 else "all"
wl = sys.argv[2] if len(sys.argv) > 2 else "bp1p3"
dofs = float(sys.argv[3]) if len(sys.argv) > 3 else 10e6


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def parity():
    from oracle import oracle as O
    bad = 0
    ceed = Ceed()
    for bp, p, nel in [(1, 3, (4, 3, 3)), (1, 3, (5, 3, 3)), (1, 3, (7, 5, 3)), (2, 3, (3, 2, 3)), (1, 1, (5, 4, 3)), (1, 2, (3, 3, 3)), (1, 4, (3, 2, 3)), (1, 5, (2, 3, 2)),
                       (1, 6, (2, 2, 2)), (1, 7, (2, 1, 2)), (2, 2, (3, 3, 2)), (2, 5, (2, 2, 1)), (2, 1, (4, 4, 3))]:
        from libceed_b200 import mesh as M
        prob = BPProblem(ceed, bp, p, nel, elem_perm=M.morton_permutation(*nel) if (bp + p) % 2 else None)
        u = seeded_uniform(prob.num_dofs, 31)
        prob.u.set_array(u)
        qd = O.bp_qdata(bp, p, prob.offsets, prob.coords)
        ref = O.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)
        for scatter in (0, 1):
            ceed.set_scatter_mode(scatter)
            op = ceed.Operator(prob.qf)
            op.set_field("u", prob.rstr_u, prob.basis_u, cm.VECTOR_ACTIVE)
            op.set_field("qdata", prob.rstr_qd, cm.BASIS_NONE, prob.qdata)
            op.set_field("v", prob.rstr_u, prob.basis_u, cm.VECTOR_ACTIVE)
            for E, warps, stage in ((3, 2, 0), (6, 4, 0), (3, 1, 32), (6, 4, 40), (5, 2, 8)):
                op.set_kernel_shape(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
                prob.v.set_value(-3.0)
                op.apply(prob.u, prob.v)
                got = op.get_kernel_shape()
                e1 = rel(prob.v.get_array_read(), ref)
                w0 = seeded_uniform(prob.num_dofs, 5)
                prob.v.set_array(w0)
                op.apply_add(prob.u, prob.v)
                e2 = rel(prob.v.get_array_read() - w0, ref)
                ok = (got["qf_mode"] == 4 or p == 7) and (got["stage_mask"] == stage or p == 7) and e1 < 1e-12 and e2 < 1e-10
                bad += not ok
                print(f"bp{bp} p={p} nel={nel} scatter={scatter} E={E} warps={warps} stage={got['stage_mask']}: layout {got['qf_mode']} epw {got['elems_per_group']} apply {e1:.1e} add {e2:.1e} "
                      f"{'ok' if ok else 'FAIL'}", flush=True)
        ceed.set_scatter_mode(0)
    print("parity:", "all ok" if not bad else f"{bad} FAILED")
    return bad


def parity_runs():
    """run scatter (multi-iteration runs) vs oracle and bitwise vs the classic owner/halo tables"""
    from oracle import oracle as O
    from libceed_b200 import mesh as M
    bad = 0
    ceed = Ceed()
    for bp, p, nel, morton in [(1, 3, (32, 31, 31), True), (1, 3, (33, 30, 29), False), (2, 2, (30, 29, 28), True), (1, 1, (50, 50, 49), True)]:
        prob = BPProblem(ceed, bp, p, nel, elem_perm=M.morton_permutation(*nel) if morton else None)
        u = seeded_uniform(prob.num_dofs, 37)
        prob.u.set_array(u)
        qd = O.bp_qdata(bp, p, prob.offsets, prob.coords)
        ref = O.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)
        os.environ["CEED_B200_NO_RUNS"] = "1"
        prob.op.set_kernel_shape(qf_mode=4, elems_per_group=4, cta_warps=4, group_warps=1, stage_mask=0)
        prob.op.apply(prob.u, prob.v)
        v_classic = prob.v.get_array_read().copy()
        os.environ.pop("CEED_B200_NO_RUNS")
        for E, warps, stage, parts in ((1, 4, 0, 0), (6, 4, 0, 0), (6, 4, 128, 8), (2, 4, 128, 3), (5, 2, 128, 16), (8, 8, 128, 5)):
            if stage:
                os.environ.pop("CEED_B200_RUNS", None); os.environ["CEED_B200_PARTS"] = str(parts)
            else:
                os.environ["CEED_B200_RUNS"] = "1"
            prob.op.set_kernel_shape(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
            prob.v.set_value(-3.0)
            prob.op.apply(prob.u, prob.v)
            v = prob.v.get_array_read().copy()
            w0 = seeded_uniform(prob.num_dofs, 5)
            prob.v.set_array(w0)
            prob.op.apply_add(prob.u, prob.v)
            e2 = rel(prob.v.get_array_read() - w0, ref)
            ok = rel(v, ref) < 1e-12 and np.array_equal(v, v_classic) and e2 < 1e-9
            bad += not ok
            print(f"runs bp{bp} p={p} nel={nel} morton={morton} E={E} warps={warps} stage={prob.op.get_kernel_shape()['stage_mask']} parts={parts}: vs oracle {rel(v, ref):.1e}, bitwise == classic tables {np.array_equal(v, v_classic)}, "
                  f"add {e2:.1e} {'ok' if ok else 'FAIL'}", flush=True)
    print("parity_runs:", "all ok" if not bad else f"{bad} FAILED")
    return bad


def sweep():
    m = re.fullmatch(r"bp(\d)p(\d)", wl)
    bp, p = int(m.group(1)), int(m.group(2))
    ceed = Ceed()
    base = BPProblem(ceed, bp, p, choose_elements(dofs, p, BP_TABLE[bp][0]))
    base.u.set_array(seeded_uniform(base.num_dofs))
    print(f"{wl}: {base.num_dofs/1e6:.2f}M DoFs, {base.num_elem} elements")
    vref = [None]

    def run(tag, scatter=0, **shape):
        ceed.set_scatter_mode(scatter)
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
            print(f"{tag:40s} FAILED {str(e)[:160]}", flush=True); return
        f, a = np.median([x[0] for x in t]), np.median([x[1] for x in t])
        i = op.kernel_info()
        v = base.v.get_array_read().copy()
        if vref[0] is None: vref[0] = v
        err = rel(v, vref[0])
        print(f"{tag:40s} {f:.3f}+{a:.3f} ms {base.num_dofs/(f+a)/1e6:6.2f} GDoF/s {base.bytes_per_apply()/(f+a)/1e6/6550.1*100:5.1f}% regs={i['regs']} "
              f"epw={i['elems_per_block']} thr={i['threads']} grid={i['grid']} smem={i['smem_bytes']} loc={i['local_bytes']} err={err:.1e}", flush=True)

    os.environ.pop("CEED_B200_NO_TUNE_TABLE", None)
    if os.environ.get("LEAN_RUNS"):
        from libceed_b200 import mesh as M
        nel = choose_elements(dofs, p, BP_TABLE[bp][0])
        probs = {"lex": base, "morton": BPProblem(ceed, bp, p, nel, elem_perm=M.morton_permutation(*nel))}
        probs["morton"].u.set_array(seeded_uniform(base.num_dofs))
        for order, prob in probs.items():
            base = prob
            vref[0] = None
            for runs in (0, 1):
                if runs: os.environ.pop("CEED_B200_NO_RUNS", None)
                else: os.environ["CEED_B200_NO_RUNS"] = "1"
                for E, warps in ((4, 4), (6, 4), (6, 2), (8, 4), (6, 8)):
                    run(f"{order} runs={runs} lean E={E} warps={warps}", qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=0)
            os.environ.pop("CEED_B200_NO_RUNS", None)
            run(f"{order} general kernel (table)")
        return
    run("table (general kernel)")
    for parts in (2, 3, 4, 6, 8, 12):
        os.environ["CEED_B200_PARTS"] = str(parts)
        for E, warps in ((6, 4), (4, 4)):
            run(f"lean fin parts={parts} E={E} warps={warps}", qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=128)
    os.environ.pop("CEED_B200_PARTS", None)
    for stage in (0,):
        for E, warps, minb in ((6, 4, 0),):
            run(f"lean stage={stage} E={E} warps={warps} minb={minb}", qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage, min_blocks_per_sm=minb)
    return
    for stage in (0, 8, 32, 40):
        for E, warps in ((4, 4), (6, 4), (6, 8), (8, 4), (3, 4), (2, 4)):
            run(f"lean stage={stage} E={E} warps={warps}", qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
    return
    for E in (4, 6, 8):
        run(f"lean E={E} warps=4", qf_mode=4, elems_per_group=E, cta_warps=4, group_warps=1, stage_mask=0)
    for E in (3, 4, 5, 6, 8):
        for warps in (2, 4, 8):
            for minb in (0,):
                run(f"lean bulk E={E} warps={warps}", qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, min_blocks_per_sm=minb, stage_mask=32)
    run("lean bulk E=6 warps=4 minb=3", qf_mode=4, elems_per_group=6, cta_warps=4, group_warps=1, min_blocks_per_sm=3, stage_mask=32)
    run("lean bulk E=6 warps=4 minb=5", qf_mode=4, elems_per_group=6, cta_warps=4, group_warps=1, min_blocks_per_sm=5, stage_mask=32)
    run("lean bulk E=4 warps=4 minb=6", qf_mode=4, elems_per_group=4, cta_warps=4, group_warps=1, min_blocks_per_sm=6, stage_mask=32)
    run("lean bulk E=6 warps=4 atomic", scatter=1, qf_mode=4, elems_per_group=6, cta_warps=4, group_warps=1, stage_mask=32)
    run("lean E=6 warps=4 atomic", scatter=1, qf_mode=4, elems_per_group=6, cta_warps=4, group_warps=1, stage_mask=0)


rc = 0
if what in ("parity", "all"): rc = parity()
if what in ("runs", "all"): rc += parity_runs()
if what in ("sweep", "all"): sweep()
sys.exit(1 if rc else 0)
