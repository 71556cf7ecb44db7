"""First GPU contact: self-consistency of the fused kernel (fused vs unfused path, three scatter modes) and a timing sweep.
usage: python scripts/gpu_first.py [quick]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libceed_b200 import Ceed, ceed as cm  # noqa: E402
from libceed_b200.bp import BPProblem, seeded_uniform  # noqa: E402
from libceed_b200.mesh import choose_elements  # noqa: E402


def apply_once(bp, p, nel, mode, fuse=True, interlaced=False):
    if not fuse:
        os.environ["CEED_B200_NO_FUSE"] = "1"
    else:
        os.environ.pop("CEED_B200_NO_FUSE", None)
    ceed = Ceed()
    ceed.set_scatter_mode(mode)
    prob = BPProblem(ceed, bp, p, nel, interlaced=interlaced)
    u = seeded_uniform(prob.num_dofs)
    prob.u.set_array(u)
    prob.op.apply(prob.u, prob.v)
    v = prob.v.get_array_read()
    # apply_add on top must give 2 v
    prob.op.apply_add(prob.u, prob.v)
    v2 = prob.v.get_array_read()
    qd = prob.qdata.get_array_read()
    fused = prob.op.is_fused
    return v, v2, qd, fused


def check():
    ok = True
    for bp, p, nel in [(1, 1, (3, 2, 2)), (1, 3, (3, 3, 2)), (3, 1, (2, 3, 4)), (3, 2, (3, 3, 3)), (3, 4, (2, 2, 3)), (5, 3, (3, 2, 2)),
                       (2, 2, (2, 2, 2)), (4, 2, (2, 3, 2)), (6, 3, (2, 2, 2)), (3, 6, (2, 2, 2)), (3, 8, (2, 1, 1))]:
        ref_v, ref_v2, ref_qd, f0 = apply_once(bp, p, nel, cm.SCATTER_DETERMINISTIC, fuse=False)
        for mode in (cm.SCATTER_DETERMINISTIC, cm.SCATTER_ATOMIC, cm.SCATTER_EVECTOR):
            v, v2, qd, fused = apply_once(bp, p, nel, mode)
            scale = np.abs(ref_v).max()
            e1 = np.abs(v - ref_v).max() / scale
            e2 = np.abs(v2 - 2 * ref_v).max() / scale
            eq = np.abs(qd - ref_qd).max() / np.abs(ref_qd).max()
            good = e1 < 1e-12 and e2 < 1e-12 and eq < 1e-12 and fused and not f0
            ok = ok and good
            print(f"BP{bp} p={p} nel={nel} mode={mode} fused={fused}/{f0}: |v| {scale:.3e} err {e1:.2e} add-err {e2:.2e} qdata-err {eq:.2e} "
                  f"{'OK' if good else 'FAIL'}", flush=True)
        if bp in (2, 4, 6):
            v, v2, qd, fused = apply_once(bp, p, nel, cm.SCATTER_DETERMINISTIC, interlaced=True)
            n = ref_v.size // 3
            vi = v.reshape(n, 3).T.reshape(-1)
            e1 = np.abs(vi - ref_v).max() / np.abs(ref_v).max()
            print(f"BP{bp} p={p} interlaced err {e1:.2e} {'OK' if e1 < 1e-12 else 'FAIL'}", flush=True)
            ok = ok and e1 < 1e-12
    return ok


def bench(bp, p, ndofs, mode=cm.SCATTER_DETERMINISTIC, reps=10):
    ceed = Ceed()
    ceed.set_scatter_mode(mode)
    ncomp = 3 if bp in (2, 4, 6) else 1
    nel = choose_elements(ndofs, p, ncomp)
    t0 = time.time()
    prob = BPProblem(ceed, bp, p, nel)
    prob.u.set_array(seeded_uniform(prob.num_dofs))
    prob.op.set_timing(True)
    for _ in range(3):
        prob.op.apply(prob.u, prob.v)
    setup_s = time.time() - t0
    ms = []
    for _ in range(reps):
        prob.op.apply(prob.u, prob.v)
        ms.append(prob.op.last_kernel_ms())
    fused = np.median([m[0] for m in ms])
    aux = np.median([m[1] for m in ms])
    tot = fused + aux
    info = prob.op.kernel_info()
    gb = prob.bytes_per_apply() / 1e9
    print(f"BP{bp} p={p} dofs={prob.num_dofs / 1e6:.2f}M nel={prob.num_elem} mode={mode} fused {fused:.3f} ms + aux {aux:.3f} ms -> "
          f"{prob.num_dofs / tot / 1e6:.2f} GDoF/s, {gb / tot * 1e3:.0f} GB/s algorithmic ({gb / tot * 1e3 / 6550.1 * 100:.1f}% of 6550) "
          f"regs={info['regs']} epb={info['elems_per_block']} thr={info['threads']} grid={info['grid']} smem={info['smem_bytes']} "
          f"local={info['local_bytes']} setup {setup_s:.1f}s", flush=True)


if __name__ == "__main__":
    ok = check()
    print("SELF-CONSISTENCY", "PASS" if ok else "FAIL", flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        sys.exit(0 if ok else 1)
    n = 10_000_000
    for mode in (cm.SCATTER_DETERMINISTIC, cm.SCATTER_ATOMIC, cm.SCATTER_EVECTOR):
        bench(3, 6, n, mode)
    bench(1, 3, n)
    for p in range(1, 9):
        bench(3, p, n)
    for p in range(4, 8):
        bench(5, p, n)
    for p in (4, 6):
        bench(6, p, n)
    sys.exit(0 if ok else 1)
