"""Tuning sweep on the GPU: kernel time of one workload under different launch shapes / generator options.
usage: python scripts/gpu_tune.py bp3p6 [dofs]"""
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libceed_b200 import Ceed  # noqa: E402
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform  # noqa: E402
from libceed_b200.mesh import choose_elements  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "bp3p6"
dofs = float(sys.argv[2]) if len(sys.argv) > 2 else 10e6
m = re.fullmatch(r"bp(\d)p(\d)", wl)
bp, p = int(m.group(1)), int(m.group(2))
nel = choose_elements(dofs, p, BP_TABLE[bp][0])
Q = p + BP_TABLE[bp][2]

ceed = Ceed()
base = BPProblem(ceed, bp, p, nel)  # builds qdata once
u = seeded_uniform(base.num_dofs)
base.u.set_array(u)


def run(tag, env, epb=0, bpsm=0, scatter=0):
    for k in list(os.environ):
        if k.startswith("CEED_B200_") and k not in ("CEED_B200_JIT_DIR",):
            del os.environ[k]
    os.environ.update(env)
    ceed.set_scatter_mode(scatter)
    from libceed_b200 import ceed as cm
    op = ceed.Operator(base.qf)
    op.set_field("u", base.rstr_u, base.basis_u, cm.VECTOR_ACTIVE)
    op.set_field("qdata", base.rstr_qd, cm.BASIS_NONE, base.qdata)
    op.set_field("v", base.rstr_u, base.basis_u, cm.VECTOR_ACTIVE)
    op.set_tuning(epb, bpsm)
    op.set_timing(True)
    try:
        for _ in range(3):
            op.apply(base.u, base.v)
        t = []
        for _ in range(10):
            op.apply(base.u, base.v)
            t.append(op.last_kernel_ms())
    except Exception as e:
        print(f"{tag:40s} FAILED {str(e)[:100]}", flush=True)
        return
    f, a = np.median([x[0] for x in t]), np.median([x[1] for x in t])
    info = op.kernel_info()
    gb = base.bytes_per_apply() / 1e9
    print(f"{tag:40s} fused {f:.3f} aux {a:.3f} ms  {base.num_dofs / (f + a) / 1e6:6.2f} GDoF/s  {gb / (f + a) * 1e3 / 6550.1 * 100:5.1f}%  "
          f"regs={info['regs']} epb={info['elems_per_block']} thr={info['threads']} grid={info['grid']} smem={info['smem_bytes']} local={info['local_bytes']}",
          flush=True)


print(f"{wl}: {base.num_dofs / 1e6:.2f}M DoFs, {base.num_elem} elements, Q={Q}")
run("block mode (async staging)", {"CEED_B200_BLOCK_MODE": "1"})
run("block mode, no async", {"CEED_B200_BLOCK_MODE": "1", "CEED_B200_NO_ASYNC": "1"})
run("warp default", {})
run("warp atomic", {}, scatter=1)
for stage in (0, 1, 3, 5, 7):
    for minb in (3, 4, 5):
        run(f"warp stage={stage} minb={minb}", {"CEED_B200_STAGE": str(stage), "CEED_B200_MINB": str(minb)})
for warps, minb in ((2, 8), (2, 6), (8, 2), (8, 1), (1, 16), (1, 12)):
    run(f"warp W={warps} minb={minb} stage=0", {"CEED_B200_STAGE": "0", "CEED_B200_WARPS": str(warps), "CEED_B200_MINB": str(minb)})
for epb in (1, 2, 3, 4):
    run(f"warp epb={epb} stage=0", {"CEED_B200_STAGE": "0"}, epb=epb)
