"""Randomized CPU emulation of generated kernels (no GPU): random BP operator, order, small mesh, element order, scatter mode, kernel shape
(layout, batch width, group width incl. odd ones, staging bits incl. the bulk-copy ones), Apply / ApplyAdd -- every case against the oracle.
A development aid on top of tests/test_kernel_emulation.py.  usage: python scripts/emu_fuzz.py [seed] [cases]"""
import os, sys, random, time
os.environ["CEED_B200_COMPILE_ONLY"] = "1"; os.environ["CEED_B200_NO_TUNE_TABLE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from libceed_b200 import Ceed, mesh as M
from libceed_b200.bp import BPProblem, seeded_uniform
from oracle import oracle as O
import kernel_emu as KE
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rnd = random.Random(seed)
bad = 0
t0 = time.time()
for case in range(n_cases):
    bp = rnd.choice([1, 2, 3, 4, 5, 6])
    p = rnd.randint(2 if bp in (5, 6) else 1, 8)
    budget = 40 if p <= 3 else (10 if p <= 5 else 4)
    while True:
        nel = (rnd.randint(1, 5), rnd.randint(1, 4), rnd.randint(1, 3))
        if nel[0] * nel[1] * nel[2] <= budget: break
    mode = rnd.choice([0, 0, 0, 1, 2])  # (deterministic, atomic, E-vector; the in-kernel ORDERED completion is not emulated)
    interlaced = bp % 2 == 0 and rnd.random() < 0.4
    morton = rnd.random() < 0.3
    ceed = Ceed(); ceed.set_scatter_mode(mode)
    prob = BPProblem(ceed, bp, p, nel, build_qdata=False, interlaced=interlaced, elem_perm=M.morton_permutation(*nel) if morton else None)
    qd = O.bp_qdata(bp, p, prob.offsets, prob.coords); prob.qdata.set_array(qd)
    u = seeded_uniform(prob.num_dofs, 31 + case); prob.u.set_array(u)
    ref = O.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u, interlaced=interlaced)
    if bp in (1, 2) and rnd.random() < 0.6:
        stage = rnd.choice([0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 32, 33, 35, 36, 39, 40, 64, 65, 68])
        shape = dict(qf_mode=4, elems_per_group=rnd.randint(1, 12), cta_warps=rnd.choice([1, 2, 3, 4, 8]), group_warps=1, stage_mask=stage)
    else:
        stage = rnd.choice([0, 1, 9, 3, 5, 7, 17, 33, 41, 257, 265, 289, 513, 521, 129])
        shape = dict(qf_mode=rnd.choice([-1, 0, 1, 2, 3]), elems_per_group=rnd.choice([0, 1, 2, 3, 5]), group_warps=rnd.choice([0, 1, 2, 3, 4, 5]), cta_warps=rnd.choice([0, 1, 2, 4, 6, 8]),
                     qf_unroll=rnd.choice([0, 1, 2, 4]), stage_mask=stage)
    add = rnd.random() < 0.3
    tag = f"case {case}: bp{bp} p={p} nel={nel} mode={mode} il={interlaced} morton={morton} add={add} shape={shape}"
    try:
        prob.op.set_kernel_shape(**shape)
        if add:
            w0 = seeded_uniform(prob.num_dofs, 5); prob.v.set_array(w0)
        KE.emulated_apply(prob.op, prob.u, prob.v, add=add)
        v = prob.v.get_array_read()
        err = np.abs((v - w0 if add else v) - ref).max() / np.abs(ref).max()
        got = prob.op.get_kernel_shape()
        ok = err < 1e-12
        print(("ok  " if ok else "FAIL") + f" {err:.1e} {tag} -> resolved qf {got['qf_mode']} E {got['elems_per_group']} gw {got['group_warps']} w {got['cta_warps']} st {got['stage_mask']}", flush=True)
        bad += not ok
    except NotImplementedError as e:
        print("skip (not emulated) " + tag, flush=True)
    except Exception as e:
        msg = str(e)
        if "not fused" in msg or "evector" in msg.lower():
            print("skip (" + msg[:80] + ") " + tag, flush=True)
        else:
            bad += 1
            print("EXC  " + msg[-400:].replace("\n", " | ") + " " + tag, flush=True)
    del prob, ceed
print(f"seed {seed}: {bad} bad of {n_cases} ({time.time()-t0:.0f} s)")
