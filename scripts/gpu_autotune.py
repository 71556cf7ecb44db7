"""Run the core's autotuner over the BP kernel shapes on a B200 and write a tuning table.
usage: CEED_B200_TUNE_SAVE=gpurun_out/sm_100a.tune python scripts/gpu_autotune.py [dofs] bp3p6 bp1p3 ...
The table is then committed as libceed_b200/tuned/sm_100a.tune (loaded by ceedb200_init)."""
import os, re, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CEED_B200_TUNE_SAVE", "gpurun_out/sm_100a.tune")
from libceed_b200 import Ceed, ceed as cm
from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
from libceed_b200.mesh import choose_elements

args = sys.argv[1:]
dofs = 10e6
if args and not args[0].startswith("bp"):
    dofs = float(args.pop(0))


def timed(op, prob, n=10):
    op.set_timing(True)
    for _ in range(3): op.apply(prob.u, prob.v)
    t = []
    for _ in range(n):
        op.apply(prob.u, prob.v); t.append(sum(op.last_kernel_ms()))
    return float(np.median(t))


for wl in args:
    m = re.fullmatch(r"bp(\d)p(\d)", wl); bp, p = int(m.group(1)), int(m.group(2))
    ceed = Ceed()
    prob = BPProblem(ceed, bp, p, choose_elements(dofs, p, BP_TABLE[bp][0]))
    prob.u.set_array(seeded_uniform(prob.num_dofs))
    t0 = timed(prob.op, prob)
    v0 = prob.v.get_array_read().copy()
    shape0 = prob.op.get_kernel_shape()
    ceed.set_autotune(2)
    op = ceed.Operator(prob.qf)
    op.set_field("u", prob.rstr_u, prob.basis_u, cm.VECTOR_ACTIVE)
    op.set_field("qdata", prob.rstr_qd, cm.BASIS_NONE, prob.qdata)
    op.set_field("v", prob.rstr_u, prob.basis_u, cm.VECTOR_ACTIVE)
    w0 = time.time()
    op.apply(prob.u, prob.v)
    tune_s = time.time() - w0
    t1 = timed(op, prob)
    err = np.abs(prob.v.get_array_read() - v0).max() / np.abs(v0).max()
    s = op.get_kernel_shape()
    gb = prob.bytes_per_apply() / 1e6
    print(f"{wl}: {prob.num_dofs/1e6:.2f}M DoFs  before {t0:.4f} ms ({prob.num_dofs/t0/1e6:.2f} GDoF/s, {gb/t0/6550.1*100:.1f}%)  "
          f"after {t1:.4f} ms ({prob.num_dofs/t1/1e6:.2f} GDoF/s, {gb/t1/6550.1*100:.1f}%)  tune {tune_s:.1f}s  err {err:.1e}\n"
          f"    {s['signature']}  before {[shape0[k] for k in op.SHAPE_KEYS]}  after {[s[k] for k in op.SHAPE_KEYS]}  {op.kernel_info()}", flush=True)
    del op, prob, ceed
