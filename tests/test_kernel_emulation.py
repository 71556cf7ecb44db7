"""CPU emulation of the GENERATED fused-operator kernels against the oracle (no GPU): tests/kernel_emu.py compiles the very source NVRTC
would compile with g++ (one OS thread per CUDA thread) and runs it on the argument block ceedb200_operator_debug_launch describes.
What this covers that the compile-only tests cannot: lane -> task maps, shared-memory plane addressing, the owner / halo scatter tables
as the kernel uses them, tail batches, element ranges of partitioned meshes, Apply vs ApplyAdd variants -- for every kernel shape
without bulk copies.  (Performance and the asynchronous-copy pipelines need the GPU: tests/test_gpu_parity.py.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = r"""
import json, os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
from libceed_b200 import Ceed, ceed as cm, mesh as M
from libceed_b200.bp import BPProblem, seeded_uniform
from oracle import oracle as O
import kernel_emu as KE
out = {}
SHARD, NSHARD = int(os.environ.get("EMU_SHARD", "0")), int(os.environ.get("EMU_NSHARD", "1"))
def mine(i): return i %% NSHARD == SHARD   # (the larger tests split their cases over a few concurrent processes)
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
def problem(ceed, bp, p, nel, **kw):
    prob = BPProblem(ceed, bp, p, nel, build_qdata=False, **kw)
    qd = O.bp_qdata(bp, p, prob.offsets, prob.coords)
    prob.qdata.set_array(qd)
    u = seeded_uniform(prob.num_dofs, 31)
    prob.u.set_array(u)
    return prob, qd, u, O.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u, interlaced=kw.get("interlaced", False))
""" % (ROOT, ROOT)


def run(body, timeout=1500, shards=1, table=False):
    env = dict(os.environ, CEED_B200_COMPILE_ONLY="1", CEED_B200_NO_TUNE_TABLE="1", EMU_NSHARD=str(shards))
    if table:
        env.pop("CEED_B200_NO_TUNE_TABLE")
    procs = [subprocess.Popen([sys.executable, "-c", PRELUDE + body + '\nprint("RESULT" + json.dumps(out))\n'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                              env=dict(env, EMU_SHARD=str(i))) for i in range(shards)]
    res = {}
    for pr in procs:
        so, se = pr.communicate(timeout=timeout)
        assert pr.returncode == 0, se[-4000:]
        res.update(json.loads([ln for ln in so.splitlines() if ln.startswith("RESULT")][0][6:]))
    return res


LEAN = r"""
for k, (bp, p, nel, morton) in enumerate(((1, 3, (5, 3, 3), False), (1, 2, (4, 3, 3), True), (2, 3, (3, 2, 3), False))):
    if not mine(k): continue
    for mode in (0, 1):
        ceed = Ceed(); ceed.set_scatter_mode(mode)
        prob, qd, u, ref = problem(ceed, bp, p, nel, elem_perm=M.morton_permutation(*nel) if morton else None)
        v_det = None
        # stage bits of the lean kernel: 1 element-interleaved columns, 2 element stride = P (mod 16), 4 quadrature data through 16-byte loads
        for E, warps, stage in ((6, 4, 0), (6, 4, 1), (8, 4, 3), (5, 2, 7), (8, 2, 4), (7, 1, 5)):
            if mode == 1 and stage not in (1, 7): continue   # (atomic scatter: two shapes)
            prob.op.set_kernel_shape(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
            KE.emulated_apply(prob.op, prob.u, prob.v)
            got = prob.op.get_kernel_shape()
            v = prob.v.get_array_read().copy()
            key = "bp%d p%d mode%d E%d w%d s%d" % (bp, p, mode, E, warps, stage)
            v_det = v if (mode == 0 and v_det is None) else v_det
            add_err = 0.0
            if stage in (3, 7):   # the accumulating kernel variant (ApplyAdd)
                w0 = seeded_uniform(prob.num_dofs, 5)
                prob.v.set_array(w0)
                KE.emulated_apply(prob.op, prob.u, prob.v, add=True)
                add_err = float(np.abs(prob.v.get_array_read() - w0 - ref).max() / np.abs(ref).max())
            out[key] = dict(layout=got["qf_mode"], stage=got["stage_mask"], err=rel(v, ref), bitwise=bool(mode == 1 or np.array_equal(v, v_det)), add_err=add_err)
"""


def test_lean_kernel_variants_emulated_against_the_oracle():
    """Every lean-kernel shape incl. the element-interleaved column map (stage bit 1), the padded element stride (2) and the 16-byte
    quadrature-data loads (4): oracle parity, bitwise equality of the deterministic variants, ApplyAdd, tails, Morton order, atomic scatter."""
    res = run(LEAN, shards=3)
    assert len(res) == 3 * (6 + 2)
    for key, v in res.items():
        assert v["layout"] == 4 and v["stage"] == int(key.rsplit("s", 1)[1]), (key, v)
        assert v["err"] < 1e-12 and v["add_err"] < 1e-12 and v["bitwise"], (key, v)


MISALIGNED = r"""
ceed = Ceed()
for bp, p, nel in ((1, 3, (5, 3, 3)), (2, 2, (4, 3, 3))):
    prob, qd, u, ref = problem(ceed, bp, p, nel)
    prob.op.set_kernel_shape(qf_mode=4, elems_per_group=4, cta_warps=2, group_warps=1, stage_mask=0)
    KE.emulated_apply(prob.op, prob.u, prob.v)
    v0 = prob.v.get_array_read().copy()
    raw = np.full(qd.size + 4, np.nan)
    base = (-(raw.ctypes.data // 8)) % 2   # index of the first 16-byte aligned double
    for shift in (0, 1):
        view = raw[base + shift: base + shift + qd.size]
        view[:] = qd
        assert view.ctypes.data % 16 == 8 * shift
        prob.qdata.set_array(view.ctypes.data, cm.MEM_DEVICE, cm.USE_POINTER)   # (compile-only mode: "device" memory is host memory)
        for E, warps, stage in ((4, 2, 4), (3, 1, 7)):
            prob.op.set_kernel_shape(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
            KE.emulated_apply(prob.op, prob.u, prob.v)
            out["bp%d p%d shift%d E%d s%d" % (bp, p, shift, E, stage)] = dict(err=rel(prob.v.get_array_read(), ref), bitwise=bool(np.array_equal(prob.v.get_array_read(), v0)))
    prob.qdata.take_array(cm.MEM_DEVICE)
"""


def test_lean_vector_qdata_loads_with_misaligned_arrays_emulated():
    """16-byte window loads (stage bit 4) with a user array 8 bytes off a 16-byte boundary: the window of the first line starts before the
    array, that of the last line may end after it, an even Q loses its alignment altogether -- all through the scalar path, same bits."""
    res = run(MISALIGNED)
    assert len(res) == 8
    for key, v in res.items():
        assert v["err"] < 1e-12 and v["bitwise"], (key, v)


GENERAL = r"""
ceed = Ceed()
shapes = (dict(stage_mask=1), dict(stage_mask=257), dict(stage_mask=513), dict(stage_mask=9, qf_mode=1, qf_unroll=2), dict(stage_mask=0, qf_mode=2, qf_unroll=2),
          dict(stage_mask=7, group_warps=2, cta_warps=4, elems_per_group=3), dict(qf_mode=3, stage_mask=17, group_warps=4, cta_warps=4))
for k, (bp, p, nel, kw) in enumerate(((3, 2, (4, 3, 2), {}), (6, 2, (3, 2, 2), {}), (4, 1, (3, 3, 2), dict(interlaced=True)), (3, 6, (2, 1, 1), {}), (1, 5, (2, 2, 1), {}))):
    if not mine(k): continue
    prob, qd, u, ref = problem(ceed, bp, p, nel, **kw)
    v0 = None
    for shape in shapes:
        prob.op.set_kernel_shape(**shape)
        KE.emulated_apply(prob.op, prob.u, prob.v)
        v = prob.v.get_array_read().copy()
        v0 = v if v0 is None else v0
        out["bp%d p%d %s" % (bp, p, sorted(shape.items()))] = dict(err=rel(v, ref), same=rel(v, v0))
"""


def test_general_kernel_shapes_emulated_against_the_oracle():
    """The general kernel (z-line / pointwise / point-pair / x-line QFunction stage; padded, swizzled and even-Q linear planes; one- to
    four-warp element groups with named barriers; cp.async staging of targets, offsets, gathered inputs and quadrature data) on BP3-BP6."""
    res = run(GENERAL, shards=3)
    assert len(res) == 35
    for key, v in res.items():
        assert v["err"] < 1e-12 and v["same"] < 1e-13, (key, v)


PARTS = r"""
from libceed_b200.mesh import Partition
ceed = Ceed()
for bp, p, nel, shape in ((1, 3, (6, 5, 4), dict(qf_mode=4, elems_per_group=6, cta_warps=2, group_warps=1, stage_mask=3)), (3, 2, (5, 4, 4), dict(stage_mask=1))):
    prob, qd, u, ref = problem(ceed, bp, p, nel)
    prob.op.set_kernel_shape(**shape)
    KE.emulated_apply(prob.op, prob.u, prob.v)
    whole = prob.v.get_array_read().copy()
    # the same mesh with a boundary-first element order and a split: part 1 then part 2 = the whole apply on that order
    ne = prob.num_elem
    perm = np.roll(np.arange(ne), 7)
    split = 2 * ne // 5
    prob2, _, _, ref2 = problem(Ceed(), bp, p, nel, elem_perm=perm, split=split)
    prob2.op.set_kernel_shape(**shape)
    d1 = KE.emulated_apply(prob2.op, prob2.u, prob2.v, part=1)
    d2 = KE.emulated_apply(prob2.op, prob2.u, prob2.v, part=2)
    out["bp%d p%d" % (bp, p)] = dict(err=rel(whole, ref), part_err=rel(prob2.v.get_array_read(), ref2), ranges=[d1.e_begin, d1.e_end, d2.e_begin, d2.e_end], ne=ne, split=split)
"""


def test_element_parts_emulated():
    """apply_part (boundary / interior element ranges of a partitioned mesh, also the chunks of the streamed apply): part 1 + part 2 of
    the lean and the general kernel reproduce the operator."""
    res = run(PARTS)
    for key, v in res.items():
        assert v["err"] < 1e-12 and v["part_err"] < 1e-12, (key, v)
        assert v["ranges"] == [0, v["split"], v["split"], v["ne"]], (key, v)


GUARD = r"""
import ctypes as C
ceed = Ceed(); lib = ceed._lib
prob, qd, u, ref = problem(ceed, 1, 3, (5, 3, 3))
off = np.ascontiguousarray(prob.offsets, dtype=np.int32).reshape(-1)
raw = np.zeros(off.size + 8, dtype=np.int32)
first = (-(raw.ctypes.data // 4)) % 4
for shift in (0, 1):   # a caller-owned "device" offset array (compile-only mode: host memory), 16-byte aligned and 4 bytes off
    view = raw[first + shift: first + shift + off.size]; view[:] = off
    assert view.ctypes.data % 16 == 4 * shift
    ptr = C.c_void_p()
    ceed._chk(lib.ceedb200_restriction_create(ceed._ptr, prob.num_elem, 64, 1, prob.num_nodes, prob.num_nodes, cm.MEM_DEVICE, cm.USE_POINTER, C.c_void_p(view.ctypes.data), C.byref(ptr)))
    rs = cm.ElemRestriction.__new__(cm.ElemRestriction); cm._Object.__init__(rs, ceed, ptr)
    op = ceed.Operator(prob.qf)
    op.set_field("u", rs, prob.basis_u, cm.VECTOR_ACTIVE)
    op.set_field("qdata", prob.rstr_qd, cm.BASIS_NONE, prob.qdata)
    op.set_field("v", rs, prob.basis_u, cm.VECTOR_ACTIVE)
    op.set_kernel_shape(qf_mode=4, elems_per_group=6, cta_warps=2, group_warps=1, stage_mask=7)
    KE.emulated_apply(op, prob.u, prob.v)
    out["shift%d" % shift] = dict(stage=op.get_kernel_shape()["stage_mask"], err=rel(prob.v.get_array_read(), ref))
    del op, rs
"""


def test_misaligned_offset_table_drops_the_vector_table_loads():
    """Stage bit 1 stages the offset table with 16-byte loads: a caller-owned offset array that is not 16-byte aligned makes the apply
    regenerate the kernel without that bit (same mechanism as the bulk copies of misaligned quadrature data) -- never a misaligned access."""
    res = run(GUARD)
    assert res["shift0"]["stage"] == 7 and res["shift1"]["stage"] == 6, res
    assert res["shift0"]["err"] < 1e-12 and res["shift1"]["err"] < 1e-12, res


BULK = r"""
ceed = Ceed()
lean = lambda E, w, st: dict(qf_mode=4, elems_per_group=E, cta_warps=w, group_warps=1, stage_mask=st)
for bp, p, nel, shapes in ((1, 3, (5, 3, 3), [lean(3, 1, 32), lean(6, 4, 40), lean(5, 2, 8), lean(4, 2, 39), lean(8, 2, 35), lean(6, 4, 12), lean(4, 4, 64)]),
                           (2, 2, (3, 3, 2), [lean(3, 2, 40)]),
                           (3, 2, (4, 3, 2), [dict(stage_mask=33), dict(stage_mask=33, group_warps=2, cta_warps=4, elems_per_group=3), dict(stage_mask=289, group_warps=2, cta_warps=2)]),
                           (3, 3, (3, 2, 2), [dict(stage_mask=33), dict(stage_mask=41, qf_mode=1, qf_unroll=2)])):
    prob, qd, u, ref = problem(ceed, bp, p, nel)
    for shape in shapes:
        prob.op.set_kernel_shape(**shape)
        KE.emulated_apply(prob.op, prob.u, prob.v)
        out["bp%d p%d %s" % (bp, p, sorted(shape.items()))] = dict(err=rel(prob.v.get_array_read(), ref), stage=prob.op.get_kernel_shape()["stage_mask"], want=shape["stage_mask"])
"""


def test_bulk_copy_pipelines_emulated():
    """cp.async.bulk + mbarrier variants (quadrature data of the general kernel; index tables / quadrature data / L2 prefetch of the lean
    kernel, alone and with the round-2c stage bits): the arm / wait / re-arm protocol and the phase parities of the generated code run against
    a host model of the mbarrier (tests/emu/b200-tma.h), odd Q^3 (8-byte source misalignment, aligned-down copies) included."""
    res = run(BULK)
    assert len(res) == 13
    for key, v in res.items():
        assert v["err"] < 1e-12 and v["stage"] == v["want"], (key, v)


TABLE = r"""
import re
ceed = Ceed()   # (the subprocess of this test runs WITHOUT CEED_B200_NO_TUNE_TABLE: the context carries the shipped table)
table = {}
for ln in open(os.path.join(%r, "libceed_b200", "tuned", "sm_100a.tune")):
    if ln.strip() and not ln.startswith("#"):
        tok = ln.split("#")[0].split()
        table[tok[0]] = [int(x) for x in tok[1:8]]
for bp in (1, 2, 3, 4, 5, 6):
    for p in range(1, 9):
        if (bp in (5, 6) and p == 1) or not mine(bp * 8 + p): continue
        nel = (3, 2, 2) if p <= 3 else ((2, 2, 1) if p <= 5 else (2, 1, 1))
        prob, qd, u, ref = problem(ceed, bp, p, nel)
        KE.emulated_apply(prob.op_setup, prob.x, prob.qdata)   # the setup operator's fused kernel (x INTERP + GRAD, weights -> strided qdata)
        qd_err = rel(prob.qdata.get_array_read(), qd)
        d = KE.emulated_apply(prob.op, prob.u, prob.v)         # ... whose output the apply operator then reads
        got = prob.op.get_kernel_shape()
        entry = table.get(got["signature"])
        out["bp%%d p%%d" %% (bp, p)] = dict(err=rel(prob.v.get_array_read(), ref), qd_err=qd_err, in_table=entry is not None, entry=entry,
                                         got=[got[k] for k in ("elems_per_group", "group_warps", "cta_warps", "min_blocks_per_sm", "qf_mode", "qf_unroll", "stage_mask")],
                                         num_elem=prob.num_elem)
""" % ROOT


def test_every_shipped_tuning_table_entry_generates_a_correct_kernel():
    """BP1-BP6, p = 1..8 with the shipped tuning table active, setup operator (quadrature data) and apply operator: the kernel each entry selects (layout, batch width, staging bits incl. the
    bulk-copy ones) is generated, emulated on the CPU and compared with the oracle; the resolved shape is the table's (the batch width is
    capped by the small mesh, the occupancy target by what shared memory allows)."""
    res = run(TABLE, shards=4, table=True)
    assert len(res) == 46
    hits = 0
    for key, v in res.items():
        assert v["err"] < 1e-12 and v["qd_err"] < 1e-12, (key, v)
        if v["in_table"]:
            hits += 1
            e, g = v["entry"], v["got"]
            assert g[0] == min(e[0], v["num_elem"]) and g[1] == e[1] and g[2] <= e[2] and g[4] == e[4] and g[5] == e[5], (key, v)
            assert g[6] == e[6] or g[4] == 4, (key, v)  # (the lean kernel masks the stage bits it does not read)
    assert hits >= 30, hits


MODES = r"""
for bp, p, nel, kw in ((1, 3, (4, 3, 2), {}), (3, 2, (3, 3, 2), {}), (6, 2, (3, 2, 2), dict(interlaced=True)), (5, 4, (2, 2, 1), {})):
    res = {}
    for mode in (0, 1, 2):   # deterministic owner / halo tables, atomics, E-vector + ordered transpose restriction
        ceed = Ceed(); ceed.set_scatter_mode(mode)
        prob, qd, u, ref = problem(ceed, bp, p, nel, **kw)
        KE.emulated_apply(prob.op, prob.u, prob.v)
        res[mode] = prob.v.get_array_read().copy()
        w0 = seeded_uniform(prob.num_dofs, 5)
        prob.v.set_array(w0)
        KE.emulated_apply(prob.op, prob.u, prob.v, add=True)
        out["bp%d p%d mode%d" % (bp, p, mode)] = dict(err=rel(res[mode], ref), add_err=float(np.abs(prob.v.get_array_read() - w0 - ref).max() / np.abs(ref).max()),
                                                      bitwise_vs_det=bool(np.array_equal(res[mode], res[0])))
"""


def test_scatter_modes_emulated_and_deterministic_equals_the_ordered_transpose():
    """The three scatter modes of the fused kernel that finish without inter-CTA flags.  The deterministic owner / halo scheme reproduces,
    bit for bit, the E-vector mode's ordered transpose (ascending E-index per node = the serial CPU order): the claim of DESIGN.md section 2,
    checked here on the generated code itself."""
    res = run(MODES)
    assert len(res) == 12
    for key, v in res.items():
        assert v["err"] < 1e-12 and v["add_err"] < 1e-12, (key, v)
        if key.endswith("mode2"):
            assert v["bitwise_vs_det"], (key, v)
