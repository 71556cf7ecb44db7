"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against
  (1) the oracle (CPU restatement of /cpu/self/ref/serial) on the same seeded inputs,
  (2) golden vectors produced by the unmodified reference (tests/golden/reference_golden.npz),
  (3) the unmodified reference library itself when oracle/_ref/ travelled to this box,
  (4) size-independent properties at BASELINE.json's full sizes (10M DoFs).
Tolerances (BASELINE.json north_star): operator output 1e-12 relative in FP64; restriction bit-exact."""
import numpy as np
import pytest

from conftest import BP_CASES, bp_case_key
from libceed_b200 import mesh as M
from libceed_b200.bp import BP_TABLE, seeded_uniform

pytestmark = pytest.mark.gpu

OP_TOL = 1e-12


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def cm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from libceed_b200 import ceed as cm
    return cm


def make_problem(cm, bp, p, nel, mode=0, interlaced=False, **kw):
    from libceed_b200.bp import BPProblem
    ceed = cm.Ceed("/gpu/cuda/b200")
    ceed.set_scatter_mode(mode)
    return BPProblem(ceed, bp, p, nel, interlaced=interlaced, **kw)


# ------------------------------------------------------------------------------------------------ vector
def test_vector_state_machine_and_blas1(cm):
    ceed = cm.Ceed()
    n = 1000
    x, y, w = ceed.Vector(n), ceed.Vector(n), ceed.Vector(n)
    a = seeded_uniform(n, 1)
    b = seeded_uniform(n, 2)
    x.set_array(a)
    y.set_array(b)
    assert np.array_equal(x.get_array_read(), a)  # H -> D -> H round trip is exact
    y.axpy(2.5, x)
    assert np.allclose(y.get_array_read(), b + 2.5 * a, rtol=0, atol=1e-15)
    y.axpby(0.5, -1.0, x)
    assert np.allclose(y.get_array_read(), 0.5 * a - (b + 2.5 * a), rtol=0, atol=1e-14)
    w.pointwise_mult(x, x)
    assert np.array_equal(w.get_array_read(), a * a)
    x.scale(-3.0)
    assert np.array_equal(x.get_array_read(), -3.0 * a)
    assert abs(x.norm(cm.NORM_1) - np.abs(3 * a).sum()) < 1e-10
    assert abs(x.norm(cm.NORM_2) - np.linalg.norm(3 * a)) < 1e-11
    assert x.norm(cm.NORM_MAX) == np.abs(3 * a).max()
    x.set_value(7.0)
    assert np.all(x.get_array_read() == 7.0)
    x.reciprocal()
    assert np.allclose(x.get_array_read(), 1.0 / 7.0)
    e = ceed.Vector(0)
    e.set_value(1.0)
    assert e.get_array_read().size == 0 and e.norm() == 0.0
    with pytest.raises(cm.CeedError):
        ceed.Vector(5).get_array_read()  # no valid data


# ------------------------------------------------------------------------------------------------ restriction
def test_restriction_bit_exact_vs_golden_and_oracle(cm, golden, oracle):
    ceed = cm.Ceed()
    nelem, esize, ncomp, lsize = [int(v) for v in golden["rstr_dims"]]
    off = golden["rstr_offsets"]
    r = ceed.ElemRestriction(nelem, esize, ncomp, lsize, ncomp * lsize, off)
    u, ev = ceed.Vector(ncomp * lsize), ceed.Vector(nelem * esize * ncomp)
    u.set_array(golden["rstr_u"])
    r.apply(u, ev)
    assert r.get_e_layout() == (1, nelem * esize, esize)  # GPU-family layout [comp][elem][node]
    e_gpu = ev.get_array_read().reshape(ncomp, nelem, esize)
    e_ref = golden["rstr_e"].reshape(nelem, ncomp, esize)  # CPU reference layout [elem][comp][node]
    assert np.array_equal(e_gpu.transpose(1, 0, 2), e_ref)
    # transpose: deterministic, ascending (elem, node) order == serial reference order => bit-exact
    w_ref = golden["rstr_w"].reshape(nelem, ncomp, esize)
    w, lt = ceed.Vector(nelem * esize * ncomp), ceed.Vector(ncomp * lsize)
    w.set_array(np.ascontiguousarray(w_ref.transpose(1, 0, 2)).reshape(-1))
    lt.set_value(0.0)
    r.T_apply(w, lt)
    assert np.array_equal(lt.get_array_read(), golden["rstr_lt"])
    assert np.array_equal(lt.get_array_read(), oracle.restriction_offset(nelem, esize, ncomp, lsize, off, 1, golden["rstr_w"], ncomp * lsize))


def test_restriction_ragged_and_edge_cases(cm, oracle):
    ceed = cm.Ceed()
    rng = np.random.default_rng(3)
    for nelem, esize, ncomp, lsize in [(1, 1, 1, 1), (7, 27, 3, 50), (300, 8, 1, 999), (2, 64, 2, 64)]:
        off = rng.integers(0, lsize, nelem * esize).astype(np.int32)
        r = ceed.ElemRestriction(nelem, esize, ncomp, lsize, ncomp * lsize, off)
        u = seeded_uniform(ncomp * lsize, 4)
        uv, ev = ceed.Vector(u.size), ceed.Vector(nelem * esize * ncomp)
        uv.set_array(u)
        r.apply(uv, ev)
        e_ref = oracle.restriction_offset(nelem, esize, ncomp, lsize, off, 0, u, nelem * esize * ncomp).reshape(nelem, ncomp, esize)
        assert np.array_equal(ev.get_array_read().reshape(ncomp, nelem, esize).transpose(1, 0, 2), e_ref)
        wv = seeded_uniform(nelem * esize * ncomp, 5)
        w, lt = ceed.Vector(wv.size), ceed.Vector(ncomp * lsize)
        w.set_array(wv)
        lt.set_array(u)  # transpose ADDS into existing data
        r.T_apply(w, lt)
        w_cpu = np.ascontiguousarray(wv.reshape(ncomp, nelem, esize).transpose(1, 0, 2)).reshape(-1)
        expect = u.copy()
        import ctypes as C
        oracle.lib().oracle_restriction_offset(nelem, esize, ncomp, lsize, off.ctypes.data_as(oracle.ip), 1, w_cpu.ctypes.data_as(oracle.dp),
                                               expect.ctypes.data_as(oracle.dp))
        assert np.array_equal(lt.get_array_read(), expect)
    with pytest.raises(cm.CeedError):
        ceed.ElemRestriction(1, 2, 1, 1, 2, np.array([0, 5], np.int32))  # out-of-range offset
    # strided
    r = ceed.StridedElemRestriction(3, 4, 2, 24, (2, 1, 8))
    u = ceed.Vector(24)
    u.set_array(np.arange(24.0))
    ev = ceed.Vector(24)
    r.apply(u, ev)
    ref = oracle.restriction_strided(3, 4, 2, (2, 1, 8), 0, np.arange(24.0), 24).reshape(3, 2, 4)
    assert np.array_equal(ev.get_array_read().reshape(2, 3, 4).transpose(1, 0, 2), ref)


# ------------------------------------------------------------------------------------------------ basis
@pytest.mark.parametrize("dim,P,Q,qm", [(1, 3, 4, 0), (2, 4, 5, 0), (3, 3, 5, 0), (3, 5, 5, 1), (3, 4, 3, 0), (2, 5, 6, 1)])
def test_standalone_basis_matches_oracle(cm, oracle, dim, P, Q, qm):
    import ctypes as C
    ceed = cm.Ceed()
    nc, nelem = 2, 3
    b = ceed.BasisTensorH1Lagrange(dim, nc, P, Q, qm)
    interp, grad, qref, qw = oracle.lagrange_1d(P, Q, qm)
    assert np.abs(b.interp_1d.reshape(-1) - interp).max() < 1e-14
    nn, nq = P ** dim, Q ** dim
    u = seeded_uniform(nc * nelem * nn, 8)
    uv = ceed.Vector(u.size)
    uv.set_array(u)
    O = oracle.lib()

    class OB(C.Structure):
        _fields_ = [("dim", C.c_int), ("nc", C.c_int), ("P", C.c_int), ("Q", C.c_int), ("interp", oracle.dp), ("grad", oracle.dp),
                    ("qw", oracle.dp), ("collo", oracle.dp), ("is_collocated", C.c_int)]

    collocated = P == Q and np.abs(interp.reshape(Q, P) - np.eye(P)).max() < 1e-15
    collo = oracle.collocated_grad(P, Q, interp, grad) if (Q >= P and not collocated) else None
    ob = OB(dim, nc, P, Q, interp.ctypes.data_as(oracle.dp), grad.ctypes.data_as(oracle.dp), qw.ctypes.data_as(oracle.dp),
            collo.ctypes.data_as(oracle.dp) if collo is not None else None, int(collocated))
    O.oracle_basis_apply.argtypes = [C.POINTER(OB), C.c_int, C.c_int, oracle.dp, oracle.dp]
    O.oracle_basis_apply.restype = None
    for emode, qcomp in ((cm.EVAL_INTERP, 1), (cm.EVAL_GRAD, dim)):
        vv = ceed.Vector(qcomp * nc * nelem * nq)
        b.apply(nelem, emode, uv, vv)
        got = vv.get_array_read().reshape(qcomp, nc, nelem, nq)
        back = ceed.Vector(u.size)
        b.apply(nelem, emode, vv, back, tmode=cm.TRANSPOSE)
        got_t = back.get_array_read().reshape(nc, nelem, nn)
        for e in range(nelem):
            ue = np.ascontiguousarray(u.reshape(nc, nelem, nn)[:, e, :]).reshape(-1)
            ve = np.zeros(qcomp * nc * nq)
            O.oracle_basis_apply(C.byref(ob), 0, emode, ue.ctypes.data_as(oracle.dp), ve.ctypes.data_as(oracle.dp))
            assert rel(got[:, :, e, :].reshape(-1), ve) < 1e-13
            te = np.zeros(nc * nn)
            O.oracle_basis_apply(C.byref(ob), 1, emode, ve.ctypes.data_as(oracle.dp), te.ctypes.data_as(oracle.dp))
            assert rel(got_t[:, e, :].reshape(-1), te) < 1e-13
    wv = ceed.Vector(nelem * nq)
    b.apply(nelem, cm.EVAL_WEIGHT, None, wv)
    w1 = qw
    full = w1
    for _ in range(dim - 1):
        full = np.multiply.outer(w1, full).reshape(-1)
    assert np.allclose(wv.get_array_read().reshape(nelem, nq), full[None, :], rtol=1e-14, atol=0)  # product order differs by an ulp


# ------------------------------------------------------------------------------------------------ qfunction
def test_standalone_qfunction_matches_oracle(cm, oracle):
    import ctypes as C
    from libceed_b200.bp import APPLY_H
    ceed = cm.Ceed()
    Q = 777
    qf = ceed.QFunction(APPLY_H, "BPDiff")
    qf.add_input("u", 3, cm.EVAL_GRAD)
    qf.add_input("qdata", 7, cm.EVAL_NONE)
    qf.add_output("v", 3, cm.EVAL_GRAD)
    ug, qd = seeded_uniform(3 * Q, 1), seeded_uniform(7 * Q, 2)
    U, D, V = ceed.Vector(3 * Q), ceed.Vector(7 * Q), ceed.Vector(3 * Q)
    U.set_array(ug)
    D.set_array(qd)
    qf.apply(Q, [U, D], [V])
    out = np.zeros(3 * Q)
    ins = (oracle.dp * 2)(ug.ctypes.data_as(oracle.dp), qd.ctypes.data_as(oracle.dp))
    outs = (oracle.dp * 1)(out.ctypes.data_as(oracle.dp))
    oracle.lib().oracle_qf_bp_diff(None, Q, ins, outs)
    assert rel(V.get_array_read(), out) < 1e-15


# ------------------------------------------------------------------------------------------------ operator
@pytest.mark.parametrize("case", BP_CASES, ids=lambda c: bp_case_key(*c))
@pytest.mark.parametrize("mode", [0, 1, 2, 3], ids=["deterministic", "atomic", "evector", "ordered"])
def test_bp_operator_vs_oracle_and_golden(cm, oracle, golden, case, mode):
    bp, p, nel, gallery, interlaced = case
    if gallery:
        pytest.skip("gallery QFunction sources live in the reference tree; covered by test_gallery_qfunctions_from_reference_install")
    prob = make_problem(cm, bp, p, nel, mode, interlaced)
    assert prob.op.is_fused and prob.op_setup.is_fused
    off, coords, nn = prob.offsets, prob.coords, prob.num_nodes
    key = bp_case_key(*case)
    # setup operator output (qdata), GPU layout [comp][elem][qpt] == fixture layout
    qd = prob.qdata.get_array_read()
    qd_or = oracle.bp_qdata(bp, p, off, coords)
    assert rel(qd, qd_or) < OP_TOL and rel(qd, golden[key + "_qdata"]) < OP_TOL
    u = seeded_uniform(prob.num_dofs)
    prob.u.set_array(u)
    prob.op.apply(prob.u, prob.v)
    v = prob.v.get_array_read()
    v_or = oracle.bp_apply(bp, p, off, nn, qd_or, u, interlaced=interlaced)
    assert rel(v, v_or) < OP_TOL
    assert rel(v, golden[key + "_v"]) < OP_TOL
    # ApplyAdd accumulates on top of existing data
    prob.op.apply_add(prob.u, prob.v)
    assert rel(prob.v.get_array_read(), 2 * v_or) < OP_TOL
    # Apply overwrites stale data
    prob.v.set_value(1e30)
    prob.op.apply(prob.u, prob.v)
    assert rel(prob.v.get_array_read(), v_or) < OP_TOL


def test_gallery_qfunctions_from_reference_install(cm, oracle, golden, refceed):
    """Gallery QFunction headers (include/ceed/jit-source/gallery) JIT-compile unchanged into the fused kernel."""
    import os
    from oracle import refceed as R
    inc = os.path.join(R.REF_DIR, "include")
    for bp, p, nel in [(1, 2, (2, 2, 2)), (3, 2, (2, 2, 2))]:
        ceed = cm.Ceed()
        ceed.add_jit_source_root(inc)
        P, Q = p + 1, p + 2
        off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
        nn, ne = coords.shape[1], off.shape[0]
        ncq = 6 if bp == 3 else 1
        rx = ceed.ElemRestriction(ne, P ** 3, 3, nn, 3 * nn, off)
        ru = ceed.ElemRestriction(ne, P ** 3, 1, nn, nn, off)
        rq = ceed.StridedElemRestriction(ne, Q ** 3, ncq, ne * Q ** 3 * ncq)
        bx, bu = ceed.BasisTensorH1Lagrange(3, 3, P, Q, cm.GAUSS), ceed.BasisTensorH1Lagrange(3, 1, P, Q, cm.GAUSS)
        gal = os.path.join(inc, "ceed", "jit-source", "gallery")
        qs = ceed.QFunction(os.path.join(gal, "ceed-poisson3dbuild.h" if bp == 3 else "ceed-mass3dbuild.h"), "Poisson3DBuild" if bp == 3 else "Mass3DBuild")
        qs.add_input("dx", 9, cm.EVAL_GRAD)
        qs.add_input("weights", 1, cm.EVAL_WEIGHT)
        qs.add_output("qdata", ncq, cm.EVAL_NONE)
        x, qd = ceed.Vector(3 * nn), ceed.Vector(ne * Q ** 3 * ncq)
        x.set_array(coords.reshape(-1))
        ops = ceed.Operator(qs)
        ops.set_field("dx", rx, bx, cm.VECTOR_ACTIVE)
        ops.set_field("weights", None, bx, cm.VECTOR_NONE)
        ops.set_field("qdata", rq, None, cm.VECTOR_ACTIVE)
        ops.apply(x, qd)
        key = bp_case_key(bp, p, nel, True, False)
        assert rel(qd.get_array_read(), golden[key + "_qdata"]) < OP_TOL
        qa = ceed.QFunction(os.path.join(gal, "ceed-poisson3dapply.h" if bp == 3 else "ceed-massapply.h"), "Poisson3DApply" if bp == 3 else "MassApply")
        names = ("du", "qdata", "dv") if bp == 3 else ("u", "qdata", "v")
        em, sz = (cm.EVAL_GRAD, 3) if bp == 3 else (cm.EVAL_INTERP, 1)
        qa.add_input(names[0], sz, em)
        qa.add_input(names[1], ncq, cm.EVAL_NONE)
        qa.add_output(names[2], sz, em)
        op = ceed.Operator(qa)
        op.set_field(names[0], ru, bu, cm.VECTOR_ACTIVE)
        op.set_field(names[1], rq, None, qd)
        op.set_field(names[2], ru, bu, cm.VECTOR_ACTIVE)
        u, v = ceed.Vector(nn), ceed.Vector(nn)
        uu = seeded_uniform(nn)
        u.set_array(uu)
        op.apply(u, v)
        assert op.is_fused
        assert rel(v.get_array_read(), golden[key + "_v"]) < OP_TOL


def test_against_live_reference_library(cm, refceed):
    """Same operator, same inputs on the unmodified reference /cpu/self/ref/serial (when oracle/_ref travelled here)."""
    rc = refceed.RefCeed("/cpu/self/ref/serial")
    for bp, p, nel in [(3, 5, (2, 2, 1)), (1, 6, (2, 1, 1)), (6, 5, (1, 1, 2)), (3, 3, (4, 3, 2))]:
        prob = make_problem(cm, bp, p, nel)
        ref = refceed.RefBP(rc, bp, p, prob.num_elem, prob.num_nodes, prob.offsets, prob.coords)
        u = seeded_uniform(prob.num_dofs, 11)
        prob.u.set_array(u)
        prob.op.apply(prob.u, prob.v)
        assert rel(prob.qdata.get_array_read(), ref.qdata_array()) < OP_TOL
        assert rel(prob.v.get_array_read(), ref.apply(u)) < OP_TOL


def test_deterministic_scatter_is_bitwise_reproducible_and_equals_serial_order(cm):
    bp, p, nel = 3, 3, (5, 4, 3)
    pa = make_problem(cm, bp, p, nel, mode=0)
    pb = make_problem(cm, bp, p, nel, mode=2)  # E-vector + ordered CSR transpose (ascending (elem,node) like the serial reference)
    u = seeded_uniform(pa.num_dofs, 13)
    outs = []
    for prob in (pa, pa, pb):
        prob.u.set_array(u)
        prob.op.apply(prob.u, prob.v)
        outs.append(prob.v.get_array_read())
    assert np.array_equal(outs[0], outs[1])  # run-to-run bitwise identical
    assert np.array_equal(outs[0], outs[2])  # owner/halo scheme sums in the same order as the ordered transpose
    # in-kernel ordered completion (mode 3, cooperative launch, per-group flags): same bits again, also across element-group
    # sizes, repeated applies (epoch counter) and ApplyAdd
    for bp, p, nel, epg in [(3, 3, (5, 4, 3), 0), (3, 2, (7, 5, 3), 4), (1, 3, (6, 5, 4), 3), (6, 2, (4, 3, 3), 2), (5, 4, (4, 4, 3), 1)]:
        pa, pc = make_problem(cm, bp, p, nel, mode=0), make_problem(cm, bp, p, nel, mode=3)
        if epg:
            pc.op.set_tuning(epg, 0)
        u = seeded_uniform(pa.num_dofs, 17)
        pa.u.set_array(u), pc.u.set_array(u)
        pa.op.apply(pa.u, pa.v)
        for _ in range(3):
            pc.v.set_value(3.0)
            pc.op.apply(pc.u, pc.v)
            assert np.array_equal(pa.v.get_array_read(), pc.v.get_array_read()), (bp, p, nel)
        pa.op.apply_add(pa.u, pa.v), pc.op.apply_add(pc.u, pc.v)  # v + (c0 + c1 + ..) vs ((v + c0) + c1) + ..: equal up to rounding
        assert rel(pc.v.get_array_read(), pa.v.get_array_read()) < 1e-14, (bp, p, nel)


def test_unfused_fallback_matches_fused(cm, monkeypatch):
    bp, p, nel = 3, 2, (3, 2, 2)
    fused = make_problem(cm, bp, p, nel)
    monkeypatch.setenv("CEED_B200_NO_FUSE", "1")
    unf = make_problem(cm, bp, p, nel)
    assert not unf.op.is_fused  # operator setup (lazy) happens here, while the switch is still set
    monkeypatch.delenv("CEED_B200_NO_FUSE")
    assert fused.op.is_fused
    u = seeded_uniform(fused.num_dofs, 17)
    for prob in (fused, unf):
        prob.u.set_array(u)
        prob.op.apply(prob.u, prob.v)
    assert rel(fused.qdata.get_array_read(), unf.qdata.get_array_read()) < OP_TOL
    assert rel(fused.v.get_array_read(), unf.v.get_array_read()) < OP_TOL


def test_tail_batches_and_tuning(cm, oracle):
    """Element counts that are not a multiple of the batch size, forced batch sizes."""
    bp, p = 3, 2
    for nel, epb in [((1, 1, 1), 0), ((5, 1, 1), 4), ((7, 3, 1), 16), ((3, 3, 3), 5)]:
        prob = make_problem(cm, bp, p, nel)
        if epb:
            prob.op.set_tuning(epb, 0)
        u = seeded_uniform(prob.num_dofs, 19)
        prob.u.set_array(u)
        prob.op.apply(prob.u, prob.v)
        qd = oracle.bp_qdata(bp, p, prob.offsets, prob.coords)
        assert rel(prob.v.get_array_read(), oracle.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)) < OP_TOL


SHAPES = [dict(group_warps=2, cta_warps=2), dict(group_warps=2, cta_warps=8), dict(group_warps=4, cta_warps=4), dict(qf_mode=1, qf_unroll=2),
          dict(qf_mode=2, qf_unroll=2), dict(qf_mode=2, group_warps=2, cta_warps=4, elems_per_group=3), dict(stage_mask=17), dict(stage_mask=9),
          dict(stage_mask=0, cta_warps=1), dict(stage_mask=19, group_warps=2, cta_warps=4), dict(qf_mode=3), dict(qf_mode=0),
          # quadrature data through cp.async.bulk (TMA) + mbarrier (stage bit 32), also with odd Q^3 (8-byte source misalignment)
          dict(stage_mask=33), dict(stage_mask=33, group_warps=2, cta_warps=4, elems_per_group=3), dict(stage_mask=41, qf_mode=1, qf_unroll=2),
          dict(stage_mask=32, group_warps=4, cta_warps=4, elems_per_group=2),
          # conflict-free swizzled plane layout (stage bit 256), alone and with the other staging options
          dict(stage_mask=257), dict(stage_mask=257, group_warps=2, cta_warps=4, elems_per_group=2), dict(stage_mask=289, group_warps=2, cta_warps=2),
          dict(stage_mask=265, group_warps=4, cta_warps=4, elems_per_group=3), dict(stage_mask=256, qf_mode=0, cta_warps=2),
          # even-Q linear layout (stage bit 512; Q odd: falls back to the padded layout): unpadded rows, z-stride = Q (mod 16), 16-byte x-lines
          dict(stage_mask=513), dict(stage_mask=512, group_warps=2, cta_warps=4, elems_per_group=3), dict(stage_mask=521, group_warps=4, cta_warps=4, elems_per_group=2),
          # lean in-place-plane kernel of gradient-free operators (layout 4; other operators fall back to their own layouts)
          dict(qf_mode=4, group_warps=1, cta_warps=4, elems_per_group=6, stage_mask=0), dict(qf_mode=4, group_warps=1, cta_warps=2, elems_per_group=3, stage_mask=40)]


@pytest.mark.parametrize("bp,p,nel", [(3, 2, (5, 3, 2)), (5, 3, (3, 3, 2)), (1, 3, (4, 3, 3)), (6, 2, (3, 2, 2)), (3, 4, (3, 2, 2)), (3, 6, (2, 2, 1)), (5, 7, (2, 1, 2)),
                                      (3, 5, (2, 2, 2)), (5, 6, (1, 2, 2))])
def test_kernel_shapes_give_identical_results(cm, oracle, monkeypatch, bp, p, nel):
    """Every kernel shape the autotuner may pick (multi-warp groups, pointwise / point-pair QFunction stage, cp.async staging
    variants incl. the quadrature-data ring) computes the same operator: compared with the oracle and with the default shape."""
    monkeypatch.setenv("CEED_B200_NO_TUNE_TABLE", "1")  # knobs a shape leaves open come from the heuristics, not the table
    prob = make_problem(cm, bp, p, nel)
    u = seeded_uniform(prob.num_dofs, 23)
    prob.u.set_array(u)
    qd = oracle.bp_qdata(bp, p, prob.offsets, prob.coords)
    ref = oracle.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)
    prob.op.apply(prob.u, prob.v)
    v0 = prob.v.get_array_read()
    assert rel(v0, ref) < OP_TOL
    for shape in SHAPES:
        prob.op.set_kernel_shape(**shape)
        prob.v.set_value(-7.0)
        prob.op.apply(prob.u, prob.v)
        got = prob.op.get_kernel_shape()
        for k, val in shape.items():
            if (k == "qf_mode" and val in (2, 3, 4)) or shape.get("qf_mode") == 4:
                continue  # point pairs need an even Q, x-line fusion / the lean kernel a gradient-free operator: otherwise they fall back
            if k == "cta_warps":
                assert got[k] <= val, (shape, got)  # groups per CTA are reduced when their shared memory would not fit
                continue
            assert got[k] == val, (shape, got)
        assert rel(prob.v.get_array_read(), ref) < OP_TOL, shape
        assert rel(prob.v.get_array_read(), v0) < 1e-13, shape


LEAN_CASES = [(1, 3, (5, 3, 3)), (1, 3, (7, 5, 3)), (2, 3, (3, 2, 3)), (1, 1, (5, 4, 3)), (1, 2, (3, 3, 3)), (1, 4, (3, 2, 3)), (1, 5, (2, 3, 2)), (2, 2, (3, 3, 2)),
              (2, 1, (4, 4, 3)), (2, 4, (2, 2, 1))]


@pytest.mark.parametrize("bp,p,nel", LEAN_CASES)
def test_lean_kernel_matches_oracle(cm, oracle, monkeypatch, bp, p, nel):
    """The lean in-place-plane kernel (QFunction layout 4, b200_opgen_lean.cpp) on BP1 / BP2: every batch width incl. partial tail
    batches, deterministic and atomic scatter, Apply and ApplyAdd, direct loads and the bulk-copy (cp.async.bulk + mbarrier) pipelines
    for the index tables (stage bit 8) and the quadrature data (bit 32) and the bulk L2 prefetch (bit 64), lexicographic and Morton
    element order -- against the oracle; the deterministic variants also bitwise against each other."""
    from libceed_b200 import mesh as M
    monkeypatch.setenv("CEED_B200_NO_TUNE_TABLE", "1")
    perm = M.morton_permutation(*nel) if (bp + p) % 2 else None
    v_det = None
    for mode in (0, 1):
        prob = make_problem(cm, bp, p, nel, mode=mode, elem_perm=perm)
        u = seeded_uniform(prob.num_dofs, 31)
        prob.u.set_array(u)
        qd = oracle.bp_qdata(bp, p, prob.offsets, prob.coords)
        ref = oracle.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)
        # (bit 1: element-interleaved node columns, bit 2: element stride = P (mod 16), bit 4: quadrature data through 16-byte loads)
        shapes = [(1, 1, 0), (3, 2, 0), (6, 4, 0), (8, 8, 0), (3, 1, 32), (6, 4, 40), (5, 2, 8), (4, 4, 64), (2, 2, 72)]
        if nel in ((7, 5, 3), (3, 2, 3), (3, 3, 3), (2, 2, 1)):  # (the new bits on four of the cases: every shape costs two NVRTC compilations)
            shapes += [(6, 4, 1), (8, 4, 3), (5, 2, 7), (8, 4, 4), (6, 4, 12), (4, 2, 39), (7, 1, 5), (8, 2, 35)] if mode == 0 else [(8, 4, 7), (3, 2, 6)]
        for E, warps, stage in shapes:
            prob.op.set_kernel_shape(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
            prob.v.set_value(-3.0)
            prob.op.apply(prob.u, prob.v)
            got = prob.op.get_kernel_shape()
            assert got["qf_mode"] == 4 and got["stage_mask"] == stage and got["elems_per_group"] == min(E, prob.num_elem), (got, E, stage)
            v = prob.v.get_array_read().copy()
            assert rel(v, ref) < OP_TOL, (mode, E, warps, stage)
            if mode == 0:
                v_det = v if v_det is None else v_det
                assert np.array_equal(v, v_det), (E, warps, stage)  # same ascending E-order whatever the shape
            w0 = seeded_uniform(prob.num_dofs, 5)
            prob.v.set_array(w0)
            prob.op.apply_add(prob.u, prob.v)
            assert np.abs(prob.v.get_array_read() - w0 - ref).max() < 1e-12 * max(1.0, np.abs(ref).max()), (mode, E, warps, stage)


@pytest.mark.parametrize("bp,p,nel", [(1, 3, (5, 3, 3)), (1, 2, (4, 3, 3)), (2, 3, (3, 3, 2)), (1, 4, (3, 2, 2))])
def test_lean_kernel_vector_qdata_loads_with_any_alignment(cm, oracle, monkeypatch, bp, p, nel):
    """Stage bit 4 of the lean kernel reads the quadrature data of an x-line through 16-byte loads from the aligned window around the line.
    A user array that starts 8 bytes off a 16-byte boundary flips which lines are aligned (odd Q) or misaligns all of them (even Q), and
    puts the first line's window before the start of the array: same bits as the 8-byte loads in every case."""
    import torch
    monkeypatch.setenv("CEED_B200_NO_TUNE_TABLE", "1")
    prob = make_problem(cm, bp, p, nel)
    u = seeded_uniform(prob.num_dofs, 41)
    prob.u.set_array(u)
    qd = oracle.bp_qdata(bp, p, prob.offsets, prob.coords)
    ref = oracle.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)
    prob.op.set_kernel_shape(qf_mode=4, elems_per_group=4, cta_warps=2, group_warps=1, stage_mask=0)
    prob.op.apply(prob.u, prob.v)
    v0 = prob.v.get_array_read().copy()
    assert rel(v0, ref) < OP_TOL
    q_host = prob.qdata.get_array_read().copy()
    n = q_host.size
    buf = torch.full((n + 3,), float("nan"), dtype=torch.float64, device="cuda")
    for shift in (0, 1):  # (torch allocations are 256-byte aligned: shift 1 = 8 bytes off)
        view = buf[shift:shift + n]
        view.copy_(torch.from_numpy(q_host))
        assert view.data_ptr() % 16 == 8 * shift
        prob.qdata.set_array(view, cm.MEM_DEVICE, cm.USE_POINTER)
        for E, warps, stage in ((4, 2, 4), (6, 4, 5), (3, 1, 7), (8, 4, 4)):
            prob.op.set_kernel_shape(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
            prob.v.set_value(-3.0)
            prob.op.apply(prob.u, prob.v)
            assert prob.op.get_kernel_shape()["stage_mask"] == stage
            assert np.array_equal(prob.v.get_array_read(), v0), (shift, E, warps, stage)
    prob.qdata.take_array(cm.MEM_DEVICE)


@pytest.mark.parametrize("bp,p,nel,morton", [(1, 3, (32, 31, 31), True), (1, 3, (33, 30, 29), False), (2, 2, (30, 29, 28), True)])
def test_lean_kernel_run_scatter_is_bitwise_equal(cm, oracle, monkeypatch, bp, p, nel, morton):
    """Experimental run scatter (CEED_B200_RUNS; B200RunScatter): warps own contiguous element runs and add the E-entries whose earlier
    touchers they processed themselves straight into v.  Same ascending E-order: bitwise equal to the owner/halo tables.  Also the in-kernel
    finalize (stage bit 128): one cooperative launch instead of fused kernel + finalize kernel, same bits."""
    from libceed_b200 import mesh as M
    monkeypatch.setenv("CEED_B200_NO_TUNE_TABLE", "1")
    perm = M.morton_permutation(*nel) if morton else None
    prob = make_problem(cm, bp, p, nel, elem_perm=perm)
    u = seeded_uniform(prob.num_dofs, 37)
    prob.u.set_array(u)
    qd = oracle.bp_qdata(bp, p, prob.offsets, prob.coords)
    ref = oracle.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)
    prob.op.set_kernel_shape(qf_mode=4, elems_per_group=4, cta_warps=4, group_warps=1, stage_mask=0)
    prob.op.apply(prob.u, prob.v)
    v_classic = prob.v.get_array_read().copy()
    assert rel(v_classic, ref) < OP_TOL
    # (stage bit 128: in-kernel finalize -- warps fold the shared nodes of finished element parts into v themselves; cooperative launch)
    for E, warps, stage, parts in ((1, 4, 0, 0), (2, 4, 0, 0), (6, 4, 0, 0), (8, 2, 0, 0), (6, 4, 128, 8), (2, 4, 128, 3), (5, 2, 128, 16), (8, 8, 128, 5)):
        if stage:
            monkeypatch.delenv("CEED_B200_RUNS", raising=False)
            monkeypatch.setenv("CEED_B200_PARTS", str(parts))
        else:
            monkeypatch.setenv("CEED_B200_RUNS", "1")
        prob.op.set_kernel_shape(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
        prob.v.set_value(-3.0)
        prob.op.apply(prob.u, prob.v)
        assert prob.op.get_kernel_shape()["stage_mask"] == stage
        assert np.array_equal(prob.v.get_array_read(), v_classic), (E, warps)
        w0 = seeded_uniform(prob.num_dofs, 5)
        prob.v.set_array(w0)
        prob.op.apply_add(prob.u, prob.v)
        assert np.abs(prob.v.get_array_read() - w0 - ref).max() < 1e-12 * max(1.0, np.abs(ref).max()), (E, warps)


@pytest.mark.parametrize("bp,p,nel", [(3, 3, (12, 12, 12)), (1, 2, (20, 20, 20))])
def test_autotuner_picks_a_shape_and_keeps_results(cm, oracle, tmp_path, monkeypatch, bp, p, nel):
    """The core's autotuner on a diffusion operator (general kernel) and on a mass operator (lean kernel candidates incl. the
    interleaved-column / vector-load stage bits): the tuned operator still matches the oracle and the table line is written."""
    monkeypatch.setenv("CEED_B200_TUNE_SAVE", str(tmp_path / "t.tune"))
    monkeypatch.setenv("CEED_B200_NO_TUNE_TABLE", "1")
    prob = make_problem(cm, bp, p, nel)
    prob.ceed.set_autotune(2)
    u = seeded_uniform(prob.num_dofs, 29)
    prob.u.set_array(u)
    prob.op.apply(prob.u, prob.v)  # tunes on this call
    qd = oracle.bp_qdata(bp, p, prob.offsets, prob.coords)
    assert rel(prob.v.get_array_read(), oracle.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)) < OP_TOL
    line = open(tmp_path / "t.tune").read().split()
    assert line[0] == prob.op.get_kernel_shape()["signature"] and len(line) >= 8


def test_interface_pack_unpack_kernels(cm):
    """ceedb200_iface_pack / ceedb200_iface_unpack_sum against the host interpretation of the same tables."""
    import torch
    from libceed_b200.parallel import CudaInterfaceKernels, build_interface_tables
    ceed = cm.Ceed("/gpu/cuda/b200")
    part = M.Partition((4, 4, 2), 2, 4, 1)
    ncomp, nloc = 2, part.num_local_nodes
    ranks, seg, send_idx, node, ptr, src = build_interface_tables(part, ncomp, nloc)
    rng = np.random.default_rng(5)
    v = rng.uniform(-1, 1, ncomp * nloc)
    recv = rng.uniform(-1, 1, send_idx.size)
    expect = v.copy()
    for i in range(node.size):
        terms = [v[node[i]] if s < 0 else recv[s] for s in src[ptr[i]:ptr[i + 1]]]
        acc = terms[0]
        for x in terms[1:]:
            acc = acc + x
        expect[node[i]] = acc
    dev = torch.device("cuda")
    k = CudaInterfaceKernels(ceed)
    v_d, send_d = torch.from_numpy(v).to(dev), torch.empty(send_idx.size, dtype=torch.float64, device=dev)
    k.pack(v_d, torch.from_numpy(send_idx).to(dev), send_d)
    ceed.synchronize()
    assert np.array_equal(send_d.cpu().numpy(), v[send_idx])
    k.unpack_sum(v_d, torch.from_numpy(node).to(dev), torch.from_numpy(ptr).to(dev), torch.from_numpy(src).to(dev), torch.from_numpy(recv).to(dev))
    ceed.synchronize()
    assert np.array_equal(v_d.cpu().numpy(), expect)  # same order of additions -> same bits


def test_device_resident_cg_converges(cm):
    """CG around the fused operator with device-resident scalars (cg.DeviceCG): solves M x = b for the BP1 mass operator."""
    import torch
    from libceed_b200.cg import DeviceCG
    prob = make_problem(cm, 1, 3, (4, 4, 3))
    n, dev = prob.num_dofs, torch.device("cuda")
    cg = DeviceCG(prob.ceed, prob.op, prob.u, prob.v, n, dev)
    x_true = torch.from_numpy(seeded_uniform(n, 31)).to(dev)
    cg.p.copy_(x_true)
    cg.apply()
    b = cg.Ap.clone()
    cg.start(b)
    r0 = cg.residual_norm2()
    cg.iterate(40)
    r40 = cg.residual_norm2()
    cg.iterate(260)
    assert r40 < 1e-2 * r0 and cg.residual_norm2() < 1e-9 * r0, (r0, r40, cg.residual_norm2())
    assert float((cg.x - x_true).abs().max()) < 1e-6 * float(x_true.abs().max())
    # the dot kernel against numpy, weighted and unweighted
    a, c, w = (torch.from_numpy(seeded_uniform(n, s)).to(dev) for s in (1, 2, 3))
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    lib, C = prob.ceed._lib, __import__("ctypes")
    prob.ceed._chk(lib.ceedb200_cg_dot(prob.ceed._ptr, C.c_void_p(a.data_ptr()), C.c_void_p(c.data_ptr()), C.c_void_p(w.data_ptr()), n, C.c_void_p(out.data_ptr())))
    assert abs(float(out.item()) - float((a * c * w).sum().item())) < 1e-12 * n


@pytest.mark.parametrize("bp,p,nel,kw", [(1, 3, (24, 23, 22), {}), (3, 2, (30, 29, 28), {}), (2, 2, (20, 21, 22), {}), (4, 1, (30, 30, 30), dict(interlaced=True)),
                                         (1, 3, (20, 20, 20), dict(morton=True)), (5, 4, (14, 13, 12), {})])
def test_streamed_host_buffer_apply_is_bitwise_equal(cm, bp, p, nel, kw):
    """ceedb200_operator_apply_streamed (chunked H2D / apply / finalize / D2H pipeline for host-resident vectors): bitwise the plain apply
    for every chunk count, blocked and interlaced components, an element order without locality (Morton); the result is on the host when
    the call returns and the vector is valid on both sides; pageable host memory and small meshes fall back to the plain apply."""
    import torch
    from libceed_b200 import mesh as M
    kw = dict(kw)
    perm = M.morton_permutation(*nel) if kw.pop("morton", False) else None
    prob = make_problem(cm, bp, p, nel, elem_perm=perm, **kw)
    n = prob.num_dofs
    u_t, v_t = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
    u_np, v_np = u_t.numpy(), v_t.numpy()
    u_np[:] = seeded_uniform(n, 41)
    prob.u.set_array(u_np.copy())
    prob.op.apply(prob.u, prob.v)
    ref = prob.v.get_array_read().copy()
    for K in (0, 2, 5, 33):
        v_np[:] = -7.0
        prob.u.set_array(u_np, cm.MEM_HOST, cm.USE_POINTER)
        prob.v.set_array(v_np, cm.MEM_HOST, cm.USE_POINTER)
        used = prob.op.apply_streamed(prob.u, prob.v, K)
        assert used == (prob.num_elem >= 4096)
        if used:
            assert np.array_equal(v_np, ref), K           # already on the host, no sync needed
        assert np.array_equal(prob.v.get_array_read(), ref), K
        prob.u.take_array(), prob.v.take_array()
    u_pg, v_pg = u_np.copy(), np.zeros(n)                  # pageable memory: plain apply, same result
    prob.u.set_array(u_pg, cm.MEM_HOST, cm.USE_POINTER)
    prob.v.set_array(v_pg, cm.MEM_HOST, cm.USE_POINTER)
    assert not prob.op.apply_streamed(prob.u, prob.v, 0)
    assert np.array_equal(prob.v.get_array_read(), ref)
    prob.u.take_array(), prob.v.take_array()
    # the restriction keeps its chunk parts: a plain apply afterwards still gives the same bits
    prob.u.set_array(u_np.copy())
    prob.op.apply(prob.u, prob.v)
    assert np.array_equal(prob.v.get_array_read(), ref)


def test_device_cg_with_dirichlet_conditions_solves_poisson(cm):
    """BP3 (Poisson) with homogeneous Dirichlet conditions on the box boundary, the reference's BP3 setting (examples/petsc/bps.c with
    DMPlex-constrained boundary DoFs): constrained rows act as identity rows (ceedb200_cg_constrain).  Without them the operator is
    singular; with them CG recovers a manufactured solution that vanishes on the boundary."""
    import torch
    from libceed_b200.cg import DeviceCG
    p, nel = 2, (6, 5, 5)
    prob = make_problem(cm, 3, p, nel)
    n, dev = prob.num_dofs, torch.device("cuda")
    nx, ny, nz = (k * p + 1 for k in nel)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    free = ((ix > 0) & (ix < nx - 1) & (iy > 0) & (iy < ny - 1) & (iz > 0) & (iz < nz - 1)).reshape(-1).astype(np.float64)
    cg = DeviceCG(prob.ceed, prob.op, prob.u, prob.v, n, dev, free_mask=free)
    x_true = torch.from_numpy(seeded_uniform(n, 33) * free).to(dev)
    cg.p.copy_(x_true)
    cg.apply()
    b = cg.Ap.clone()
    assert float((b * torch.from_numpy(1.0 - free).to(dev)).abs().max()) == 0.0  # identity rows: b = x_true = 0 on the boundary
    cg.start(b)
    r0 = cg.residual_norm2()
    cg.iterate(400)
    assert cg.residual_norm2() < 1e-10 * r0, (r0, cg.residual_norm2())
    assert float((cg.x - x_true).abs().max()) < 1e-7 * float(x_true.abs().max())
    assert float((cg.x * torch.from_numpy(1.0 - free).to(dev)).abs().max()) == 0.0   # the iterates never leave the constrained subspace


@pytest.mark.parametrize("bp,p", [(1, 3), (3, 6), (5, 7), (6, 4)])
def test_full_size_vs_reference_cpu_backend(cm, refceed, bp, p):
    """BASELINE.json sizes (10M DoFs) against the UNMODIFIED reference's /cpu/self/opt/blocked on the same mesh and input: the
    persistent grid-stride path with its full grid, tail groups and the tuned kernel shape, checked value by value (<= 1e-12)."""
    ncomp = BP_TABLE[bp][0]
    nel = M.choose_elements(10_000_000, p, ncomp)
    prob = make_problem(cm, bp, p, nel)
    rc = refceed.RefCeed("/cpu/self/opt/blocked")
    ref = refceed.RefBP(rc, bp, p, prob.num_elem, prob.num_nodes, prob.offsets, prob.coords)
    u = seeded_uniform(prob.num_dofs, 41)
    prob.u.set_array(u)
    prob.op.apply(prob.u, prob.v)
    v = prob.v.get_array_read()
    v_ref = ref.apply(u)
    assert rel(v, v_ref) < OP_TOL, (bp, p, rel(v, v_ref))
    # a second apply into a poisoned vector: overwrite semantics and bitwise reproducibility at full size
    prob.v.set_value(-3.0)
    prob.op.apply(prob.u, prob.v)
    assert np.array_equal(prob.v.get_array_read(), v)


@pytest.mark.parametrize("bp,p", [(1, 3), (3, 6), (5, 7), (6, 4)])
def test_full_size_properties(cm, bp, p):
    """BASELINE.json sizes (10M DoFs): properties that need no CPU oracle."""
    ncomp = BP_TABLE[bp][0]
    nel = M.choose_elements(10_000_000, p, ncomp)
    prob = make_problem(cm, bp, p, nel)
    ceed = prob.ceed
    n = prob.num_dofs
    u1, u2 = seeded_uniform(n, 1), seeded_uniform(n, 2)

    def A(x):
        prob.u.set_array(x)
        prob.op.apply(prob.u, prob.v)
        return prob.v.get_array_read()

    v1, v2 = A(u1), A(u2)
    assert rel(A(2.0 * u1 - 0.5 * u2), 2.0 * v1 - 0.5 * v2) < 1e-12  # linearity
    assert abs(u2 @ v1 - u1 @ v2) < 1e-10 * abs(u1 @ v1)  # symmetry
    assert u1 @ v1 > 0  # positive (semi-)definite
    ones = A(np.ones(n))
    if BP_TABLE[bp][1] == "diff":
        assert np.abs(ones).max() < 1e-10 * np.abs(v1).max()  # constants in the kernel
    else:
        assert abs(ones.sum() / ncomp - 1.0) < 1e-6  # sum of mass-matrix rows = volume of the (smoothly mapped) unit cube, up to quadrature error
    assert np.array_equal(A(u1), v1)  # idempotent / reproducible
    del ceed


def test_fallback_ladder_never_fails_a_valid_operator(cm, oracle, monkeypatch, tmp_path):
    """Try-compile -> fall back (backends/cuda-gen/ceed-cuda-gen-operator.c:291-298): a fused kernel that cannot be built first drops
    to a conservative shape, then to the unfused kernels -- the apply still succeeds with the same result.  Build failures are
    injected with the CEED_B200_FAIL_FUSED_BUILD test hook; a register-hungry user QFunction exercises the spill rung for real."""
    bp, p, nel = 3, 2, (3, 3, 2)
    ref_prob = make_problem(cm, bp, p, nel)
    u = seeded_uniform(ref_prob.num_dofs, 37)
    ref_prob.u.set_array(u)
    ref_prob.op.apply(ref_prob.u, ref_prob.v)
    v_ref = ref_prob.v.get_array_read().copy()
    for nfail, expect_fused in ((1, True), (2, False)):
        # fresh process-wide counter per case: the hook counts failures since library load, so raise the limit cumulatively
        prob = make_problem(cm, bp, p, nel)  # qdata is computed by the (already cached) setup kernel
        prob.u.set_array(u)
        monkeypatch.setenv("CEED_B200_FAIL_FUSED_BUILD", str({1: 1, 2: 3}[nfail]))
        prob.op.set_kernel_shape(elems_per_group=2, group_warps=1, cta_warps=2)  # a shape of its own: not in the module cache
        prob.op.apply(prob.u, prob.v)
        monkeypatch.delenv("CEED_B200_FAIL_FUSED_BUILD")
        assert rel(prob.v.get_array_read(), v_ref) < OP_TOL
        assert prob.op.is_fused == expect_fused
        prob.op.apply_add(prob.u, prob.v)
        assert rel(prob.v.get_array_read(), 2 * v_ref) < OP_TOL
    # a QFunction that needs far more registers than any shape provides: whatever rung it lands on, the result is right
    src = tmp_path / "heavy.h"
    src.write_text('''
#include <ceed/types.h>
CEED_QFUNCTION(HeavyDiff)(void *ctx, const CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out) {
  const CeedScalar(*ug)[CEED_Q_VLA] = (const CeedScalar(*)[CEED_Q_VLA])in[0];
  const CeedScalar(*qd)[CEED_Q_VLA] = (const CeedScalar(*)[CEED_Q_VLA])in[1];
  CeedScalar(*vg)[CEED_Q_VLA]       = (CeedScalar(*)[CEED_Q_VLA])out[0];
  for (CeedInt i = 0; i < Q; i++) {
    CeedScalar w[160];
    for (int k = 0; k < 160; k++) w[k] = qd[k % 7][i] * (1.0 + 1e-3 * k);
    CeedScalar s = 0.0;
    for (int k = 0; k < 160; k++) s += w[(k * 37 + (int)(ug[0][i] > 0)) % 160] - w[(k * 37) % 160];   /* == 0, keeps w alive and dynamically indexed */
    const CeedScalar u0 = ug[0][i], u1 = ug[1][i], u2 = ug[2][i];
    vg[0][i] = qd[1][i] * u0 + qd[2][i] * u1 + qd[3][i] * u2 + 0.0 * s;
    vg[1][i] = qd[2][i] * u0 + qd[4][i] * u1 + qd[5][i] * u2;
    vg[2][i] = qd[3][i] * u0 + qd[5][i] * u1 + qd[6][i] * u2;
  }
  return 0;
}
''')
    ceed = ref_prob.ceed
    qf = ceed.QFunction(str(src), "HeavyDiff")
    qf.add_input("u", 3, cm.EVAL_GRAD)
    qf.add_input("qdata", 7, cm.EVAL_NONE)
    qf.add_output("v", 3, cm.EVAL_GRAD)
    op = ceed.Operator(qf)
    op.set_field("u", ref_prob.rstr_u, ref_prob.basis_u, cm.VECTOR_ACTIVE)
    op.set_field("qdata", ref_prob.rstr_qd, cm.BASIS_NONE, ref_prob.qdata)
    op.set_field("v", ref_prob.rstr_u, ref_prob.basis_u, cm.VECTOR_ACTIVE)
    v2 = ceed.Vector(ref_prob.num_dofs)
    op.apply(ref_prob.u, v2)
    assert rel(v2.get_array_read(), v_ref) < OP_TOL


@pytest.mark.parametrize("bp,p,nel", [(3, 7, (2, 1, 1)), (3, 8, (1, 1, 2)), (1, 7, (1, 2, 1)), (4, 7, (1, 1, 1))])
def test_swizzled_layout_row_width_16(cm, oracle, monkeypatch, bp, p, nel):
    """Q = 9, 10: the conflict-free swizzled planes with 16-wide rows (stage bit 256) against the oracle and the padded layout."""
    monkeypatch.setenv("CEED_B200_NO_TUNE_TABLE", "1")
    prob = make_problem(cm, bp, p, nel)
    u = seeded_uniform(prob.num_dofs, 43)
    prob.u.set_array(u)
    qd = oracle.bp_qdata(bp, p, prob.offsets, prob.coords)
    ref = oracle.bp_apply(bp, p, prob.offsets, prob.num_nodes, qd, u)
    prob.op.apply(prob.u, prob.v)
    v0 = prob.v.get_array_read().copy()
    assert rel(v0, ref) < OP_TOL
    for shape in (dict(stage_mask=257), dict(stage_mask=257, group_warps=4, cta_warps=4), dict(stage_mask=265, group_warps=2, cta_warps=2),
                  dict(stage_mask=256, group_warps=1, cta_warps=2)):
        prob.op.set_kernel_shape(**shape)
        prob.v.set_value(-7.0)
        prob.op.apply(prob.u, prob.v)
        assert prob.op.is_fused and prob.op.get_kernel_shape()["stage_mask"] == shape["stage_mask"]
        assert rel(prob.v.get_array_read(), ref) < OP_TOL, shape
        assert rel(prob.v.get_array_read(), v0) < 1e-13, shape
        prob.op.apply_add(prob.u, prob.v)
        assert rel(prob.v.get_array_read(), 2 * ref) < OP_TOL, shape
