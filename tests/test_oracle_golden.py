"""CPU tests (no GPU): pin the oracle (oracle/ceed_oracle.c) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and, when oracle/_ref is present, against the reference run live."""
import numpy as np
import pytest

from conftest import BP_CASES, bp_case_key
from libceed_b200 import mesh as M
from libceed_b200.bp import seeded_uniform

TOL = 1e-13  # oracle vs reference: same algorithm and order, differences only from compiler FMA contraction


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("P", range(2, 10))
def test_basis_matrices_match_reference(golden, oracle, P):
    for Q, qm in ((P + 1, 0), (P, 1), (P + 2, 1), (max(2, P - 1), 0)):
        interp, grad, q_ref, q_w = oracle.lagrange_1d(P, Q, qm)
        k = f"basis_P{P}_Q{Q}_m{qm}_"
        assert rel(interp, golden[k + "interp_1d"]) < TOL
        assert rel(grad, golden[k + "grad_1d"]) < TOL
        assert np.abs(q_ref - golden[k + "q_ref_1d"]).max() < 1e-15
        assert rel(q_w, golden[k + "q_weight_1d"]) < TOL
        if Q >= P:
            assert rel(oracle.collocated_grad(P, Q, interp, grad), golden[k + "collo_grad_1d"]) < 1e-12


def test_quadrature_known_answers(oracle):
    # known-answer values also asserted by the reference's own t3xx tests: weights sum to 2, Gauss integrates x^(2Q-1) exactly
    for Q in range(1, 12):
        x, w = oracle.gauss(Q)
        assert abs(w.sum() - 2.0) < 1e-14
        for k in range(0, 2 * Q, 2):
            assert abs((w * x ** k).sum() - 2.0 / (k + 1)) < 1e-13
    for Q in range(2, 12):
        x, w = oracle.lobatto(Q)
        assert x[0] == -1.0 and x[-1] == 1.0 and abs(w.sum() - 2.0) < 1e-14
        for k in range(0, 2 * Q - 2, 2):
            assert abs((w * x ** k).sum() - 2.0 / (k + 1)) < 1e-13


def test_restriction_bit_exact(golden, oracle):
    nelem, esize, ncomp, lsize = [int(v) for v in golden["rstr_dims"]]
    off = golden["rstr_offsets"]
    e = oracle.restriction_offset(nelem, esize, ncomp, lsize, off, 0, golden["rstr_u"], nelem * esize * ncomp)
    assert tuple(golden["rstr_layout"]) == (1, esize, esize * ncomp)  # CPU reference E-layout [elem][comp][node]
    assert np.array_equal(e, golden["rstr_e"])  # gather is exact
    lt = oracle.restriction_offset(nelem, esize, ncomp, lsize, off, 1, golden["rstr_w"], ncomp * lsize)
    assert np.array_equal(lt, golden["rstr_lt"])  # scatter-add in ascending (elem, comp, node) order: bit-exact


def test_restriction_edge_cases(oracle):
    # empty restriction, single element, all entries hitting one node
    assert oracle.restriction_offset(0, 4, 1, 1, np.zeros(0, np.int32), 0, np.ones(3), 0).size == 0
    off = np.zeros(8, np.int32)
    lt = oracle.restriction_offset(2, 4, 1, 3, off, 1, np.arange(8.0), 3)
    assert lt[0] == sum(range(8)) and lt[1] == 0 and lt[2] == 0
    s = oracle.restriction_strided(2, 3, 2, (1, 6, 3), 0, np.arange(12.0), 12)  # [elem][comp][node] from [comp][elem][node]
    assert list(s[:6]) == [0, 1, 2, 6, 7, 8]


@pytest.mark.parametrize("case", BP_CASES, ids=lambda c: bp_case_key(*c))
def test_bp_operator_matches_golden(golden, oracle, case):
    bp, p, nel, gallery, interlaced = case
    off = M.hex_offsets(*nel, p)
    coords = M.hex_coords(*nel, p)
    nn = coords.shape[1]
    key = bp_case_key(*case)
    qd = oracle.bp_qdata(bp, p, off, coords, gallery=gallery)
    assert rel(qd, golden[key + "_qdata"]) < TOL
    ncomp = oracle.bp_sizes(bp, gallery, p)[2]
    v = oracle.bp_apply(bp, p, off, nn, qd, seeded_uniform(ncomp * nn), gallery=gallery, interlaced=interlaced)
    assert rel(v, golden[key + "_v"]) < TOL


def test_oracle_apply_add_and_linearity(oracle):
    bp, p, nel = 3, 2, (2, 2, 2)
    off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
    nn = coords.shape[1]
    qd = oracle.bp_qdata(bp, p, off, coords)
    u1, u2 = seeded_uniform(nn, 1), seeded_uniform(nn, 2)
    v1, v2 = oracle.bp_apply(bp, p, off, nn, qd, u1), oracle.bp_apply(bp, p, off, nn, qd, u2)
    assert rel(oracle.bp_apply(bp, p, off, nn, qd, 2 * u1 - 3 * u2), 2 * v1 - 3 * v2) < 1e-13
    assert rel(oracle.bp_apply(bp, p, off, nn, qd, u1, v0=v2), v1 + v2) < 1e-14
    assert abs(u2 @ v1 - u1 @ v2) < 1e-12 * abs(u1 @ v1)  # symmetric operator
    assert np.abs(oracle.bp_apply(bp, p, off, nn, qd, np.ones(nn))).max() < 1e-12  # constants are in the kernel of diffusion


def test_oracle_against_live_reference(refceed, oracle):
    rc = refceed.RefCeed("/cpu/self/ref/serial")
    for bp, p, nel in [(3, 5, (1, 2, 1)), (1, 4, (2, 1, 1)), (6, 4, (1, 1, 1)), (3, 7, (1, 1, 1))]:
        off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
        nn = coords.shape[1]
        ref = refceed.RefBP(rc, bp, p, off.shape[0], nn, off, coords)
        u = seeded_uniform(ref.ncomp * nn, 3)
        qd = oracle.bp_qdata(bp, p, off, coords)
        assert rel(qd, ref.qdata_array()) < TOL
        assert rel(oracle.bp_apply(bp, p, off, nn, qd, u), ref.apply(u)) < TOL


def test_reference_backends_agree(refceed):
    """/cpu/self/opt/blocked and /cpu/self/avx/blocked (the CPU baselines that get timed) match the oracle backend."""
    bp, p, nel = 3, 3, (3, 2, 2)
    off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
    nn = coords.shape[1]
    u = seeded_uniform(nn, 5)
    res = {}
    for r in ("/cpu/self/ref/serial", "/cpu/self/opt/blocked", "/cpu/self/avx/blocked"):
        res[r] = refceed.RefBP(refceed.RefCeed(r), bp, p, off.shape[0], nn, off, coords).apply(u)
    for r in res:
        assert rel(res[r], res["/cpu/self/ref/serial"]) < 1e-13
