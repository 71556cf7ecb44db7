"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/ceed_b200.h declares, the host-side
math of the product matches the reference's golden matrices, the mesh/partition host logic is consistent, and the
fused-kernel generator + NVRTC produce sm_100a code for every BP configuration (compile-only, nothing is executed)."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from libceed_b200 import _lib as L
from libceed_b200 import mesh as M


def test_library_exports_every_declared_symbol():
    names = L.declared_symbols()
    assert len(names) >= 70
    lib = C.CDLL(L.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.lib()
    ceed = C.c_void_p()
    code = lib.ceedb200_init(0, C.byref(ceed))
    assert code != 0 and not ceed.value
    assert b"CUDA" in lib.ceedb200_last_error(None) or b"device" in lib.ceedb200_last_error(None)


@pytest.mark.parametrize("P", range(2, 10))
def test_host_basis_math_matches_reference_golden(golden, P):
    lib = L.lib()
    for Q, qm in ((P + 1, 0), (P, 1), (P + 2, 1), (max(2, P - 1), 0)):
        interp, grad, qref, qw = np.zeros(P * Q), np.zeros(P * Q), np.zeros(Q), np.zeros(Q)
        d = lambda a: a.ctypes.data_as(L.c_scalar_p)  # noqa: E731
        assert lib.ceedb200_host_lagrange_1d(P, Q, qm, d(interp), d(grad), d(qref), d(qw)) == 0
        k = f"basis_P{P}_Q{Q}_m{qm}_"
        assert np.abs(interp - golden[k + "interp_1d"]).max() < 1e-13
        assert np.abs(grad - golden[k + "grad_1d"]).max() < 1e-12
        assert np.abs(qref - golden[k + "q_ref_1d"]).max() < 1e-15
        assert np.abs(qw - golden[k + "q_weight_1d"]).max() < 1e-14
        if Q >= P:
            cg = np.zeros(Q * Q)
            assert lib.ceedb200_host_collocated_grad_1d(P, Q, d(interp), d(grad), d(cg)) == 0
            assert np.abs(cg - golden[k + "collo_grad_1d"]).max() < 1e-11 * max(1.0, np.abs(cg).max())


def test_mesh_builder():
    for p, nel in [(1, (2, 3, 4)), (3, (2, 2, 2)), (6, (1, 2, 1))]:
        off = M.hex_offsets(*nel, p)
        nn = (nel[0] * p + 1) * (nel[1] * p + 1) * (nel[2] * p + 1)
        assert off.shape == (nel[0] * nel[1] * nel[2], (p + 1) ** 3)
        assert off.min() == 0 and off.max() == nn - 1 and len(np.unique(off)) == nn
        # x fastest inside the element, elements x fastest
        assert off[0, 1] - off[0, 0] == 1 and (off[1, 0] - off[0, 0] == p if nel[0] > 1 else True)
        c = M.hex_coords(*nel, p, perturb=False)
        assert c.shape == (3, nn) and c.min() == 0.0 and abs(c.max() - 1.0) < 1e-15
    assert M.choose_elements(10_000_000, 6) == (36, 36, 35) or abs(np.prod([n * 6 + 1 for n in M.choose_elements(10_000_000, 6)]) - 1e7) < 3e5


def test_partition_interfaces_are_consistent():
    """Every pair of neighbouring ranks lists the same shared nodes in the same (global) order; owned nodes tile the mesh."""
    p, ng = 2, (4, 3, 2)
    for size in (1, 2, 4, 8):
        parts = [M.Partition(ng, p, size, r) for r in range(size)]
        total_owned = sum(int(pt.owned_mask().sum()) for pt in parts)
        assert total_owned == (ng[0] * p + 1) * (ng[1] * p + 1) * (ng[2] * p + 1)
        assert sum(np.prod(pt.n_local) for pt in parts) == np.prod(ng)
        for pt in parts:
            gid = pt.global_node_ids()
            for nbr, idx in pt.neighbors:
                other = parts[nbr]
                back = [i for (r, i) in other.neighbors if r == pt.rank]
                assert len(back) == 1
                assert np.array_equal(gid[idx], other.global_node_ids()[back[0]])


COMPILE_ONLY = r"""
import json, os, sys
sys.path.insert(0, %r)
from libceed_b200 import Ceed
from libceed_b200.bp import BPProblem
out = {}
for bp, p in [(1, 3), (3, 1), (3, 6), (5, 7), (6, 4), (2, 2)]:
    ceed = Ceed()
    prob = BPProblem(ceed, bp, p, (2, 2, 2), build_qdata=False)
    info = prob.op.kernel_info()
    src = prob.op.kernel_source()
    out[f"{bp}_{p}"] = dict(info=info, fused=prob.op.is_fused, setup_fused=prob.op_setup.is_fused, has_qf="BP" in src, nsrc=len(src))
    try:
        import numpy as np
        prob.u.set_array(np.ones(prob.num_dofs))
        prob.qdata.set_array(np.ones(len(prob.qdata)))
        prob.op.apply(prob.u, prob.v)
        out[f"{bp}_{p}"]["ran"] = True
    except Exception as e:
        out[f"{bp}_{p}"]["ran"] = False
        out[f"{bp}_{p}"]["err"] = str(e)
print("RESULT" + json.dumps(out))
""" % ROOT


def test_fused_kernels_generate_and_compile_for_sm100a_without_gpu():
    env = dict(os.environ, CEED_B200_COMPILE_ONLY="1")
    r = subprocess.run([sys.executable, "-c", COMPILE_ONLY], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    for key, val in res.items():
        assert val["fused"] and val["setup_fused"] and val["has_qf"], (key, val)
        assert val["info"]["smem_bytes"] <= 227 * 1024 and val["info"]["threads"] % 32 == 0
        # compile-only mode never computes anything: applying must fail loudly, not fall back to the CPU
        assert val["ran"] is False and "COMPILE_ONLY" in val["err"], (key, val)


TUNE_LOOKUP = r"""
import json, os, sys
sys.path.insert(0, %r)
from libceed_b200 import Ceed
from libceed_b200.bp import BPProblem
out = {}
for key, (bp, p, scatter) in {"exact": (3, 6, 0), "reduced": (4, 2, 0), "other_scatter": (4, 2, 1), "shapes": (3, 4, 0)}.items():
    c = Ceed(); c.set_scatter_mode(scatter)
    pr = BPProblem(c, bp, p, (3, 3, 3), build_qdata=False)
    if key == "shapes":
        res = {}
        for name, shape in {"groups": dict(group_warps=2, cta_warps=4), "pairs": dict(qf_mode=2, qf_unroll=2), "ring": dict(stage_mask=17),
                            "ordered": None}.items():
            if shape is None:
                c.set_scatter_mode(3); pr = BPProblem(c, bp, p, (3, 3, 3), build_qdata=False)
            else:
                pr.op.set_kernel_shape(**shape)
            src = pr.op.kernel_source()
            res[name] = dict(shape=pr.op.get_kernel_shape(), bar="bar.sync" in src, vla2="#define CEED_Q_VLA 2" in src, ring="b200_ring_prologue" in src,
                             ordered="b200_ordered_complete" in src and "st.release.gpu" in src)
        out[key] = res
    else:
        pr.op.kernel_source()
        out[key] = pr.op.get_kernel_shape()
print("RESULT" + json.dumps(out))
""" % ROOT


def test_tuning_table_lookup_and_kernel_shapes_without_gpu(tmp_path):
    """Kernel shapes are data: exact table hit, reduced-signature hit (other QFunction / other number of quadrature-data
    components), other scatter mode falling back to the default mode's entry; every shape variant generates and compiles."""
    tune = tmp_path / "extra.tune"
    tune.write_text("# test entry: 3-component diffusion, P=3, Q=4, 5 streamed components, QFunction Foo\n"
                    "Q4|sc0|in:P3n3goN5|out:P3n3go|Foo 2 2 2 3 1 2 0\n")
    env = dict(os.environ, CEED_B200_COMPILE_ONLY="1", CEED_B200_TUNE_FILE=str(tune))
    r = subprocess.run([sys.executable, "-c", TUNE_LOOKUP], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    shipped = [ln.split() for ln in open(os.path.join(ROOT, "libceed_b200", "tuned", "sm_100a.tune")) if ln.startswith("Q8|sc0|in:P7n1goN7|out:P7n1go|BPDiff")][0]
    keys = ["elems_per_group", "group_warps", "cta_warps", "min_blocks_per_sm", "qf_mode", "qf_unroll", "stage_mask"]
    assert [res["exact"][k] for k in keys] == [int(x) for x in shipped[1:8]]
    for which in ("reduced", "other_scatter"):
        got = res[which]
        assert (got["elems_per_group"], got["group_warps"], got["qf_mode"], got["qf_unroll"], got["stage_mask"]) == (2, 2, 1, 2, 0), (which, got)
    s = res["shapes"]
    assert s["groups"]["bar"] and s["groups"]["shape"]["group_warps"] == 2
    assert s["pairs"]["vla2"] and s["pairs"]["shape"]["qf_mode"] == 2
    assert s["ring"]["ring"] and s["ordered"]["ordered"]


SCATTER_TABLES = r"""
import ctypes as C, json, sys
import numpy as np
sys.path.insert(0, %r)
from libceed_b200 import Ceed, mesh as M

def check(nel, p, group_elems, scramble):
    ceed = Ceed()
    off = M.hex_offsets(*nel, p).astype(np.int32)
    if scramble:  # an "unstructured" element order: the tables must not rely on lexicographic numbering
        off = off[np.random.default_rng(3).permutation(off.shape[0])]
    num_elem, es = off.shape
    nnodes = int(off.max()) + 1
    r = ceed.ElemRestriction(num_elem, es, 1, 1, nnodes, off.reshape(-1))
    flat = off.reshape(-1).astype(np.int64)
    lib = ceed._lib
    # touch lists per L-node in ascending E-order = the order the serial reference adds in
    order = np.argsort(flat, kind="stable")
    nodes, starts, counts = np.unique(flat[order], return_index=True, return_counts=True)
    out = {}
    # ---- two-pass deterministic tables
    tgt = np.zeros(flat.size, dtype=np.int32); cnt = np.zeros(5, dtype=np.int64)
    assert lib.ceedb200_restriction_debug_scatter_tables(r._ptr, 0, 1, tgt.ctypes.data, cnt.ctypes.data, None, None, 0) == 0
    slot_seen = np.zeros(int(cnt[1]), dtype=bool)
    next_slot = 0
    for node, s, c in zip(nodes, starts, counts):
        entries = order[s:s + c]
        assert tgt[entries[0]] == node                      # the first E-entry owns the store
        for e in entries[1:]:                                # the others: consecutive halo slots, ascending E-order
            slot = ~int(tgt[e]); assert slot == next_slot and not slot_seen[slot]
            slot_seen[slot] = True; next_slot += 1
    assert next_slot == cnt[1] == int((counts - 1).sum()) and cnt[0] == int((counts > 1).sum())
    out["det"] = [int(x) for x in cnt[:2]]
    # ---- ordered (in-kernel completion) tables
    ngroups = (num_elem + group_elems - 1) // group_elems
    pred_ptr = np.zeros(ngroups + 1, dtype=np.int32); pred_idx = np.zeros(64 * ngroups + 64, dtype=np.int32)
    assert lib.ceedb200_restriction_debug_scatter_tables(r._ptr, 3, group_elems, tgt.ctypes.data, cnt.ctypes.data, pred_ptr.ctypes.data,
                                                         pred_idx.ctypes.data, pred_idx.size) == 0
    assert cnt[4] == 1 and cnt[2] == ngroups
    expect_pred = [set() for _ in range(ngroups)]
    slot = 0
    for node, s, c in zip(nodes, starts, counts):
        entries = order[s:s + c]
        if c == 1:
            assert tgt[entries[0]] == node; continue
        slot += 1                                            # id slot of the node
        g_last = int(entries[-1]) // (group_elems * es)
        for k, e in enumerate(entries):
            x = ~int(tgt[e]); assert x >= 0
            assert (x & 0x7ffffff) == slot, (node, k); slot += 1
            is_last = k == c - 1
            assert ((x >> 27) & 1) == int(is_last)
            if is_last: assert ((x >> 28) & 7) + 2 == c
            else:
                g = int(e) // (group_elems * es)
                if g != g_last: expect_pred[g_last].add(g)
    assert slot == cnt[1] == int((counts[counts > 1] + 1).sum())
    for g in range(ngroups):
        got = sorted(int(x) for x in pred_idx[pred_ptr[g]:pred_ptr[g + 1]])
        assert got == sorted(expect_pred[g]) and all(q < g for q in got), g   # waits only ever point to earlier groups
    out["ordered"] = [int(x) for x in cnt[:4]]
    # ---- owner/halo tables with K element parts (streamed host-buffer apply / multi-GPU split): shared nodes ordered by the part of their
    # last toucher, halo slots consecutive in that node order and ascending E-order inside a node
    out["parts"] = []
    for K in (1, 2, 3, 5):
        prefix = np.zeros(K + 2, dtype=np.int32); hnode = np.zeros(flat.size, dtype=np.int32)
        assert lib.ceedb200_restriction_debug_scatter_tables(r._ptr, 5, K, tgt.ctypes.data, cnt.ctypes.data, prefix.ctypes.data, hnode.ctypes.data, hnode.size) == 0
        assert cnt[2] == K + 1 and prefix[0] == 0 and prefix[K] == cnt[0]
        ends = [num_elem * c // K for c in range(1, K + 1)]
        last_part = {}
        for node, s, c in zip(nodes, starts, counts):
            entries = order[s:s + c]
            assert tgt[entries[0]] == node
            if c > 1:
                e_last = int(entries[-1]) // es
                last_part[int(node)] = min(K - 1, int(np.searchsorted(ends, e_last, side="right")))
        slot = 0
        for k in range(K):
            seg = [int(x) for x in hnode[prefix[k]:prefix[k + 1]]]
            assert seg == sorted(seg) and all(last_part[x] == k for x in seg), (K, k)
            for node in seg:
                i = int(np.searchsorted(nodes, node)); entries = order[starts[i]:starts[i] + counts[i]]
                for e in entries[1:]:
                    assert ~int(tgt[e]) == slot; slot += 1
        assert slot == cnt[1] and len(last_part) == cnt[0]
        out["parts"].append([int(x) for x in prefix[:K + 1]])
    # ---- run scatter tables (B200RunScatter): simulate the warps walking their runs and check that every direct entry (store / read-modify-
    # write) finds exactly the ascending-E prefix of its node already in v, and that the halo entries are the remaining suffix in order
    RMW = 1 << 30
    out["runs"] = []
    for num_groups, E in ((1, 2), (3, 1), (4, 3), (7, 2)):
        assert lib.ceedb200_restriction_debug_scatter_tables(r._ptr, 4, E, tgt.ctypes.data, cnt.ctypes.data, None, None, num_groups) == 0
        applied = {}                                          # node -> list of E-entries already added into v, in order of arrival
        for w in range(num_groups):                           # groups are independent: a direct entry only ever follows entries of its own group
            s0, s1 = w * num_elem // num_groups, (w + 1) * num_elem // num_groups
            nb = (s1 - s0 + E - 1) // E if s1 > s0 else 0
            for it in range(nb):
                batch = [s0 + it + k * nb for k in range(E) if s0 + it + k * nb < s1]
                touched = set()
                for e in batch:
                    for n in range(es):
                        g = int(tgt[e * es + n])
                        if g < 0: continue
                        node = g & (RMW - 1)
                        assert node == flat[e * es + n]
                        assert node not in touched, "two direct entries of one node in the same iteration would race"
                        touched.add(node)
                        if g & RMW: assert node in applied
                        else: assert node not in applied
                        applied.setdefault(node, []).append(e * es + n)
        n_rmw = 0; next_slot = 0
        for node, s, c in zip(nodes, starts, counts):
            entries = [int(x) for x in order[s:s + c]]
            direct = applied[int(node)]
            assert direct == entries[:len(direct)]             # ascending E-order prefix, owner first
            n_rmw += len(direct) - 1
            for e in entries[len(direct):]:                    # halo suffix: consecutive slots in ascending E-order
                assert ~int(tgt[e]) == next_slot; next_slot += 1
        assert next_slot == cnt[1] and n_rmw == cnt[3] and cnt[2] == num_groups
        out["runs"].append([int(cnt[1]), int(cnt[3])])
    return out

res = {"%%dx%%dx%%d p%%d g%%d s%%d" %% (*nel, p, ge, sc): check(nel, p, ge, sc)
       for nel, p, ge, sc in [((3, 2, 2), 2, 1, 0), ((4, 3, 2), 1, 5, 0), ((2, 2, 3), 3, 2, 1), ((5, 1, 1), 4, 3, 1)]}
print("RESULT" + json.dumps(res))
""" % ROOT


def test_scatter_tables_match_serial_order_model_without_gpu():
    """Owner/halo tables (two-pass deterministic scatter) and ordered tables (in-kernel completion): slots, owners, completing
    entries and predecessor groups against a numpy model of 'add in ascending E-order', structured and scrambled element order."""
    env = dict(os.environ, CEED_B200_COMPILE_ONLY="1")
    r = subprocess.run([sys.executable, "-c", SCATTER_TABLES], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    assert len(res) == 4 and all(v["det"][1] > 0 and v["ordered"][2] > 0 for v in res.values())
    assert all(len(v["parts"]) == 4 and v["parts"][0][-1] == v["det"][0] for v in res.values())
    for v in res.values():  # one group walking everything one element at a time needs no halo at all; more groups need more
        assert len(v["runs"]) == 4 and all(h + m == v["det"][1] for h, m in v["runs"]) and v["runs"][1][0] >= 0


LEAN_SHAPES = r"""
import json, sys
sys.path.insert(0, %r)
from libceed_b200 import Ceed
from libceed_b200.bp import BPProblem
ceed = Ceed()
out = {}
for bp, p in ((1, 3), (2, 2), (3, 3)):   # (tests/test_kernel_emulation.py also RUNS these kernels)
    prob = BPProblem(ceed, bp, p, (3, 3, 2), build_qdata=False)
    for E, warps, stage in ((6, 4, 0), (3, 2, 40), (4, 4, 8), (5, 1, 96), (8, 4, 1), (6, 4, 7), (5, 2, 12), (4, 2, 38), (4, 2, 35)):
        prob.op.set_kernel_shape(qf_mode=4, elems_per_group=E, cta_warps=warps, group_warps=1, stage_mask=stage)
        src = prob.op.kernel_source()
        got = prob.op.get_kernel_shape()
        out["bp%%d p%%d E%%d s%%d" %% (bp, p, E, stage)] = dict(layout=got["qf_mode"], stage=got["stage_mask"], lean="b200_lean_z" in src, bulk="b200_lean_copy" in src,
                                                               prefetch="b200_lean_prefetch_qd" in src, smem=prob.op.kernel_info()["smem_bytes"],
                                                               interleaved=("/ %%d) %%%% %%d, ij" %% (p + 1, min(E, 18))) in src, vector_qd="const double2 *const wp" in src)
print("RESULT" + json.dumps(out))
""" % ROOT


STREAM_PLAN = r"""
import ctypes as C, json, sys
import numpy as np
sys.path.insert(0, %r)
from libceed_b200 import Ceed, mesh as M
from libceed_b200.bp import BPProblem
ceed = Ceed(); lib = ceed._lib
out = {}
for name, bp, p, nel, kw in (("bp1", 1, 2, (6, 5, 7), {}), ("bp2 blocked", 2, 1, (5, 5, 6), {}), ("bp4 interlaced", 4, 1, (4, 5, 6), dict(interlaced=True)),
                            ("bp3 morton", 3, 2, (6, 6, 6), dict(morton=True))):
    morton = kw.pop("morton", False)
    prob = BPProblem(ceed, bp, p, nel, build_qdata=False, elem_perm=M.morton_permutation(*nel) if morton else None, **kw)
    off = prob.offsets.astype(np.int64) * (prob.ncomp if kw.get("interlaced") else 1)
    for K in (2, 5, 8):
        ends = np.zeros(K, dtype=np.int32); in_hi = np.zeros(K, dtype=np.int64); out_done = np.zeros(K, dtype=np.int64); pc = C.c_int()
        assert lib.ceedb200_operator_debug_stream_plan(prob.op._ptr, K, ends.ctypes.data, in_hi.ctypes.data, out_done.ctypes.data, C.byref(pc)) == 0
        assert ends[-1] == prob.num_elem and np.all(np.diff(ends) >= 0)
        lo = 0
        for c in range(K):
            chunk = off[lo:ends[c]]
            # everything chunk c gathers lies below in_hi[c]; nothing a later chunk touches lies below out_done[c]
            assert chunk.size == 0 or chunk.max() < in_hi[c]
            later = off[ends[c]:]
            assert later.size == 0 or later.min() >= out_done[c] or c == K - 1
            if c < K - 1 and later.size: assert out_done[c] == later.min()
            lo = ends[c]
        assert np.all(np.diff(in_hi) >= 0) and np.all(np.diff(out_done[:-1]) >= 0)
        out["%%s K%%d" %% (name, K)] = dict(per_comp=pc.value, first_in=int(in_hi[0]), nodes=int(prob.num_nodes))
print("RESULT" + json.dumps(out))
""" % ROOT


def test_streamed_apply_chunk_tables_without_gpu():
    """Chunk tables of ceedb200_operator_apply_streamed against a numpy restatement: a chunk only gathers what has arrived, only final entries
    of v are copied back; blocked components stream per component, a Morton element order needs (almost) the whole input up front."""
    env = dict(os.environ, CEED_B200_COMPILE_ONLY="1")
    r = subprocess.run([sys.executable, "-c", STREAM_PLAN], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    assert res["bp2 blocked K5"]["per_comp"] == 1 and res["bp1 K5"]["per_comp"] == 0 and res["bp4 interlaced K5"]["per_comp"] == 0
    lex, mor = res["bp1 K8"]["first_in"] / res["bp1 K8"]["nodes"], res["bp3 morton K8"]["first_in"] / res["bp3 morton K8"]["nodes"]
    assert lex < 0.3 and mor > 1.3 * lex, (lex, mor)  # lexicographic order streams; a space-filling order touches far-away nodes early


def test_lean_kernel_generates_and_compiles_without_gpu():
    """The lean in-place-plane kernel (layout 4) is generated and NVRTC-compiled for sm_100a for BP1 / BP2 shapes incl. the bulk pipelines;
    operators with gradients keep their own layout."""
    env = dict(os.environ, CEED_B200_COMPILE_ONLY="1", CEED_B200_NO_TUNE_TABLE="1")
    r = subprocess.run([sys.executable, "-c", LEAN_SHAPES], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    for key, v in res.items():
        if key.startswith("bp3"):
            assert v["layout"] != 4 and not v["lean"], (key, v)
            continue
        assert v["layout"] == 4 and v["lean"], (key, v)
        want = int(key.rsplit("s", 1)[1]) & (1 | 2 | 4 | 8 | 32 | 64)
        assert v["stage"] == want and v["bulk"] == bool(want & 40) and v["prefetch"] == bool((want & 64) and not (want & 32)), (key, v)
        # bit 1: (i, element, j) column decode; bit 4: 16-byte window loads of the directly loaded quadrature data (not with the bulk copy, bit 32)
        assert v["interleaved"] == bool(want & 1) and v["vector_qd"] == bool((want & 4) and not (want & 32)), (key, v)


VECTOR_STATE = r"""
import ctypes as C, json, sys
import numpy as np
sys.path.insert(0, %r)
from libceed_b200 import Ceed, ceed as cm
ceed = Ceed(); lib = ceed._lib
out = {}
def flags(v):
    a, bh, bd = C.c_int(), C.c_int(), C.c_int()
    lib.ceedb200_vector_has_valid_array(v._ptr, C.byref(a))
    lib.ceedb200_vector_has_borrowed_array_of_type(v._ptr, cm.MEM_HOST, C.byref(bh))
    lib.ceedb200_vector_has_borrowed_array_of_type(v._ptr, cm.MEM_DEVICE, C.byref(bd))
    return [a.value, bh.value, bd.value]
def dev_as_numpy(v, n):   # compile-only mode: "device" allocations are host memory, so the mirror can be inspected
    return np.ctypeslib.as_array(C.cast(v.get_array_read(cm.MEM_DEVICE), C.POINTER(C.c_double)), shape=(n,)).copy()
n = 11
v = ceed.Vector(n)
out["fresh"] = flags(v)                                   # no valid data yet
a = np.arange(n, dtype=np.float64)
v.set_array(a, cm.MEM_HOST, cm.COPY_VALUES)
out["after_copy"] = flags(v)
out["mirror"] = dev_as_numpy(v, n).tolist()               # lazy host -> device sync on first device access
b = np.linspace(-1, 1, n)
v.set_array(b, cm.MEM_HOST, cm.USE_POINTER)
out["after_borrow"] = flags(v)
out["borrowed_read"] = v.get_array_read().tolist()
out["borrowed_mirror"] = dev_as_numpy(v, n).tolist()
v.take_array(cm.MEM_HOST)
out["after_take"] = flags(v)                              # the device mirror is still valid after the host array was taken
out["after_take_read"] = v.get_array_read().tolist()
w = ceed.Vector(n)
code = lib.ceedb200_vector_sync_array(w._ptr, cm.MEM_HOST)  # nothing valid to sync: an error, not silence
out["sync_empty_code"] = code
code = lib.ceedb200_vector_set_value(v._ptr, C.c_double(1.0))  # a kernel: must fail loudly in compile-only mode
out["kernel_code"] = code; out["kernel_msg"] = lib.ceedb200_last_error(ceed._ptr).decode()
print("RESULT" + json.dumps(out))
""" % ROOT


def test_vector_mirrors_and_state_flags_without_gpu():
    """CeedVector semantics of the core (owned / borrowed host and device mirrors, lazy sync, take) exercised in compile-only mode,
    where device allocations are host memory and any kernel launch fails loudly (backends/cuda-ref/ceed-cuda-ref-vector.c:21-228)."""
    env = dict(os.environ, CEED_B200_COMPILE_ONLY="1")
    r = subprocess.run([sys.executable, "-c", VECTOR_STATE], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][0][6:])
    n = 11
    assert res["fresh"] == [0, 0, 0]
    assert res["after_copy"] == [1, 0, 0] and res["mirror"] == list(map(float, range(n)))
    assert res["after_borrow"] == [1, 1, 0]
    assert np.allclose(res["borrowed_read"], np.linspace(-1, 1, n)) and res["borrowed_mirror"] == res["borrowed_read"]
    assert res["after_take"][0] == 1 and res["after_take"][1] == 0 and res["after_take_read"] == res["borrowed_read"]
    assert res["sync_empty_code"] != 0
    assert res["kernel_code"] != 0 and "COMPILE_ONLY" in res["kernel_msg"]
