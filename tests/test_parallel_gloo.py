"""Multi-rank logic on CPU (gloo, world_size 2 and 4): element partition + interface-DoF sum.

Every rank applies the BP operator on its own element box with the CPU oracle (tests may use oracle/), the interface
contributions are exchanged with libceed_b200.parallel.InterfaceExchange over gloo, and the result is compared with a
single-rank apply on the global mesh.  On the GPU box the same classes run over NCCL (bench.py --gpus N).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libceed_b200 import mesh  # noqa: E402
from libceed_b200.bp import seeded_uniform  # noqa: E402


def _global_problem(bp, p, n_global):
    from oracle import oracle

    off = mesh.hex_offsets(*n_global, p)
    coords = mesh.hex_coords(*n_global, p)
    num_nodes = coords.shape[1]
    _, _, nc, _ = oracle.bp_sizes(bp, False, p)
    u = seeded_uniform(nc * num_nodes)
    qd = oracle.bp_qdata(bp, p, off, coords)
    v = oracle.bp_apply(bp, p, off, num_nodes, qd, u)
    return u, v, num_nodes, nc


class HostKernels:
    """Test-side executor of the exchange tables (the product runs ceedb200_iface_pack / ceedb200_iface_unpack_sum on the GPU;
    tests/test_gpu_parity.py checks those kernels against exactly this interpretation)."""

    def pack(self, v, idx, send):
        send.copy_(v[idx])

    def unpack_sum(self, v, node, ptr, src, recv):
        vn, rn = v.numpy(), recv.numpy()
        node, ptr, src = node.numpy(), ptr.numpy(), src.numpy()
        for i in range(node.size):
            terms = [vn[node[i]] if s < 0 else rn[s] for s in src[ptr[i]:ptr[i + 1]]]
            acc = terms[0]
            for x in terms[1:]:
                acc = acc + x
            vn[node[i]] = acc


def _worker(rank, world, port, bp, p, n_global, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from libceed_b200.parallel import InterfaceExchange

        part = mesh.Partition(n_global, p, world, rank)
        off = mesh.hex_offsets(*part.n_local, p)
        coords = mesh.hex_coords(*part.n_local, p, n_global=n_global, e0=part.e0)
        nloc = part.num_local_nodes
        assert coords.shape[1] == nloc
        _, _, nc, _ = oracle.bp_sizes(bp, False, p)
        u_glob, _, n_glob, _ = _global_problem(bp, p, n_global)
        gid = part.global_node_ids()
        u_loc = np.concatenate([u_glob[gid + c * n_glob] for c in range(nc)])
        qd = oracle.bp_qdata(bp, p, off, coords)
        v_loc = oracle.bp_apply(bp, p, off, nloc, qd, u_loc)
        ex = InterfaceExchange(part, nc, nloc, torch.device("cpu"), kernels=HostKernels())
        v_t = torch.from_numpy(v_loc.copy())
        ex.sum_interfaces(v_t)
        # second exchange on fresh data must give the same bits (deterministic order)
        v_t2 = torch.from_numpy(v_loc.copy())
        ex.sum_interfaces(v_t2)
        assert torch.equal(v_t, v_t2)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), v=v_t.numpy(), gid=gid, owned=part.owned_mask(), nc=nc)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,bp,p,n_global", [(2, 1, 2, (4, 2, 2)), (2, 3, 2, (3, 2, 2)), (4, 3, 1, (4, 4, 2))])
def test_partitioned_apply_matches_single_rank(tmp_path, oracle, world, bp, p, n_global):
    port = 29500 + (os.getpid() + world * 7 + bp) % 2000
    mp.spawn(_worker, args=(world, port, bp, p, n_global, str(tmp_path)), nprocs=world, join=True)
    _, v_glob, n_glob, nc = _global_problem(bp, p, n_global)
    seen = np.zeros(n_glob, dtype=int)
    scale = np.abs(v_glob).max()
    shared_vals = {}
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        v, gid, owned = d["v"], d["gid"], d["owned"]
        nloc = gid.size
        seen[gid[owned]] += 1
        for c in range(nc):
            # EVERY local copy (owned or not) holds the fully summed value
            np.testing.assert_allclose(v[c * nloc:(c + 1) * nloc], v_glob[gid + c * n_glob], rtol=0, atol=1e-12 * scale)
        for g, val in zip(gid[~owned], v[:nloc][~owned]):
            shared_vals.setdefault(int(g), []).append(val)
        for g, val in zip(gid[owned], v[:nloc][owned]):
            shared_vals.setdefault(int(g), []).append(val)
    assert (seen == 1).all(), "every global node has exactly one owner"
    # copies of a shared node carry identical bits on all ranks (fixed rank-ordered sum)
    for g, vals in shared_vals.items():
        assert all(x == vals[0] for x in vals), g


def test_exchange_requires_cuda_library():
    """No silent host fallback in the product path."""
    part = mesh.Partition((2, 1, 1), 1, 2, 0)
    from libceed_b200.parallel import InterfaceExchange
    with pytest.raises(RuntimeError):
        InterfaceExchange(part, 1, part.num_local_nodes, torch.device("cpu"))


def test_partition_tables_consistent():
    """Neighbour lists of both sides of an interface name the same global nodes in the same order."""
    for world, n_global, p in [(2, (4, 3, 2), 2), (4, (4, 4, 3), 1), (8, (4, 4, 4), 3)]:
        parts = [mesh.Partition(n_global, p, world, r) for r in range(world)]
        total_owned = 0
        for a in parts:
            total_owned += int(a.owned_mask().sum())
            ga = a.global_node_ids()
            for rank_b, idx_a in a.neighbors:
                b = parts[rank_b]
                idx_b = dict(b.neighbors)[a.rank]
                assert np.array_equal(ga[idx_a], b.global_node_ids()[idx_b])
        gx, gy, gz = (n * p + 1 for n in n_global)
        assert total_owned == gx * gy * gz


def test_partition_and_exchange_tables_random_grids():
    """Property test over random process counts / mesh sizes / degrees: ownership is a partition of the global nodes, both sides
    of every interface agree, and the exchange tables of all ranks reproduce the global sum when executed on the host."""
    from hypothesis import given, settings, strategies as st
    from libceed_b200.parallel import build_interface_tables

    @settings(max_examples=25, deadline=None)
    @given(world=st.sampled_from([1, 2, 3, 4, 6, 8, 12]), p=st.integers(1, 3), ex=st.integers(0, 2), ey=st.integers(0, 2), ez=st.integers(0, 1),
           ncomp=st.integers(1, 2), seed=st.integers(0, 1000))
    def run(world, p, ex, ey, ez, ncomp, seed):
        grid = mesh.split3(world)
        n_global = (grid[0] + ex, grid[1] + ey, grid[2] + ez)   # at least one element per rank in every direction
        parts = [mesh.Partition(n_global, p, world, r) for r in range(world)]
        n_glob = int(np.prod([n * p + 1 for n in n_global]))
        owners = np.zeros(n_glob, dtype=int)
        rng = np.random.default_rng(seed)
        # every rank holds a random partial value for each of its local (node, comp) entries
        partial = [rng.uniform(-1, 1, ncomp * pt.num_local_nodes) for pt in parts]
        expect = np.zeros(ncomp * n_glob)
        for pt, val in zip(parts, partial):
            gid, nloc = pt.global_node_ids(), pt.num_local_nodes
            owners[gid[pt.owned_mask()]] += 1
            for c in range(ncomp):
                np.add.at(expect, gid + c * n_glob, val[c * nloc:(c + 1) * nloc])
        assert (owners == 1).all()
        tables = [build_interface_tables(pt, ncomp, pt.num_local_nodes) for pt in parts]
        send = [val[t[2]] if t[2].size else np.zeros(0) for val, t in zip(partial, tables)]
        for r, (pt, t) in enumerate(zip(parts, tables)):
            ranks, seg, send_idx, node, ptr, src = t
            recv = np.zeros(send_idx.size)
            for k, nb in enumerate(ranks):           # what neighbour nb sends to r: its segment addressed to r
                tn = tables[nb]
                kk = tn[0].index(r)
                recv[seg[k]:seg[k + 1]] = send[nb][tn[1][kk]:tn[1][kk + 1]]
            v = partial[r].copy()
            for i in range(node.size):
                terms = [partial[r][node[i]] if s < 0 else recv[s] for s in src[ptr[i]:ptr[i + 1]]]
                v[node[i]] = sum(terms[1:], terms[0])
            gid, nloc = pt.global_node_ids(), pt.num_local_nodes
            for c in range(ncomp):
                np.testing.assert_allclose(v[c * nloc:(c + 1) * nloc], expect[gid + c * n_glob], rtol=0, atol=1e-13)

    run()


def test_boundary_first_permutation_splits_interface_elements():
    """The overlapped multi-GPU step applies the elements that touch a rank interface first: the split must contain exactly
    the elements holding an interface node, in lexicographic order within both classes."""
    for world, n_global, p in [(2, (4, 3, 2), 2), (4, (5, 4, 3), 1), (8, (4, 4, 4), 3), (6, (6, 4, 2), 2)]:
        for r in range(world):
            part = mesh.Partition(n_global, p, world, r)
            perm, split = part.boundary_first_permutation()
            nel = int(np.prod(part.n_local))
            assert sorted(perm.tolist()) == list(range(nel))
            assert (np.diff(perm[:split]) > 0).all() and (np.diff(perm[split:]) > 0).all()
            iface = np.zeros(part.num_local_nodes, dtype=bool)
            for _, idx in part.neighbors:
                iface[idx] = True
            off = mesh.hex_offsets(*part.n_local, p)
            touches = iface[off].any(axis=1)
            assert np.array_equal(np.sort(perm[:split]), np.nonzero(touches)[0])
