"""Multi-rank results ON HARDWARE (pytest -m gpu; needs >= 2 GPUs, skips cleanly otherwise): N ranks (one per GPU, NCCL for the
bootstrap / all-reduce, NVLink peer memory or NCCL send/recv for the interface sum) apply a BP operator on their element boxes of
a global mesh -- serial and overlapped (boundary elements first) variants -- and the summed result is compared with
  * the single-rank apply of the whole mesh on one GPU: BITWISE on nodes that are not on a rank interface (serial variant: same
    element order), <= 1e-12 on interface nodes and for the overlapped variant (elements reordered boundary-first),
  * the unmodified reference /cpu/self/ref/serial when oracle/_ref travelled here,
and all copies of a shared node must carry identical bits on every rank.  A distributed CG (dot products all-reduced) must converge.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from libceed_b200 import mesh as M
from libceed_b200.bp import BP_TABLE, seeded_uniform

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kind_is_mass(bp):
    return BP_TABLE[bp][1] == "mass"


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def run_ranks(world, out_dir, bp, p, n_global, cg_iters=0, transport="auto"):
    port = 29600 + (os.getpid() + 13 * world + bp) % 1500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_gpu_worker.py"), str(out_dir), str(bp), str(p), *[str(n) for n in n_global], str(cg_iters)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, CEED_B200_TEST_TRANSPORT=transport))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("world,bp,p,n_global,cg,transport", [(2, 3, 3, (6, 4, 4), 0, "auto"), (2, 1, 3, (8, 5, 4), 250, "auto"), (2, 3, 2, (5, 4, 3), 40, "nccl"),
                                                              (8, 3, 3, (6, 6, 6), 0, "auto"), (8, 1, 3, (8, 6, 6), 0, "auto"), (4, 6, 2, (6, 6, 3), 0, "auto"),
                                                              (4, 5, 4, (4, 4, 2), 0, "nccl")])
def test_partitioned_apply_on_gpus(tmp_path, world, bp, p, n_global, cg, transport):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs, found {_gpus()}")
    run_ranks(world, tmp_path, bp, p, n_global, cg, transport)
    single = np.load(tmp_path / "single.npz")
    v_glob, n_glob = single["v"], int(single["n_glob"])
    ncomp = BP_TABLE[bp][0]
    scale = np.abs(v_glob).max()
    # interface nodes: shared by at least two ranks
    count = np.zeros(n_glob, dtype=int)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for d in ranks:
        count[d["gid"]] += 1
    copies = {}
    for r, d in enumerate(ranks):
        gid, nloc = d["gid"], d["gid"].size
        iface = count[gid] > 1
        for key in ("v_serial", "v_overlap"):
            v = d[key]
            for c in range(ncomp):
                got, want = v[c * nloc:(c + 1) * nloc], v_glob[gid + c * n_glob]
                assert np.abs(got - want).max() <= 1e-12 * scale, (key, r, c)
                if key == "v_serial":
                    assert np.array_equal(got[~iface], want[~iface]), ("interior nodes must be bitwise equal to the single-rank apply", r, c)
                for g, val in zip(gid[iface] + c * n_glob, got[iface]):
                    copies.setdefault((key, int(g)), []).append(val)
    for (key, g), vals in copies.items():
        assert all(x == vals[0] for x in vals), ("copies of a shared node differ between ranks", key, g)
    # the unmodified reference on the whole mesh
    from oracle import refceed as R
    if R.available():
        rc = R.RefCeed("/cpu/self/ref/serial")
        off, coords = M.hex_offsets(*n_global, p), M.hex_coords(*n_global, p)
        ref = R.RefBP(rc, bp, p, off.shape[0], coords.shape[1], off, coords)
        v_ref = ref.apply(seeded_uniform(ncomp * n_glob))
        assert np.abs(v_glob - v_ref).max() <= 1e-12 * scale
        for d in ranks:
            gid, nloc = d["gid"], d["gid"].size
            for c in range(ncomp):
                assert np.abs(d["v_overlap"][c * nloc:(c + 1) * nloc] - v_ref[gid + c * n_glob]).max() <= 1e-12 * scale
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "multirank_transport.txt"), "a") as f:
        f.write(f"world {world} bp{bp} p{p} mesh {n_global}: interface exchange transport = {str(ranks[0]['transport'])}\n")
    if transport != "auto":
        assert str(ranks[0]["transport"]) == transport
    if cg and kind_is_mass(bp):
        for d in ranks:
            assert d["cg_r"] < 1e-8 * d["cg_r0"], (float(d["cg_r0"]), float(d["cg_r"]))
            assert np.abs(d["cg_x"] - d["u"]).max() < 1e-5 * np.abs(d["u"]).max()
    elif cg:
        for d in ranks:  # diffusion operator (singular without boundary conditions): the residual must still fall, identically on all ranks
            assert d["cg_r"] < 0.5 * d["cg_r0"] and d["cg_r"] == ranks[0]["cg_r"]
