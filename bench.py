#!/usr/bin/env python
"""bench.py -- CeedOperatorApply throughput (GDoF/s) of the /gpu/cuda/b200 path on synthetic structured hex meshes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload bp1p3|bp3p6|bp5p7|...] [--dofs D] [--impl reference]

One "step" = one CeedOperatorApply (v = A u) on the rank's mesh, plus -- for N > 1 -- the interface-DoF sum over NVLink,
overlapped with the interior elements.  Prints ONE JSON line (rank 0).
N = 1: `value` / `e2e` are measured THROUGH libCEED's public C API on the registered /gpu/cuda/b200 plugin (eager launches; the
unmodified reference library of oracle/_ref/lib-cuda hosts the plugin), device-resident resp. with pinned HOST buffers (H2D of u
and D2H of v inside the timed region); `gpu_baseline` is the reference's own /gpu/cuda/gen on the same mesh in the same library;
the C-ABI and CUDA-graph-replay figures are secondary keys.  N > 1: the multi-GPU layer sits above libCEED (as in the reference's
examples), so the step runs through the C ABI (parallel.DistributedOperator); `value` = owned DoFs of all ranks / max-over-ranks
device time, `checksum_*` verify the distributed result against a single-GPU apply of the same global mesh.
`roofline` = algorithmic bytes of one apply / CUDA-event duration of the fused kernel (+ finalize kernel), against the measured
HBM peak in MEASURED_PEAKS.json.  `scaling_case` = a second timed case, BP3 p=6 at 50M DoFs per GPU (BASELINE.json configs[4]).
`cpu_baseline` / `--impl reference` = the unmodified reference's /cpu/self/avx/blocked on the host cores on the same mesh.
"""
import argparse
import json
import multiprocessing as mp
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD_RE = re.compile(r"bp(\d)p(\d)")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="bp1p3", help="bp<B>p<P>, e.g. bp1p3 (BASELINE configs[1]), bp3p6, bp5p7")
    ap.add_argument("--dofs", type=float, default=10e6, help="DoFs per GPU (weak scaling)")
    ap.add_argument("--scatter", default="deterministic", choices=["deterministic", "atomic", "evector"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="skip the software-pipelined end-to-end measurement")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-sweep", action="store_true", help="skip the `sweep` key (kernel numbers of BP1 p=2..5, BP2 p=3, BP3 p=1..8, BP5 p=4..7, BP6 p=4,6)")
    ap.add_argument("--no-libceed", action="store_true", help="do not measure through libCEED's public API (oracle/_ref/lib-cuda + plugin)")
    ap.add_argument("--no-scaling-case", action="store_true", help="skip the second timed case (BP3 p=6, 50M DoFs per GPU: BASELINE configs[4])")
    ap.add_argument("--scaling-dofs", type=float, default=50e6, help="DoFs per GPU of the second timed case")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: apply, then exchange (no boundary-first overlap)")
    ap.add_argument("--timeline", action="store_true", help="N > 1: add per-rank phase timestamps of the overlapped step (`timeline_ms`)")
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "nccl"], help="N > 1: interface exchange over NVLink peer memory or NCCL send/recv")
    return ap.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
    bp, p, nel, resource, seconds, steps, warmup, barrier_dir, wid, use_ref = args
    from libceed_b200 import mesh as M
    from libceed_b200.bp import seeded_uniform
    off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
    nn = coords.shape[1]
    if use_ref:
        from oracle import refceed as R
        rc = R.RefCeed(resource)
        prob = R.RefBP(rc, bp, p, off.shape[0], nn, off, coords)
        ncomp = prob.ncomp
        rc.set_array(prob.u, seeded_uniform(ncomp * nn))
        apply = lambda: rc.op_apply(prob.op, prob.u, prob.v)
    else:
        from oracle import oracle as O
        _, _, ncomp, _ = O.bp_sizes(bp, False, p)
        qd = O.bp_qdata(bp, p, off, coords)
        u = seeded_uniform(ncomp * nn)
        apply = lambda: O.bp_apply(bp, p, off, nn, qd, u)
    for _ in range(max(1, warmup)):
        apply()
    # crude barrier through the file system so that all workers time the same interval
    open(os.path.join(barrier_dir, f"ready{wid}"), "w").close()
    while not os.path.exists(os.path.join(barrier_dir, "go")):
        time.sleep(0.001)
    t0 = time.perf_counter()
    n = 0
    while True:
        apply()
        n += 1
        dt = time.perf_counter() - t0
        if n >= max(1, steps) and dt >= seconds:   # at least `steps` applies AND at least `seconds` of work
            break
    return ncomp * nn, n, dt


def cpu_reference(bp, p, dofs, seconds, steps=0, warmup=1, resource="/cpu/self/avx/blocked", max_cores=0):
    """Reference CPU backend on all host cores, on the SAME mesh as the GPU arm: the `dofs`-sized structured hex mesh is cut into
    `cores` z-slabs, one forked single-threaded worker per slab (libCEED CPU backends are single-threaded per Ceed; SURVEY.md
    section 8(d)), every worker applies its operator repeatedly for at least `seconds`; aggregate = sum(DoFs * applies) / max(time)."""
    from libceed_b200 import mesh as M
    from libceed_b200.bp import BP_TABLE
    from oracle import refceed as R
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    if max_cores:
        cores = min(cores, max_cores)
    nx, ny, nz = M.choose_elements(dofs, p, BP_TABLE[bp][0])
    cores = max(1, min(cores, nz))
    slabs = [(nx, ny, M.block_range(nz, cores, w)[1] - M.block_range(nz, cores, w)[0]) for w in range(cores)]
    # the unmodified reference (oracle/_ref) when it travelled with the repository, else the C restatement of its ref backend
    have_ref = R.available() and not os.environ.get("CEED_B200_BENCH_FORCE_PORT")
    kind = "reference" if have_ref else "port"
    if not have_ref:
        resource = "oracle/ceed_oracle.c (restatement of /cpu/self/ref/serial)"
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            res = pool.map_async(_cpu_worker, [(bp, p, slabs[w], resource, seconds, steps, warmup, d, w, have_ref) for w in range(cores)])
            t_start = time.time()
            while len([f for f in os.listdir(d) if f.startswith("ready")]) < cores and time.time() - t_start < 900:
                time.sleep(0.01)
            open(os.path.join(d, "go"), "w").close()
            out = res.get()
    dofs_applied = sum(o[0] * o[1] for o in out)
    tmax = max(o[2] for o in out)
    applies = min(o[1] for o in out)
    mesh_dofs = sum(o[0] for o in out)
    return dict(value=dofs_applied / tmax / 1e9, unit="GDoF/s", cores=cores, kind=kind,
                sample=f"{resource}: the {nx}x{ny}x{nz}-element mesh ({mesh_dofs} DoFs incl. duplicated slab faces) cut into {cores} z-slabs, one forked "
                       f"worker each, >= {applies} applies per worker in {tmax:.3f} s",
                ms_per_step=mesh_dofs / (dofs_applied / tmax) * 1e3, applies=applies, seconds=tmax, mesh_dofs=mesh_dofs)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if not self.proc:
            return out
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ main
# ------------------------------------------------------------------------------------------------ through libCEED (product boundary)
def libceed_available():
    from oracle import refceed as R
    return R.available(cuda=True) and os.path.exists(os.path.join(ROOT, "libceed_b200", "lib", "libceed_b200_backend.so"))


_LIBCEED_LOADED = False


def _load_libceed_with_plugin():
    """One libceed.so per process: the UNMODIFIED reference built with its CUDA backends (oracle/_ref/lib-cuda) + the b200 plugin."""
    global _LIBCEED_LOADED
    import ctypes as C
    from oracle import refceed as R
    if not _LIBCEED_LOADED:
        lib = C.CDLL(os.path.join(R.REF_DIR, "lib-cuda", "libceed.so"), mode=C.RTLD_GLOBAL)
        R.RefCeed._libs["lib-cuda"] = R.RefCeed._libs["lib"] = lib
        C.CDLL(os.path.join(ROOT, "libceed_b200", "lib", "libceed_b200_backend.so"), mode=C.RTLD_GLOBAL)  # constructor registers /gpu/cuda/b200
        _LIBCEED_LOADED = True
    return R


def through_libceed(resource, bp, p, nel, steps, warmup, host_buffers=False, streamed=True):
    """GDoF/s of CeedOperatorApply(op, u, v, CEED_REQUEST_IMMEDIATE) called through libCEED's public C API on `resource`
    (eager launches, the call every C / Fortran / Python / Julia / Rust user makes).  Device-resident vectors, or -- host_buffers --
    the caller's pinned HOST arrays every step: CeedVectorSetArray(u, HOST, USE_POINTER), apply, CeedVectorSyncArray(v, HOST),
    i.e. the H2D copy of u and the D2H copy of v are inside the timed region (streamed=False: CEED_B200_NO_STREAMED, the plugin then
    copies, applies and copies back one after the other like the reference's backends).  Timed with CUDA events on the default stream
    (the stream libCEED's CUDA backends launch on), synchronised on both sides."""
    import ctypes as C
    import torch
    from libceed_b200 import mesh as M
    from libceed_b200.bp import seeded_uniform
    R = _load_libceed_with_plugin()
    off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
    nn = coords.shape[1]
    rc = R.RefCeed(resource, cuda=True)
    prob = R.RefBP(rc, bp, p, off.shape[0], nn, off, coords)
    n = prob.ncomp * nn
    u_host = torch.from_numpy(seeded_uniform(n)).pin_memory()
    v_host = torch.empty(n, dtype=torch.float64).pin_memory()
    lib = rc.lib
    if streamed:
        os.environ.pop("CEED_B200_NO_STREAMED", None)
    else:
        os.environ["CEED_B200_NO_STREAMED"] = "1"
    if host_buffers:
        def step():
            rc._chk(lib.CeedVectorSetArray(prob.u, R.MEM_HOST, R.USE_POINTER, C.c_void_p(u_host.data_ptr())))
            rc._chk(lib.CeedVectorSetArray(prob.v, R.MEM_HOST, R.USE_POINTER, C.c_void_p(v_host.data_ptr())))
            rc.op_apply(prob.op, prob.u, prob.v)
            rc._chk(lib.CeedVectorSyncArray(prob.v, R.MEM_HOST))
    else:
        rc.set_array(prob.u, u_host.numpy())

        def step():
            rc.op_apply(prob.op, prob.u, prob.v)
    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    ms = ev0.elapsed_time(ev1) / steps
    if host_buffers:
        ms = max(ms, wall * 1e3)  # the copies into / out of host memory are synchronous host calls: wall clock is the honest figure
    v = v_host.numpy().copy() if host_buffers else rc.get_array(prob.v, n)
    os.environ.pop("CEED_B200_NO_STREAMED", None)
    return dict(value=n / ms / 1e6, ms_per_step=ms, wall_ms_per_step=wall * 1e3, dofs=n, checksum=float(np.linalg.norm(v))), v


# ------------------------------------------------------------------------------------------------ one timed case on this rank's GPU(s)
def run_case(args, ceed, cm, bp, p, dofs, rank, world, local_rank, dev, steps, with_graph, with_e2e):
    """Builds the (partitioned) problem and times `steps` steps.  A step = CeedOperatorApply on the rank's element box + -- for
    world > 1 -- the interface-DoF sum, overlapped with the interior elements (parallel.DistributedOperator)."""
    import torch
    import torch.distributed as dist
    from libceed_b200 import mesh as M
    from libceed_b200.bp import BP_TABLE, BPProblem, seeded_uniform
    ncomp = BP_TABLE[bp][0]
    nel = M.choose_elements(dofs, p, ncomp)
    stream = torch.cuda.current_stream()
    out = dict(nel=nel)
    if world > 1:
        from libceed_b200.parallel import DistributedOperator
        grid = M.split3(world)
        n_global = tuple(nel[d] * grid[d] for d in range(3))
        part = M.Partition(n_global, p, world, rank)
        dop = DistributedOperator(ceed, bp, p, part, dev, overlap=not args.no_overlap, transport=args.transport)
        prob, u_dev, v_dev = dop.prob, dop.u_t, dop.v_t
        owned_mask = torch.from_numpy(np.tile(part.owned_mask(), ncomp)).to(dev)
        t = torch.tensor([int(owned_mask.sum().item())], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        total_dofs = int(t.item())
        step = dop.apply
        out.update(transport=dop.transport, overlap=dop.overlap, n_global=n_global)
    else:
        dop, part, owned_mask, n_global = None, None, None, nel
        prob = BPProblem(ceed, bp, p, nel)
        u_dev = torch.zeros(prob.num_dofs, dtype=torch.float64, device=dev)
        v_dev = torch.zeros(prob.num_dofs, dtype=torch.float64, device=dev)
        prob.u.set_array(u_dev, cm.MEM_DEVICE, cm.USE_POINTER)
        prob.v.set_array(v_dev, cm.MEM_DEVICE, cm.USE_POINTER)
        total_dofs = prob.num_dofs
        step = lambda: prob.op.apply(prob.u, prob.v)
    n_local = prob.num_dofs
    # the same global input on every rank count: u(global node) from a fixed seed
    if world > 1:
        n_glob = int(np.prod([n * p + 1 for n in n_global]))
        gid = part.global_node_ids()
        # hashed values instead of a global random array (400M entries at 8 x 50M): reproducible per global node
        def u_of(g):
            x = (g.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(11)
            return x.astype(np.float64) / float(1 << 53) * 2.0 - 1.0
        u_loc = np.concatenate([u_of(gid + c * n_glob) for c in range(ncomp)])
    else:
        n_glob = prob.num_nodes
        g = np.arange(ncomp * n_glob, dtype=np.uint64)
        u_loc = ((g * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0
    u_host = torch.from_numpy(u_loc).pin_memory()
    v_host = torch.empty(n_local, dtype=torch.float64).pin_memory()
    u_dev.copy_(u_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    l0 = ceed.launch_count()
    step()
    launches_per_step = ceed.launch_count() - l0
    barrier()

    def timed(fn, nsteps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(nsteps):
            fn()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / nsteps

    ms_eager = timed(step, steps)
    if dop is not None and args.timeline:
        # per-rank event timeline of the overlapped step, gathered on rank 0
        tl = dop.record_timeline()
        rows = [None] * world
        dist.all_gather_object(rows, tl)
        out["timeline_ms"] = rows
    out.update(ms_per_step=ms_eager, value=total_dofs / ms_eager / 1e6, total_dofs=total_dofs, n_local=n_local, launches_per_step=launches_per_step)
    # verified checksum of the distributed result: ||v||_2 over owned DoFs
    vv = v_dev if owned_mask is None else v_dev * owned_mask
    nrm2 = (vv * vv).sum().reshape(1)
    if world > 1:
        dist.all_reduce(nrm2)
    out["checksum_norm2"] = float(nrm2.sqrt().item())
    # CUDA-graph replay of the same step (secondary figure: no libCEED caller gets this)
    if with_graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ceed.set_stream(torch.cuda.current_stream().cuda_stream)
                step()
            ok = 1
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed: {exc}", file=sys.stderr)
            g, ok = None, 0
        ceed.set_stream(stream.cuda_stream)
        okt = torch.tensor([ok], device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()):
            for _ in range(3):
                g.replay()
            ms_graph = timed(g.replay, steps)
            out.update(graph_replay_ms_per_step=ms_graph, graph_replay_value=total_dofs / ms_graph / 1e6)
        del g
    # kernel-level timing (CUDA events inside the library, around the fused kernel and the finalize pass)
    prob.op.set_timing(True)
    kms = []
    for _ in range(max(5, min(steps, 20))):
        prob.op.apply(prob.u, prob.v)
        kms.append(prob.op.last_kernel_ms())
    prob.op.set_timing(False)
    out.update(fused_ms=float(np.median([k[0] for k in kms])), aux_ms=float(np.median([k[1] for k in kms])), alg_bytes=prob.bytes_per_apply(),
               kernel_info=prob.op.kernel_info())
    if with_e2e:
        # end to end with HOST buffers through the C ABI: H2D(u), the step (incl. the interface sum), D2H(v), every step
        uh, vh = ceed.Vector(n_local), ceed.Vector(n_local)
        e2e_steps = max(3, min(steps, 10))

        def e2e_step():
            u_dev.copy_(u_host, non_blocking=True)
            step()
            v_host.copy_(v_dev, non_blocking=True)
            stream.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        out["e2e_cabi"] = dict(value=total_dofs / e2e_s / 1e9, unit="GDoF/s", h2d_bytes_per_step=8 * n_local, d2h_bytes_per_step=8 * n_local, ms_per_step=e2e_s * 1e3,
                               mode="C ABI, pinned host buffers: H2D(u), step (apply + interface sum), D2H(v) one after the other, per rank; max over ranks")
        del uh, vh
        # ONE step, streamed inside the call (single GPU): ceedb200_operator_apply_streamed cuts the elements into chunks, applies a chunk as
        # soon as its part of u has arrived and copies the finished part of v back while the next chunks run -- same bytes over PCIe
        # inside the timed region, one public call per step, results bitwise equal
        if world == 1:
            try:
                us, vs = ceed.Vector(n_local), ceed.Vector(n_local)
                v_host_s = torch.empty(n_local, dtype=torch.float64).pin_memory()
                used = [False]

                def streamed_step():
                    us.set_array(u_host.numpy(), cm.MEM_HOST, cm.USE_POINTER)
                    vs.set_array(v_host_s.numpy(), cm.MEM_HOST, cm.USE_POINTER)
                    used[0] = prob.op.apply_streamed(us, vs, 0)
                    vs.sync_array(cm.MEM_HOST)
                    us.take_array(), vs.take_array()

                streamed_step(), streamed_step()
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    streamed_step()
                torch.cuda.synchronize()
                st_s = (time.perf_counter() - t0) / e2e_steps
                out["e2e_cabi"].update(streamed_value=total_dofs / st_s / 1e9, streamed_ms_per_step=st_s * 1e3, streamed_path_taken=bool(used[0]),
                                       streamed_matches_serial=bool(torch.equal(v_host_s, v_host)),
                                       streamed_mode="C ABI, one call per step: ceedb200_operator_apply_streamed (chunked H2D / apply / D2H on two copy streams)")
                del us, vs
            except Exception as exc:  # noqa: BLE001
                out["e2e_cabi"]["streamed_error"] = str(exc)[:200]
        # The same end-to-end step software-pipelined over steps (single GPU): every step still copies its own u from pinned host
        # memory and its own v back, but on separate copy streams with double-buffered device vectors, so the H2D of step i+1 and the
        # D2H of step i-1 overlap the kernels of step i (PCIe is full duplex).  Results must be bitwise equal to the serial path.
        if world == 1 and not args.no_e2e_pipeline:
            try:
                nbuf = 2
                s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
                u_d = [torch.empty(n_local, dtype=torch.float64, device=dev) for _ in range(nbuf)]
                v_d = [torch.empty(n_local, dtype=torch.float64, device=dev) for _ in range(nbuf)]
                v_h = [torch.empty(n_local, dtype=torch.float64).pin_memory() for _ in range(nbuf)]
                uvec, vvec = [ceed.Vector(n_local) for _ in range(nbuf)], [ceed.Vector(n_local) for _ in range(nbuf)]
                for k in range(nbuf):
                    uvec[k].set_array(u_d[k], cm.MEM_DEVICE, cm.USE_POINTER)
                    vvec[k].set_array(v_d[k], cm.MEM_DEVICE, cm.USE_POINTER)
                ev_in, ev_done, ev_out = ([torch.cuda.Event() for _ in range(nbuf)] for _ in range(3))

                def pipelined(nsteps):
                    for i in range(nsteps):
                        k = i % nbuf
                        if i >= nbuf:
                            s_in.wait_event(ev_out[k])            # buffer k is free once the D2H of step i - nbuf has finished
                        with torch.cuda.stream(s_in):
                            u_d[k].copy_(u_host, non_blocking=True)
                            ev_in[k].record(s_in)
                        stream.wait_event(ev_in[k])
                        prob.op.apply(uvec[k], vvec[k])           # on the backend's stream (= `stream`)
                        ev_done[k].record(stream)
                        s_out.wait_event(ev_done[k])
                        with torch.cuda.stream(s_out):
                            v_h[k].copy_(v_d[k], non_blocking=True)
                            ev_out[k].record(s_out)
                    s_out.synchronize()
                    stream.synchronize()

                pipelined(2)
                barrier()
                t0 = time.perf_counter()
                pipelined(e2e_steps * 2)
                pipe_s = (time.perf_counter() - t0) / (e2e_steps * 2)
                out["e2e_cabi"].update(pipelined_value=total_dofs / pipe_s / 1e9, pipelined_ms_per_step=pipe_s * 1e3,
                                       pipelined_matches_serial=bool(torch.equal(v_h[0], v_host)),
                                       pipelined_mode="C ABI, double-buffered: H2D of step i+1 and D2H of step i-1 overlap the kernels of step i; every step moves its own u and v")
                del uvec, vvec
            except Exception as exc:  # noqa: BLE001
                out["e2e_cabi"]["pipelined_error"] = str(exc)[:200]
    out["prob"], out["dop"], out["part"], out["u_loc"], out["v_dev"], out["owned_mask"] = prob, dop, part, u_loc, v_dev, owned_mask
    return out


def verify_against_single_gpu(ceed, cm, bp, p, case, dev):
    """N > 1: rank 0 applies the SAME operator on the whole global mesh on its own GPU and compares ||v||_2 (over all DoFs) with the
    distributed checksum (over owned DoFs)."""
    import torch
    from libceed_b200.bp import BP_TABLE, BPProblem
    ncomp = BP_TABLE[bp][0]
    n_global = case["n_global"]
    prob = BPProblem(ceed, bp, p, n_global)
    g = np.arange(ncomp * prob.num_nodes, dtype=np.uint64)
    u = ((g * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0
    u_t = torch.from_numpy(u).to(dev)
    v_t = torch.zeros_like(u_t)
    prob.u.set_array(u_t, cm.MEM_DEVICE, cm.USE_POINTER)
    prob.v.set_array(v_t, cm.MEM_DEVICE, cm.USE_POINTER)
    prob.op.apply(prob.u, prob.v)
    torch.cuda.synchronize()
    ref = float(torch.linalg.vector_norm(v_t).item())
    del prob
    return ref


def main():
    args = parse_args()
    # the contract is ONE JSON line on stdout: route everything libraries print to fd 1 (e.g. the NCCL version banner) to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    m = WORKLOAD_RE.fullmatch(args.workload)
    assert m, "workload must look like bp3p6"
    bp, p = int(m.group(1)), int(m.group(2))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from libceed_b200.bp import BP_TABLE
    ncomp, kind, q_extra = BP_TABLE[bp][0], BP_TABLE[bp][1], BP_TABLE[bp][2]
    workload = f"BP{bp} {'mass' if kind == 'mass' else 'diffusion'} p={p} q={p + q_extra} ncomp={ncomp}, {args.dofs / 1e6:.0f}M DoFs per GPU, structured hex mesh"
    config = dict(workload=workload, bp=bp, p=p, q=p + q_extra, dofs_per_gpu=args.dofs, scatter=args.scatter,
                  l2="inputs (u, v, qdata, offsets) exceed the 126 MB L2; no flush between steps",
                  partition="3-D element blocks, one local L-vector per GPU, interface sum over NVLink overlapped with the interior elements" if world > 1 else "single GPU")

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's own CPU implementation on the host cores, on the SAME mesh (cut into one slab per core), >= 2 s of applies
        r = cpu_reference(bp, p, args.dofs, max(2.0, min(args.cpu_seconds, 20.0)), steps=args.steps, warmup=args.warmup)
        line = dict(metric="CeedOperatorApply throughput", value=r["value"], unit="GDoF/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    impl="reference", config=config, applies_timed=r["applies"], seconds_timed=r["seconds"],
                    cpu_baseline=dict(value=r["value"], unit="GDoF/s", cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    e2e=dict(value=r["value"], unit="GDoF/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line), file=json_out, flush=True)
        return

    import torch
    import torch.distributed as dist
    from libceed_b200 import ceed as cm
    from libceed_b200 import mesh as M
    from libceed_b200.bp import BPProblem, seeded_uniform
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))  # a hang aborts instead of blocking the box
    ceed = cm.Ceed(f"/gpu/cuda/b200:device_id={local_rank}")
    ceed.set_scatter_mode({"deterministic": 0, "atomic": 1, "evector": 2}[args.scatter])
    stream = torch.cuda.current_stream()
    ceed.set_stream(stream.cuda_stream)
    peak, peak_src = hbm_peak()

    # ---- headline case (BASELINE.json configs[1] by default): C ABI, eager launches, device-resident
    sampler = ClockSampler(local_rank) if rank == 0 else None
    case = run_case(args, ceed, cm, bp, p, args.dofs, rank, world, local_rank, dev, args.steps, with_graph=not args.no_graph, with_e2e=True)
    clocks = None
    if sampler:
        t_end = time.time() + 1.0   # keep the GPU busy a little longer so that the sampler sees the loaded state even for sub-ms steps
        while time.time() < t_end:
            case["prob"].op.apply(case["prob"].u, case["prob"].v)  # local work only
        torch.cuda.synchronize()
        clocks = sampler.stop()
    fused_ms, aux_ms, alg_bytes, info = case["fused_ms"], case["aux_ms"], case["alg_bytes"], case["kernel_info"]
    achieved = alg_bytes / ((fused_ms + aux_ms) * 1e-3) / 1e9
    traffic, traffic_src, finalize_traffic = None, None, None
    for tname in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if traffic is None and os.path.exists(tpath) and world == 1 and abs(args.dofs - 10e6) < 1 and args.scatter == "deterministic":
            t = json.load(open(tpath)).get(args.workload)
            if t:
                traffic, traffic_src = t["traffic"], t["source"]
                finalize_traffic = t.get("finalize_traffic")
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, traffic_source=traffic_src, peak_source=peak_src,
                    kernel="b200_operator (+ halo finalize)", fused_kernel_ms=fused_ms, finalize_ms=aux_ms, algorithmic_bytes=alg_bytes,
                    kernel_share_of_step=(fused_ms + aux_ms) / case["ms_per_step"], regs=info["regs"], elems_per_block=info["elems_per_block"],
                    threads=info["threads"], grid=info["grid"], smem_bytes=info["smem_bytes"])
    if traffic is not None and finalize_traffic is not None:
        # what the DRAM actually moves per step (ncu capture of both kernels) against the same peak: the deterministic scatter's tables, halo
        # buffer and second pass are traffic the algorithmic figure does not count
        roofline["finalize_traffic"] = finalize_traffic
        roofline["actual_dram_frac"] = (traffic + finalize_traffic) / ((fused_ms + aux_ms) * 1e-3) / 1e9 / peak
    value, ms_per_step, total_dofs = case["value"], case["ms_per_step"], case["total_dofs"]
    e2e = case["e2e_cabi"]
    config["boundary"] = "C ABI (libceed_b200.so), eager launches"
    extra = dict(cabi_value=case["value"], cabi_ms_per_step=case["ms_per_step"], graph_replay_value=case.get("graph_replay_value"),
                 graph_replay_ms_per_step=case.get("graph_replay_ms_per_step"), checksum_norm2=case["checksum_norm2"])
    if world > 1:
        extra.update(transport=case["transport"], overlap=case["overlap"])
        if "timeline_ms" in case:
            extra["timeline_ms"] = case["timeline_ms"]
        # verified checksum: the same operator on the whole global mesh on ONE GPU (rank 0)
        if rank == 0 and total_dofs <= 110e6:
            try:
                ref = verify_against_single_gpu(ceed, cm, bp, p, case, dev)
                extra["checksum_single_gpu_norm2"] = ref
                extra["checksum_rel_diff"] = abs(ref - case["checksum_norm2"]) / ref
                extra["checksum_ok"] = bool(extra["checksum_rel_diff"] < 1e-12)
            except Exception as exc:  # noqa: BLE001
                extra["checksum_error"] = str(exc)[:200]
        dist.barrier()
    gpu_launches = case["launches_per_step"] * args.steps
    n_local = case["n_local"]
    del case

    # ---- N = 1: the same workload THROUGH libCEED's public API on the registered plugin (the product boundary), the reference's
    # own /gpu/cuda/gen on the same mesh, and the CPU reference
    gpu_baseline, cpu = None, None
    if world == 1 and rank == 0:
        nel = M.choose_elements(args.dofs, p, ncomp)
        if libceed_available() and not args.no_libceed:
            try:
                dev_res, v_b200 = through_libceed("/gpu/cuda/b200", bp, p, nel, args.steps, args.warmup)
                host_res, v_host = through_libceed("/gpu/cuda/b200", bp, p, nel, max(3, min(args.steps, 10)), args.warmup, host_buffers=True)
                serial_res, v_serial = through_libceed("/gpu/cuda/b200", bp, p, nel, max(3, min(args.steps, 10)), args.warmup, host_buffers=True, streamed=False)
                extra["libceed_matches_host_path"] = bool(np.array_equal(v_b200, v_host) and np.array_equal(v_b200, v_serial))
                value, ms_per_step = dev_res["value"], dev_res["ms_per_step"]
                config["boundary"] = "CeedOperatorApply on /gpu/cuda/b200 through libCEED's public C API (oracle/_ref/lib-cuda/libceed.so + plugin), eager launches"
                cabi = e2e
                e2e = dict(value=host_res["value"], unit="GDoF/s", h2d_bytes_per_step=8 * n_local, d2h_bytes_per_step=8 * n_local, ms_per_step=host_res["ms_per_step"],
                           mode="through libCEED, ONE step at a time: CeedVectorSetArray(u, HOST, USE_POINTER) -> CeedOperatorApply -> CeedVectorSyncArray(v, HOST), pinned "
                                "host buffers; inside the call the plugin streams the step (chunked H2D / apply / D2H on two copy streams, ceedb200_operator_apply_streamed), "
                                "results bitwise equal to the device-resident apply",
                           serial_through_libceed=serial_res["value"], serial_ms_per_step=serial_res["ms_per_step"],
                           serial_mode="the same calls with CEED_B200_NO_STREAMED=1: whole-vector H2D, apply, whole-vector D2H one after the other", cabi=cabi)
                extra["libceed_checksum_norm2"] = dev_res["checksum"]
                try:
                    gen_res, v_gen = through_libceed("/gpu/cuda/gen", bp, p, nel, args.steps, args.warmup)
                    gpu_baseline = dict(resource="/gpu/cuda/gen (the reference's fused CUDA backend, same libceed.so, same mesh and input)", value=gen_res["value"],
                                        unit="GDoF/s", ms_per_step=gen_res["ms_per_step"], rel_diff_to_b200=float(np.abs(v_gen - v_b200).max() / np.abs(v_b200).max()),
                                        speedup_b200=dev_res["value"] / gen_res["value"])
                except Exception as exc:  # noqa: BLE001
                    gpu_baseline = dict(resource="/gpu/cuda/gen", value=None, error=str(exc)[:200])
            except Exception as exc:  # noqa: BLE001
                extra["libceed_error"] = str(exc)[:300]
        if not args.no_cpu_baseline:
            try:
                r = cpu_reference(bp, p, args.dofs, args.cpu_seconds)
                cpu = dict(value=r["value"], unit="GDoF/s", cores=r["cores"], kind=r["kind"], sample=r["sample"])
                others = []
                for res_name, mc in (("/cpu/self/opt/blocked", 0), ("/cpu/self/avx/blocked", 1)):
                    try:
                        o = cpu_reference(bp, p, args.dofs if mc == 0 else args.dofs / 16, 3.0, resource=res_name, max_cores=mc)
                        others.append(dict(value=o["value"], unit="GDoF/s", cores=o["cores"], kind=o["kind"], sample=o["sample"]))
                    except Exception as exc:  # noqa: BLE001
                        others.append(dict(value=None, sample=f"{res_name}: unavailable: {exc}"))
                cpu["others"] = others
            except Exception as exc:  # the reference build did not travel: say so instead of inventing a number
                cpu = dict(value=None, unit="GDoF/s", cores=0, kind="reference", sample=f"unavailable: {exc}")

    # ---- second timed case in the same line: the north-star weak-scaling configuration (BASELINE.json configs[4]): BP3 p=6, 50M DoFs / GPU
    scaling_case = None
    if not args.no_scaling_case:
        try:
            sc = run_case(args, ceed, cm, 3, 6, args.scaling_dofs, rank, world, local_rank, dev, max(3, min(args.steps, 10)), with_graph=False, with_e2e=False)
            k = sc["fused_ms"] + sc["aux_ms"]
            scaling_case = dict(workload=f"BP3 diffusion p=6 q=8, {args.scaling_dofs / 1e6:.0f}M DoFs per GPU (weak scaling)", value=sc["value"], unit="GDoF/s",
                                ms_per_step=sc["ms_per_step"], total_dofs=sc["total_dofs"], kernel_ms=k, kernel_share_of_step=k / sc["ms_per_step"],
                                roofline_frac=sc["alg_bytes"] / (k * 1e-3) / 1e9 / peak, checksum_norm2=sc["checksum_norm2"], transport=sc.get("transport"),
                                overlap=sc.get("overlap"), timeline_ms=sc.get("timeline_ms"))
            if world > 1 and rank == 0 and sc["total_dofs"] <= 110e6:
                ref = verify_against_single_gpu(ceed, cm, 3, 6, sc, dev)
                scaling_case["checksum_single_gpu_norm2"] = ref
                scaling_case["checksum_ok"] = bool(abs(ref - sc["checksum_norm2"]) / ref < 1e-12)
            del sc
        except Exception as exc:  # noqa: BLE001
            scaling_case = dict(error=str(exc)[:300])
        if world > 1:
            dist.barrier()

    sweep = None
    if not args.no_sweep and rank == 0 and world == 1:
        sweep = []
        for sbp, ps in ((1, (2, 3, 4, 5)), (2, (3,)), (3, range(1, 9)), (5, range(4, 8)), (6, (4, 6))):
            for sp in ps:
                sc = BP_TABLE[sbp][0]
                sprob = BPProblem(ceed, sbp, sp, M.choose_elements(args.dofs, sp, sc))
                sprob.u.set_array(seeded_uniform(sprob.num_dofs))
                sprob.op.set_timing(True)
                t = []
                for i in range(8):
                    sprob.op.apply(sprob.u, sprob.v)
                    if i >= 3:
                        t.append(sum(sprob.op.last_kernel_ms()))
                ms = float(np.median(t))
                sweep.append(dict(bp=sbp, p=sp, dofs=sprob.num_dofs, ms=ms, gdofs=sprob.num_dofs / ms / 1e6,
                                  frac=sprob.bytes_per_apply() / (ms * 1e-3) / 1e9 / peak))
                del sprob
        # the headline operator with the ATOMIC scatter (the reference's /gpu/cuda/gen semantics: not reproducible run to run), for comparison
        # with the deterministic default: memset + fused kernel, no finalize pass
        try:
            aceed = cm.Ceed(f"/gpu/cuda/b200:device_id={local_rank}")
            aceed.set_scatter_mode(1)
            aprob = BPProblem(aceed, bp, p, M.choose_elements(args.dofs, p, ncomp))
            aprob.u.set_array(seeded_uniform(aprob.num_dofs))
            aprob.op.set_timing(True)
            t = []
            for i in range(8):
                aprob.op.apply(aprob.u, aprob.v)
                if i >= 3:
                    t.append(sum(aprob.op.last_kernel_ms()))
            ms = float(np.median(t))
            extra["atomic_scatter"] = dict(ms=ms, gdofs=aprob.num_dofs / ms / 1e6, frac=aprob.bytes_per_apply() / (ms * 1e-3) / 1e9 / peak,
                                           note="opt-in (CEED_B200_SCATTER=atomic): zero v, accumulate with red.global.add.f64; the default is the deterministic owner/halo scatter")
            del aprob, aceed
        except Exception as exc:  # noqa: BLE001
            extra["atomic_scatter"] = dict(error=str(exc)[:200])

    if rank == 0:
        line = dict(metric="CeedOperatorApply throughput", value=value, unit="GDoF/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=config,
                    roofline=roofline, cpu_baseline=cpu, gpu_baseline=gpu_baseline, e2e=e2e, gpu_launches=int(gpu_launches), clocks=clocks, total_dofs=total_dofs,
                    scaling_case=scaling_case, **extra)
        if sweep:
            line["sweep"] = sweep
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        # Everything is measured and printed: synchronise, meet the other ranks once more and leave without running destructors
        # (peer-mapped IPC buffers / captured graphs make an orderly teardown of 8 processes slow).
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        json_out.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
