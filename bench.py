#!/usr/bin/env python
"""bench.py -- CeedOperatorApply throughput (GDoF/s) of the /gpu/cuda/b200 path on synthetic structured hex meshes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload bp1p3|bp3p6|bp5p7|...] [--dofs D] [--impl reference]

One "step" = one CeedOperatorApply (v = A u) on the rank's mesh, followed -- for N > 1 -- by the interface-DoF sum over
NCCL.  Prints ONE JSON line (rank 0).  `value` = total DoFs of all ranks / max-over-ranks device time, inputs resident in
HBM.  `e2e` = same metric through the C ABI with HOST buffers (pinned): H2D of u and D2H of v inside the timed region.
`roofline` = algorithmic bytes of one apply / CUDA-event duration of the fused kernel (+ finalize kernel), against the
measured HBM peak in MEASURED_PEAKS.json.  `cpu_baseline` = the unmodified reference's /cpu/self/avx/blocked on the host
cores, on a bounded sample of the same workload.  `--impl reference` runs only that CPU arm.
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
    ap.add_argument("--no-sweep", action="store_true", help="skip the `sweep` key (kernel numbers of BP1 p=3, BP3 p=1..8, BP5 p=4..7, BP6 p=4,6)")
    return ap.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
    bp, p, nel, resource, seconds, steps, warmup, barrier_dir, wid = args
    from libceed_b200 import mesh as M
    from libceed_b200.bp import seeded_uniform
    from oracle import refceed as R
    off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
    nn = coords.shape[1]
    rc = R.RefCeed(resource)
    prob = R.RefBP(rc, bp, p, off.shape[0], nn, off, coords)
    rc.set_array(prob.u, seeded_uniform(prob.ncomp * nn))
    for _ in range(max(1, warmup)):
        rc.op_apply(prob.op, prob.u, prob.v)
    # crude barrier through the file system so that all workers time the same interval
    open(os.path.join(barrier_dir, f"ready{wid}"), "w").close()
    while not os.path.exists(os.path.join(barrier_dir, "go")):
        time.sleep(0.001)
    t0 = time.perf_counter()
    n = 0
    while True:
        rc.op_apply(prob.op, prob.u, prob.v)
        n += 1
        if (steps and n >= steps) or (not steps and time.perf_counter() - t0 > seconds):
            break
    dt = time.perf_counter() - t0
    return prob.ncomp * nn, n, dt


def _cpu_port_worker(args):
    bp, p, nel, seconds, steps, warmup, barrier_dir, wid = args
    from libceed_b200 import mesh as M
    from libceed_b200.bp import seeded_uniform
    from oracle import oracle as O
    off, coords = M.hex_offsets(*nel, p), M.hex_coords(*nel, p)
    nn = coords.shape[1]
    _, _, nc, _ = O.bp_sizes(bp, False, p)
    qd = O.bp_qdata(bp, p, off, coords)
    u = seeded_uniform(nc * nn)
    for _ in range(max(1, warmup)):
        O.bp_apply(bp, p, off, nn, qd, u)
    open(os.path.join(barrier_dir, f"ready{wid}"), "w").close()
    while not os.path.exists(os.path.join(barrier_dir, "go")):
        time.sleep(0.001)
    t0, n = time.perf_counter(), 0
    while True:
        O.bp_apply(bp, p, off, nn, qd, u)
        n += 1
        if (steps and n >= steps) or (not steps and time.perf_counter() - t0 > seconds):
            break
    return nc * nn, n, time.perf_counter() - t0


def cpu_reference(bp, p, seconds, steps=0, warmup=1, resource="/cpu/self/avx/blocked", dofs_per_worker=150_000, max_cores=0):
    """Reference CPU backend on all host cores: `cores` forked single-threaded workers (libCEED CPU backends are
    single-threaded per Ceed), each owning its own slab of the workload mesh; aggregate = sum(DoFs * applies) / max(time)."""
    from libceed_b200 import mesh as M
    from libceed_b200.bp import BP_TABLE
    from oracle import refceed as R
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    if max_cores:
        cores = min(cores, max_cores)
    nel = M.choose_elements(dofs_per_worker, p, BP_TABLE[bp][0])
    # the unmodified reference (oracle/_ref) when it travelled with the repository, else the C restatement of its ref backend
    have_ref = R.available() and not os.environ.get("CEED_B200_BENCH_FORCE_PORT")
    kind = "reference" if have_ref else "port"
    if not have_ref:
        resource = "oracle/ceed_oracle.c (restatement of /cpu/self/ref/serial)"
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            if have_ref:
                jobs = pool.map_async(_cpu_worker, [(bp, p, nel, resource, seconds, steps, warmup, d, w) for w in range(cores)])
            else:
                jobs = pool.map_async(_cpu_port_worker, [(bp, p, nel, seconds, steps, warmup, d, w) for w in range(cores)])
            res = jobs
            t_start = time.time()
            while len([f for f in os.listdir(d) if f.startswith("ready")]) < cores and time.time() - t_start < 600:
                time.sleep(0.01)
            open(os.path.join(d, "go"), "w").close()
            out = res.get()
    dofs = sum(o[0] * o[1] for o in out)
    tmax = max(o[2] for o in out)
    applies = sum(o[1] for o in out)
    return dict(value=dofs / tmax / 1e9, unit="GDoF/s", cores=cores, kind=kind,
                sample=f"{resource}: {cores} forked workers x ({nel[0]}x{nel[1]}x{nel[2]} elements, {out[0][0]} DoFs), "
                       f"{applies} applies in {tmax:.3f} s", ms_per_step=tmax / max(1, out[0][1]) * 1e3, applies=applies)


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
                  partition="3-D element blocks, one local L-vector per GPU, NCCL interface sum" if world > 1 else "single GPU")

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(bp, p, args.cpu_seconds, steps=args.steps, warmup=args.warmup)
        line = dict(metric="CeedOperatorApply throughput", value=r["value"], unit="GDoF/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    impl="reference", config=config,
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
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))  # a hang aborts quickly
    ceed = cm.Ceed(f"/gpu/cuda/b200:device_id={local_rank}")
    ceed.set_scatter_mode({"deterministic": 0, "atomic": 1, "evector": 2}[args.scatter])
    stream = torch.cuda.current_stream()
    ceed.set_stream(stream.cuda_stream)

    # ---- problem: weak scaling -> per-GPU element box of ~dofs, global box = process grid x local box
    nel = M.choose_elements(args.dofs, p, ncomp)
    part = None
    if world > 1:
        grid = M.split3(world)
        part = M.Partition(tuple(nel[d] * grid[d] for d in range(3)), p, world, rank)
    prob = BPProblem(ceed, bp, p, nel, part=part)
    n_local = prob.num_dofs
    u_host = torch.from_numpy(seeded_uniform(n_local, 0x5EED + rank)).pin_memory()
    v_host = torch.empty(n_local, dtype=torch.float64).pin_memory()
    u_dev = u_host.to(dev)
    v_dev = torch.zeros(n_local, dtype=torch.float64, device=dev)
    prob.u.set_array(u_dev, cm.MEM_DEVICE, cm.USE_POINTER)
    prob.v.set_array(v_dev, cm.MEM_DEVICE, cm.USE_POINTER)
    exch = None
    if world > 1:
        from libceed_b200.parallel import InterfaceExchange
        exch = InterfaceExchange(part, ncomp, prob.num_nodes, dev, ceed=ceed)
        owned = int(part.owned_mask().sum()) * ncomp
        t = torch.tensor([owned], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        total_dofs = int(t.item())
    else:
        total_dofs = n_local

    def step():
        prob.op.apply(prob.u, prob.v)
        if exch is not None:
            exch.sum_interfaces(v_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = ceed.launch_count()
    step()
    launches_per_step = ceed.launch_count() - launches0
    barrier()
    # The step (apply kernels + pack + NCCL send/recv + unpack) is captured once in a CUDA graph and replayed: the per-step host
    # work (ctypes calls, building the grouped P2P ops) leaves the timed region.  Falls back to eager launches if capture fails.
    graph = None
    if not args.no_graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ceed.set_stream(torch.cuda.current_stream().cuda_stream)
                step()
            graph = g
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed, running eagerly: {exc}", file=sys.stderr)
            graph = None
        ceed.set_stream(stream.cuda_stream)
        ok = torch.tensor([1 if graph is not None else 0], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks replay, or none
        if int(ok.item()) == 0:
            graph = None
    run_step = graph.replay if graph is not None else step
    config["cuda_graph"] = graph is not None
    for _ in range(3):
        run_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        run_step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    gpu_launches = launches_per_step * args.steps
    # keep the GPU busy a little longer so that the clock sampler sees the loaded state even for sub-ms steps
    if sampler:
        t_end = time.time() + 1.0
        while time.time() < t_end:
            prob.op.apply(prob.u, prob.v)  # local work only: the other ranks do not take part in this extra second
        torch.cuda.synchronize()
        clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = total_dofs / (ms_per_step * 1e-3) / 1e9

    # ---- kernel-level timing for the roofline (CUDA events on the launching stream, inside the library)
    prob.op.set_timing(True)
    kms = []
    for _ in range(max(5, min(args.steps, 20))):
        prob.op.apply(prob.u, prob.v)
        kms.append(prob.op.last_kernel_ms())
    prob.op.set_timing(False)
    fused_ms = float(np.median([k[0] for k in kms]))
    aux_ms = float(np.median([k[1] for k in kms]))
    peak, peak_src = hbm_peak()
    alg_bytes = prob.bytes_per_apply()
    achieved = alg_bytes / ((fused_ms + aux_ms) * 1e-3) / 1e9
    info = prob.op.kernel_info()
    # DRAM bytes of one launch of the fused kernel from the committed ncu capture of this workload (profiles/), when there is one
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tpath) and world == 1 and abs(args.dofs - 10e6) < 1 and args.scatter == "deterministic":
        t = json.load(open(tpath)).get(args.workload)
        if t:
            traffic, traffic_src = t["traffic"], t["source"]
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, traffic_source=traffic_src, peak_source=peak_src,
                    kernel="b200_operator (+ halo finalize)", fused_kernel_ms=fused_ms, finalize_ms=aux_ms, algorithmic_bytes=alg_bytes,
                    kernel_share_of_step=(fused_ms + aux_ms) / ms_per_step, regs=info["regs"], elems_per_block=info["elems_per_block"],
                    threads=info["threads"], grid=info["grid"], smem_bytes=info["smem_bytes"])

    # ---- end to end through the C ABI with host buffers: H2D(u) + apply + D2H(v) every step
    e2e_steps = max(3, min(args.steps, 10))
    uh, vh = ceed.Vector(n_local), ceed.Vector(n_local)

    def e2e_step():
        uh.set_array(u_host, cm.MEM_HOST, cm.USE_POINTER)   # host pointer becomes the only valid copy
        vh.set_array(v_host, cm.MEM_HOST, cm.USE_POINTER)
        prob.op.apply(uh, vh)                                # syncs u to the device, runs the kernels
        vh.sync_array(cm.MEM_HOST)                           # D2H into the caller's (pinned) buffer
        if exch is not None:
            pass  # interface sum is part of the device-resident step; host round trip measured per rank only

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = dict(value=total_dofs / e2e_s / 1e9, unit="GDoF/s", h2d_bytes_per_step=8 * n_local, d2h_bytes_per_step=8 * n_local,
               ms_per_step=e2e_s * 1e3, mode="serial: H2D(u), apply, D2H(v) one after the other")

    # ---- the same end-to-end step, software-pipelined over steps: every step still copies its own u from pinned host memory and
    # its own v back, but on separate copy streams with double-buffered device vectors, so the H2D of step i+1 and the D2H of
    # step i-1 overlap the kernels of step i (PCIe is full duplex).  Reported as e2e.pipelined_value; any failure keeps the serial one.
    if not args.no_e2e_pipeline:
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
            ev_in = [torch.cuda.Event() for _ in range(nbuf)]
            ev_done = [torch.cuda.Event() for _ in range(nbuf)]
            ev_out = [torch.cuda.Event() for _ in range(nbuf)]

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
            ok = bool(torch.equal(v_h[0], v_host)) if exch is None else True   # same bits as the serial path
            if world > 1:
                t = torch.tensor([pipe_s], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                pipe_s = float(t.item())
            e2e["pipelined_value"] = total_dofs / pipe_s / 1e9
            e2e["pipelined_ms_per_step"] = pipe_s * 1e3
            e2e["pipelined_matches_serial"] = ok
            e2e["pipelined_mode"] = "double-buffered: H2D of step i+1 and D2H of step i-1 overlap the kernels of step i"
            if ok and e2e["pipelined_value"] > e2e["value"]:
                # headline = the pipelined figure (every step still moves its own u and v over PCIe inside the timed region);
                # the strictly serial figure stays next to it
                e2e["serial_value"], e2e["serial_ms_per_step"] = e2e["value"], e2e["ms_per_step"]
                e2e["value"], e2e["ms_per_step"], e2e["mode"] = e2e["pipelined_value"], e2e["pipelined_ms_per_step"], e2e["pipelined_mode"]
        except Exception as exc:  # noqa: BLE001
            e2e["pipelined_value"] = None
            e2e["pipelined_error"] = str(exc)[:200]

    sweep = None
    if not args.no_sweep and rank == 0 and world == 1:
        sweep = []
        del prob
        for sbp, ps in ((1, (3,)), (3, range(1, 9)), (5, range(4, 8)), (6, (4, 6))):
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference(bp, p, args.cpu_seconds)
            cpu = dict(value=r["value"], unit="GDoF/s", cores=r["cores"], kind=r["kind"], sample=r["sample"])
            # also the opt/blocked backend on all cores and the one-core figure (short samples)
            others = []
            for res_name, mc in (("/cpu/self/opt/blocked", 0), ("/cpu/self/avx/blocked", 1)):
                try:
                    o = cpu_reference(bp, p, 4.0, resource=res_name, max_cores=mc)
                    others.append(dict(value=o["value"], unit="GDoF/s", cores=o["cores"], kind=o["kind"], sample=o["sample"]))
                except Exception as exc:  # noqa: BLE001
                    others.append(dict(value=None, sample=f"{res_name}: unavailable: {exc}"))
            cpu["others"] = others
        except Exception as exc:  # the reference build did not travel: say so instead of inventing a number
            cpu = dict(value=None, unit="GDoF/s", cores=0, kind="reference", sample=f"unavailable: {exc}")

    if rank == 0:
        line = dict(metric="CeedOperatorApply throughput", value=value, unit="GDoF/s", n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=config,
                    roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(gpu_launches), clocks=clocks, total_dofs=total_dofs)
        if sweep:
            line["sweep"] = sweep
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        # A captured graph holds NCCL work: tearing the process group down with it alive can block.  Everything is measured and
        # printed at this point, so synchronise, meet the other ranks once more and leave without running the destructors.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        json_out.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
