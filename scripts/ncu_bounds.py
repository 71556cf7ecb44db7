"""Which unit bounds a kernel: one table row per ncu capture (L1 data pipe = shared-memory + global load/store wavefronts, DRAM, issue
slots, FP64 pipe, resident warps).  usage: python scripts/ncu_bounds.py label=prof.ncu-rep [...]"""
import csv, io, subprocess, sys
KEYS = [("L1 data pipe %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("of it shared mem %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("FP64 pipe %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("warps/sched", "smsp__warps_active.avg.per_cycle_active"),
        ("regs", "launch__registers_per_thread"),
        ("us (under ncu)", "gpu__time_duration.sum"),
        ("smem conflicts / wavefronts", None)]
print("| kernel | " + " | ".join(k for k, _ in KEYS) + " |")
print("|---|" + "---|" * len(KEYS))
for arg in sys.argv[1:]:
    label, path = arg.rsplit("=", 1)
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "b200_operator" not in d.get("Kernel Name", ""): continue
        cells = []
        for name, key in KEYS:
            if key is None:
                c, w = float(d["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"].replace(",", "")), float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",", ""))
                cells.append(f"{100 * c / w:.0f} %")
            else:
                v = float(d[key].replace(",", ""))
                cells.append(f"{v:.0f}" if name in ("regs",) else f"{v:.1f}")
        print(f"| {label} ({d['Kernel Name'].split('_')[-1]}, grid {d['Grid Size']} x {d['Block Size']}) | " + " | ".join(cells) + " |")
