"""Per-kernel summary of an .ncu-rep (`ncu -i X --page raw --csv`): duration, DRAM bytes and throughput, L2 hit rate, occupancy, issue-slot and pipe use,
top stall reasons.  Usage: python tools/ncu_summary.py X.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]; pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; units = rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
        ("lts__t_sector_hit_rate.pct", "l2_hit%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%peak")]
seen = collections.OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if pat and not pat.search(name): continue
    key = re.sub(r"\(.*", "", name)[:60]
    seen.setdefault(key, []).append(r)
for key, rs in seen.items():
    r = min(rs, key=lambda x: float(x[col["gpu__time_duration.sum"]].replace(",", "")) if "gpu__time_duration.sum" in col else 0)
    parts = []
    for m, lab in want:
        if m in col and r[col[m]] not in ("", "n/a"):
            parts.append(f"{lab}={r[col[m]]}{units[col[m]] if lab in ('time', 'dram_rd', 'dram_wr') else ''}")
    print(f"{key}  x{len(rs)}\n    " + "  ".join(parts))
