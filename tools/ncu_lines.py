"""Aggregate ncu warp-stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`."""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1], errors="ignore"))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, cur_line, cur_src = "?", None, ""
agg = collections.Counter(); src = {}; stall = collections.defaultdict(collections.Counter)
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur_line = (cur_file, r[0]); src[cur_line] = r[1][:100]; continue
    try:
        n = int(r[hdr.index("# Samples")])
    except Exception:
        continue
    agg[cur_line] += n
    for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_selected", "stall_not_selected", "stall_branch_resolving"):
        try:
            stall[cur_line][k] += int(r[hdr.index(k)])
        except Exception:
            pass
tot = sum(agg.values())
print("total samples", tot)
for k, n in agg.most_common(top):
    s = ", ".join(f"{a[6:]}={b}" for a, b in stall[k].most_common(3))
    print(f"{n:6d} {100*n/tot:5.1f}% {k[0]}:{k[1]:>4}  [{s}]  {src.get(k,'')}")
