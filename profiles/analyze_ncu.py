"""Summarise an ncu report exported as CSV (raw page + source page): key counters, stall mix, opcode mix.
usage: python profiles/analyze_ncu.py raw.csv sass.csv [rows_per_launch]"""
import collections
import csv
import sys

raw, sass = sys.argv[1], sys.argv[2]
nrows = float(sys.argv[3]) if len(sys.argv) > 3 else 1048576
rows = list(csv.reader(open(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for w in KEYS:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:75s} {r[i]:>16s} {units[i]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(r[i]) > 0.15:
        print(f"  stall {h[34:-28]:30s} {r[i]}")
rows = list(csv.reader(open(sass)))
h = rows[1]
c = {n: i for i, n in enumerate(h)}
data = []
for rr in rows[2:]:
    if len(rr) < len(h):
        continue
    try:
        data.append((int(rr[c["Instructions Executed"]]), int(rr[c["Warp Stall Sampling (All Samples)"]]), rr[c["Source"]].strip()))
    except Exception:
        pass
tot = sum(d[0] for d in data)
ts = sum(d[1] for d in data)
op, ops = collections.Counter(), collections.Counter()
for n, s, src in data:
    t = src.split()
    o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    op[o] += n
    ops[o] += s
print("total warp inst", tot, " per row:", tot / nrows)
for o, n in op.most_common(18):
    print(f"  {o:10s} {n / tot * 100:5.1f}% inst ({n / nrows:6.2f}/row) {ops[o] / max(ts, 1) * 100:5.1f}% stall")
print("top stalls:")
for d in sorted(data, key=lambda x: -x[1])[:8]:
    print(f"  {d[1] / max(ts, 1) * 100:5.1f}%  {d[2][:80]}")
