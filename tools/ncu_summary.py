"""Summarise an .ncu-rep here (no GPU): key raw metrics + instruction share per CUDA source line.
usage: python tools/ncu_summary.py <report.ncu-rep> [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tex.sum", "l1tex__texin_requests.sum",
    "l1tex__t_sectors_pipe_tex.sum", "l1tex__t_requests_pipe_tex.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__cycles_active.avg", "l1tex__data_pipe_tex_wavefronts.sum", "l1tex__f_tex2sm_cycles_active.avg.pct_of_peak_sustained_elapsed",
]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for k, vals in enumerate(rows[2:]):
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    print(f"== launch {k}: {name[:100]}")
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"  {h} [{units[i]}] = {vals[i]}")
    stall = {h: float(vals[i] or 0) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(stall.values()) or 1
    print("  stall samples: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {v / tot * 100:.1f}%" for h, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows[:6]) if "Source" in r]
if not hi:
    sys.exit(0)
hdr = rows[hi[0]]
iS, iN, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg = collections.OrderedDict()
cur = None
tot_i = tot_s = 0
for r in rows[hi[0] + 1:]:
    if len(r) <= iT:
        continue
    if r[0].strip():
        cur = (r[0], r[1].strip()[:105])
        agg.setdefault(cur, [0, 0, 0])
        continue
    if cur is None:
        continue
    try:
        s, n, t = int(r[iS] or 0), int(r[iN] or 0), int(r[iT] or 0)
    except ValueError:
        continue
    agg[cur][0] += s
    agg[cur][1] += n
    agg[cur][2] += t
    tot_i += n
    tot_s += s
print(f"== per source line (share of {tot_i} warp instructions, {tot_s} stall samples)")
for (ln, text), (s, n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print(f"{n / max(tot_i, 1) * 100:5.1f}% inst {s / max(tot_s, 1) * 100:5.1f}% samp  thr/inst {t / max(n, 1):4.1f}  L{ln}: {text}")
