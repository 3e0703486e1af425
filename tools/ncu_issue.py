"""Extracts the counters bench.py quotes from an .ncu-rep (run here, no GPU) into profiles/<round>/issue.json and traffic.json.
usage: python tools/ncu_issue.py <report.ncu-rep> <round dir> <workload> <spp> <kernel label> [pt_mode]"""
import csv
import io
import json
import os
import subprocess
import sys

rep, outdir, workload, spp, kernel = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
pt_mode = int(sys.argv[6]) if len(sys.argv) > 6 else 2
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda name: float(vals[hdr.index(name)].replace(",", "")) if name in hdr else None
unit = lambda name: units[hdr.index(name)]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram = get("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")] + get("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
entry = {
    "workload": workload, "spp": spp, "kernel": kernel, "report": os.path.basename(rep),
    "warp_inst_per_launch": int(get("smsp__inst_executed.sum")),
    "issue_slot_util": round(get("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0, 4),
    "lanes_per_inst": round(get("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
    "warps_active_frac": round(get("sm__warps_active.avg.pct_of_peak_sustained_active") / 100.0, 4),
    "l1tex_hit": round(get("l1tex__t_sector_hit_rate.pct") / 100.0, 4), "l2_hit": round(get("lts__t_sector_hit_rate.pct") / 100.0, 4),
    "dram_bytes_per_launch": int(dram), "registers": int(get("launch__registers_per_thread")),
    "ms_under_ncu": round(get("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(unit("gpu__time_duration.sum"), 1), 4),
}
os.makedirs(outdir, exist_ok=True)
path = os.path.join(outdir, "issue.json")
entries = json.load(open(path)) if os.path.exists(path) else []
entries = [e for e in entries if not (e["workload"] == workload and e["spp"] == spp and e["kernel"] == kernel)] + [entry]
json.dump(entries, open(path, "w"), indent=1)
if workload == "C3":
    json.dump({"workload": workload, "spp": spp, "pt_mode": pt_mode, "kernel": kernel, "dram_bytes_per_launch": int(dram), "report": os.path.basename(rep)},
              open(os.path.join(outdir, "traffic.json"), "w"), indent=1)
print(json.dumps(entry))
