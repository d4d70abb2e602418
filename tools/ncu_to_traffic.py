#!/usr/bin/env python3
"""profiles/traffic.json from an ncu report, by script (not by hand):

    python tools/ncu_to_traffic.py gpurun_out/r2_prof_pooled_mesh1m.ncu-rep mesh1m [summary.txt]

Takes the longest launch of the dominant kernel (trace_pooled_kernel<0,...> on the mesh, trace_flat_kernel<0,...> on
cornell_box) out of an `ncu --set full --clock-control none` capture and writes the per-launch figures bench.py copies into
its JSON line (a bench never runs under the profiler): DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), duration,
instructions, lanes per instruction, issue-slot / l1tex / LSU-data-pipe / lts utilisation, L1 and L2 hit rates."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload = sys.argv[1], sys.argv[2]
kernel_key = {"mesh1m": "trace_pooled_kernel<0", "cornell": "trace_flat_kernel<0"}[workload]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]


def col(r, name, default=None):
    if name not in hdr:
        return default
    v = r[hdr.index(name)].replace(",", "")
    try:
        x = float(v)
    except ValueError:
        return default
    u = units[hdr.index(name)]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return x * scale


cands = [r for r in rows[2:] if kernel_key in r[hdr.index("Kernel Name")].replace("(int)", "").replace(" ", "")
         or kernel_key in r[hdr.index("Kernel Name")]]
if not cands:
    cands = [r for r in rows[2:] if kernel_key.split("<")[0] in r[hdr.index("Kernel Name")] and "0" in r[hdr.index("Kernel Name")].split("<")[1][:8]]
best = max(cands, key=lambda r: col(r, "gpu__time_duration.sum", 0.0))
ent = {
    "kernel": kernel_key + ">",
    "trace_closest_dram_bytes_per_launch": int(col(best, "dram__bytes_read.sum", 0) + col(best, "dram__bytes_write.sum", 0)),
    "ncu": {
        "source": os.path.relpath(sys.argv[3], ROOT) if len(sys.argv) > 3 else os.path.basename(rep),
        "launch_ms": col(best, "gpu__time_duration.sum"),
        "warp_inst": col(best, "smsp__inst_executed.sum"),
        "lanes_per_inst": col(best, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": col(best, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l1tex_throughput_pct": col(best, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        "l1tex_lsu_data_pipe_pct": col(best, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "lts_throughput_pct": col(best, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1_hit_pct": col(best, "l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": col(best, "lts__t_sector_hit_rate.pct"),
        "dram_pct": col(best, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": col(best, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "long_scoreboard_per_issue": col(best, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "registers": col(best, "launch__registers_per_thread"),
        "note": "one launch under ncu (cold caches, serialised): compare shares and counters, not absolute times",
    },
}
path = os.path.join(ROOT, "profiles", "traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data["_comment"] = ("per-launch ncu figures of the dominant kernel, written by tools/ncu_to_traffic.py from the .ncu-rep named in "
                    "'source' and read by bench.py (a bench never runs under the profiler)")
data[workload] = ent
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(ent, indent=1))
