#!/usr/bin/env python3
"""Build the committed profiles/ artefacts from what tools/final_measure.sh brought back in gpurun_out/.
usage: tools/make_profiles.py [tag]      (tag defaults to r1)"""
import csv, io, json, os, shutil, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"

shutil.copy(os.path.join(G, "final_bench.json"), os.path.join(P, f"{tag}_bench.json"))
shutil.copy(os.path.join(G, "final_ref.json"), os.path.join(P, f"{tag}_bench_reference_arm.json"))
for src, dst in (("final_bench_spec.json", "bench_pose_spread_spec.json"), ("final_bench_xyz.json", "bench_scan_format_xyz.json"),
                 ("final_bench_offroad.json", "bench_offroad_rig.json"), ("final_next_rows.jsonl", "next_rows.jsonl")):
    if os.path.isfile(os.path.join(G, src)) and os.path.getsize(os.path.join(G, src)) > 0:
        shutil.copy(os.path.join(G, src), os.path.join(P, f"{tag}_{dst}"))

# ---- launch list (ncu --metrics gpu__time_duration.sum --clock-control none) + per-kernel shares
rows = [r for r in csv.reader(open(os.path.join(G, "final_launches.csv"))) if r and r[0].isdigit()]
shutil.copy(os.path.join(G, "final_launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows:
    k = r[4].split("(")[0]; tot[k] += float(r[-1].replace(",", "")) / 1e6; cnt[k] += 1
allms = sum(tot.values())
with open(os.path.join(P, f"{tag}_launches_summary.txt"), "w") as f:
    f.write(f"# bench.py --steps 2 --warmup 1 --no-cpu under ncu (serialised, cold caches): share of the step per kernel, {len(rows)} launches, {allms:.1f} ms\n")
    for k in sorted(tot, key=lambda k: -tot[k]):
        f.write(f"{k:22s} launches {cnt[k]:3d}  total {tot[k]:9.3f} ms  share {100 * tot[k] / allms:5.1f} %\n")

# ---- every kernel, 200-frame step (SpeedOfLight / memory / occupancy sections)
raw = list(csv.reader(open(os.path.join(G, "final_all_raw.csv"))))
H, U = raw[0], raw[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
idx = [H.index(w) for w in want if w in H]; ki = H.index("Kernel Name")
seen = set()
with open(os.path.join(P, f"{tag}_all_kernels_ncu.csv"), "w") as f:
    f.write("kernel," + ",".join(f"{H[i]} [{U[i]}]" for i in idx) + "\n")
    for r in raw[2:]:
        k = r[ki].split("(")[0]
        if k in seen: continue
        seen.add(k); f.write(k + "," + ",".join(r[i].replace(",", "") for i in idx) + "\n")

# ---- the dominant kernel at the bench size
rep = os.path.join(G, "final_icp.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out))); h, u, v = r[0], r[1], r[-1]
keep = ("dram__", "gpu__", "l1tex__t_sector_hit", "lts__t_sector_hit", "launch__", "sm__warps_active", "smsp__issue_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst", "smsp__average_warps_issue_stalled", "sm__inst_executed_pipe", "sm__throughput", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
with open(os.path.join(P, f"{tag}_icp_pass_ncu_raw.csv"), "w") as f:
    f.write("metric,unit,value\n")
    for i, k in enumerate(h):
        if k.startswith(keep): f.write(f"{k},{u[i]},{v[i].replace(',', '')}\n")
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_icp_pass_ncu_details.csv"), "w").write(det)
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "45"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_icp_pass_hot_lines.txt"), "w").write(lines)
m = {k: v[i].replace(",", "") for i, k in enumerate(h)}; mu = {k: u[i] for i, k in enumerate(h)}
def to_bytes(k):
    s = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[mu[k]]; return float(m[k]) * s
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
json.dump({"kernel": "icp_pass", "frames": 1000, "pose_spread": "tight", "dram_bytes_per_launch": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
           "warp_instructions": int(float(m["smsp__inst_executed.sum"])),
           "source": f"profiles/{tag}_icp_pass_ncu_raw.csv (ncu --set full --clock-control none, bench.py --frames 1000 --steps 1 --warmup 0)"},
          open(os.path.join(P, "traffic.json"), "w"))
print(open(os.path.join(P, f"{tag}_launches_summary.txt")).read()); print(lines[:1500]); print(open(os.path.join(P, "traffic.json")).read())
