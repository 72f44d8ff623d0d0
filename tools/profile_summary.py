"""profiles/ summary of one bench configuration: kernel shares from the ncu launch list
(--metrics gpu__time_duration.sum; cold-cache, serialised) + the --set full metrics of the dominant kernel.

  python tools/profile_summary.py <launches.csv> <k1.ncu-rep> <out.json> [command string]
"""
import csv, json, subprocess, sys
from collections import defaultdict

launches, rep, out = sys.argv[1:4]
cmd = sys.argv[4] if len(sys.argv) > 4 else ""
rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = defaultdict(list)
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    name = r[ix["Kernel Name"]].split("(")[0]
    if "at::native" in name or "vectorized_elementwise" in name or "elementwise_kernel" in name:
        name = "torch fill/copy (bench harness: L2 flush, result check)"
    per[name].append(us)
ours = {k: v for k, v in per.items() if not k.startswith("torch")}
tot = sum(sum(v) for v in ours.values())
shares = {k: {"launches": len(v), "mean_us": sum(v) / len(v), "share_of_step": sum(v) / tot} for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1]))}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
m = dict(zip(r[0], zip(r[1], r[2])))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
full = {"kernel": m["Kernel Name"][1]}
for k in keys:
    if k in m:
        full[k] = " ".join(x for x in (m[k][1], m[k][0]) if x)
stalls = {k.split("issue_stalled_")[1].split("_per_")[0]: float(val[1]) for k, val in m.items()
          if k.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in k and k.endswith("per_issue_active.ratio")}
full["warp_stalls_per_issue"] = {k: round(x, 3) for k, x in sorted(stalls.items(), key=lambda kv: -kv[1]) if x > 0.05}
def to_bytes(s):
    v, u = s.split()[:2]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
full["dram_bytes_per_launch"] = to_bytes(full["dram__bytes_read.sum"]) + to_bytes(full["dram__bytes_write.sum"])
json.dump({"command": cmd, "launch_list_shares_cold_cache_serialised": shares, "k1_ncu_set_full": full}, open(out, "w"), indent=1)
print(open(out).read())
