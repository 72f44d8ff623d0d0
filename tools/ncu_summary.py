"""Summarise an .ncu-rep (first kernel): key metrics + stall breakdown + top stalled instructions."""
import csv, subprocess, sys, json
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, u, v = r[0], r[1], r[2]
m = dict(zip(h, zip(u, v)))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__inst_executed_pipe_tensor_op_dmma.sum", "sm__inst_executed_pipe_tensor_op_dmma.sum", "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
out = {}
for k in keys:
    if k in m:
        out[k] = " ".join(x for x in (m[k][1], m[k][0]) if x)
stalls = {k.split("issue_stalled_")[1].split("_per_")[0]: float(val[1]) for k, val in m.items()
          if k.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in k and k.endswith("per_issue_active.ratio")}
out["stalls_per_issue"] = {k: round(x, 3) for k, x in sorted(stalls.items(), key=lambda kv: -kv[1]) if x > 0.01}
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]; data = rows[2:]
    ix = {hh: i for i, hh in enumerate(hdr)}
    def f(rr, k):
        try: return float(rr[ix[k]])
        except Exception: return 0.0
    tot = sum(f(rr, "# Samples") for rr in data)
    agg = defaultdict(lambda: defaultdict(float))
    for rr in data:
        t = rr[ix["Source"]].split()
        if not t: continue
        op = t[1] if t[0].startswith("@") else t[0]
        op = op.split(".")[0]
        for k in ["# Samples", "stall_wait", "stall_math", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_mio", "stall_lg", "stall_branch_resolving", "stall_no_inst"]:
            agg[op][k] += f(rr, k)
    print("total samples", tot)
    for op, d in sorted(agg.items(), key=lambda x: -x[1]["# Samples"])[:14]:
        print(op, {k: int(x) for k, x in d.items() if x > 0})
    print("--- top instructions")
    for rr in sorted(data, key=lambda rr: -f(rr, "# Samples"))[:int(sys.argv[2])]:
        print(rr[ix["Address"]][-5:], rr[ix["Source"]][:70], int(f(rr, "# Samples")), {k[6:]: int(f(rr, k)) for k in ["stall_wait", "stall_math", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_lg", "stall_mio"] if f(rr, k) > 0})
