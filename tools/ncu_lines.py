"""Per-CUDA-source-line stall samples of an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; cur = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; ns = len(r); continue
    if r[0] in ("Function Name",): continue
    if hdr is None: continue
    if r[0] != "":  # a source line row (its text may contain commas -> variable length); start a new group
        cur = (fname, r[0], ",".join(r[1:len(r) - (ns - 2) + 0])[:90] if len(r) > ns else r[1][:90]); agg.setdefault(cur, [0, 0, 0, 0, 0]); continue
    if cur is None or len(r) < ns: continue
    def f(k):
        try: return float(r[hdr[k]])
        except Exception: return 0.0
    a = agg[cur]
    a[0] += f("# Samples"); a[1] += f("stall_barrier"); a[2] += f("stall_long_sb"); a[3] += f("stall_short_sb"); a[4] += f("Instructions Executed")
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:>4} samples={int(a[0]):7d} ({100*a[0]/tot:4.1f}%) barrier={int(a[1]):6d} long_sb={int(a[2]):6d} short_sb={int(a[3]):6d} inst={int(a[4]):9d} | {k[2]}")
