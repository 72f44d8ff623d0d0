"""SASS evidence per kernel of libdto_b200.so: instruction mnemonics that show which hardware paths a kernel uses
(DMMA = FP64 tensor pipe, DFMA = FP64 FMA pipe, LDGSTS = cp.async, LDS/STS = shared memory, UTMALDG/UTMASTG/UBLKCP = TMA,
UTCxMMA = tcgen05, SHFL, BAR, ATOM/RED, LDL/STL = local-memory spills).  Writes a table to stdout.
Usage: python tools/sass_counts.py [path/to/lib.so] > profiles/r02_sass_counts.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "directtrajopt.jl_b200/lib/libdto_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["DMMA", "DFMA", "DMUL", "DADD", "LDGSTS", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR", "ATOM", "RED", "UTMALDG", "UTMASTG",
        "UBLKCP", "UTCHMMA", "UTCQMMA", "MUFU", "CS2R"]
kern, counts, arch = None, {}, None
for line in out.splitlines():
    m = re.match(r"\s*arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        counts[kern]["arch"] = arch
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["total"] += 1
        for k in keys:
            if op == k or op.startswith(k + "."):
                counts[kern][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}: static SASS instruction counts per kernel (cuobjdump -sass); all cubins {sorted(set(c['arch'] for c in counts.values()))}")
print("kernel".ljust(72) + "".join(k.rjust(8) for k in ["total"] + keys))
for k, nm in sorted(zip(counts, names), key=lambda kv: -kv[0].__len__() * 0 - counts[kv[0]]["DMMA"]):
    short = re.sub(r"\(anonymous namespace\)::", "", nm)
    short = re.sub(r"\(.*", "", short)
    print(short[:71].ljust(72) + "".join(str(counts[k][x]).rjust(8) for x in ["total"] + keys))
tot = collections.Counter()
for c in counts.values():
    for k in keys:
        tot[k] += c[k]
print("ALL KERNELS".ljust(72) + " " * 8 + "".join(str(tot[k]).rjust(8) for k in keys))
