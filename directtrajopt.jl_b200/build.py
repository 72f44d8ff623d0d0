"""Build libdto_b200.so (sm_100a) in-tree with nvcc.  Called by __graft_entry__.build().

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(LIBDIR, "libdto_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

# DTO_EXTRA_NVCC_FLAGS: extra flags for debug builds (e.g. -DDTO_TDB_PROFILE); part of the rebuild key
NVCC_FLAGS = os.environ.get("DTO_EXTRA_NVCC_FLAGS", "").split() + [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # NOT -split-compile: it cuts the build from 4 min to under 2 but the kernels come out slower (measured on B200:
    # K1 at c2 0.273 -> 0.403 ms, K7 at c3 8.6 -> 9.7 ms)
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_hash():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/dto_b200.h"]:
        p = os.path.normpath(os.path.join(CSRC, f))
        if os.path.isfile(p) and (p.endswith((".h", ".cuh"))):
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, dephash, verbose):
    # the object is named by the hash of everything that went into it (source, headers, flags): an object from another
    # checkout can never be mistaken for a current one, and nothing about the build cache is tracked by git
    key = hashlib.sha256(open(src, "rb").read() + dephash.encode()).hexdigest()[:16]
    stem = os.path.basename(src)[:-3]
    obj = os.path.join(OBJDIR, f"{stem}.{key}.o")
    if os.path.exists(obj):
        return obj, ""
    for f in os.listdir(OBJDIR):  # older objects of the same source
        if f.startswith(stem + ".") and f.endswith(".o"):
            os.remove(os.path.join(OBJDIR, f))
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    os.replace(obj + ".tmp", obj)
    return obj, r.stderr


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    dephash = _deps_hash()
    srcs = _sources()
    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = []
        for obj, log in ex.map(lambda s: _compile(s, dephash, verbose), srcs):
            objs.append(obj)
            logs.append(log)
    # relink when the set of objects changed (their names carry their content keys)
    manifest = os.path.join(OBJDIR, "linked.txt")
    want = "\n".join(sorted(os.path.basename(o) for o in objs))
    have = open(manifest).read() if os.path.exists(manifest) else ""
    if force or not os.path.exists(LIB) or have != want:
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        open(manifest, "w").write(want)
    if verbose:
        sys.stderr.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
