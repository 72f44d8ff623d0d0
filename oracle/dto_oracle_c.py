"""ctypes wrapper of oracle/dto_oracle.c (the C restatement of the reference's ForwardDiff-through-expv
algorithm).  TEST INFRASTRUCTURE ONLY: checker for tests/ and the timed CPU baseline of bench.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libdto_oracle.so")
_lib = None


def build():
    src = os.path.join(HERE, "dto_oracle.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "_build/libdto_oracle.so"], check=True, capture_output=True)
    return SO


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        lib = C.CDLL(SO)
        dp = C.POINTER(C.c_double)
        lib.dto_port_bilinear_interval.restype = C.c_int
        lib.dto_port_bilinear_interval.argtypes = [C.c_int] * 6 + [dp, dp, dp, dp, C.c_int, C.c_int, dp, dp, dp]
        lib.dto_port_eval_intervals.restype = C.c_double
        lib.dto_port_eval_intervals.argtypes = [C.c_int] * 6 + [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.dto_port_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def bilinear_interval(spec, it, zk, zk1, mu=None, all_dirs=True):
    """Residual, Jacobian block (n x 2z) and Hessian block (2z x 2z) of one BilinearIntegrator interval
    by forward-mode jets through the truncated Taylor expv."""
    lib = load()
    z = spec["z"]
    x_off, n = spec["components"][it["x"]]
    u_off, m = spec["components"][it["u"]]
    dt_off = spec["components"][spec["timestep"]][0]
    G = np.ascontiguousarray(np.asarray(it["G"], float))
    zk, zk1 = np.ascontiguousarray(zk, float), np.ascontiguousarray(zk1, float)
    r, Jb = np.zeros(n), np.zeros((n, 2 * z))
    lib.dto_port_bilinear_interval(n, m, z, x_off, u_off, dt_off, _p(G), _p(zk), _p(zk1), None, 1, int(all_dirs), _p(r), _p(Jb), None)
    Hb = None
    if mu is not None:
        mu = np.ascontiguousarray(mu, float)
        Hb = np.zeros((2 * z, 2 * z))
        lib.dto_port_bilinear_interval(n, m, z, x_off, u_off, dt_off, _p(G), _p(zk), _p(zk1), _p(mu), 2, int(all_dirs), None, None, _p(Hb))
    return r, Jb, Hb


def time_port(prob, sample_intervals, threads, all_dirs=True):
    """Time residual + Jacobian + Hessian passes over `sample_intervals` knot intervals of the
    problem's first BilinearIntegrator; return (evaluations/s of the WHOLE problem, info)."""
    lib = load()
    spec = prob.to_spec()
    it = next(i for i in spec["integrators"] if i["kind"] == "bilinear")
    z, N = spec["z"], spec["N"]
    x_off, n = spec["components"][it["x"]]
    u_off, m = spec["components"][it["u"]]
    dt_off = spec["components"][spec["timestep"]][0]
    G = np.ascontiguousarray(np.asarray(it["G"], float))
    Z = np.ascontiguousarray(prob.trajectory.datavec, float)
    count = int(min(sample_intervals, N - 1))
    threads = int(min(threads, lib.dto_port_max_threads(), count))
    mu = np.random.default_rng(0).random(count * n)
    t0 = time.perf_counter()
    chk = lib.dto_port_eval_intervals(n, m, z, x_off, u_off, dt_off, _p(G), _p(Z), _p(mu), 0, count, int(all_dirs), threads)
    dt = time.perf_counter() - t0
    per_interval = dt / count  # wall seconds per interval with `threads` threads working in parallel
    rate = 1.0 / (per_interval * (N - 1))
    info = {"threads": threads, "seconds": dt, "checksum": chk,
            "sample": f"{count} of {N - 1} knot intervals (residual + Jacobian + Hessian passes, "
                      f"{'all 2z' if all_dirs else 'active'} directions), {dt:.2f} s wall on {threads} threads, scaled linearly in N"}
    return rate, info
