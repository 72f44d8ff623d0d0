#!/usr/bin/env python
"""bench.py -- full NLP evaluations per second of the B200 evaluator (BASELINE.json metric).

One *step* = one full evaluation of one iterate of the workload problem: constraint residual +
sparse constraint Jacobian + sparse Hessian of the Lagrangian (objective and gradient are computed
in the same pass and are part of the step).  Workload at N=1: BASELINE.json configs[1] ("c2"):
bilinear isomorphic-state quantum gate problem, state dim 32, 4 drives, N=2000 knots, free dt +
MinimumTimeObjective.  With --gpus N > 1 every rank evaluates its own independent problem of the
same shape (problem-parallel, no data-path collective): weak scaling.

  value      whole-job evaluations/s with Z, mu and all outputs resident in HBM (dto_eval_all_dev)
  e2e        the same metric through the host-pointer C-ABI call (dto_eval_all) with pinned HOST
             buffers: H2D of Z and mu and D2H of all five outputs inside the timed region
  roofline   dominant kernel (K1, the bilinear interval kernel): algorithmic FP64 flops per launch
             (SURVEY.md section 8d canonical count C1) / CUDA-event duration measured live
  cpu_baseline   the CPU oracle port timed on the host cores on a bounded sample of the same workload

`--impl reference` times the reference's CPU algorithm for the path (the oracle port; the reference
is Julia and neither Julia nor its packages exist here or on the GPU box) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] (default, the configuration the metric is quoted on): one problem per GPU, weak scaling
    "c2": dict(kind="gate", N=2000, levels=16, n_drives=4),
    # configs[3]: ONE long trajectory, knot ranges sharded over the ranks with a one-knot NVLink halo (strong scaling)
    "c4": dict(kind="scaled", N=100000, state_dim=16, n_controls=2, generator_scale=0.25),
    # configs[4]: 4096 independent 8-state problems, problem-parallel over the ranks (strong scaling)
    "c5": dict(kind="scaled", N=200, state_dim=8, n_controls=2, generator_scale=0.35, batch=4096),
}
FP64_PEAK_TFLOPS = 37.1  # measured on this pool's B200: DMMA m8n8k4 saturation (profiles/r01_fp64_peaks_and_box_probe.log)
TAYLOR_T = 20            # canonical term count of SURVEY.md section 8d (C1)


def canonical_flops_per_interval(n, m, T=TAYLOR_T, s=0):
    """SURVEY.md section 8d, count C1: (i) full propagator by Pade-13 scaling and squaring,
    (ii) forward second-order directional propagation, (iii) adjoint first-order propagation."""
    p = m + 1
    i = 2 * n**3 * (6 + s) + (8.0 / 3.0) * n**3
    ii = (1 + 2 * p + 3 * p * (p + 1) / 2) * T * 2 * n * n
    iii = (1 + 2 * p) * T * 2 * n * n
    return i + ii + iii


def recorded_traffic(workload, variant):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload and kernel variant (profiles/); None when no capture matches."""
    path = os.path.join(ROOT, "profiles", f"r01_ncu_summary_{workload}.json")
    try:
        full = json.load(open(path))["k1_ncu_set_full"]
        if variant in full["kernel"]:
            return float(full["dram_bytes_per_launch"]), os.path.relpath(path, ROOT)
    except Exception:
        pass
    return None, None


def build_problem(workload, seed):
    import dto_b200 as dto

    kw = dict(WORKLOADS[workload])
    kind = kw.pop("kind")
    kw.pop("batch", None)
    if kind == "gate":
        return dto.problem_templates.quantum_gate_problem(seed=seed, **kw)
    return dto.problem_templates.scaled_problem(seed=seed, **kw)


def workload_config(workload, prob, mode):
    """The `config` both arms print (identical strings: the driver pairs the two lines by it).  Sizes are the
    closed-form counts of SURVEY.md section 8a (the bench workloads have no knot constraints)."""
    t = prob.trajectory
    n, m, z, N = t.dims["x"], t.dims["u"], t.dim, t.N
    dsum = sum(i.x_dim for i in prob.integrators)
    rows, jac = (N - 1) * dsum, (N - 1) * dsum * 2 * z
    hess = N * z * (z + 1) // 2 + (N - 1) * z * z
    batch = WORKLOADS[workload].get("batch")
    return {
        "workload": f"{workload}: " + ("bilinear quantum gate" if mode == "replicas" else "random dense bilinear (make_scaled_problem)") +
                    f", state dim {n}, {m} drives, N={N}" + (f", batch {batch}" if batch else "") +
                    f" (z={z}; per problem: {N * z} vars, {rows} rows, {jac} Jac nnz, {hess} Hess nnz)",
        "per_rank": {"replicas": "one independent problem per GPU (problem-parallel, no collective)",
                     "knot_shards": "contiguous knot range of ONE trajectory per GPU; one-knot halo read through a CUDA-IPC peer pointer (NVLink) inside the kernels; NCCL all-reduce of 2 scalars (objective, violation) per step",
                     "batch_split": "contiguous block of the problem batch per GPU (no collective)"}[mode],
        "step": "objective+gradient+constraint+Jacobian+Hessian of one iterate",
    }


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_port_rate(prob, sample_intervals, threads):
    """Oracle port timed on the host: full constraint+Jacobian+Hessian work of `sample_intervals`
    knot intervals (the reference's cost is linear in N: serial knot loop,
    src/integrators/bilinear_integrator.jl:99,113,142), scaled to evaluations/s of the whole problem."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dto_oracle_c as oc

    return oc.time_port(prob, sample_intervals, threads)


def run_reference(args, rank, world):
    if rank != 0:
        return
    prob = build_problem(args.workload, seed=42)
    threads = os.cpu_count() or 1
    vals = []
    sample = args.ref_sample
    for i in range(args.warmup + args.steps):
        rate, info = cpu_port_rate(prob, sample, threads)
        if i >= args.warmup:
            vals.append(rate)
    v = float(np.mean(vals))
    t = prob.trajectory
    line = {
        "impl": "reference", "metric": "full NLP evals/sec (constraint+Jacobian+Hessian)", "value": v, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, prob, {"c2": "replicas", "c4": "knot_shards", "c5": "batch_split"}[args.workload]),
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": info["threads"], "kind": "port", "sample": info["sample"]},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Julia (absent here and on the GPU box); this arm times oracle/dto_oracle.c, a C restatement of the reference's ForwardDiff-through-expv algorithm, POSIX threads over knot intervals, all host cores",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="knot intervals timed for cpu_baseline (about 30 CPU-seconds at c2)")
    ap.add_argument("--ref-sample", type=int, default=64, help="knot intervals per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-register", action="store_true", help="e2e with plain (unregistered) output buffers")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    # host threads of the handle (they write the Hessian's structural zeros into the caller's buffer on the host-pointer
    # path): share the box's cores between the ranks; below 3 the library delivers the whole array over PCIe instead
    os.environ.setdefault("DTO_B200_HOST_THREADS", str(max(1, min(4, (os.cpu_count() or 4) // (2 * world)))))

    import torch
    import torch.distributed as dist

    import dto_b200 as dto

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the evaluator has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from dto_b200.sharding import ShardedEvaluator, split_batch

    mode = {"c2": "replicas", "c4": "knot_shards", "c5": "batch_split"}[args.workload]
    prob = build_problem(args.workload, seed=42 + (rank if mode == "replicas" else 0))
    t = prob.trajectory
    n, m = t.dims["x"], t.dims["u"]
    rng = np.random.default_rng(1234 + rank)
    sharded = None
    batch_local = 1
    if mode == "replicas":  # an independent problem per rank
        ev = dto.Evaluator(prob, device=local_rank)
        Z = t.datavec.copy()
    elif mode == "knot_shards":
        sharded = ShardedEvaluator(prob, rank, world, device=local_rank, dist=dist if world > 1 else None)
        ev = sharded.local
        Z = sharded.local_slice(t.datavec)
    else:
        b0, b1 = split_batch(WORKLOADS[args.workload]["batch"], world)[rank]
        batch_local = b1 - b0
        ev = dto.Evaluator(prob, device=local_rank, batch=batch_local)
        Z = np.tile(t.datavec, batch_local) + 0.01 * rng.standard_normal(batch_local * ev.n_vars)
    mu = rng.random(batch_local * ev.n_constraints)
    sigma = 1.0
    n_grad = batch_local * (ev.shard_layout.z_end - ev.shard_layout.z_begin)

    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(ev.stream, device=dev)
    dZ = torch.from_numpy(Z).to(dev)
    dmu = torch.from_numpy(mu).to(dev)
    dJ = torch.empty(batch_local, dtype=torch.float64, device=dev)
    dviol = torch.zeros(batch_local, dtype=torch.float64, device=dev)
    dgrad = torch.empty(n_grad, dtype=torch.float64, device=dev)
    dg = torch.empty(batch_local * ev.n_constraints, dtype=torch.float64, device=dev)
    djac = torch.empty(batch_local * ev.nnz_jacobian, dtype=torch.float64, device=dev)
    dhess = torch.empty(batch_local * ev.nnz_hessian, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2
    zptr = dZ.data_ptr()
    if sharded is not None:
        # the shard's iterate lives in the evaluator's own buffer: that is what the left neighbour's kernels
        # read the halo knot from (CUDA-IPC mapped peer pointer over NVLink)
        zptr = ev.local_Z_ptr
        ev.eval_objective(Z)  # uploads Z into the resident buffer
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    def step_dev():
        ev.eval_all_dev(zptr, sigma, dmu.data_ptr(), dJ.data_ptr(), dgrad.data_ptr(), dg.data_ptr(), djac.data_ptr(), dhess.data_ptr())
        if sharded is not None and world > 1:
            # the only collective of the sharded path: objective (sum) and violation (max), two scalars
            ev.violation_dev(dg.data_ptr(), dviol.data_ptr())
            dist.all_reduce(dJ, op=dist.ReduceOp.SUM)
            dist.all_reduce(dviol, op=dist.ReduceOp.MAX)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            flush.zero_()
            step_dev()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ev.launch_count
    ev.kernel_timing(True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for e0, e1 in evs:
            flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
            e0.record(stream)
            step_dev()
            e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    k1_ms, k1_n = ev.kernel_time_ms()
    ev.kernel_timing(False)
    launches = ev.launch_count - l0
    dev_ms = float(np.sum(step_ms))

    # ---- end-to-end through the host-pointer C ABI ------------------------------------------------
    # Host buffers as a solver holds them: page-locked Z / mu inputs, its own Jacobian / Hessian value arrays registered
    # once with the handle (dto_register_outputs: structural constants written once, value-dependent entries per call).
    # Every timed step is a NEW iterate (three iterates in rotation), so the upload of Z and mu is part of every step.
    n_iter = 3
    hZs = [torch.from_numpy(Z + 1e-3 * i * rng.standard_normal(Z.size)).pin_memory() for i in range(n_iter)]
    hmus = [torch.from_numpy(rng.random(mu.size)).pin_memory() for _ in range(n_iter)]
    hZs[0].copy_(torch.from_numpy(Z))
    hmus[0].copy_(torch.from_numpy(mu))
    hJ = torch.empty(batch_local, dtype=torch.float64).pin_memory()
    hgrad = torch.empty(n_grad, dtype=torch.float64).pin_memory()
    hg = torch.empty(batch_local * ev.n_constraints, dtype=torch.float64).pin_memory()
    hjac = torch.empty(batch_local * ev.nnz_jacobian, dtype=torch.float64).pin_memory()
    hhess = torch.empty(batch_local * ev.nnz_hessian, dtype=torch.float64).pin_memory()
    nzs, nmus = [a.numpy() for a in hZs], [a.numpy() for a in hmus]
    outs = [a.numpy() for a in (hJ, hgrad, hg, hjac, hhess)]
    registered = False
    if batch_local == 1 and not args.no_register:
        ev.register_outputs(outs[3], outs[4])
        registered = True

    def step_fused(i):
        ev.eval_all(nzs[i % n_iter], sigma, nmus[i % n_iter], *outs)  # synchronises the stream before returning

    def step_sequence(i):
        # the five MOI callbacks as Ipopt / MadNLP issue them on one iterate (src/solvers/ipopt_solver/solver.jl:85)
        zi, mi = nzs[i % n_iter], nmus[i % n_iter]
        if batch_local == 1:
            outs[0][0] = ev.eval_objective(zi)
        else:
            outs[0][:] = ev.eval_objective(zi)
        ev.eval_objective_gradient(outs[1], zi)
        ev.eval_constraint(outs[2], zi)
        ev.eval_constraint_jacobian(outs[3], zi)
        ev.eval_hessian_lagrangian(outs[4], zi, sigma, mi)

    def time_host(step):
        for i in range(args.warmup):
            step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(i)
        barrier()
        return time.perf_counter() - t0

    e2e_fused_s = time_host(step_fused)
    d2h_fused = ev.last_d2h_bytes  # what crossed PCIe in the last fused call
    l_seq0 = ev.launch_count
    e2e_seq_s = time_host(step_sequence)
    seq_launches = (ev.launch_count - l_seq0) / (args.steps + args.warmup)
    # bytes of one sequence step: measured callback by callback on one more new iterate
    zi = nzs[0] + 1e-6
    d2h_seq = 0
    ev.eval_objective(zi)
    d2h_seq += ev.last_d2h_bytes
    ev.eval_objective_gradient(outs[1], zi)
    d2h_seq += ev.last_d2h_bytes
    ev.eval_constraint(outs[2], zi)
    d2h_seq += ev.last_d2h_bytes  # with registered outputs this includes the Jacobian that leaves on the side
    ev.eval_constraint_jacobian(outs[3], zi)
    d2h_seq += ev.last_d2h_bytes
    ev.eval_hessian_lagrangian(outs[4], zi, sigma, nmus[0])
    d2h_seq += ev.last_d2h_bytes
    sampler.stop_flag.set()
    sampler.join()
    h2d = 8 * (Z.size + mu.size)
    d2h_outputs = 8 * (hJ.numel() + hgrad.numel() + hg.numel() + hjac.numel() + hhess.numel())

    # parity guard on the timed outputs: the device-resident and host paths (fused and callback sequence) must agree bit for bit
    step_sequence(0)
    same_seq = bool(np.array_equal(outs[3], djac.cpu().numpy()) and np.array_equal(outs[4], dhess.cpu().numpy()) and
                    np.array_equal(outs[2], dg.cpu().numpy()))
    step_fused(0)
    same = bool(np.array_equal(outs[3], djac.cpu().numpy()) and np.array_equal(outs[4], dhess.cpu().numpy()))
    if registered:
        ev.unregister_outputs()

    # ---- max over ranks ---------------------------------------------------------------------------
    agg = torch.tensor([dev_ms, e2e_seq_s, k1_ms / max(k1_n, 1), e2e_fused_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max, k1_avg_ms, e2e_fused_s_max = agg.tolist()

    if rank == 0:
        # units: problem evaluations.  replicas: one problem per rank per step (weak); knot_shards: the ranks
        # share ONE problem (strong); batch_split: the ranks share the batch (strong)
        per_step = {"replicas": world, "knot_shards": 1, "batch_split": WORKLOADS[args.workload].get("batch", 1)}[mode]
        evals = per_step * args.steps
        value = evals / (dev_ms_max * 1e-3)
        e2e_value = evals / e2e_s_max
        # algorithmic flops of ONE launch of the dominant kernel on this rank
        intervals_per_launch = {"replicas": t.N - 1, "knot_shards": int(ev.n_dynamics_constraints // max(1, sum(i.x_dim for i in prob.integrators))),
                                "batch_split": batch_local * (t.N - 1)}[mode]
        flops_launch = canonical_flops_per_interval(n, m) * intervals_per_launch
        achieved = flops_launch / (k1_avg_ms * 1e-3) * 1e-12
        traffic, traffic_src = recorded_traffic(args.workload, ev.kernel_variant(0)) if world == 1 or mode == "replicas" else (None, None)
        line = {
            "metric": "full NLP evals/sec (constraint+Jacobian+Hessian)", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak" if mode == "replicas" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.workload, prob, mode),
                           l2="256 MB memset between timed steps (outside the per-step CUDA events)",
                           outputs="value: outputs left in HBM (dto_eval_all_dev); e2e: the five callbacks (and, beside it, the fused dto_eval_all) with "
                                   "page-locked host buffers; the Jacobian / Hessian value arrays are registered with the handle once "
                                   "(dto_register_outputs: structural constants written once, only value-dependent entries cross PCIe per call; "
                                   "--no-register: plain buffers, host threads write the Hessian's structural zeros); knot-range pipeline: D2H of "
                                   "finished ranges overlaps the next range",
                           kernel_variant=ev.kernel_variant(0)),
            # headline e2e: the call sequence a solver makes (five separate C-ABI callbacks per iterate); the fused
            # single call (dto_eval_all, what benchmark/benchmarks.jl's loop amounts to) beside it
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_seq,
                    "ms_per_step": e2e_s_max / args.steps * 1e3, "matches_device_path": same_seq, "output_bytes_per_step": d2h_outputs,
                    "call": "dto_eval_objective + dto_eval_gradient + dto_eval_constraint + dto_eval_jacobian + dto_eval_hessian on a new iterate every step",
                    "outputs_registered": registered, "gpu_launches_per_step": seq_launches,
                    "fused": {"call": "dto_eval_all", "value": evals / e2e_fused_s_max, "ms_per_step": e2e_fused_s_max / args.steps * 1e3,
                              "d2h_bytes_per_step": d2h_fused, "matches_device_path": same},
                    "sequence_over_fused": e2e_s_max / e2e_fused_s_max},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write, ncu --set full)", "traffic_source": traffic_src,
                "kernel": "bilinear interval kernel (K1)", "kernel_ms": k1_avg_ms,
                "kernel_share_of_step": k1_avg_ms * max(1, sum(1 for i in prob.integrators if type(i).__name__ != "DerivativeIntegrator")) / (dev_ms_max / args.steps),
                "algorithmic_flops_per_launch": flops_launch,
                "peak_source": "FP64 DMMA m8n8k4 saturation measured on this pool (profiles/r01_fp64_peaks_and_box_probe.log); "
                               "MEASURED_PEAKS.json carries no FP64 figure; cuBLAS DGEMM 8192^3 measured 35.5 TFLOP/s",
                "hbm_gbs_step": 8 * (2 * Z.size + 2 * mu.size + ev.n_constraints + ev.nnz_jacobian + ev.nnz_hessian) / (dev_ms_max / args.steps * 1e-3) * 1e-9,
            },
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is timed on rank 0 of the single-GPU run only
            try:
                threads = os.cpu_count() or 1
                rate, info = cpu_port_rate(prob, args.cpu_sample, threads)
                line["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": info["threads"], "kind": "port", "sample": info["sample"]}
            except Exception as e:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "evals/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ev.close()


if __name__ == "__main__":
    main()
