"""Device-resident evaluation time of every BASELINE.json configuration on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt

which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
dev = torch.device("cuda")


def run(name, prob, batch=1, reps=5):
    t0 = time.time()
    ev = dto.Evaluator(prob, batch=batch)
    t_create = time.time() - t0
    rng = np.random.default_rng(0)
    Z = np.tile(prob.trajectory.datavec, batch) + 0.01 * rng.standard_normal(batch * ev.n_vars)
    dZ = torch.from_numpy(Z).to(dev)
    dmu = torch.rand(batch * ev.n_constraints, dtype=torch.float64, device=dev)
    dJ = torch.empty(batch, dtype=torch.float64, device=dev)
    dgrad = torch.empty(batch * ev.n_vars, dtype=torch.float64, device=dev)
    dg = torch.empty(batch * ev.n_constraints, dtype=torch.float64, device=dev)
    djac = torch.empty(batch * ev.nnz_jacobian, dtype=torch.float64, device=dev)
    dhess = torch.empty(batch * ev.nnz_hessian, dtype=torch.float64, device=dev)
    stream = torch.cuda.ExternalStream(ev.stream)
    def step():
        ev.eval_all_dev(dZ.data_ptr(), 1.0, dmu.data_ptr(), dJ.data_ptr(), dgrad.data_ptr(), dg.data_ptr(), djac.data_ptr(), dhess.data_ptr())
    step(); ev.synchronize()
    ev.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            step()
        e1.record(stream)
    ev.synchronize()
    ms = e0.elapsed_time(e1) / reps
    k1, nk = ev.kernel_time_ms()
    nbytes = 8 * batch * (ev.n_constraints + ev.nnz_jacobian + ev.nnz_hessian)
    print(f"{name}: batch={batch} vars={ev.n_vars} rows={ev.n_constraints} jac={ev.nnz_jacobian} hess={ev.nnz_hessian} "
          f"create={t_create:.2f}s eval={ms:.3f} ms ({batch/ms*1e3:.1f} problem-evals/s) interval-kernels={k1/max(nk,1)*len(prob.integrators):.3f} ms "
          f"out={nbytes/1e6:.1f} MB -> {nbytes/ms/1e6:.0f} GB/s variant={ev.kernel_variant(0)}", flush=True)
    ev.close()


if "c1" in which:
    run("c1 README n=2 N=50", pt.readme_problem(N=50), reps=50)
if "c2" in which:
    run("c2 gate n=32 m=4 N=2000", pt.quantum_gate_problem(N=2000, levels=16, n_drives=4), reps=10)
if "c3" in which:
    run("c3 TDBI n=64 m=2 N=1000", pt.carrier_problem(N=1000, state_dim=64, n_drives=2), reps=2)
if "c4" in which:
    run("c4 long n=16 m=2 N=100000", pt.scaled_problem(N=100000, state_dim=16, n_controls=2, generator_scale=0.25), reps=3)
if "c5" in which:
    run("c5 batch 4096 x (n=8, N=200)", pt.scaled_problem(N=200, state_dim=8, n_controls=2, generator_scale=0.35), batch=4096, reps=3)
