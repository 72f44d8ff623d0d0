"""c5 shape (n = 8, N = 200) at the per-rank batch sizes of a 1/2/4/8-GPU split, octet phase split on/off: K1 time per problem."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt

dev = torch.device("cuda")
prob = pt.scaled_problem(N=200, state_dim=8, n_controls=2, generator_scale=0.35)
for split in (None, "0", "1"):
    if split is None: os.environ.pop("DTO_B200_OCTET_SPLIT", None)
    else: os.environ["DTO_B200_OCTET_SPLIT"] = split
    for B in (512, 1024, 2048, 4096):
        ev = dto.Evaluator(prob, batch=B)
        rng = np.random.default_rng(0)
        Z = np.tile(prob.trajectory.datavec, B) + 0.01 * rng.standard_normal(B * ev.n_vars)
        dZ = torch.from_numpy(Z).to(dev)
        dmu = torch.rand(B * ev.n_constraints, dtype=torch.float64, device=dev)
        outs = [torch.empty(B * k, dtype=torch.float64, device=dev) for k in (1, ev.n_vars, ev.n_constraints, ev.nnz_jacobian, ev.nnz_hessian)]
        flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
        stream = torch.cuda.ExternalStream(ev.stream)
        step = lambda: ev.eval_all_dev(dZ.data_ptr(), 1.0, dmu.data_ptr(), *[o.data_ptr() for o in outs])
        with torch.cuda.stream(stream):
            for _ in range(3): step()
        ev.synchronize()
        ev.kernel_timing(True)
        tot = 0.0
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                flush.zero_(); e0.record(stream); step(); e1.record(stream)
            ev.synchronize(); tot += e0.elapsed_time(e1)
        k1, nk = ev.kernel_time_ms()
        print(f"split={split} batch={B}: step {tot/8*1e3:.0f} us  K1 {k1/nk*1e3:.0f} us  K1/problem {k1/nk/B*1e6:.1f} ns", flush=True)
        ev.close(); del dZ, dmu, outs, flush; torch.cuda.empty_cache()
