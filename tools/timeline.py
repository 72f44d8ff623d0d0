"""Device timeline (CUPTI through torch.profiler) of one steady-state device-resident evaluation of a BASELINE config:
every kernel with its start offset and duration -- where the step's time outside the interval kernel goes.
usage: python tools/timeline.py c2|c3|c4|c5"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dto_b200 as dto
from dto_b200 import problem_templates as pt

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
batch = 1
if which == "c2": prob = pt.quantum_gate_problem(N=2000, levels=16, n_drives=4)
elif which == "c3": prob = pt.carrier_problem(N=1000, state_dim=64, n_drives=2)
elif which == "c4": prob = pt.scaled_problem(N=100000, state_dim=16, n_controls=2, generator_scale=0.25)
else: prob, batch = pt.scaled_problem(N=200, state_dim=8, n_controls=2, generator_scale=0.35), 4096
dev = torch.device("cuda")
ev = dto.Evaluator(prob, batch=batch)
rng = np.random.default_rng(0)
dZ = torch.from_numpy(np.tile(prob.trajectory.datavec, batch) + 0.01 * rng.standard_normal(batch * ev.n_vars)).to(dev)
dmu = torch.rand(batch * ev.n_constraints, dtype=torch.float64, device=dev)
out = [torch.empty(batch * k, dtype=torch.float64, device=dev) for k in (1, ev.n_vars, ev.n_constraints, ev.nnz_jacobian, ev.nnz_hessian)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.ExternalStream(ev.stream)
step = lambda: ev.eval_all_dev(dZ.data_ptr(), 1.0, dmu.data_ptr(), *[o.data_ptr() for o in out])
with torch.cuda.stream(stream):
    for _ in range(4):
        flush.zero_(); step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    with torch.cuda.stream(stream):
        for _ in range(4):
            flush.zero_(); step()
    torch.cuda.synchronize()
kern = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
fills = [i for i, e in enumerate(kern) if "FillFunctor" in e.name]
a, b = fills[2], fills[3]
t0 = kern[a].time_range.end
print(f"{which}: kernels of one evaluation (offsets from the end of the L2 flush)")
for e in kern[a + 1:b]:
    print(f"   +{e.time_range.start - t0:8.1f} us  {e.time_range.end - e.time_range.start:8.1f} us  {e.name[:100]}")
print(f"   span {max(e.time_range.end for e in kern[a + 1:b]) - t0:.1f} us")
ev.close()
