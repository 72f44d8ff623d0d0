"""Device timeline of the linked knot-shard step (config c4 shape) on every rank: which kernels and gaps make up a step.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port 29533 tools/shard_trace.py [N]
Prints, for rank 0 and the last rank, the kernels of one steady-state step with start offsets and durations (CUPTI via
torch.profiler), and the per-step time (CUDA events, max over ranks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import dto_b200 as dto
from dto_b200 import problem_templates as pt
from dto_b200.sharding import ShardedEvaluator

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
prob = pt.scaled_problem(N=N, state_dim=16, n_controls=2, generator_scale=0.25)
sh = ShardedEvaluator(prob, rank, world, device=lr, dist=dist if world > 1 else None)
ev = sh.local
dev = torch.device("cuda", lr)
Z = sh.local_slice(prob.trajectory.datavec)
rng = np.random.default_rng(rank)
dZs = [torch.from_numpy(Z).to(dev), torch.from_numpy(Z + 1e-3 * rng.standard_normal(Z.size)).to(dev)]
dmu = torch.rand(ev.n_constraints, dtype=torch.float64, device=dev)
n_grad = ev.shard_layout.z_end - ev.shard_layout.z_begin
out = [torch.empty(k, dtype=torch.float64, device=dev) for k in (1, n_grad, ev.n_constraints, ev.nnz_jacobian, ev.nnz_hessian)]
dviol = torch.zeros(1, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.ExternalStream(ev.stream, device=dev)


def step(i):
    ev.upload_dev(dZs[i % 2].data_ptr())
    ev.eval_all_dev(ev.local_Z_ptr, 1.0, dmu.data_ptr(), *[o.data_ptr() for o in out])
    if world > 1:
        ev.shard_scalars_dev(out[2].data_ptr(), out[0].data_ptr(), dviol.data_ptr())


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


with torch.cuda.stream(stream):
    for i in range(5):
        flush.zero_()
        step(i)
sync()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
with torch.cuda.stream(stream):
    for i, (a, b) in enumerate(evs):
        flush.zero_()
        a.record(stream)
        step(i)
        b.record(stream)
sync()
ms = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in evs]))], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"N={N} world={world}: {ms.item():.4f} ms per step (max over ranks)", flush=True)

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    with torch.cuda.stream(stream):
        for i in range(4):
            flush.zero_()
            step(i)
    sync()
kern = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
# the third step: from the third flush fill to the fourth
fills = [i for i, e in enumerate(kern) if "FillFunctor" in e.name]
lines = []
if len(fills) >= 4:
    a, b = fills[2], fills[3]
    t0 = kern[a].time_range.end
    for e in kern[a + 1:b]:
        lines.append(f"   +{e.time_range.start - t0:8.1f} us  {e.time_range.end - e.time_range.start:8.1f} us  {e.name[:90]}")
    lines.append(f"   step span {kern[b - 1].time_range.end - t0:.1f} us")
for r in (0, world - 1):
    if world > 1:
        dist.barrier()
    if rank == r:
        print(f"rank {rank}:", flush=True)
        print("\n".join(lines), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
ev.close()
