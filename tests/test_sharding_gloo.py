"""world_size-2 (and 3) gloo runs of the knot-range sharding host logic on CPU: the partition tiles the
trajectory, every row/knot has exactly one owner, and the only collective -- the all-reduce of the
objective (sum) and violation (max) -- reproduces the whole-problem values.  The local evaluator is an
oracle-backed stand-in (the CUDA evaluator needs a GPU; the same ShardedEvaluator drives it in
tests/test_multi_gpu.py and bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dto_b200 as dto
import dto_oracle as orc
from dto_b200 import problem_templates as pt
from dto_b200.sharding import ShardedEvaluator, knot_ranges, split_batch


class OracleShard:
    """Stand-in local evaluator: shard-local objective and violation from the CPU oracle."""

    def __init__(self, prob, shard, device):
        self.spec, self.k0, self.k1 = prob.to_spec(), shard[0], shard[1]

    def objective_and_violation(self, Z):
        per_knot = orc.objective_by_knot(self.spec, Z)
        g = orc.eval_constraint(self.spec, Z)
        v, _ = orc.violation(self.spec, g)
        own = orc.row_owner_knot(self.spec)
        mine = (own >= self.k0) & (own <= self.k1)
        return per_knot[self.k0 - 1 : self.k1].sum(), (v[mine].max() if mine.any() else 0.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = pt.standard_problem(N=11, seed=1)
    Z = prob.trajectory.datavec + 0.01 * np.random.default_rng(0).standard_normal(prob.trajectory.datavec.size)
    sh = ShardedEvaluator(prob, rank, world, dist=dist, evaluator_factory=lambda p, shard, device: OracleShard(p, shard, device))
    Jl, vl = sh.local.objective_and_violation(Z)
    J, v = sh.reduce_scalars(Jl, vl)
    sl = sh.local_slice(Z)
    q.put((rank, sh.k0, sh.k1, J, v, sl.size, sh.z_begin, sh.z_halo_end))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scalar_reduction_under_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob = pt.standard_problem(N=11, seed=1)
    spec = prob.to_spec()
    Z = prob.trajectory.datavec + 0.01 * np.random.default_rng(0).standard_normal(prob.trajectory.datavec.size)
    Jref = orc.eval_objective(spec, Z)
    vref = orc.violation(spec, orc.eval_constraint(spec, Z))[0].max()
    z = spec["z"]
    k = 1
    for rank, k0, k1, J, v, nslice, zb, zh in res:
        assert k0 == k and k1 >= k0
        k = k1 + 1
        assert abs(J - Jref) <= 1e-12 * max(1, abs(Jref)) and v == vref  # every rank holds the reduced scalars
        assert zb == (k0 - 1) * z and zh == min(k1 + 1, 11) * z and nslice == zh - zb
    assert k == 12


def test_knot_ranges_balance_and_tile():
    for N, world in [(2000, 8), (100000, 8), (11, 3), (9, 8), (50, 1)]:
        r = knot_ranges(N, world)
        assert r[0][0] == 1 and r[-1][1] == N
        for (a0, a1), (b0, b1) in zip(r[:-1], r[1:]):
            assert b0 == a1 + 1
        owned_intervals = [min(k1, N - 1) - k0 + 1 for k0, k1 in r]
        assert sum(owned_intervals) == N - 1 and max(owned_intervals) - min(owned_intervals) <= 1
    with pytest.raises(ValueError):
        knot_ranges(5, 5)


def test_split_batch():
    assert split_batch(4096, 8) == [(i * 512, (i + 1) * 512) for i in range(8)]
    s = split_batch(10, 4)
    assert s[0] == (0, 3) and s[-1][1] == 10 and sum(b - a for a, b in s) == 10
