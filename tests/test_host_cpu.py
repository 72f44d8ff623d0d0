"""CPU-only checks: the C-ABI library loads and exports every symbol include/dto_b200.h declares
(no compute calls without a GPU), and the host-side mirror lowers problems correctly."""
import ctypes
import os
import re

import numpy as np
import pytest

import dto_b200 as dto
from dto_b200 import problem_templates as pt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dto_b200.h")).read()
    declared = set(re.findall(r"\b(dto_[a-zA-Z_]+)\s*\(", header))
    declared -= {"dto_status"}
    lib = dto._lib.load()
    bound = {name for name, _, _ in dto._lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dto_abi_version() == dto._lib.ABI_VERSION


def test_create_without_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dto.DtoError) as e:
        dto.Evaluator(pt.readme_problem(N=4))
    assert "no CPU fallback" in str(e.value) or e.value.code == dto._lib.DTO_ERR_CUDA


def test_struct_layout_matches_header():
    # sizes the C compiler gives the descriptor structs (x86-64 SysV) -- guards the ctypes mirror
    L = dto._lib
    assert ctypes.sizeof(L.IntegratorDesc) == 8 * 4 + 8 + 8 + 7 * 8 + 8
    assert ctypes.sizeof(L.ObjectiveDesc) == 8 + 8 + 8 + 4 * 8 + 8 + 8 + 2 * 8
    assert ctypes.sizeof(L.ConstraintDesc) == 16 + 16 + 8 + 8
    assert ctypes.sizeof(L.SizeInfo) == 48


def test_generator_lowering_and_spec():
    prob = pt.standard_problem(N=6)
    it = prob.integrators[0]
    assert it.G.shape == (3, 4, 4)
    assert np.allclose(it.G[0], 0.1 * pt.GZ) and np.allclose(it.G[1], pt.GX) and np.allclose(it.G[2], pt.GY)
    spec = prob.to_spec()
    assert spec["components"] == {"x": (0, 4), "u": (4, 2), "du": (6, 2), "ddu": (8, 2), "dt": (10, 1)}
    assert [o["kind"] for o in spec["objectives"]] == ["knot", "quadreg", "quadreg", "mintime"]
    assert spec["constraints"][0]["times"] == list(range(2, 6))


def test_objective_algebra_flattens_like_the_reference():
    _, traj = pt.bilinear_dynamics_and_trajectory(N=4)
    a, b = dto.QuadraticRegularizer("u", traj, 1.0), dto.MinimumTimeObjective(traj)
    c = 2.0 * (a + b) + 0.5 * a
    assert isinstance(c, dto.CompositeObjective)
    assert c.weights == [2.0, 2.0, 0.5] and len(c.objectives) == 3


def test_trajectory_layout_is_knot_major():
    _, traj = pt.bilinear_dynamics_and_trajectory(N=4)
    Z = traj.datavec
    assert Z.size == traj.dim * traj.N
    assert np.array_equal(Z[traj.dim : 2 * traj.dim], traj.data[:, 1])
    assert traj[2].timestep == traj.data[traj.components["dt"].start, 1]
