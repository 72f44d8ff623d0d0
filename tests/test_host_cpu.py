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


def test_struct_layout_matches_header(tmp_path):
    # sizes and field offsets the C compiler gives the descriptor structs of include/dto_b200.h -- guards the ctypes mirror
    import subprocess

    L = dto._lib
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "dto_b200.h"\n'
        "int main(void) { printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(dto_integrator_desc), sizeof(dto_objective_desc),"
        " sizeof(dto_constraint_desc), sizeof(dto_problem_desc), sizeof(dto_size_info), sizeof(dto_shard_layout),"
        " offsetof(dto_objective_desc, gvar_offs), offsetof(dto_constraint_desc, gvar_offs), offsetof(dto_problem_desc, global_dim));"
        " return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(L.IntegratorDesc), ctypes.sizeof(L.ObjectiveDesc), ctypes.sizeof(L.ConstraintDesc), ctypes.sizeof(L.ProblemDesc),
            ctypes.sizeof(L.SizeInfo), ctypes.sizeof(L.ShardLayout), L.ObjectiveDesc.gvar_offs.offset, L.ConstraintDesc.gvar_offs.offset,
            L.ProblemDesc.global_dim.offset]
    assert got == want


def test_global_trajectory_and_components():
    """Global components sit after the knots in vec(traj); the global terms lower to one catalogue call on
    [knot vars; global vars] (global_objectives.jl, global_constraint.jl, global_knot_point_constraint.jl)."""
    prob = pt.global_problem(N=5)
    t = prob.trajectory
    assert t.global_names == ("g", "x_goal") and t.global_dim == 5 + 4
    assert t.vec().size == t.dim * t.N + t.global_dim and np.array_equal(t.vec()[t.dim * t.N :], t.global_data)
    spec = prob.to_spec()
    assert spec["global_dim"] == 9 and spec["global_components"] == {"g": (0, 5), "x_goal": (5, 4)}
    kinds = [o["kind"] for o in spec["objectives"]]
    assert kinds.count("global_knot") == 3
    gc = prob.nonlinear_constraints()[2]
    assert isinstance(gc, dto.NonlinearGlobalConstraint) and gc.dim == 1 and gc.var_dim == 0 and gc.global_dim == 5
    gk = prob.nonlinear_constraints()[1]
    assert gk.dim == 2 * t.N and gk.combined_dim == 2 + 5
    auto = dto.GlobalKnotPointObjective(dto.NormSqPlus(0.0), ["u"], None, t, times=[1])
    assert auto.global_names == ["g", "x_goal"]
    t2 = t.copy_with(np.arange(t.vec().size, dtype=float))
    assert np.array_equal(t2.global_data, np.arange(t.dim * t.N, t.vec().size))
    # KnotHVP carriers: the trait defaults to nothing (knot_hvp.jl:148)
    ob = dto.KnotPointObjective(dto.NormSqPlus(0.0), "u", t)
    assert dto.knot_hvp(ob, t) is None
    ob.knot_hvp = dto.ConstantLowRankHVP(np.eye(2), "neg2_sign")
    assert dto.knot_hvp(ob, t).core == "neg2_sign" and isinstance(dto.knot_hvp(ob, t), dto.KnotHVP)


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


def _build_c_smoke(tmp_path):
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib_dir = os.path.join(ROOT, "directtrajopt.jl_b200", "lib")
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_smoke.c"),
                    "-L", lib_dir, "-ldto_b200", "-lm", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    return exe


def test_plain_c_consumer_links_and_fails_loudly_without_a_device(tmp_path):
    """tests/c_abi_smoke.c: a C99 program builds the descriptor by hand and calls the ABI -- no Python, no C++.  Without a
    CUDA device dto_create must fail with DTO_ERR_CUDA (exit code 77: there is no CPU fallback); with one, the closed-form
    checks must pass (exit code 0)."""
    import subprocess

    import torch

    exe = _build_c_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == (0 if torch.cuda.is_available() else 77), r.stdout + r.stderr


@pytest.mark.gpu
def test_plain_c_consumer_on_the_gpu(tmp_path):
    import subprocess

    exe = _build_c_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "worst abs error" in r.stdout
