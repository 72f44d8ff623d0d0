"""Problem recipes: the reference's test/benchmark fixtures and the BASELINE.json configurations,
built through this package's mirror of the reference API with synthetic (seeded NumPy PCG64) data.

  readme_problem            README.md:68-95 (BASELINE config c1)
  standard_problem          test/test_snippets.jl:27-54 + test/test_utils.jl:113-178 (evaluator test problem)
  evaluator_test_problem    src/solvers/evaluator.jl:656-684
  linear_regularizer_problem  standard problem + LinearRegularizer terms (src/objectives/regularizers.jl:207-313)
  bilinear_benchmark        benchmark/problem_utils.jl:10-42  (published micro-benchmark shape, N=51)
  scaled_problem            benchmark/problem_utils.jl:49-77  (BASELINE config c5 shape, c4 shape)
  quantum_gate_problem      BASELINE config c2: isomorphic state dim 32, 4 drives, free dt + MinimumTime
  carrier_problem           BASELINE config c3: TimeDependentBilinear + Derivative chain + knot constraints
  global_problem            the standard problem with a global component and the reference's four global test
                            terms (global_objectives.jl:393-475, global_constraint.jl:162-180,
                            global_knot_point_constraint.jl:260-288)
"""
from __future__ import annotations

import numpy as np

from .components import (BilinearIntegrator, CarrierGenerator, DerivativeIntegrator, GlobalKnotPointObjective, GlobalObjective,
                         LinearRegularizer, MinimumTimeObjective, NonlinearGlobalConstraint, NonlinearGlobalKnotPointConstraint,
                         NonlinearKnotPointConstraint, NormMinus, NormProduct, NormSqPlus, QuadraticRegularizer, SplitSqDist, SqDist,
                         TerminalObjective, TimeDependentBilinearIntegrator)
from .evaluator import DirectTrajOptProblem
from .trajectory import NamedTrajectory

GX = np.array([[0, 0, 0, 1], [0, 0, 1, 0], [0, -1, 0, 0], [-1, 0, 0, 0]], float)
GY = np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 0, -1], [0, 0, 1, 0]], float)
GZ = np.array([[0, 0, 1, 0], [0, 0, 0, -1], [-1, 0, 0, 0], [0, 1, 0, 0]], float)


def readme_problem(N=50, seed=42):
    rng = np.random.default_rng(seed)
    traj = NamedTrajectory({"x": rng.standard_normal((2, N)), "u": rng.standard_normal((1, N)), "dt": np.full(N, 0.1)},
                           timestep="dt", controls=("u",), initial={"x": [0.0, 0.0]}, final={"x": [1.0, 0.0]})
    G_drift = np.array([[-0.1, 1.0], [-1.0, -0.1]])
    G_drives = [np.array([[0.0, 1.0], [1.0, 0.0]])]
    G = lambda u: G_drift + sum(ui * Gi for ui, Gi in zip(u, G_drives))
    integrator = BilinearIntegrator(G, "x", "u", traj)
    obj = QuadraticRegularizer("u", traj, 1.0)
    return DirectTrajOptProblem(traj, obj, integrator)


def bilinear_dynamics_and_trajectory(N=10, dt=0.1, u_bound=0.1, omega=0.1, add_time=False, seed=0, add_global=False, global_dim=None):
    rng = np.random.default_rng(seed)
    G = lambda u: omega * GZ + u[0] * GX + u[1] * GY
    comps = {
        "x": 2 * rng.random((4, N)) - 1,
        "u": u_bound * (2 * rng.random((2, N)) - 1),
        "du": rng.standard_normal((2, N)),
        "ddu": rng.standard_normal((2, N)),
        "dt": np.full(N, dt),
    }
    if add_time:
        comps["t"] = np.arange(N) * dt
    glob = None
    if add_global:  # add_component(traj, :g, randn(N), type = :global)  (test/test_utils.jl:173-175)
        glob = {"g": rng.standard_normal(N if global_dim is None else global_dim)}
    traj = NamedTrajectory(comps, controls=("ddu", "dt"), timestep="dt", bounds={"u": u_bound, "dt": (0.01, 0.5)},
                           initial={"x": [1.0, 0, 0, 0], "u": [0, 0]}, final={"u": [0, 0]}, goal={"x": [0.0, 1.0, 0, 0]},
                           global_components=glob)
    return G, traj


def standard_problem(N=10, seed=0):
    G, traj = bilinear_dynamics_and_trajectory(N=N, seed=seed)
    integrators = [BilinearIntegrator(G, "x", "u", traj), DerivativeIntegrator("u", "du", traj), DerivativeIntegrator("du", "ddu", traj)]
    J = TerminalObjective(SqDist(traj.goal["x"]), "x", traj)
    J = J + QuadraticRegularizer("u", traj, 1.0)
    J = J + QuadraticRegularizer("du", traj, 1.0)
    J = J + MinimumTimeObjective(traj)
    g_u_norm = NonlinearKnotPointConstraint(NormMinus(1.0), "u", traj, times=range(2, traj.N), equality=False)
    return DirectTrajOptProblem(traj, J, integrators, constraints=[g_u_norm])


def global_problem(N=10, seed=0, global_dim=None, with_goal=True):
    """The standard problem plus a global component ``g`` and every kind of global term, with the functions of the
    reference's own tests: GlobalObjective norm(g)^2 (Q = 2), GlobalKnotPointObjective norm(u)^2 + norm(g)^2 at
    times [1, N] (Qs [1, 2]), NonlinearGlobalConstraint norm(g) - 1 <= 0, NonlinearGlobalKnotPointConstraint
    [norm(u) - 1; norm(u) norm(g) - 1] <= 0 at every knot; and the documented terminal objective
    norm(x_N - x_goal)^2 with the goal held in a second global (global_objectives.jl:364-369)."""
    rng = np.random.default_rng(seed + 100)
    G, traj = bilinear_dynamics_and_trajectory(N=N, seed=seed, add_global=True, global_dim=global_dim)
    if with_goal:
        glob = {"g": traj.global_data.copy(), "x_goal": rng.standard_normal(4)}
        _, traj = bilinear_dynamics_and_trajectory(N=N, seed=seed)
        traj = NamedTrajectory({n: traj.data[traj.components[n], :] for n in traj.names}, controls=("ddu", "dt"), timestep="dt",
                               goal=traj.goal, global_components=glob)
    integrators = [BilinearIntegrator(G, "x", "u", traj), DerivativeIntegrator("u", "du", traj), DerivativeIntegrator("du", "ddu", traj)]
    J = QuadraticRegularizer("u", traj, 1.0) + MinimumTimeObjective(traj)
    J = J + GlobalObjective(NormSqPlus(0.0), "g", traj, Q=2.0)
    J = J + 0.5 * GlobalKnotPointObjective(NormSqPlus(0.0), ["u"], ["g"], traj, times=[1, traj.N], Qs=[1.0, 2.0])
    if with_goal:
        J = J + TerminalObjective(SplitSqDist(), "x", traj, global_names="x_goal", Q=100.0)
    cons = [
        NonlinearKnotPointConstraint(NormMinus(1.0), "u", traj, times=range(2, traj.N), equality=False),
        NonlinearGlobalKnotPointConstraint(NormProduct(2, 1.0, 1.0), ["u"], ["g"], traj, equality=False),
        NonlinearGlobalConstraint(NormMinus(1.0), "g", traj, equality=False),
    ]
    return DirectTrajOptProblem(traj, J, integrators, constraints=cons)


def global_goal_problem(N=11, seed=0):
    """A well-posed problem with global variables for the solver-in-the-loop check: the state has to reach a goal that is
    itself a decision variable on the unit sphere (global ``x_goal``: terminal objective norm(x_N - x_goal)^2 as in
    global_objectives.jl:364-369, pulled towards a target by a GlobalObjective, norm(x_goal)^2 = 1 as a
    NonlinearGlobalConstraint), and the drives stay in a ball whose radius is set by a second global ``gam``
    (NonlinearGlobalKnotPointConstraint norm([u; gam])^2 <= c)."""
    from .components import NormSqMinus

    G, traj0 = bilinear_dynamics_and_trajectory(N=N, seed=seed)
    glob = {"x_goal": np.array([0.1, 0.9, 0.1, 0.05]), "gam": np.array([0.1])}  # no exact zeros: the stored Jacobian pattern is value-dependent
    traj = NamedTrajectory({n: traj0.data[traj0.components[n], :] for n in traj0.names}, controls=("ddu", "dt"), timestep="dt",
                           bounds=traj0.bounds, initial=traj0.initial, final=traj0.final, goal=traj0.goal, global_components=glob)
    integrators = [BilinearIntegrator(G, "x", "u", traj), DerivativeIntegrator("u", "du", traj), DerivativeIntegrator("du", "ddu", traj)]
    J = QuadraticRegularizer("u", traj, 1.0) + QuadraticRegularizer("du", traj, 1.0) + QuadraticRegularizer("ddu", traj, 1.0)
    J = J + TerminalObjective(SplitSqDist(), "x", traj, global_names="x_goal", Q=100.0)
    J = J + GlobalObjective(SqDist(np.array([0.0, 0.8, 0.6, 0.0])), "x_goal", traj, Q=10.0)
    J = J + GlobalObjective(SqDist(np.array([0.3])), "gam", traj, Q=1.0)
    cons = [
        NonlinearGlobalConstraint(NormSqMinus(1.0), "x_goal", traj, equality=True),
        NonlinearGlobalKnotPointConstraint(NormSqMinus(0.06), ["u"], ["gam"], traj, times=range(2, traj.N), equality=False),
    ]
    return DirectTrajOptProblem(traj, J, integrators, constraints=cons)


def linear_regularizer_problem(N=8, seed=1):
    """The standard problem with LinearRegularizer terms (regularizers.jl:207-313): vector R, a subset of
    times, a scaled term inside the composite."""
    G, traj = bilinear_dynamics_and_trajectory(N=N, seed=seed)
    integrators = [BilinearIntegrator(G, "x", "u", traj), DerivativeIntegrator("u", "du", traj), DerivativeIntegrator("du", "ddu", traj)]
    J = TerminalObjective(SqDist(traj.goal["x"]), "x", traj)
    J = J + LinearRegularizer("u", traj, [0.3, -0.7])
    J = J + 2.5 * LinearRegularizer("ddu", traj, 0.4, times=[1, 3, N])
    J = J + QuadraticRegularizer("du", traj, 1.0)
    return DirectTrajOptProblem(traj, J, integrators)


def evaluator_test_problem(N=10, seed=0):
    G, traj = bilinear_dynamics_and_trajectory(N=N, seed=seed)
    integrators = [BilinearIntegrator(G, "x", "u", traj), DerivativeIntegrator("u", "du", traj), DerivativeIntegrator("du", "ddu", traj)]
    J = TerminalObjective(SqDist(traj.goal["x"]), "x", traj)
    J = J + QuadraticRegularizer("u", traj, 2.0e-1)
    J = J + QuadraticRegularizer("du", traj, 3.0e-1)
    J = J + QuadraticRegularizer("ddu", traj, 4.0e-1)
    J = J + MinimumTimeObjective(traj)
    g_u_norm = NonlinearKnotPointConstraint(NormMinus(1.0), "u", traj, times=range(2, traj.N), equality=False)
    return DirectTrajOptProblem(traj, J, integrators, constraints=[g_u_norm])


def bilinear_benchmark(N=51, seed=42):
    G, traj = bilinear_dynamics_and_trajectory(N=N, seed=seed)
    integrators = [BilinearIntegrator(G, "x", "u", traj), DerivativeIntegrator("u", "du", traj), DerivativeIntegrator("du", "ddu", traj)]
    J = QuadraticRegularizer("u", traj, 1.0) + QuadraticRegularizer("du", traj, 1.0)
    return DirectTrajOptProblem(traj, J, integrators)


def scaled_problem(N, state_dim, n_controls=2, seed=42, generator_scale=1.0):
    """make_scaled_problem: random dense generators, x, u, du, dt; Bilinear + Derivative(u, du); QuadReg(u)."""
    rng = np.random.default_rng(seed)
    G_drift = generator_scale * rng.standard_normal((state_dim, state_dim))
    G_drives = [generator_scale * rng.standard_normal((state_dim, state_dim)) for _ in range(n_controls)]
    x_init = np.zeros(state_dim)
    x_init[0] = 1.0
    x_goal = np.zeros(state_dim)
    x_goal[min(1, state_dim - 1)] = 1.0
    traj = NamedTrajectory(
        {"x": rng.standard_normal((state_dim, N)), "u": 0.1 * rng.standard_normal((n_controls, N)),
         "du": rng.standard_normal((n_controls, N)), "dt": np.full(N, 0.1)},
        controls=("du", "dt"), timestep="dt", bounds={"u": 1.0, "dt": (0.01, 0.5)},
        initial={"x": x_init, "u": np.zeros(n_controls)}, final={"u": np.zeros(n_controls)}, goal={"x": x_goal})
    integrators = [BilinearIntegrator((G_drift, G_drives), "x", "u", traj), DerivativeIntegrator("u", "du", traj)]
    J = QuadraticRegularizer("u", traj, 1.0)
    return DirectTrajOptProblem(traj, J, integrators)


def iso_generator(H):
    """Real isomorphism of -iH acting on [re; im]: G = [[Im H, Re H], [-Re H, Im H]]."""
    return np.block([[H.imag, H.real], [-H.real, H.imag]])


def quantum_gate_problem(N=2000, levels=16, n_drives=4, seed=42, dt=0.1):
    """BASELINE config c2: a `levels`-level complex system in the real isomorphism
    (state dim 2*levels), `n_drives` drives + drift, Hermitian random Hamiltonians scaled so that
    ||dt*G(u)||_1 ~ 1; x ~ U(-1,1), u ~ 0.1 U(-1,1), free dt; QuadReg(u) + MinimumTime + terminal
    ||x - x_goal||^2."""
    rng = np.random.default_rng(seed)
    n = 2 * levels

    def herm():
        M = rng.standard_normal((levels, levels)) + 1j * rng.standard_normal((levels, levels))
        return (M + M.conj().T) / 2

    G0 = iso_generator(herm())
    Gd = [iso_generator(herm()) for _ in range(n_drives)]
    scale = 1.0 / (dt * np.abs(G0).sum(axis=0).max())
    G0 = G0 * scale
    Gd = [g * scale for g in Gd]
    x_goal = np.zeros(n)
    x_goal[1] = 1.0
    traj = NamedTrajectory(
        {"x": 2 * rng.random((n, N)) - 1, "u": 0.1 * (2 * rng.random((n_drives, N)) - 1), "dt": np.full(N, dt)},
        controls=("u", "dt"), timestep="dt", bounds={"u": 1.0, "dt": (0.01, 0.5)}, goal={"x": x_goal})
    integ = BilinearIntegrator((G0, Gd), "x", "u", traj)
    J = QuadraticRegularizer("u", traj, 1.0) + MinimumTimeObjective(traj, D=1.0) + TerminalObjective(SqDist(x_goal), "x", traj)
    return DirectTrajOptProblem(traj, J, [integ])


def carrier_problem(N=1000, state_dim=64, n_drives=2, seed=42, dt=0.05, spline_order=1, steps=0):
    """BASELINE config c3: TimeDependentBilinearIntegrator with carrier-modulated drives,
    DerivativeIntegrator chain (u, du, ddu), knot constraint ||u|| - 1 <= 0 at knots 2..N-1."""
    rng = np.random.default_rng(seed)
    n = state_dim

    def skew():
        M = rng.standard_normal((n, n))
        return (M - M.T) / 2

    G0 = skew()
    G0 /= np.abs(G0).sum(axis=0).max()
    A = np.stack([skew() for _ in range(n_drives)])
    B = np.stack([skew() for _ in range(n_drives)])
    A /= np.abs(A).sum(axis=1).max()
    B /= np.abs(B).sum(axis=1).max()
    omega = 1.0 + rng.random(n_drives)
    gen = CarrierGenerator(G0, A, B, omega, np.zeros(n_drives))
    comps = {"x": 2 * rng.random((n, N)) - 1, "u": 0.3 * (2 * rng.random((n_drives, N)) - 1),
             "du": rng.standard_normal((n_drives, N)), "ddu": rng.standard_normal((n_drives, N)),
             "t": np.arange(N) * dt, "dt": np.full(N, dt)}
    traj = NamedTrajectory(comps, controls=("ddu", "dt"), timestep="dt")
    integrators = [TimeDependentBilinearIntegrator(gen, "x", "u", "t", traj, spline_order=spline_order, steps=steps),
                   DerivativeIntegrator("u", "du", traj), DerivativeIntegrator("du", "ddu", traj)]
    J = QuadraticRegularizer("u", traj, 1.0) + QuadraticRegularizer("du", traj, 1.0) + MinimumTimeObjective(traj)
    con = NonlinearKnotPointConstraint(NormMinus(1.0), "u", traj, times=range(2, N), equality=False)
    return DirectTrajOptProblem(traj, J, integrators, constraints=[con])
