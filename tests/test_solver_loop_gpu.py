"""BASELINE.json: "Ipopt/MadNLP must converge to the same objective within 1e-6 with the same iteration count
+-2".  Neither solver (nor Julia) exists here, so the same interior-point loop (SciPy trust-constr,
tests/helpers/nlp_loop.py) is driven once by the CPU oracle and once by the CUDA evaluator through the C
ABI, on BASELINE config c1 (README example, N=50, max_iter=100) and on the reference's benchmark problem."""
import numpy as np
import pytest

from dto_b200 import problem_templates as pt
from helpers import nlp_loop

pytestmark = pytest.mark.gpu

CASES = {
    "readme_c1": (lambda: pt.readme_problem(N=50), 100),
    "bilinear_benchmark": (lambda: pt.bilinear_benchmark(N=21), 300),
    # global variables (goal on the unit sphere + a drive radius as decision variables): 60 iterations of the same path
    "global_goal": (lambda: pt.global_goal_problem(N=11), 60),
}


@pytest.mark.parametrize("name", list(CASES))
def test_same_objective_and_iteration_count(name):
    make, max_iter = CASES[name]
    prob = make()
    Zo, o = nlp_loop.solve(prob, nlp_loop.OracleCallbacks(prob), max_iter=max_iter)
    dev = nlp_loop.DeviceCallbacks(prob)
    try:
        Zd, d = nlp_loop.solve(prob, dev, max_iter=max_iter)
    finally:
        dev.close()
    assert abs(o["iterations"] - d["iterations"]) <= 2, (o, d)
    assert abs(o["objective"] - d["objective"]) <= 1e-6 * max(1.0, abs(o["objective"])), (o, d)
    assert abs(o["violation"] - d["violation"]) <= 1e-6, (o, d)
    assert np.abs(Zo - Zd).max() <= 1e-5 * max(1.0, np.abs(Zo).max()), (o, d)
