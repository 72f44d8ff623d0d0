# DTOB200CUDAExt.jl -- package extension (weak dependency on CUDA.jl): the MOI callbacks of `B200Evaluator` for arguments
# that already live on the GPU.  This is the consumer half of the device-resident hand-off: with
# `solve!(prob; options = MadNLPOptions(...), array_type = CuArray)` (ext/MadNLPSolverExt/utils.jl:11-110,
# src/solvers/madnlp_solver/options.jl:10-15) MadNLP's primal vector, multipliers and value arrays are `CuArray`s; these
# methods hand their device pointers straight to `dto_eval_all_dev`, so an interior-point iteration moves no
# Jacobian / Hessian bytes over PCIe at all.
#
# STATUS: unexecuted (no Julia here); the C entry point it calls is exercised from Python with torch device
# buffers (tests/test_fullsize_parity_gpu.py: dev == host bit for bit; bench.py `value`).
module DTOB200CUDAExt

import MathOptInterface as MOI
using CUDA
using ..DTOB200: B200Evaluator, DevPtr, eval_all_dev!, synchronize

dev(x::CuArray{Float64}) = reinterpret(DevPtr, pointer(x))   # CuPtr{Float64} -> raw device address
const NULL = DevPtr(C_NULL)

# The handle has its own stream.  CUDA.jl's task-local stream must have produced Z (and mu) before the kernels read them,
# and must not read the outputs before they are written: synchronise on both sides (two host waits of ~5 us; a finer
# version records a CuEvent on each stream and uses cuStreamWaitEvent).
function fenced(f, e::B200Evaluator)
    CUDA.synchronize()
    f()
    synchronize(e)
end

function MOI.eval_objective(e::B200Evaluator, Z::CuArray{Float64})
    J = CUDA.zeros(Float64, 1)
    fenced(() -> eval_all_dev!(e, dev(Z), 0.0, NULL, dev(J), NULL, NULL, NULL, NULL), e)
    return Array(J)[1]
end
MOI.eval_objective_gradient(e::B200Evaluator, g::CuArray{Float64}, Z::CuArray{Float64}) =
    fenced(() -> eval_all_dev!(e, dev(Z), 0.0, NULL, NULL, dev(g), NULL, NULL, NULL), e)
MOI.eval_constraint(e::B200Evaluator, g::CuArray{Float64}, Z::CuArray{Float64}) =
    fenced(() -> eval_all_dev!(e, dev(Z), 0.0, NULL, NULL, NULL, dev(g), NULL, NULL), e)
MOI.eval_constraint_jacobian(e::B200Evaluator, J::CuArray{Float64}, Z::CuArray{Float64}) =
    fenced(() -> eval_all_dev!(e, dev(Z), 0.0, NULL, NULL, NULL, NULL, dev(J), NULL), e)
MOI.eval_hessian_lagrangian(e::B200Evaluator, H::CuArray{Float64}, Z::CuArray{Float64}, σ::Float64, μ::CuArray{Float64}) =
    fenced(() -> eval_all_dev!(e, dev(Z), σ, dev(μ), NULL, NULL, NULL, NULL, dev(H)), e)

"""The whole iterate in one pass (what MadNLP's `eval_*_wrapper`s amount to when called back to back): objective,
gradient, constraint, Jacobian and Hessian values of `Z`, all `CuArray`s."""
function eval_all!(e::B200Evaluator, Z::CuArray{Float64}, σ::Float64, μ::CuArray{Float64}, J::CuArray{Float64}, ∇::CuArray{Float64},
                   g::CuArray{Float64}, ∂::CuArray{Float64}, H::CuArray{Float64})
    fenced(() -> eval_all_dev!(e, dev(Z), σ, dev(μ), dev(J), dev(∇), dev(g), dev(∂), dev(H)), e)
end

end # module
