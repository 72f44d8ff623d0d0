# DTOB200.jl -- reference-side binding of libdto_b200.so: a drop-in `MOI.AbstractNLPEvaluator`
# for DirectTrajOpt.jl's `Solvers.Evaluator` (src/solvers/evaluator.jl:66-456).
#
# STATUS: written against include/dto_b200.h, NOT executed -- there is no Julia toolchain in the build
# container or on the GPU box (profiles/r01_fp64_peaks_and_box_probe.log).  The same ABI is exercised
# end to end through the Python ctypes mirror (directtrajopt.jl_b200/_lib.py, tests/).
#
# Usage inside DirectTrajOpt.jl (the only two construction sites of the evaluator):
#   src/solvers/ipopt_solver/solver.jl:68-69   evaluator = DTOB200.B200Evaluator(prob; eval_hessian = options.eval_hessian)
#   ext/MadNLPSolverExt/solver.jl:81           evaluator = DTOB200.B200Evaluator(prob; eval_hessian = true)
# Everything above `MOI.NLPBlockData(nl_cons, evaluator, true)` stays the reference's Julia.
module DTOB200

import MathOptInterface as MOI
using NamedTrajectories
using DirectTrajOpt

const LIB = get(ENV, "DTO_B200_LIB", joinpath(@__DIR__, "..", "directtrajopt.jl_b200", "lib", "libdto_b200.so"))
const ABI_VERSION = Cint(3)

# ---- mirror of the C descriptor structs (include/dto_b200.h) -------------------------------------
struct IntegratorDesc
    kind::Cint; x_off::Cint; x_dim::Cint; u_off::Cint; u_dim::Cint; t_off::Cint; spline_order::Cint; n_carrier::Cint
    G::Ptr{Cdouble}; G_batch_stride::Int64
    A::Ptr{Cdouble}; B::Ptr{Cdouble}; omega::Ptr{Cdouble}; phi::Ptr{Cdouble}
    D::Ptr{Cdouble}; omega_d::Ptr{Cdouble}; phi_d::Ptr{Cdouble}
    tdb_steps::Cint; _pad::Cint
end
struct ObjectiveDesc
    kind::Cint; fn::Cint; weight::Cdouble; n_vars::Cint; n_times::Cint
    var_offs::Ptr{Cint}; times::Ptr{Cint}; R::Ptr{Cdouble}; baseline::Ptr{Cdouble}; D::Cdouble
    n_params::Cint; _pad::Cint; params::Ptr{Cdouble}; Qs::Ptr{Cdouble}
    n_gvars::Cint; _pad2::Cint; gvar_offs::Ptr{Cint}
end
struct ConstraintDesc
    fn::Cint; equality::Cint; n_vars::Cint; n_times::Cint
    var_offs::Ptr{Cint}; times::Ptr{Cint}; g_dim::Cint; n_params::Cint; params::Ptr{Cdouble}
    n_gvars::Cint; _pad::Cint; gvar_offs::Ptr{Cint}
end
struct ProblemDesc
    abi_version::Cint; N::Cint; z::Cint; dt_off::Cint; batch::Cint; eval_hessian::Cint
    shard_k0::Cint; shard_k1::Cint; device::Cint; n_integrators::Cint; n_objectives::Cint; n_constraints::Cint
    integrators::Ptr{IntegratorDesc}; objectives::Ptr{ObjectiveDesc}; constraints::Ptr{ConstraintDesc}; Z0::Ptr{Cdouble}
    global_dim::Cint; _pad::Cint
end
struct SizeInfo
    n_vars::Int64; n_dynamics_cons::Int64; n_nonlinear_cons::Int64; n_cons::Int64; nnz_jac::Int64; nnz_hess::Int64
end

# ---- device catalogue of knot functions (the Julia closures g, l cannot cross a C ABI) -------------
abstract type KnotFunction end
# constraints g(v; p)  (ids: include/dto_b200.h DTO_G_*)
struct NormMinus <: KnotFunction; c::Float64; end            # g(v) = [norm(v) - c]
struct NormSqMinus <: KnotFunction; c::Float64; end          # g(v) = [norm(v)^2 - c]
struct SqDistMinus <: KnotFunction; target::Vector{Float64}; c::Float64; end  # g(v) = [norm(v - target)^2 - c]
struct LinearMap <: KnotFunction; A::Matrix{Float64}; b::Vector{Float64}; end # g(v) = A v - b
struct NormProduct <: KnotFunction; n1::Int; c1::Float64; c2::Float64; end  # g(v) = [norm(v1) - c1; norm(v1) norm(v2) - c2]
# objectives l(v; p)  (ids: DTO_L_*)
struct NormSqPlus <: KnotFunction; p::Float64; end           # l(v) = norm(v)^2 + p
struct SqDist <: KnotFunction; target::Vector{Float64}; end  # l(v) = norm(v - target)^2
struct LinearCost <: KnotFunction; c::Vector{Float64}; end   # l(v) = c'v
struct IsoInfidelity <: KnotFunction; goal::Vector{Float64}; end
struct SplitSqDist <: KnotFunction; end                      # l(v) = norm(v[1:h] - v[h+1:2h])^2 (state vs a goal held in a global)
"""
    PerTime(build)

The reference hands the i-th listed time its own parameter object, `g(v, params[i])` / `l(v, params[i])`
(knot_point_constraint.jl:76-83, knot_point_objectives.jl:65-72).  `build(params[i])` returns the catalogue entry for that
time (same type for every time), e.g. `PerTime(p -> SqDist(p))` for a tracking cost with a per-knot target.
"""
struct PerTime <: KnotFunction; build::Function; end
cfun_id(::NormMinus) = Cint(1); cfun_id(::NormSqMinus) = Cint(2); cfun_id(::SqDistMinus) = Cint(3); cfun_id(::LinearMap) = Cint(4)
cfun_id(::NormProduct) = Cint(5)
lfun_id(::NormSqPlus) = Cint(1); lfun_id(::SqDist) = Cint(2); lfun_id(::LinearCost) = Cint(3); lfun_id(::IsoInfidelity) = Cint(4)
lfun_id(::SplitSqDist) = Cint(5)
params(f::NormMinus) = [f.c]; params(f::NormSqMinus) = [f.c]; params(f::SqDistMinus) = vcat(f.c, f.target)
params(f::LinearMap) = vcat(Float64(size(f.A, 1)), vec(f.A), f.b)   # [g_dim, A column-major, b]
params(f::NormProduct) = [f.c1, f.c2, Float64(f.n1)]
params(f::NormSqPlus) = [f.p]; params(f::SqDist) = f.target; params(f::LinearCost) = f.c; params(f::IsoInfidelity) = f.goal
params(::SplitSqDist) = [0.0]
gdim(f::KnotFunction) = f isa LinearMap ? size(f.A, 1) : (f isa NormProduct ? 2 : 1)
"""(id function applied to the entry of time 1, `n_times x n_params` row-major parameter table) of a catalogue entry;
`ref_params` is the reference object's own per-time parameter vector (used by `PerTime`)."""
function param_table(f::KnotFunction, ref_params, n_times::Int)
    entries = f isa PerTime ? [f.build(ref_params[i]) for i = 1:n_times] : fill(f, n_times)
    all(e -> typeof(e) == typeof(entries[1]), entries) || error("PerTime: every time must map to the same catalogue function")
    rows = [params(e) for e in entries]
    all(r -> length(r) == length(rows[1]), rows) || error("PerTime: parameter blocks of different length")
    return entries[1], vcat(rows...), length(rows[1])
end

"""
    CarrierGenerator(G0, A, B, omega, phi, D, omega_d, phi_d)

Data form of the time-dependent generator the device integrates,
`G(u, t) = G0 + sum_i u_i (cos(w_i t + phi_i) A_i + sin(w_i t + phi_i) B_i) + sum_j cos(wd_j t + phd_j) D_j`
(replaces the closure `G(u, t)` of time_dependent_bilinear_integrator.jl:70-78, which cannot cross a C ABI).  Pass one per
`TimeDependentBilinearIntegrator` through the `carrier_generators` keyword; it is checked against the closure at
construction (three random (u, t) probes) so that the device never integrates a different system.
"""
struct CarrierGenerator
    G0::Matrix{Float64}; A::Vector{Matrix{Float64}}; B::Vector{Matrix{Float64}}; omega::Vector{Float64}; phi::Vector{Float64}
    D::Vector{Matrix{Float64}}; omega_d::Vector{Float64}; phi_d::Vector{Float64}
end
CarrierGenerator(G0, A, B, omega; phi = zeros(length(A)), D = Matrix{Float64}[], omega_d = Float64[], phi_d = zeros(length(D))) =
    CarrierGenerator(G0, A, B, omega, phi, D, omega_d, phi_d)
(g::CarrierGenerator)(u, t) = g.G0 + sum(u[i] * (cos(g.omega[i] * t + g.phi[i]) * g.A[i] + sin(g.omega[i] * t + g.phi[i]) * g.B[i]) for i in eachindex(g.A); init = zero(g.G0)) +
                              sum(cos(g.omega_d[j] * t + g.phi_d[j]) * g.D[j] for j in eachindex(g.D); init = zero(g.G0))

# 0-based positions inside traj.global_data of the listed global components
goffs(traj, names) = isempty(names) ? Cint[] : Cint.(vcat([collect(traj.global_components[n]) for n in names]...) .- 1)
koffs(traj, names) = isempty(names) ? Cint[] : Cint.(vcat([collect(traj.components[n]) for n in names]...) .- 1)

"""Lower `G::Function` of a BilinearIntegrator to (G_drift, G_drives) by probing; error if G is not affine in u
(no CPU fallback: unsupported components raise at construction, never mid-solve)."""
function lower_generator(G, m::Int)
    G0 = Matrix{Float64}(G(zeros(m)))
    drives = [Matrix{Float64}(G([i == j ? 1.0 : 0.0 for j = 1:m])) - G0 for i = 1:m]
    u = randn(m)
    Gt = G0 + sum(u[i] * drives[i] for i = 1:m; init = zero(G0))
    isapprox(Matrix{Float64}(G(u)), Gt; rtol = 1e-12, atol = 1e-12 * (1 + maximum(abs, Gt))) ||
        error("BilinearIntegrator: G(u) is not affine in u and cannot be lowered to the device")
    return G0, drives
end

mutable struct B200Evaluator <: MOI.AbstractNLPEvaluator
    handle::Ptr{Cvoid}
    trajectory::NamedTrajectory
    objective::Any
    integrators::Vector
    constraints::Vector
    jacobian_structure::Vector{Tuple{Int,Int}}
    hessian_structure::Vector{Tuple{Int,Int}}
    n_dynamics_constraints::Int
    n_nonlinear_constraints::Int
    n_constraints::Int
    eval_hessian::Bool
    reg_jac::Ptr{Cdouble}     # the solver's value arrays currently registered with the handle (dto_register_outputs)
    reg_hess::Ptr{Cdouble}
end

check(rc::Cint, h) = rc == 0 || error("libdto_b200: " * unsafe_string(ccall((:dto_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

"""
    B200Evaluator(prob::DirectTrajOptProblem; eval_hessian=true, knot_functions=Dict(), carrier_generators=Dict())

`knot_functions` maps each `NonlinearKnotPointConstraint` / `KnotPointObjective` (and their global variants) of `prob` to
its catalogue entry (e.g. `g_u_norm => NormMinus(1.0)`, `J_terminal => SqDist(traj.goal.x)`, `J_track => PerTime(p -> SqDist(p))`
when the reference object carries per-time `params`); `carrier_generators` maps each `TimeDependentBilinearIntegrator`
to its `CarrierGenerator`.
"""
function B200Evaluator(prob; eval_hessian = true, knot_functions = Dict(), carrier_generators = Dict())
    traj = prob.trajectory
    keep = Any[]  # GC roots for every array the descriptor points into (only needed until dto_create returns)
    off(name) = Cint(first(traj.components[name]) - 1)
    ints = IntegratorDesc[]
    for it in prob.integrators
        if it isa BilinearIntegrator
            m = traj.dims[it.u_name]
            # it.f closes over G; DirectTrajOpt keeps G reachable as `it.f.G` (closure field)
            G0, drives = lower_generator(it.f.G, m)
            Gs = vcat(vec(G0), (vec(d) for d in drives)...)   # column-major matrices, drift first
            push!(keep, Gs)
            push!(ints, IntegratorDesc(1, off(it.x_name), it.x_dim, off(it.u_name), m, -1, 0, 0, pointer(Gs), 0,
                                       C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0, 0))
        elseif it isa DerivativeIntegrator
            push!(ints, IntegratorDesc(2, off(it.x_name), it.x_dim, off(it.ẋ_name), it.x_dim, -1, 0, 0, C_NULL, 0,
                                       C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0, 0))
        elseif it isa TimeDependentBilinearIntegrator
            cg = get(carrier_generators, it, nothing)
            cg isa CarrierGenerator || error("TimeDependentBilinearIntegrator needs its generator as a CarrierGenerator in `carrier_generators`")
            m = traj.dims[it.u_name]
            for _ = 1:3  # the data form must be the closure the reference would integrate (it.f.G, as for the bilinear case)
                u, t = randn(m), 10 * rand()
                Gref = Matrix{Float64}(it.f.G(u, t))
                isapprox(Gref, cg(u, t); rtol = 1e-12, atol = 1e-12 * (1 + maximum(abs, Gref))) ||
                    error("CarrierGenerator does not reproduce the integrator's G(u, t)")
            end
            cat(Ms, n) = isempty(Ms) ? Float64[] : vcat((vec(M) for M in Ms)...)   # column-major matrices back to back
            n = it.x_dim
            arrs = (vec(copy(cg.G0)), cat(cg.A, n), cat(cg.B, n), copy(cg.omega), copy(cg.phi), cat(cg.D, n), copy(cg.omega_d), copy(cg.phi_d))
            append!(keep, arrs)
            ptr(a) = isempty(a) ? Ptr{Cdouble}(C_NULL) : pointer(a)
            # spline order 1 (linear interpolation of u over the interval) is the reference's default (:76); tdb_steps = 0 lets
            # the kernels choose the macro steps per interval
            push!(ints, IntegratorDesc(3, off(it.x_name), n, off(it.u_name), m, off(it.t_name), it.spline_order, length(cg.D), ptr(arrs[1]), 0,
                                       ptr(arrs[2]), ptr(arrs[3]), ptr(arrs[4]), ptr(arrs[5]), ptr(arrs[6]), ptr(arrs[7]), ptr(arrs[8]), 0, 0))
        else
            error("integrator $(typeof(it)) has no device lowering")
        end
    end
    objs = ObjectiveDesc[]
    terms = prob.objective isa DirectTrajOpt.Objectives.CompositeObjective ?
            collect(zip(prob.objective.objectives, prob.objective.weights)) : [(prob.objective, 1.0)]
    for (ob, w) in terms
        if ob isa QuadraticRegularizer
            vo = Cint.(collect(traj.components[ob.name]) .- 1); tm = Cint.(ob.times); R = copy(ob.R); base = copy(ob.baseline)
            append!(keep, (vo, tm, R, base))
            push!(objs, ObjectiveDesc(1, 0, w, length(vo), length(tm), pointer(vo), pointer(tm), pointer(R),
                                      any(!iszero, base) ? pointer(base) : C_NULL, 0.0, 0, 0, C_NULL, C_NULL, 0, 0, C_NULL))
        elseif ob isa LinearRegularizer
            vo = Cint.(collect(traj.components[ob.name]) .- 1); tm = Cint.(ob.times); R = copy(ob.R)
            append!(keep, (vo, tm, R))
            push!(objs, ObjectiveDesc(5, 0, w, length(vo), length(tm), pointer(vo), pointer(tm), pointer(R), C_NULL, 0.0, 0, 0, C_NULL, C_NULL, 0, 0, C_NULL))
        elseif ob isa MinimumTimeObjective
            push!(objs, ObjectiveDesc(2, 0, w, 0, 0, C_NULL, C_NULL, C_NULL, C_NULL, ob.D, 0, 0, C_NULL, C_NULL, 0, 0, C_NULL))
        elseif ob isa KnotPointObjective
            f = get(knot_functions, ob, nothing)
            f === nothing && error("KnotPointObjective needs a catalogue entry in `knot_functions`")
            vo = Cint.(vcat([collect(traj.components[n]) for n in ob.var_names]...) .- 1); tm = Cint.(ob.times)
            f1, pr, npar = param_table(f, ob.params, length(tm)); Qs = copy(ob.Qs)   # params[i] per listed time (knot_point_objectives.jl:65-72)
            append!(keep, (vo, tm, pr, Qs))
            push!(objs, ObjectiveDesc(3, lfun_id(f1), w, length(vo), length(tm), pointer(vo), pointer(tm), C_NULL, C_NULL, 0.0,
                                      npar, 0, pointer(pr), pointer(Qs), 0, 0, C_NULL))
        elseif ob isa GlobalKnotPointObjective || ob isa GlobalObjective
            # J = sum_i Q_i l([knot vars; global vars]) (global_objectives.jl:139-341); a GlobalObjective is the same term
            # without knot variables, listed once with Qs = [Q] (global_objectives.jl:35-130)
            f = get(knot_functions, ob, nothing)
            f === nothing && error("$(typeof(ob)) needs a catalogue entry in `knot_functions`")
            isknot = ob isa GlobalKnotPointObjective
            vo = isknot ? koffs(traj, ob.var_names) : Cint[]; go = goffs(traj, ob.global_names)
            tm = isknot ? Cint.(ob.times) : Cint[1]; Qs = isknot ? copy(ob.Qs) : [ob.Q]
            f1, pr, npar = param_table(f, isknot ? ob.params : [nothing], length(tm))
            append!(keep, (vo, go, tm, pr, Qs))
            push!(objs, ObjectiveDesc(6, lfun_id(f1), w, length(vo), length(tm), isempty(vo) ? C_NULL : pointer(vo), pointer(tm), C_NULL, C_NULL,
                                      0.0, npar, 0, pointer(pr), pointer(Qs), length(go), 0, pointer(go)))
        elseif ob isa NullObjective
            push!(objs, ObjectiveDesc(4, 0, w, 0, 0, C_NULL, C_NULL, C_NULL, C_NULL, 0.0, 0, 0, C_NULL, C_NULL, 0, 0, C_NULL))
        else
            error("objective $(typeof(ob)) has no device lowering")
        end
    end
    nl = filter(c -> c isa DirectTrajOpt.Constraints.AbstractNonlinearConstraint, prob.constraints)
    cons = ConstraintDesc[]
    for c in nl
        f = get(knot_functions, c, nothing)
        f !== nothing || error("constraint $(typeof(c)) needs a catalogue entry")
        if c isa NonlinearKnotPointConstraint
            vo = koffs(traj, c.var_names); go = Cint[]; tm = Cint.(c.times); gd = c.g_dim
        elseif c isa NonlinearGlobalKnotPointConstraint   # global_knot_point_constraint.jl:30-256
            vo = koffs(traj, c.var_names); go = goffs(traj, c.global_names); tm = Cint.(c.times); gd = c.g_dim
        elseif c isa NonlinearGlobalConstraint            # global_constraint.jl:24-159: no knot variables, listed once
            vo = Cint[]; go = goffs(traj, c.global_names); tm = Cint[1]; gd = c.dim
        else
            error("constraint $(typeof(c)) has no device lowering")
        end
        # params[i] of the i-th listed time (knot_point_constraint.jl:76-83); global constraints have one block
        f1, pr, npar = param_table(f, hasproperty(c, :params) ? c.params : [nothing], length(tm)); append!(keep, (vo, go, tm, pr))
        gdim(f1) == gd || error("catalogue entry has g_dim $(gdim(f1)), the constraint $(gd)")
        push!(cons, ConstraintDesc(cfun_id(f1), c.equality, length(vo), length(tm), isempty(vo) ? C_NULL : pointer(vo), pointer(tm), gd,
                                   npar, pointer(pr), length(go), 0, isempty(go) ? C_NULL : pointer(go)))
    end
    Z0 = vcat(collect(traj.datavec), collect(traj.global_data))   # the solver's vector: [knots; globals]
    desc = Ref(ProblemDesc(ABI_VERSION, traj.N, traj.dim, off(traj.timestep), 1, eval_hessian, 0, 0, -1,
                           length(ints), length(objs), length(cons),
                           pointer(ints), isempty(objs) ? C_NULL : pointer(objs), isempty(cons) ? C_NULL : pointer(cons), pointer(Z0),
                           traj.global_dim, 0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep ints objs cons Z0 begin
        rc = ccall((:dto_create, LIB), Cint, (Ref{ProblemDesc}, Ref{Ptr{Cvoid}}), desc, h)
    end
    rc == 0 || error("libdto_b200: " * unsafe_string(ccall((:dto_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    si = Ref(SizeInfo(0, 0, 0, 0, 0, 0))
    check(ccall((:dto_sizes, LIB), Cint, (Ptr{Cvoid}, Ref{SizeInfo}), h[], si), h[])
    jr = Vector{Int64}(undef, si[].nnz_jac); jc = similar(jr)
    check(ccall((:dto_jac_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h[], jr, jc), h[])
    hr = Vector{Int64}(undef, si[].nnz_hess); hc = similar(hr)
    check(ccall((:dto_hess_structure, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h[], hr, hc), h[])
    ev = B200Evaluator(h[], traj, prob.objective, prob.integrators, collect(nl), collect(zip(jr, jc)), collect(zip(hr, hc)),
                       si[].n_dynamics_cons, si[].n_nonlinear_cons, si[].n_cons, eval_hessian, C_NULL, C_NULL)
    finalizer(e -> ccall((:dto_destroy, LIB), Cvoid, (Ptr{Cvoid},), e.handle), ev)
    return ev
end

# ---- the MOI interface, signature for signature (src/solvers/evaluator.jl:291-456) ------------------
MOI.initialize(::B200Evaluator, features) = nothing
MOI.features_available(e::B200Evaluator) = e.eval_hessian ? [:Grad, :Jac, :Hess] : [:Grad, :Jac]
MOI.jacobian_structure(e::B200Evaluator) = e.jacobian_structure
MOI.hessian_lagrangian_structure(e::B200Evaluator) = e.hessian_structure

dense(v::AbstractVector) = v isa Vector{Float64} ? v : collect(Float64, v)   # Z may be any AbstractVector

function MOI.eval_objective(e::B200Evaluator, Z::AbstractVector)
    J = Ref(0.0)
    check(ccall((:dto_eval_objective, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}), e.handle, dense(Z), J), e.handle)
    return J[]
end
function MOI.eval_objective_gradient(e::B200Evaluator, g::AbstractVector, Z::AbstractVector)
    check(ccall((:dto_eval_gradient, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), e.handle, dense(Z), g), e.handle)
    return nothing
end
function MOI.eval_constraint(e::B200Evaluator, g::AbstractVector, Z::AbstractVector)
    check(ccall((:dto_eval_constraint, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), e.handle, dense(Z), g), e.handle)
    return nothing
end
"""Ipopt and MadNLP hand the SAME value arrays to every Jacobian / Hessian callback of a solve (their own storage, wrapped by
`unsafe_wrap` in Ipopt.jl's C callbacks): the first time a pointer is seen it is registered with the handle
(`dto_register_outputs`: page-locked, structural constants written once, later calls move only the value-dependent
entries); a different pointer re-registers.  Set `ENV["DTO_B200_REGISTER"] = "0"` to keep plain buffers."""
function maybe_register!(e::B200Evaluator, J::Union{Nothing,Vector{Float64}}, H::Union{Nothing,Vector{Float64}})
    get(ENV, "DTO_B200_REGISTER", "1") == "0" && return
    pj = J === nothing ? e.reg_jac : pointer(J); ph = H === nothing ? e.reg_hess : pointer(H)
    (pj == e.reg_jac && ph == e.reg_hess) && return
    check(ccall((:dto_register_outputs, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), e.handle, pj, ph), e.handle)
    e.reg_jac, e.reg_hess = pj, ph
    return
end

function MOI.eval_constraint_jacobian(e::B200Evaluator, J::AbstractVector, Z::AbstractVector)
    J isa Vector{Float64} && maybe_register!(e, J, nothing)
    check(ccall((:dto_eval_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), e.handle, dense(Z), J), e.handle)
    return nothing
end
function MOI.eval_hessian_lagrangian(e::B200Evaluator, H::AbstractVector{T}, Z::AbstractVector{T}, σ::T, μ::AbstractVector{T}) where {T}
    H isa Vector{Float64} && maybe_register!(e, nothing, H)
    check(ccall((:dto_eval_hessian, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, dense(Z), σ, dense(μ), H), e.handle)
    return nothing
end
function MOI.eval_constraint_jacobian_product(e::B200Evaluator, y::AbstractVector{T}, x::AbstractVector{T}, w::AbstractVector{T}) where {T}
    check(ccall((:dto_eval_jacobian_product, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, dense(x), dense(w), y), e.handle)
    return nothing
end
function MOI.eval_constraint_jacobian_transpose_product(e::B200Evaluator, y::AbstractVector{T}, x::AbstractVector{T}, w::AbstractVector{T}) where {T}
    check(ccall((:dto_eval_jacobian_transpose_product, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, dense(x), dense(w), y), e.handle)
    return nothing
end

# ---- device-resident hand-off (MadNLP `array_type = CuArray`, ext/MadNLPSolverExt/utils.jl:11-110) ---------------------
# With `array_type` set, MadNLP keeps x, the multipliers and the Jacobian / Hessian value arrays on the GPU and its MOI
# wrapper round-trips every callback through host copies.  The methods below take raw DEVICE pointers instead (a
# `CuArray`'s `pointer(x)` reinterpreted, see DTOB200CUDAExt.jl): the iterate is evaluated where it lives, the values land
# where the KKT assembly (cuDSS) reads them, nothing crosses PCIe.  They enqueue on the handle's stream; `synchronize`
# (or a CUDA.jl event on `stream(e)`) orders them against the solver's own stream.
const DevPtr = Ptr{Cdouble}   # a device address
stream(e::B200Evaluator) = ccall((:dto_stream, LIB), Ptr{Cvoid}, (Ptr{Cvoid},), e.handle)
synchronize(e::B200Evaluator) = check(ccall((:dto_synchronize, LIB), Cint, (Ptr{Cvoid},), e.handle), e.handle)
"""One fused pass with device pointers; any output may be `C_NULL`.  (`dto_eval_all_dev`)"""
function eval_all_dev!(e::B200Evaluator, dZ::DevPtr, σ::Float64, dμ::DevPtr, dJ::DevPtr, d∇::DevPtr, dg::DevPtr, d∂::DevPtr, dH::DevPtr)
    check(ccall((:dto_eval_all_dev, LIB), Cint, (Ptr{Cvoid}, DevPtr, Cdouble, DevPtr, DevPtr, DevPtr, DevPtr, DevPtr, DevPtr),
                e.handle, dZ, σ, dμ, dJ, d∇, dg, d∂, dH), e.handle)
    return nothing
end

# ---- one long trajectory on several GPUs (INTEGRATION.md section 4b; include/dto_b200.h "knot-range sharding") --------
# One Julia process per GPU (MPI.jl or Distributed): every rank builds its evaluator with `shard_k0/shard_k1` in the
# descriptor, exchanges the 64-byte window handles once and links; per iterate every rank uploads its slice exactly once.
"""64-byte CUDA-IPC handle of this shard's exchange window (`dto_shard_export`); all-gather these in rank order."""
function shard_export(e::B200Evaluator)
    buf = Vector{UInt8}(undef, 64)
    check(ccall((:dto_shard_export, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), e.handle, buf), e.handle)
    return buf
end
"""Map the windows of all `world` shards; `handles` = the gathered 64-byte handles, rank order (`dto_shard_link`)."""
function shard_link!(e::B200Evaluator, rank::Integer, world::Integer, handles::Vector{Vector{UInt8}})
    flat = reduce(vcat, handles)
    check(ccall((:dto_shard_link, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), e.handle, rank, world, flat), e.handle)
    return nothing
end
"""New iterate from host memory: uploads this rank's slice `[z_begin, z_halo_end)` and publishes its first knot (`dto_upload`)."""
upload!(e::B200Evaluator, Zslice::Vector{Float64}) =
    (check(ccall((:dto_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), e.handle, Zslice), e.handle); nothing)
"""New iterate that already lives on the device (`dto_upload_dev`): copy + publish + wait for the neighbour's knot, one kernel."""
upload_dev!(e::B200Evaluator, dZ::DevPtr) =
    (check(ccall((:dto_upload_dev, LIB), Cint, (Ptr{Cvoid}, DevPtr), e.handle, dZ), e.handle); nothing)
"""Device address of this shard's resident Z (what the `_dev` callbacks are then given) (`dto_local_Z`)."""
local_Z(e::B200Evaluator) = ccall((:dto_local_Z, LIB), DevPtr, (Ptr{Cvoid},), e.handle)
"""Violation of this shard's residuals `dg` and the (sum, max) exchange of objective / violation with all shards in one
kernel, in place on the device, no collective library (`dto_shard_scalars_dev`)."""
function shard_scalars_dev!(e::B200Evaluator, dg::DevPtr, dJ::DevPtr, dviol::DevPtr)
    check(ccall((:dto_shard_scalars_dev, LIB), Cint, (Ptr{Cvoid}, DevPtr, DevPtr, DevPtr), e.handle, dg, dJ, dviol), e.handle)
    return nothing
end

end # module
