# ProbNumDiffEqB200.jl -- Julia side of the drop-in boundary (source only: Julia is not installed in
# the build image, so this file is syntax-reviewed, not executed; the same C ABI is exercised by the
# Python mirror in ../api.py and by tests/).
#
# It overloads DiffEqBase.__solve for the reference's algorithm types (src/algorithms.jl:23-51) so that
#     solve(prob, EK1(order=3); abstol, reltol)                      and
#     solve(EnsembleProblem(prob; prob_func), EK1(order=3), EnsembleB200(devices=0:7); trajectories=N, ...)
# run the whole OrdinaryDiffEq solve! loop inside libpnde.so (include/pnde.h, ABI version 2) and rebuild the
# fields of ProbODESolution (src/solution.jl:8-24) as LAZY views over the structure-of-arrays buffers the library
# fills: no per-state factorisation, no per-state allocation.
module ProbNumDiffEqB200

using LinearAlgebra
using ProbNumDiffEq
using ProbNumDiffEq: AbstractEK, EK0, EK1, IEKS, SRMatrix, ProbODESolution
using DiffEqBase
using GaussianDistributions: Gaussian

const libpnde = get(ENV, "PNDE_LIB", joinpath(@__DIR__, "..", "libpnde.so"))
const PNDE_MAX_DEVICES = 16

# ---- mirror of `struct pnde_config` (include/pnde.h); field order and types must match ----------
Base.@kwdef mutable struct PndeConfig
    abi_version::Int32 = 2
    alg::Int32 = 1
    order::Int32 = 3
    d::Int32 = 0
    vf_kind::Int32 = 0
    diffusion::Int32 = 0
    smooth::Int32 = 0
    adaptive::Int32 = 1
    save_mode::Int32 = 0
    save_stride::Int32 = 1
    device::Int32 = -1
    ieks_iterations::Int32 = 0
    abstol::Float64 = 1e-6
    reltol::Float64 = 1e-3
    dt::Float64 = 0.0
    t0::Float64 = 0.0
    t1::Float64 = 1.0
    qmin::Float64 = 1 / 5
    qmax::Float64 = 10.0
    gamma::Float64 = 9 / 10
    qsteady_min::Float64 = 1.0
    qsteady_max::Float64 = 1.0
    qoldinit::Float64 = 1e-4
    beta1::Float64 = 0.0
    beta2::Float64 = 0.0
    dtmin::Float64 = 0.0
    dtmax::Float64 = 0.0
    maxiters::Int64 = 100000
    max_saved::Int64 = 0
    # ABI version 2
    n_devices::Int32 = 0
    flags::Int32 = 0                       # 1: PNDE_FLAG_REFERENCE_QUIRKS, 2: PNDE_FLAG_ONE_THREAD
    device_list::NTuple{PNDE_MAX_DEVICES,Int32} = ntuple(_ -> Int32(0), PNDE_MAX_DEVICES)
end

const DIFFUSIONS = Dict(:dynamic => 0, :fixed => 1, :fixedMAP => 2, :dynamicMV => 3, :fixedMV => 4)  # src/caches.jl:89-96
const RETCODES = Dict(0 => :Success, 1 => :MaxIters, 2 => :DtNaN, 3 => :Unstable, 4 => :Failure, 5 => :DtLessThanMin,
                      6 => :Failure)

# ---- vector fields: Julia closures cannot run on the device ---------------------------------------------------
"""A field of the built-in catalogue (`PNDE_VF_*`, include/pnde.h); `f` is kept for host-side use."""
struct CatalogueFunction
    kind::Int32
    f
end

"""Any autonomous user ODE as CUDA C++ statement lists, compiled at run time (NVRTC) into the same kernels
(`pnde_create_custom`).  `custom_function(sys)` builds one from a ModelingToolkit system -- the reference already
derives its Jacobian from there (src/jacobian.jl:6-22)."""
struct CustomFunction
    d::Int32
    n_params::Int32
    f_body::String
    jac_body::String
    f
end

# build_function(...; target = CTarget()) emits `du[i] = ...;` lines over `u[]`, `p[]` (RHS1, RHS2 argument names);
# the Jacobian comes out as a flat column-major `J[k]`, rewritten here to the `J[i][j]` form pnde_create_custom takes.
function custom_function(sys, f = nothing)
    MTK = Base.require(Base.PkgId(Base.UUID("961ee093-0014-501f-94e3-6117800e7a78"), "ModelingToolkit"))
    eqs, sts, ps = MTK.equations(sys), MTK.states(sys), MTK.parameters(sys)
    d = length(sts)
    rhs = [eq.rhs for eq in eqs]
    cbody(ex) = replace(replace(string(MTK.build_function(ex, sts, ps; target = MTK.CTarget(), lhsname = :du,
                                                          rhsnames = [:u, :p], fname = :f)),
                                r"^.*?\{"s => ""), r"\}\s*$" => "")
    f_body = cbody(rhs)
    jac = MTK.calculate_jacobian(sys)
    jflat = replace(replace(string(MTK.build_function(vec(jac), sts, ps; target = MTK.CTarget(), lhsname = :Jf,
                                                      rhsnames = [:u, :p], fname = :j)),
                            r"^.*?\{"s => ""), r"\}\s*$" => "")
    jac_body = replace(jflat, r"Jf\[(\d+)\]" => m -> (k = parse(Int, match(r"\d+", m).match); "J[$(k % d)][$(k ÷ d)]"))
    return CustomFunction(d, length(ps), f_body, jac_body, f)
end

"""Ensemble algorithm: the whole ensemble in ONE library call, sharded over `devices` (contiguous blocks, one host
thread + stream per GPU inside the call, results written into disjoint slices of the output arrays)."""
struct EnsembleB200 <: DiffEqBase.EnsembleAlgorithm
    devices::Vector{Int32}
end
EnsembleB200(; devices = Int32[]) = EnsembleB200(collect(Int32, devices))

last_error(h) = unsafe_string(ccall((:pnde_last_error, libpnde), Cstring, (Ptr{Cvoid},), h))
check(rc, h) = rc == 0 || error(last_error(h))

function _create(cfg::PndeConfig, f)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = if f isa CustomFunction
        cfg.vf_kind = 100
        ccall((:pnde_create_custom, libpnde), Cint, (Ref{PndeConfig}, Int32, Int32, Cstring, Cstring, Ref{Ptr{Cvoid}}),
              cfg, f.d, f.n_params, f.f_body, f.jac_body, h)
    else
        ccall((:pnde_create, libpnde), Cint, (Ref{PndeConfig}, Ref{Ptr{Cvoid}}), cfg, h)
    end
    rc == 0 || error(last_error(C_NULL))
    return h[]
end

function _config(prob, alg::AbstractEK; abstol=1e-6, reltol=1e-3, adaptive=true, dt=nothing,
                 save_everystep=true, maxiters=100000, max_saved=0, device=-1, devices=Int32[], ieks_iterations=0,
                 reference_quirks=false, kwargs...)
    !adaptive && dt === nothing && error("Fixed timestep methods require a choice of dt")
    f = prob.f.f
    f isa Union{CatalogueFunction,CustomFunction} ||
        error("ProbNumDiffEqB200 needs a CatalogueFunction or a CustomFunction (see include/pnde.h)")
    length(devices) <= PNDE_MAX_DEVICES || error("at most $PNDE_MAX_DEVICES devices")
    smooth = alg.smooth && save_everystep
    dl = ntuple(i -> i <= length(devices) ? Int32(devices[i]) : Int32(0), PNDE_MAX_DEVICES)
    cfg = PndeConfig(alg = alg isa IEKS ? 2 : alg isa EK1 ? 1 : 0, ieks_iterations = ieks_iterations, order = alg.order,
                     vf_kind = f isa CatalogueFunction ? f.kind : 100, d = length(prob.u0),
                     diffusion = DIFFUSIONS[alg.diffusionmodel], smooth = smooth, adaptive = adaptive,
                     save_mode = save_everystep ? 1 : 0, device = device, abstol = abstol, reltol = reltol,
                     dt = dt === nothing ? 0.0 : dt, t0 = prob.tspan[1], t1 = prob.tspan[2],
                     maxiters = maxiters, max_saved = max_saved == 0 && adaptive && save_everystep ? 8192 : max_saved,
                     n_devices = length(devices), flags = reference_quirks ? 1 : 0, device_list = dl)
    return cfg, f
end

"""Run one ensemble (n trajectories; u0 is n x d, p is n x n_params: Julia's column-major layout of those IS the
library's structure-of-arrays layout with the trajectory index fastest)."""
function _run(cfg::PndeConfig, f, u0::Matrix{Float64}, p::Matrix{Float64})
    h = _create(cfg, f)
    n = size(u0, 1)
    check(ccall((:pnde_solve_ensemble, libpnde), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), h, n, u0, p), h)
    return h, n
end

function _counts(h, n)
    na, nr, nf, nj, ns = (zeros(Int64, n) for _ in 1:5)
    rc = zeros(Int32, n)
    check(ccall((:pnde_get_counts, libpnde), Cint,
                (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}), h, na, nr, nf, nj, rc, ns), h)
    return na, nr, nf, nj, rc, ns
end

# ---- lazy StructArray-like view: Gaussian{Vector,SRMatrix} per saved state over SoA buffers -------------------
# (src/solution.jl:60-64 builds StructArray{Gaussian}; consumers index it, broadcast mean/var over it
#  (src/ProbNumDiffEq.jl:61-66) and multiply single elements by SolProj -- all of which this type serves without
#  materialising N Julia objects.)  mean: D x N, cov: packed lower D(D+1)/2 x N, sqrt: D x D x N ROW-major per state.
struct SoAGaussians <: AbstractVector{Gaussian{Vector{Float64},SRMatrix{Float64}}}
    mean::Matrix{Float64}
    cov::Matrix{Float64}
    sqrt::Array{Float64,3}
    rows::UnitRange{Int}      # 1:D for states, 1:d for sol.pu (SolProj * x)
end
Base.size(x::SoAGaussians) = (size(x.mean, 2),)
function _mat(x::SoAGaussians, k)
    D = size(x.mean, 1)
    M = zeros(length(x.rows), length(x.rows))
    for i in x.rows, j in 1:i
        M[i, j] = M[j, i] = x.cov[(i - 1) * i ÷ 2 + j, k]
    end
    return M
end
# the library returns S row-major: sqrt[:, :, k] read column-major is S'; the factor of SolProj * x is rows 1:d of S
Base.getindex(x::SoAGaussians, k::Int) =
    Gaussian(x.mean[x.rows, k], SRMatrix(collect(transpose(view(x.sqrt, :, x.rows, k))), _mat(x, k)))
ProbNumDiffEq.mean(x::SoAGaussians) = [x.mean[x.rows, k] for k in 1:length(x)]
ProbNumDiffEq.var(x::SoAGaussians) = [[x.cov[(i - 1) * i ÷ 2 + i, k] for i in x.rows] for k in 1:length(x)]
ProbNumDiffEq.std(x::SoAGaussians) = [sqrt.(v) for v in ProbNumDiffEq.var(x)]
Base.getproperty(x::SoAGaussians, s::Symbol) =
    s === :μ ? ProbNumDiffEq.mean(x) : s === :Σ ? [x[k].Σ for k in 1:length(x)] : getfield(x, s)

"""History of trajectory i (0-based): t, lazy states, diffusions -- three library calls, no host factorisation
(pnde_get_history_sqrt returns the factors the kernels carry)."""
function _history(h, which, i, nsaved, D)
    off = zeros(Int64, 2)
    t = zeros(nsaved); mean = zeros(D, nsaved); cov = zeros(D * (D + 1) ÷ 2, nsaved); diff = zeros(nsaved)
    sq = zeros(D, D, nsaved)
    check(ccall((:pnde_get_history, libpnde), Cint,
                (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                h, which, i, i + 1, off, t, mean, cov, diff), h)
    check(ccall((:pnde_get_history_sqrt, libpnde), Cint,
                (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int64}, Ptr{Float64}), h, which, i, i + 1, off, sq), h)
    return t, SoAGaussians(mean, cov, sq, 1:D), diff[2:end]
end

"""solve_ieks (src/ieks.jl:53-61) as ONE library call: all iterates (EK1 solves linearised at the previous iterate's
dense output, src/perform_step.jl:111-113) run on the device; `alg.linearize_at` is not used."""
solve_ieks_b200(prob::DiffEqBase.AbstractODEProblem, alg::IEKS, args...; iterations=10, kwargs...) =
    DiffEqBase.solve(prob, alg, args...; ieks_iterations=iterations, kwargs...)

function DiffEqBase.__solve(prob::DiffEqBase.AbstractODEProblem, alg::AbstractEK; kwargs...)
    cfg, f = _config(prob, alg; kwargs...)
    u0 = reshape(collect(Float64, prob.u0), 1, :)
    p = reshape(collect(Float64, prob.p), 1, :)
    h, n = _run(cfg, f, u0, p)
    try
        na, nr, nf, nj, rc, ns = _counts(h, n)
        d = length(prob.u0)
        D = d * (alg.order + 1)
        t, x_filt, diffusions = _history(h, 0, 0, ns[1], D)
        x_smooth = cfg.smooth == 1 ? _history(h, 1, 0, ns[1], D)[2] : x_filt
        src = cfg.smooth == 1 ? x_smooth : x_filt
        pu = SoAGaussians(src.mean, src.cov, src.sqrt, 1:d)           # SolProj * x (src/integrator_utils.jl:45, :22)
        sol = DiffEqBase.build_solution(prob, alg, t, ProbNumDiffEq.mean(pu); retcode = RETCODES[rc[1]],
                                        destats = DiffEqBase.DEStats(0))
        sol.destats.naccept, sol.destats.nreject, sol.destats.nf, sol.destats.njacs = na[1], nr[1], nf[1], nj[1]
        sol.pu, sol.x_filt, sol.x_smooth, sol.diffusions = pu, x_filt, x_smooth, diffusions
        ll = zeros(1)
        check(ccall((:pnde_get_final, libpnde), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    h, C_NULL, C_NULL, C_NULL, ll), h)
        sol.log_likelihood = ll[1]
        return sol
    finally
        ccall((:pnde_destroy, libpnde), Cint, (Ptr{Cvoid},), h)
    end
end

function DiffEqBase.__solve(eprob::DiffEqBase.AbstractEnsembleProblem, alg::AbstractEK, ens::EnsembleB200;
                            trajectories, save_everystep=false, kwargs...)
    probs = [eprob.prob_func(eprob.prob, i, 1) for i in 1:trajectories]     # SURVEY App. B.5
    u0 = permutedims(reduce(hcat, [collect(Float64, pr.u0) for pr in probs]))   # n x d  == [d][n]
    p = permutedims(reduce(hcat, [collect(Float64, pr.p) for pr in probs]))
    cfg, f = _config(eprob.prob, alg; save_everystep = save_everystep, devices = ens.devices, kwargs...)
    h, n = _run(cfg, f, u0, p)
    D = size(u0, 2) * (alg.order + 1)
    mean = zeros(n, D); cov = zeros(n, D * (D + 1) ÷ 2); tf = zeros(n); ll = zeros(n)
    check(ccall((:pnde_get_final, libpnde), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                h, mean, cov, tf, ll), h)
    na, nr, nf, nj, rc, ns = _counts(h, n)
    ccall((:pnde_destroy, libpnde), Cint, (Ptr{Cvoid},), h)
    # final filtering states of all trajectories as ONE lazy view (mean / packed covariance; the factor of a final
    # state is available through save_everystep = true and _history)
    us = SoAGaussians(permutedims(mean), permutedims(cov), zeros(D, D, 0), 1:D)
    return DiffEqBase.EnsembleSolution(us, 0.0, all(rc .== 0))
end

"""perform_step! once from given states (pnde_step_from_state): what a step!/callback driver calls
(examples/fitzhughnagumo_animation.jl:23-26).  xs: Vector of Gaussian{Vector,SRMatrix}; returns the filtered means,
packed covariances, local diffusions and error estimates."""
function perform_step_b200(prob, alg::AbstractEK, xs, ts, dts, ps, uprev; kwargs...)
    cfg, f = _config(prob, alg; save_everystep = false, kwargs...)
    h = _create(cfg, f)
    n = length(xs); d = length(prob.u0); D = d * (alg.order + 1)
    mean = permutedims(reduce(hcat, [x.μ for x in xs]))                              # n x D
    sq = permutedims(reduce(hcat, [vec(permutedims(x.Σ.squareroot)) for x in xs]))   # n x D*D, row-major per state
    mo = zeros(n, D); co = zeros(n, D * (D + 1) ÷ 2); s2 = zeros(n, alg isa EK1 ? 1 : d); ee = zeros(n)
    uo = zeros(n, d); ql = zeros(n, 2); st = zeros(Int32, n)
    try
        check(ccall((:pnde_step_from_state, libpnde), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                    h, n, mean, sq, collect(Float64, ts), collect(Float64, dts), permutedims(reduce(hcat, ps)),
                    permutedims(reduce(hcat, uprev)), mo, co, s2, ee, uo, ql, st), h)
    finally
        ccall((:pnde_destroy, libpnde), Cint, (Ptr{Cvoid},), h)
    end
    return (mean = mo, cov = co, sigma2 = s2, EEst = ee, u = uo, status = st)
end

export EnsembleB200, CatalogueFunction, CustomFunction, custom_function, solve_ieks_b200, perform_step_b200

end # module
