# ProbNumDiffEqB200.jl -- Julia side of the drop-in boundary (source only: Julia is not installed in
# the build image, so this file is syntax-reviewed, not executed; the same C ABI is exercised by the
# Python mirror in ../api.py and by tests/).
#
# It overloads DiffEqBase.__solve for the reference's algorithm types (src/algorithms.jl:23-51) so that
#     solve(prob, EK1(order=3); abstol, reltol)                      and
#     solve(EnsembleProblem(prob; prob_func), EK1(order=3), EnsembleB200(); trajectories=N, ...)
# run the whole OrdinaryDiffEq solve! loop inside libpnde.so (include/pnde.h) and rebuild the fields of
# ProbODESolution (src/solution.jl:8-24).
module ProbNumDiffEqB200

using ProbNumDiffEq
using ProbNumDiffEq: AbstractEK, EK0, EK1, IEKS, SRMatrix, ProbODESolution
using DiffEqBase
using GaussianDistributions: Gaussian
using StructArrays

const libpnde = get(ENV, "PNDE_LIB", joinpath(@__DIR__, "..", "libpnde.so"))

# ---- mirror of `struct pnde_config` (include/pnde.h); field order and types must match ----------
Base.@kwdef mutable struct PndeConfig
    abi_version::Int32 = 1
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
end

const DIFFUSIONS = Dict(:dynamic => 0, :fixed => 1, :fixedMAP => 2, :dynamicMV => 3, :fixedMV => 4)  # src/caches.jl:89-96
const RETCODES = Dict(0 => :Success, 1 => :MaxIters, 2 => :DtNaN, 3 => :Unstable, 4 => :Failure, 5 => :DtLessThanMin)

"""Vector fields cannot cross the C ABI as closures: problems name a catalogue entry (include/pnde.h)."""
struct CatalogueFunction
    kind::Int32          # PNDE_VF_*
    f                    # the Julia function, kept for host-side use (plotting, analytic errors)
end

struct EnsembleB200 <: DiffEqBase.EnsembleAlgorithm end

check(rc, h) = rc == 0 || error(unsafe_string(ccall((:pnde_last_error, libpnde), Cstring, (Ptr{Cvoid},), h)))

function _create(cfg::PndeConfig)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:pnde_create, libpnde), Cint, (Ref{PndeConfig}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error(unsafe_string(ccall((:pnde_last_error, libpnde), Cstring, (Ptr{Cvoid},), C_NULL)))
    return h[]
end

function _config(prob, alg::AbstractEK; abstol=1e-6, reltol=1e-3, adaptive=true, dt=nothing,
                 save_everystep=true, maxiters=100000, max_saved=0, device=-1, ieks_iterations=0, kwargs...)
    !adaptive && dt === nothing && error("Fixed timestep methods require a choice of dt")
    f = prob.f.f
    f isa CatalogueFunction || error("ProbNumDiffEqB200 needs a catalogue vector field (see include/pnde.h)")
    smooth = alg.smooth && save_everystep
    PndeConfig(alg = alg isa IEKS ? 2 : alg isa EK1 ? 1 : 0, ieks_iterations = ieks_iterations, order = alg.order, vf_kind = f.kind,
               diffusion = DIFFUSIONS[alg.diffusionmodel], smooth = smooth, adaptive = adaptive,
               save_mode = save_everystep ? 1 : 0, device = device, abstol = abstol, reltol = reltol,
               dt = dt === nothing ? 0.0 : dt, t0 = prob.tspan[1], t1 = prob.tspan[2],
               maxiters = maxiters, max_saved = max_saved == 0 && adaptive && save_everystep ? 8192 : max_saved)
end

"""Run one ensemble (n trajectories, SoA inputs u0[d, n]' / p[np, n]' with the trajectory index fastest)."""
function _run(cfg::PndeConfig, u0::Matrix{Float64}, p::Matrix{Float64})
    h = _create(cfg)
    n = size(u0, 1)   # u0 is n x d in Julia's column-major layout == [d][n] with n fastest
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

unpack_lower(v, D) = (M = zeros(D, D); k = 0; for i in 1:D, j in 1:i; k += 1; M[i, j] = v[k]; M[j, i] = v[k]; end; M)

"""History of trajectory i (0-based) as StructArray{Gaussian{Vector,SRMatrix}} (src/solution.jl:60-64)."""
function _history(h, which, i, nsaved, D)
    off = zeros(Int64, 2)
    t = zeros(nsaved); mean = zeros(D, nsaved); cov = zeros(D * (D + 1) ÷ 2, nsaved); diff = zeros(nsaved)
    check(ccall((:pnde_get_history, libpnde), Cint,
                (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                h, which, i, i + 1, off, t, mean, cov, diff), h)
    xs = StructArray([Gaussian(mean[:, k], _srmatrix(unpack_lower(view(cov, :, k), D))) for k in 1:nsaved])
    return t, xs, diff[2:end]
end

# SRMatrix(squareroot, mat): the reference stores both (src/squarerootmatrix.jl:10-16); any S with S*S' == mat is valid
function _srmatrix(mat)
    F = eigen(Symmetric(mat))
    S = F.vectors * Diagonal(sqrt.(max.(F.values, 0)))
    return SRMatrix(S, mat)
end

"""solve_ieks (src/ieks.jl:53-61) as ONE library call: all iterates (EK1 solves linearised at the previous iterate's
dense output, src/perform_step.jl:111-113) run on the device; `alg.linearize_at` is not used."""
solve_ieks_b200(prob::DiffEqBase.AbstractODEProblem, alg::IEKS, args...; iterations=10, kwargs...) =
    DiffEqBase.solve(prob, alg, args...; ieks_iterations=iterations, kwargs...)

function DiffEqBase.__solve(prob::DiffEqBase.AbstractODEProblem, alg::AbstractEK; kwargs...)
    cfg = _config(prob, alg; kwargs...)
    u0 = reshape(collect(Float64, prob.u0), 1, :)
    p = reshape(collect(Float64, prob.p), 1, :)
    h, n = _run(cfg, u0, p)
    try
        na, nr, nf, nj, rc, ns = _counts(h, n)
        D = length(prob.u0) * (alg.order + 1)
        t, x_filt, diffusions = _history(h, 0, 0, ns[1], D)
        x_smooth = cfg.smooth == 1 ? _history(h, 1, 0, ns[1], D)[2] : copy(x_filt)
        src = cfg.smooth == 1 ? x_smooth : x_filt
        d = length(prob.u0)
        E0 = [I(d) zeros(d, D - d)]
        pu = StructArray([E0 * x for x in src])                      # src/integrator_utils.jl:45, :22
        sol = DiffEqBase.build_solution(prob, alg, t, [x.μ for x in pu]; retcode = RETCODES[rc[1]],
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

function DiffEqBase.__solve(eprob::DiffEqBase.AbstractEnsembleProblem, alg::AbstractEK, ::EnsembleB200;
                            trajectories, save_everystep=false, kwargs...)
    probs = [eprob.prob_func(eprob.prob, i, 1) for i in 1:trajectories]     # SURVEY App. B.5
    u0 = permutedims(reduce(hcat, [collect(Float64, pr.u0) for pr in probs]))   # n x d  == [d][n]
    p = permutedims(reduce(hcat, [collect(Float64, pr.p) for pr in probs]))
    cfg = _config(eprob.prob, alg; save_everystep = save_everystep, kwargs...)
    h, n = _run(cfg, u0, p)
    D = size(u0, 2) * (alg.order + 1)
    mean = zeros(n, D); cov = zeros(n, D * (D + 1) ÷ 2); tf = zeros(n); ll = zeros(n)
    check(ccall((:pnde_get_final, libpnde), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                h, mean, cov, tf, ll), h)
    na, nr, nf, nj, rc, ns = _counts(h, n)
    ccall((:pnde_destroy, libpnde), Cint, (Ptr{Cvoid},), h)
    # final filtering Gaussians per trajectory; full histories via save_everystep=true and _history
    us = [Gaussian(mean[i, :], _srmatrix(unpack_lower(view(cov, i, :), D))) for i in 1:n]
    return DiffEqBase.EnsembleSolution(us, 0.0, all(rc .== 0))
end

export EnsembleB200, CatalogueFunction

end # module
