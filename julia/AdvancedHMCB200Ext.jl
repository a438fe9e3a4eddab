# AdvancedHMCB200Ext.jl -- the reference-side binding of libahmc_b200 (include/ahmc_b200.h).
#
# STATUS: WRITTEN, NOT EXECUTED.  No `julia` binary exists in the build image (SURVEY.md section 8c), so this
# file has never been parsed or run.  It shows the `ccall` stubs a maintainer would add as a package
# extension next to ext/AdvancedHMCCUDAExt.jl; the same entry points are exercised from Python
# (advancedhmc.jl_b200/core.py) by the test-suite.  Citations are relative to the AdvancedHMC.jl checkout.
#
# Plug-in points used (SURVEY.md section 8b):
#   * a new integrator type `B200Leapfrog{T} <: AbstractLeapfrog{T}` with its own `step` method -- the
#     mechanism ext/AdvancedHMCOrdinaryDiffEqSymplecticRKExt.jl:6-13 uses for `DiffEqIntegrator`;
#   * whole-transition overrides `transition(rng, h, tau, z)` (src/trajectory.jl:271-276, 677-681) when the
#     integrator is a `B200Leapfrog`, so refresh + trajectory + MH / NUTS tree run as ONE kernel;
#   * `Hamiltonian`, `AbstractMetric`, adaptors and `sample` (src/sampler.jl:159-248) are untouched.
module AdvancedHMCB200Ext

using AdvancedHMC
using AdvancedHMC: AbstractLeapfrog, Hamiltonian, PhasePoint, DualValue, Trajectory, Transition,
    UnitEuclideanMetric, DiagEuclideanMetric, DenseEuclideanMetric, EndPointTS, MultinomialTS,
    FixedNSteps, FixedIntegrationTime, GeneralisedNoUTurn, step_size, nom_step_size, nsteps
using CUDA

const libahmc = get(ENV, "AHMC_B200_LIB", "libahmc_b200.so")

# ---- C structs (include/ahmc_b200.h) --------------------------------------------------------------
struct CMetric
    kind::Int32
    Minv::Ptr{Float64}
    chain_stride::Int64
    cholU::Ptr{Float64}
end
struct CPhasePoint
    theta::Ptr{Float64}
    r::Ptr{Float64}
    lp_value::Ptr{Float64}
    lp_gradient::Ptr{Float64}
    lk_value::Ptr{Float64}
    lk_gradient::Ptr{Float64}
    ld::Int64
end
struct CStats
    n_steps::Ptr{Int32}
    is_accept::Ptr{UInt8}
    acceptance_rate::Ptr{Float64}
    log_density::Ptr{Float64}
    hamiltonian_energy::Ptr{Float64}
    hamiltonian_energy_error::Ptr{Float64}
    max_hamiltonian_energy_error::Ptr{Float64}
    tree_depth::Ptr{Int32}
    numerical_error::Ptr{UInt8}
end
struct CRng
    seed::UInt64
    offset::UInt64
    normal_tape::Ptr{Float64}
    exp_tape::Ptr{Float64}
    exp_stride::Int64
    dir_tape::Ptr{UInt8}
    dir_stride::Int64
    partial_refresh_alpha::Float64
end

const FLAG_ASYNC = 0x4 % UInt32
const FLAG_NUTS_SLICE_TS = 0x20 % UInt32
const FLAG_NUTS_CLASSIC = 0x40 % UInt32
const FLAG_NUTS_STRICT = 0x80 % UInt32

# ---- context / models -------------------------------------------------------------------------------
mutable struct B200Context
    h::Ptr{Cvoid}
end
const CTX = Ref{Union{Nothing,B200Context}}(nothing)

function context()
    if CTX[] === nothing
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ahmc_create, libahmc), Cint, (Ref{Ptr{Cvoid}}, Int32, Ptr{Cvoid}), out, CUDA.deviceid(), CUDA.stream().handle)
        rc == 0 || error("ahmc_create failed ($rc)")
        CTX[] = B200Context(out[])
    end
    return CTX[]
end

check(rc) = rc == 0 ? nothing :
    (msg = unsafe_string(ccall((:ahmc_last_error, libahmc), Cstring, (Ptr{Cvoid},), context().h));
     rc == -1 ? throw(ArgumentError(msg)) : error("ahmc error $rc: $msg"))

"Built-in target (replaces the `lp` / `dlp/dtheta` closures for the fused kernels)."
struct B200Target
    handle::Ptr{Cvoid}
    D::Int
end
function B200Target(kind::Integer, D::Integer; p0=C_NULL, p1=C_NULL, c0=0.0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ahmc_model_create, libahmc), Cint,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Ptr{Cvoid}}),
                context().h, kind, D, p0, p1, c0, out))
    return B200Target(out[], D)
end

# ---- the integrator plug-in (src/integrator.jl:49-60 interface) -------------------------------------
struct B200Leapfrog{T<:AdvancedHMC.AbstractScalarOrVec{<:AbstractFloat}} <: AbstractLeapfrog{T}
    ϵ::T
    target::B200Target
end
AdvancedHMC.update_nom_step_size(lf::B200Leapfrog, ϵ) = B200Leapfrog(ϵ, lf.target)

cmetric(m::UnitEuclideanMetric, N) = CMetric(0, C_NULL, 0, C_NULL)
cmetric(m::DiagEuclideanMetric, N) =
    CMetric(1, pointer(m.M⁻¹), ndims(m.M⁻¹) == 2 ? size(m.M⁻¹, 1) : 0, C_NULL)
cmetric(m::DenseEuclideanMetric, N) = CMetric(2, pointer(m.M⁻¹), 0, pointer(CuArray(Matrix(m.cholM⁻¹))))

cpp(z::PhasePoint{<:CuMatrix}) = CPhasePoint(pointer(z.θ), pointer(z.r), pointer(z.ℓπ.value), pointer(z.ℓπ.gradient),
                                            pointer(z.ℓκ.value), pointer(z.ℓκ.gradient), size(z.θ, 1))

refresh_alpha(::AdvancedHMC.FullMomentumRefreshment) = 0.0
refresh_alpha(r::AdvancedHMC.PartialMomentumRefreshment) = Float64(r.α)

eps_args(ϵ::AbstractFloat) = (Float64(ϵ), Ptr{Float64}(C_NULL))
eps_args(ϵ::CuVector{Float64}) = (0.0, pointer(ϵ))

"`step` for CuArray phase points: replaces src/integrator.jl:216-265 (about 10 broadcast kernels and 4
host-syncing `all(isfinite)` per step) with ONE fused kernel for all n_steps."
function AdvancedHMC.step(lf::B200Leapfrog, h::Hamiltonian, z::PhasePoint{<:CuMatrix{Float64}}, n_steps::Int=1;
                          fwd::Bool=n_steps > 0, full_trajectory::Val{FullTraj}=Val(false)) where {FullTraj}
    FullTraj && error("full_trajectory is a 'next' row (SURVEY 8f-1); use the stock Leapfrog for MultinomialTS static")
    D, N = size(z.θ)
    zout = PhasePoint(similar(z.θ), similar(z.r), DualValue(similar(z.ℓπ.value), similar(z.ℓπ.gradient)),
                      DualValue(similar(z.ℓκ.value), similar(z.ℓκ.gradient)))
    ϵ, ϵp = eps_args(step_size(lf))
    n = fwd ? abs(n_steps) : -abs(n_steps)
    md = Ref(cmetric(h.metric, N)); zi = Ref(cpp(z)); zo = Ref(cpp(zout))
    GC.@preserve z zout h lf begin
        check(ccall((:ahmc_leapfrog_f64, libahmc), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CMetric}, Int32, Int64, Float64, Ptr{Float64}, Int32, Float64,
                     Ref{CPhasePoint}, Ref{CPhasePoint}, Ptr{UInt32}, Ptr{Int32}, UInt32),
                    context().h, lf.target.handle, md, D, N, ϵ, ϵp, n, 0.0, zi, zo, C_NULL, C_NULL, 0))
    end
    return zout
end

"Static HMC transition (src/trajectory.jl:271-300) as ONE kernel: refresh + L steps + MH + revert + flip."
function AdvancedHMC.transition(rng, h::Hamiltonian, κ::AdvancedHMC.HMCKernel{R,<:Trajectory{EndPointTS,<:B200Leapfrog}},
                                z::PhasePoint{<:CuMatrix{Float64}}) where {R}
    τ = κ.τ; lf = τ.integrator
    D, N = size(z.θ)
    zout = PhasePoint(similar(z.θ), similar(z.r), DualValue(similar(z.ℓπ.value), similar(z.ℓπ.gradient)),
                      DualValue(similar(z.ℓκ.value), similar(z.ℓκ.gradient)))
    acc = CUDA.zeros(UInt8, N); α = CUDA.zeros(Float64, N); H = CUDA.zeros(Float64, N); dH = CUDA.zeros(Float64, N)
    nerr = CUDA.zeros(UInt8, N)
    st = Ref(CStats(C_NULL, pointer(acc), pointer(α), C_NULL, pointer(H), pointer(dH), C_NULL, C_NULL, pointer(nerr)))
    rg = Ref(CRng(rand(rng, UInt64), 0, C_NULL, C_NULL, 0, C_NULL, 0, refresh_alpha(κ.refreshment)))   # Philox key drawn from the Julia rng
    ϵ, ϵp = eps_args(step_size(lf))
    md = Ref(cmetric(h.metric, N)); zi = Ref(cpp(z)); zo = Ref(cpp(zout))
    zo[] = CPhasePoint(zo[].theta, zo[].r, zo[].lp_value, zo[].lp_gradient, zo[].lk_value, C_NULL, zo[].ld)
    GC.@preserve z zout acc α H dH nerr begin
        check(ccall((:ahmc_hmc_transition_f64, libahmc), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CMetric}, Int32, Int64, Float64, Ptr{Float64}, Int32, Ref{CRng},
                     Ref{CPhasePoint}, Ref{CPhasePoint}, Ref{CStats}, UInt32),
                    context().h, lf.target.handle, md, D, N, ϵ, ϵp, nsteps(τ), rg, zi, zo, st, 0))
    end
    tstat = merge((n_steps=nsteps(τ), is_accept=acc .== 1, acceptance_rate=α, log_density=zout.ℓπ.value,
                   hamiltonian_energy=H, hamiltonian_energy_error=dH, numerical_error=any(nerr .== 1)),
                  AdvancedHMC.stat(lf))
    return Transition(zout, tstat)
end

"Many-chain NUTS (MultinomialTS + GeneralisedNoUTurn; the reference's src/trajectory.jl:677-742 is scalar-only)."
function AdvancedHMC.transition(rng, h::Hamiltonian,
                                κ::AdvancedHMC.HMCKernel{R,<:Trajectory{TS,<:B200Leapfrog,TC}},
                                z::PhasePoint{<:CuMatrix{Float64}}) where {R,TS<:Union{MultinomialTS,SliceTS},
                                                                          TC<:AdvancedHMC.DynamicTerminationCriterion}
    τ = κ.τ; lf = τ.integrator; tc = τ.termination_criterion
    # sampler / criterion variants are flag bits of the same entry point (trajectory.jl:102-109, 551-557, 579-613)
    flags = (TS <: SliceTS ? FLAG_NUTS_SLICE_TS : 0x0 % UInt32) |
            (TC <: ClassicNoUTurn ? FLAG_NUTS_CLASSIC : TC <: StrictGeneralisedNoUTurn ? FLAG_NUTS_STRICT : 0x0 % UInt32)
    D, N = size(z.θ)
    zout = PhasePoint(similar(z.θ), similar(z.r), DualValue(similar(z.ℓπ.value), similar(z.ℓπ.gradient)),
                      DualValue(similar(z.ℓκ.value), similar(z.ℓκ.gradient)))
    ns = CUDA.zeros(Int32, N); α = CUDA.zeros(Float64, N); H = CUDA.zeros(Float64, N); dH = CUDA.zeros(Float64, N)
    mx = CUDA.zeros(Float64, N); td = CUDA.zeros(Int32, N); nerr = CUDA.zeros(UInt8, N)
    st = Ref(CStats(pointer(ns), C_NULL, pointer(α), C_NULL, pointer(H), pointer(dH), pointer(mx), pointer(td), pointer(nerr)))
    rg = Ref(CRng(rand(rng, UInt64), 0, C_NULL, C_NULL, 0, C_NULL, 0, refresh_alpha(κ.refreshment)))
    ϵ, ϵp = eps_args(step_size(lf))
    md = Ref(cmetric(h.metric, N)); zi = Ref(cpp(z)); zo = Ref(cpp(zout))
    zo[] = CPhasePoint(zo[].theta, zo[].r, zo[].lp_value, zo[].lp_gradient, zo[].lk_value, C_NULL, zo[].ld)
    GC.@preserve z zout ns α H dH mx td nerr begin
        check(ccall((:ahmc_nuts_transition_f64, libahmc), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CMetric}, Int32, Int64, Float64, Ptr{Float64}, Int32, Float64, Ref{CRng},
                     Ref{CPhasePoint}, Ref{CPhasePoint}, Ref{CStats}, UInt32),
                    context().h, lf.target.handle, md, D, N, ϵ, ϵp, tc.max_depth, tc.Δ_max, rg, zi, zo, st, flags))
    end
    tstat = merge((n_steps=ns, is_accept=true, acceptance_rate=α, log_density=zout.ℓπ.value, hamiltonian_energy=H,
                   hamiltonian_energy_error=dH, max_hamiltonian_energy_error=mx, tree_depth=td, numerical_error=nerr .== 1),
                  AdvancedHMC.stat(lf))
    return Transition(zout, tstat)
end

struct CAdaptCfg
    n_adapts::Int32; init_buffer::Int32; term_buffer::Int32; window_size::Int32
    delta::Float64; gamma::Float64; t0::Float64; kappa::Float64
    adapt_metric::Int32; n_min::Int32
    eps_chain::Ptr{Float64}; Minv_chain::Ptr{Float64}; eps_trace::Ptr{Float64}
end

"""
`sample(rng, h, κ, θ, n_samples, adaptor, n_adapts)` (src/sampler.jl:159-248) for many-chain NUTS with the reference's
vectorised adaptors -- `StanHMCAdaptor(WelfordVar((D, N)), NesterovDualAveraging(δ, ϵ::Vector))` -- as ONE launch
(ahmc_nuts_adapt_sample_f64): every chain adapts its own ϵ and diagonal M⁻¹ and never waits for another chain.
Returns (θ draws D×N×n_samples, final per-chain ϵ, per-chain M⁻¹ D×N).
"""
function b200_sample_nuts(rng, h::Hamiltonian, lf::B200Leapfrog, tc::GeneralisedNoUTurn, θ::CuMatrix{Float64},
                          n_samples::Int, n_adapts::Int; δ=0.8, adapt_metric=true, init_buffer=75, term_buffer=50,
                          window_size=25)
    D, N = size(θ)
    z = b200_phasepoint(lf.target, h, θ, CUDA.zeros(Float64, D, N))
    zout = PhasePoint(similar(z.θ), similar(z.r), DualValue(similar(z.ℓπ.value), similar(z.ℓπ.gradient)),
                      DualValue(similar(z.ℓκ.value), similar(z.ℓκ.gradient)))
    ϵ = CUDA.fill(Float64(first(step_size(lf))), N); Minv = CUDA.ones(Float64, D, N)
    draws = CUDA.zeros(Float64, D, N, n_samples)
    α = CUDA.zeros(Float64, N * n_samples)
    st = Ref(CStats(C_NULL, C_NULL, pointer(α), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL))
    cfg = Ref(CAdaptCfg(n_adapts, init_buffer, term_buffer, window_size, δ, 0.05, 10.0, 0.75, adapt_metric ? 1 : 0, 10,
                        pointer(ϵ), pointer(Minv), C_NULL))
    rg = Ref(CRng(rand(rng, UInt64), 0, C_NULL, C_NULL, 0, C_NULL, 0, 0.0))
    md = Ref(cmetric(h.metric, N)); zi = Ref(cpp(z)); zo = Ref(cpp(zout))
    zo[] = CPhasePoint(zo[].theta, zo[].r, zo[].lp_value, zo[].lp_gradient, zo[].lk_value, C_NULL, zo[].ld)
    GC.@preserve z zout ϵ Minv draws α begin
        check(ccall((:ahmc_nuts_adapt_sample_f64, libahmc), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CMetric}, Int32, Int64, Int32, Float64, Int32, Ref{CAdaptCfg}, Ref{CRng},
                     Ref{CPhasePoint}, Ref{CPhasePoint}, Ptr{Float64}, Ref{CStats}, UInt32),
                    context().h, lf.target.handle, md, D, N, tc.max_depth, tc.Δ_max, n_samples, cfg, rg, zi, zo,
                    pointer(draws), st, 0))
    end
    return draws, ϵ, Minv
end

"`phasepoint(h, θ, r)` (src/hamiltonian.jl:115-119) for a B200 target."
function b200_phasepoint(t::B200Target, h::Hamiltonian, θ::CuMatrix{Float64}, r::CuMatrix{Float64})
    D, N = size(θ)
    z = PhasePoint(θ, r, DualValue(CUDA.zeros(Float64, N), similar(θ)), DualValue(CUDA.zeros(Float64, N), similar(θ)))
    md = Ref(cmetric(h.metric, N)); zc = Ref(cpp(z))
    GC.@preserve z check(ccall((:ahmc_phasepoint_f64, libahmc), Cint,
                               (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CMetric}, Int32, Int64, Ref{CPhasePoint}, UInt32),
                               context().h, t.handle, md, D, N, zc, 0))
    return z
end

"Pooled adaptor record of one iteration: [N, sum min(1,α), mean(θ), M2(θ)] (ahmc_adapt_summary_f64)."
function b200_adapt_summary(θ::CuMatrix{Float64}, α::CuVector{Float64})
    D, N = size(θ)
    out = CUDA.zeros(Float64, 2 + 2D)
    GC.@preserve θ α out check(ccall((:ahmc_adapt_summary_f64, libahmc), Cint,
                                     (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, UInt32),
                                     context().h, D, N, pointer(θ), D, pointer(α), pointer(out), 0))
    return out
end

end # module
