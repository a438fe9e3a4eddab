# gen_from_reference.jl -- golden vectors produced by THE REFERENCE ITSELF (AdvancedHMC.jl, unmodified).
#
#   julia --project=/path/to/AdvancedHMC.jl tests/golden/gen_from_reference.jl        (needs JSON.jl in that environment)
#
# STATUS: WRITTEN, NOT EXECUTED -- no `julia` binary exists in the build image (SURVEY.md 8c), so the three files this
# script writes (leapfrog_ref.json, hmc_ref.json, nuts_ref.json) are not committed yet and parity stays pinned by the
# 50-digit restatements only.  Whoever has Julia runs it once and commits the outputs; tests/helpers.py picks them up
# automatically (cases named "ref:<name>") and both the oracle (CPU suite) and the CUDA path (-m gpu) are then compared
# with numbers the reference produced.
#
# What it does: for every case of the committed *_mp50.json fixtures it rebuilds the SAME inputs (target, metric, theta0, r0,
# step size, number of steps, random tapes) and drives the real entry points
#     AdvancedHMC.step(lf, h, z, n_steps)                      src/integrator.jl:216-265   (matrix mode, all chains at once)
#     AdvancedHMC.transition(rng, h, tau::Trajectory, z)       src/trajectory.jl:271-300   (static, EndPointTS / MultinomialTS)
#     AdvancedHMC.transition(rng, h, tau::Trajectory, z)       src/trajectory.jl:677-742   (NUTS, one chain at a time: the
#                                                                                            reference's NUTS is scalar-only)
# with a TAPE rng (below) that hands the reference the fixture's own normals / exponentials / uniforms / direction bits in
# the order it asks for them (`rand(rng, Bool)` :693, `randexp` :145,193,205,858, `rand(rng)` :202, `randn` utilities.jl:5-23,
# `rand_coupled` :371-373).  Output schema = the mp50 files' (same case records, `expect` replaced by what the reference
# returned, `generator` names the AdvancedHMC version).
using AdvancedHMC
using AdvancedHMC: Hamiltonian, PhasePoint, phasepoint, Leapfrog, TemperedLeapfrog, Trajectory, EndPointTS, MultinomialTS, SliceTS,
    FixedNSteps, GeneralisedNoUTurn, ClassicNoUTurn, StrictGeneralisedNoUTurn, UnitEuclideanMetric, DiagEuclideanMetric,
    DenseEuclideanMetric, FullMomentumRefreshment, refresh, transition
using JSON
using LinearAlgebra
using Random

const HERE = @__DIR__

# ---------------------------------------------------------------------------------------------------------------- tape rng
mutable struct TapeRNG <: AbstractRNG
    normals::Vector{Float64}
    variates::Vector{Float64}   # exponentials and uniforms, in the order the reference draws them
    dirs::Vector{Bool}
    n_fwd::Int
    in::Int
    iv::Int
    id::Int
end
TapeRNG(; normals=Float64[], variates=Float64[], dirs=Bool[], n_fwd=0) = TapeRNG(normals, variates, dirs, n_fwd, 0, 0, 0)
next_variate(rng::TapeRNG) = (rng.iv += 1; rng.variates[rng.iv])
Random.rand(rng::TapeRNG, ::Random.SamplerType{Bool}) = (rng.id += 1; rng.dirs[rng.id])
Random.rand(rng::TapeRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}}) = next_variate(rng)
Random.randexp(rng::TapeRNG) = next_variate(rng)
Random.randexp(rng::TapeRNG, ::Type{Float64}) = next_variate(rng)
Random.randn(rng::TapeRNG, ::Type{Float64}) = (rng.in += 1; rng.normals[rng.in])
Random.randn(rng::TapeRNG) = randn(rng, Float64)
AdvancedHMC.rand_coupled(rng::TapeRNG, args...) = rng.n_fwd   # trajectory.jl:371-373: one draw shared by all chains

# ---------------------------------------------------------------------------------------------------------------- targets
# value and PLUS gradient of log pi for a vector or a D x N matrix (what the reference's `∂ℓπ∂θ` closure returns)
mat(rows) = rows === nothing ? nothing : Float64.(reduce(hcat, [Float64.(r) for r in rows]))   # JSON [N][D] -> D x N
vecf(x) = x === nothing ? nothing : Float64.(x)

function target(kind::String, D::Int, p0, p1, c0::Float64)
    if kind == "std_normal"
        lp = θ -> θ isa AbstractVector ? c0 - sum(abs2, θ) / 2 : c0 .- vec(sum(abs2, θ; dims=1)) ./ 2
        return lp, θ -> (lp(θ), -θ)
    elseif kind == "diag_gauss"          # p0 = mean, p1 = standard deviations
        m, w = vecf(p0), 1.0 ./ (vecf(p1) .^ 2)
        lp = θ -> θ isa AbstractVector ? c0 - sum(abs2.(θ .- m) .* w) / 2 : c0 .- vec(sum(abs2.(θ .- m) .* w; dims=1)) ./ 2
        return lp, θ -> (lp(θ), -(θ .- m) .* w)
    elseif kind == "dense_gauss"         # p0 = mean, p1 = precision matrix
        m, P = vecf(p0), mat(p1)
        lp = θ -> θ isa AbstractVector ? c0 - dot(θ .- m, P * (θ .- m)) / 2 : c0 .- vec(sum((θ .- m) .* (P * (θ .- m)); dims=1)) ./ 2
        return lp, θ -> (lp(θ), -(P * (θ .- m)))
    elseif kind == "funnel"              # SURVEY 8c: lp = c0 - v^2/18 - (sum_{i>=2} x_i^2 e^{-v} + (D-1) v)/2
        function lpg(θ::AbstractVector)
            v = θ[1]; x = θ[2:end]; ev = exp(-v); S = sum(abs2, x) * ev
            g = similar(θ); g[1] = -v / 9 + (S - (D - 1)) / 2; g[2:end] .= -x .* ev
            return c0 - v^2 / 18 - (S + (D - 1) * v) / 2, g
        end
        function lpg(θ::AbstractMatrix)
            vals = zeros(size(θ, 2)); G = similar(θ)
            for c in axes(θ, 2)
                vals[c], G[:, c] = lpg(θ[:, c])
            end
            return vals, G
        end
        return θ -> lpg(θ)[1], lpg
    end
    error("unknown target $kind")
end

function metric(kind::String, D::Int, N::Int, Minv; matrix_mode::Bool)
    if kind == "unit"
        return matrix_mode ? UnitEuclideanMetric((D, N)) : UnitEuclideanMetric(D)
    elseif kind == "diag"
        # per-chain form (metric.jl:64): the fixture stores it as [D][N] rows, `mat` turns inner lists into columns
        M = Minv[1] isa AbstractVector ? permutedims(mat(Minv)) : vecf(Minv)
        return DiagEuclideanMetric(M)
    else
        return DenseEuclideanMetric(mat(Minv))
    end
end

rows(A::AbstractMatrix) = [collect(A[:, c]) for c in axes(A, 2)]   # D x N -> JSON [N][D]

# ---------------------------------------------------------------------------------------------------------------- step
function run_leapfrog(c)
    D, N = c["D"], c["N"]
    lp, dlp = target(c["model"], D, c["p0"], c["p1"], Float64(c["c0"]))
    h = Hamiltonian(metric(c["metric"], D, N, c["Minv"]; matrix_mode=true), lp, dlp)
    θ0, r0 = mat(c["theta0"]), mat(c["r0"])
    ϵ = c["eps_chain"] === nothing ? Float64(c["eps"]) : vecf(c["eps_chain"])
    lf = c["temper_alpha"] === nothing ? Leapfrog(ϵ) : TemperedLeapfrog(ϵ, Float64(c["temper_alpha"]))
    z1 = AdvancedHMC.step(lf, h, phasepoint(h, θ0, r0), c["n_steps"])
    out = copy(c)
    out["expect"] = Dict("theta" => rows(z1.θ), "r" => rows(z1.r), "lp_gradient" => rows(z1.ℓπ.gradient),
                         "lp_value" => collect(z1.ℓπ.value), "lk_value" => collect(z1.ℓκ.value))
    return out
end

# ---------------------------------------------------------------------------------------------------------------- static transition
function run_hmc(c)
    D, N = c["D"], c["N"]
    lp, dlp = target(c["model"], D, c["p0"], c["p1"], Float64(c["c0"]))
    ts = c["sampler"] == "multinomial" ? MultinomialTS : EndPointTS
    e = Dict("theta" => [], "r" => [], "lp_gradient" => [], "lp_value" => Float64[], "lk_value" => Float64[],
             "is_accept" => Bool[], "acceptance_rate" => Float64[], "hamiltonian_energy_error" => Float64[])
    for ch in 1:N   # vector mode, one chain at a time: each chain owns its tape, as the fixture defines it
        h = Hamiltonian(metric(c["metric"], D, 1, c["Minv"]; matrix_mode=false), lp, dlp)
        rng = TapeRNG(normals=vecf(c["normals"][ch]), variates=[Float64(c["variates"][ch])],
                      n_fwd=c["n_fwd"] === nothing ? 0 : c["n_fwd"])
        z = phasepoint(h, vecf(c["theta0"][ch]), zeros(D))
        z = refresh(rng, FullMomentumRefreshment(), h, z)                       # sampler.jl:48-58
        τ = Trajectory{ts}(Leapfrog(Float64(c["eps"])), FixedNSteps(c["n_steps"]))
        t = transition(rng, h, τ, z)
        push!(e["theta"], collect(t.z.θ)); push!(e["r"], collect(t.z.r)); push!(e["lp_gradient"], collect(t.z.ℓπ.gradient))
        push!(e["lp_value"], t.z.ℓπ.value); push!(e["lk_value"], t.z.ℓκ.value)
        push!(e["is_accept"], t.stat.is_accept); push!(e["acceptance_rate"], t.stat.acceptance_rate)
        push!(e["hamiltonian_energy_error"], t.stat.hamiltonian_energy_error)
    end
    out = copy(c)
    out["expect"] = e
    return out
end

# ---------------------------------------------------------------------------------------------------------------- NUTS
criterion(name, max_depth, Δ) = name == "classic" ? ClassicNoUTurn(max_depth, Δ) :
                                name == "strict" ? StrictGeneralisedNoUTurn(max_depth, Δ) : GeneralisedNoUTurn(max_depth, Δ)

function run_nuts(c)
    D, N = c["D"], c["N"]
    lp, dlp = target(c["model"], D, c["p0"], c["p1"], Float64(c["c0"]))
    ts = c["sampler"] == "slice" ? SliceTS : MultinomialTS
    e = Dict("theta" => [], "r" => [], "lp_gradient" => [], "lp_value" => Float64[], "lk_value" => Float64[], "n_steps" => Int[],
             "tree_depth" => Int[], "numerical_error" => Bool[], "acceptance_rate" => Float64[],
             "hamiltonian_energy_error" => Float64[], "max_hamiltonian_energy_error" => Float64[], "variates_used" => Int[])
    for ch in 1:N
        h = Hamiltonian(metric(c["metric"], D, 1, c["Minv"]; matrix_mode=false), lp, dlp)
        rng = TapeRNG(variates=vecf(c["variates"][ch]), dirs=Bool.(c["dirs"][ch] .!= 0))
        z = phasepoint(h, vecf(c["theta0"][ch]), vecf(c["r0"][ch]))
        τ = Trajectory{ts}(Leapfrog(Float64(c["eps"])), criterion(c["criterion"], c["max_depth"], Float64(c["delta_max"])))
        t = transition(rng, h, τ, z)                                            # trajectory.jl:677-742, no refresh: r0 is given
        push!(e["theta"], collect(t.z.θ)); push!(e["r"], collect(t.z.r)); push!(e["lp_gradient"], collect(t.z.ℓπ.gradient))
        push!(e["lp_value"], t.z.ℓπ.value); push!(e["lk_value"], t.z.ℓκ.value)
        push!(e["n_steps"], t.stat.n_steps); push!(e["tree_depth"], t.stat.tree_depth)
        push!(e["numerical_error"], t.stat.numerical_error); push!(e["acceptance_rate"], t.stat.acceptance_rate)
        push!(e["hamiltonian_energy_error"], t.stat.hamiltonian_energy_error)
        push!(e["max_hamiltonian_energy_error"], t.stat.max_hamiltonian_energy_error)
        push!(e["variates_used"], rng.iv)
    end
    out = copy(c)
    out["expect"] = e
    return out
end

# ---------------------------------------------------------------------------------------------------------------- main
function regenerate(src::String, dst::String, runner)
    d = JSON.parsefile(joinpath(HERE, src))
    out = Dict("generator" => "AdvancedHMC.jl $(pkgversion(AdvancedHMC)) (tests/golden/gen_from_reference.jl), Julia $(VERSION)",
               "inputs_from" => src, "cases" => [runner(c) for c in d["cases"]])
    open(joinpath(HERE, dst), "w") do io
        JSON.print(io, out)
    end
    println("wrote ", dst, ": ", length(out["cases"]), " cases")
end

regenerate("leapfrog_mp50.json", "leapfrog_ref.json", run_leapfrog)
regenerate("hmc_mp50.json", "hmc_ref.json", run_hmc)
regenerate("nuts_mp50.json", "nuts_ref.json", run_nuts)
