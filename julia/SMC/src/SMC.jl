# SMC.jl shim over libsmcb200 -- keeps the reference's module name, `smc(...)` signature, `Cloud` type, accessors and
# exports (src/SMC.jl:14-17, src/smc_main.jl:118-161, src/particle.jl:31-532 of FRBNY-DSGE/SMC.jl v0.1.15) and forwards
# the stage loop (src/smc_main.jl:377-497) and the stage-0 evaluators to the CUDA engine through `ccall`.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  Every ccall below is mirrored one-to-one by the
# ctypes binding in smc_jl_b200/_lib.py and the driver in smc_jl_b200/driver.py, which ARE tested against the same
# library (tests/test_gpu_smc.py); the files this module writes with JLD2.jl / HDF5.jl are reproduced byte for byte by
# smc_jl_b200/jld2.py (tests/test_file_formats.py).
#
# Out of scope, as in SURVEY.md section 8: regime-switching parameters, split_cloud / join_cloud /
# add_parameters_to_cloud (file surgery on finished clouds; they do not touch the engine and can be `include`d from the
# reference unchanged -- they only need the `Cloud` struct below).
module SMC

using Libdl, Random, LinearAlgebra, Distributed, Distributions, JLD2, FileIO, HDF5, Printf, Dates

export smc, Cloud, resample, mutation, mvnormal_mixture_draw, initial_draw!, get_cloud, cloud_isempty,
       get_weights, get_vals, get_loglh, get_logprior, get_old_loglh, get_logpost, get_accept,
       get_likeliest_particle_value, get_highest_posterior_particle_value,
       update_draws!, update_weights!, set_weights!, update_loglh!, update_logprior!, update_old_loglh!,
       normalize_weights!, reset_weights!, zero_bad_loglh_weights!, update_cloud!, update_acceptance_rate!,
       weighted_mean, weighted_std, weighted_cov, weighted_quantile,
       DeviceLogLik, GaussRegLogLik, LinearGaussianLogLik, LinearEquationsLogLik, CAPMLogLik, AnSchorfheideLogLik

const LIB = get(ENV, "SMCB200_LIB", "libsmcb200.so")
const VERBOSITY = Dict(:none => 0, :low => 1, :high => 2)          # src/SMC.jl:19

# ---- status codes -> the exceptions the reference raises (SURVEY 8(b)) ---------------------------------
function check(h::Ptr{Cvoid}, st::Int32)
    st == 0 && return
    msg = h == C_NULL ? "" : unsafe_string(ccall((:smcb200_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    st == 1 && throw(AssertionError(msg))                       # check_nan_ess, src/helpers.jl:301
    st == 2 && throw("Invalid resampler in SMC. Options are :systematic, :multinomial, or :polyalgo")
    st == 3 && throw(DomainError(msg))                           # src/smc_main.jl:331, src/particle.jl:239,435
    st == 4 && throw(LinearAlgebra.PosDefException(1))           # MvNormal(...) at src/mutation.jl:81
    st == 7 && throw(ArgumentError(msg))                         # no device kernel; there is no CPU fallback
    error("smcb200 status $st: $msg")
end

# ---- Cloud: identical fields and layout (src/particle.jl:31-63) -----------------------------------------
mutable struct Cloud
    particles::Matrix{Float64}      # n_parts x (n_para+5), column-major == device struct-of-arrays
    tempering_schedule::Vector{Float64}
    ESS::Vector{Float64}
    stage_index::Int
    n_Φ::Int
    resamples::Int
    c::Float64
    accept::Float64
    total_sampling_time::Float64
end
Cloud(n_params::Int, n_parts::Int) =
    Cloud(Matrix{Float64}(undef, n_parts, n_params + 5), zeros(1), zeros(1), 1, 0, 0, 0., 0.25, 0.)
Base.length(c::Cloud) = size(c.particles, 1)

# column map (src/particle.jl:58-63): parameters, then loglh, logprior, old_loglh, accept, weight
ind_para_end(N::Int) = N - 5; ind_loglh(N::Int) = N - 4; ind_logprior(N::Int) = N - 3
ind_old_loglh(N::Int) = N - 2; ind_accept(N::Int) = N - 1; ind_weight(N::Int) = N
_ncol(c::Matrix{Float64}) = size(c, 2)

# accessors on the particle matrix and on the Cloud (src/particle.jl:71-259)
get_weights(c::Matrix{Float64}) = c[:, ind_weight(_ncol(c))]
get_loglh(c::Matrix{Float64}) = c[:, ind_loglh(_ncol(c))]
get_logprior(c::Matrix{Float64}) = c[:, ind_logprior(_ncol(c))]
get_old_loglh(c::Matrix{Float64}) = c[:, ind_old_loglh(_ncol(c))]
get_accept(c::Matrix{Float64}) = c[:, ind_accept(_ncol(c))]
get_logpost(c::Matrix{Float64}) = get_loglh(c) .+ get_logprior(c)
function get_vals(c::Matrix{Float64}; transpose::Bool = true)
    v = c[:, 1:ind_para_end(_ncol(c))]
    return transpose ? Matrix{Float64}(v') : v
end
cloud_isempty(c::Matrix{Float64}) = isempty(c)
get_likeliest_particle_value(c::Matrix{Float64}) = c[argmax(get_loglh(c)), 1:ind_para_end(_ncol(c))]
get_highest_posterior_particle_value(c::Matrix{Float64}) = c[argmax(get_logpost(c)), 1:ind_para_end(_ncol(c))]
for f in (:get_weights, :get_loglh, :get_logprior, :get_old_loglh, :get_accept, :get_logpost, :cloud_isempty,
          :get_likeliest_particle_value, :get_highest_posterior_particle_value)
    @eval $f(c::Cloud) = $f(c.particles)
end
get_vals(c::Cloud; transpose::Bool = true) = get_vals(c.particles; transpose = transpose)

function _setcol!(c::Matrix{Float64}, col::Int, v::Vector{Float64})
    length(v) == size(c, 1) || error("Dimensions of the cloud and of the vector do not match")
    c[:, col] = v
    return nothing
end
function update_draws!(c::Cloud, draws::Matrix{Float64})        # n_para x n_parts, or its transpose (src/particle.jl:226-247)
    I, J = size(draws); n_parts = length(c); n_para = ind_para_end(_ncol(c.particles))
    if (I, J) == (n_parts, n_para)
        c.particles[:, 1:n_para] = draws
    elseif (I, J) == (n_para, n_parts)
        c.particles[:, 1:n_para] = draws'
    else
        error("update_draws!: draws are neither n_para x n_parts nor n_parts x n_para")
    end
    return nothing
end
update_weights!(c::Matrix{Float64}, inc::Vector{Float64}) = (c[:, ind_weight(_ncol(c))] .*= inc; nothing)
set_weights!(c::Cloud, w::Vector{Float64}) = _setcol!(c.particles, ind_weight(_ncol(c.particles)), w)
update_loglh!(c::Matrix{Float64}, v::Vector{Float64}) = _setcol!(c, ind_loglh(_ncol(c)), v)
update_logprior!(c::Matrix{Float64}, v::Vector{Float64}) = _setcol!(c, ind_logprior(_ncol(c)), v)
update_old_loglh!(c::Matrix{Float64}, v::Vector{Float64}) = _setcol!(c, ind_old_loglh(_ncol(c)), v)
function normalize_weights!(c::Matrix{Float64})                  # weights sum to n_parts (src/particle.jl:362-369)
    col = ind_weight(_ncol(c))
    c[:, col] .*= size(c, 1)
    c[:, col] ./= sum(c[:, col])
    return nothing
end
reset_weights!(c::Matrix{Float64}) = (c[:, ind_weight(_ncol(c))] .= 1.0; nothing)
function zero_bad_loglh_weights!(c::Matrix{Float64})             # src/particle.jl:392-399
    c[get_loglh(c) .== -Inf, ind_weight(_ncol(c))] .= 0.0
    return nothing
end
for f in (:update_weights!, :update_loglh!, :update_logprior!, :update_old_loglh!)
    @eval $f(c::Cloud, v::Vector{Float64}) = $f(c.particles, v)
end
for f in (:normalize_weights!, :reset_weights!, :zero_bad_loglh_weights!)
    @eval $f(c::Cloud) = $f(c.particles)
end
function update_cloud!(cloud::Cloud, new_particles::Matrix{Float64})      # src/particle.jl:426-437
    I, J = size(new_particles)
    if (I, J) == size(cloud.particles)
        cloud.particles = new_particles
    elseif (J, I) == size(cloud.particles)
        cloud.particles = Matrix{Float64}(new_particles')
    else
        error("update_cloud!: the new particles do not match the cloud")
    end
    return nothing
end
update_acceptance_rate!(c::Cloud) = (c.accept = sum(get_accept(c)) / length(c); nothing)
# host-side moments of a finished cloud (src/particle.jl:481-532); inside smc() they are computed on the device
weighted_mean(c::Matrix{Float64}) = vec(get_vals(c; transpose = false)' * get_weights(c)) ./ sum(get_weights(c))
function weighted_cov(c::Matrix{Float64})
    X = get_vals(c; transpose = false); w = get_weights(c) ./ sum(get_weights(c)); m = X' * w
    Xc = X .- m'
    return Xc' * (Xc .* w)
end
weighted_std(c::Matrix{Float64}) = sqrt.(diag(weighted_cov(c)))
function weighted_quantile(c::Matrix{Float64}, i::Int64)
    x = c[:, i]; w = get_weights(c); p = sortperm(x); cw = cumsum(w[p]) ./ sum(w)
    q(α) = x[p][min(searchsortedfirst(cw, α), length(x))]
    return q(0.05), q(0.95)
end
for f in (:weighted_mean, :weighted_cov, :weighted_std)
    @eval $f(c::Cloud) = $f(c.particles)
end
weighted_quantile(c::Cloud, i::Int64) = weighted_quantile(c.particles, i)
get_cloud(path::String) = load(path, "cloud")                    # src/util.jl:113-115

# ---- likelihood descriptors: valid `loglikelihood` arguments that name a device functor ------------------
# The reference takes `loglikelihood::Function` and calls it per particle on the CPU.  A Julia closure cannot run inside a
# CUDA kernel (and this build forbids device code generation), so the likelihood is named by a descriptor: a callable
# (it evaluates the same formula on the host, so it still works with the reference's own host-side helpers) that carries
# the data of one of the engine's functor families.  Anything else raises ArgumentError: there is no CPU fallback.
abstract type DeviceLogLik <: Function end
struct GaussRegLogLik <: DeviceLogLik              # SMCB200_LIK_GAUSSREG
    iparams::Vector{Int32}          # n_eq, k, stride, coef_off, sig_off (-1: sigma known)
    dparams::Vector{Float64}        # per equation: T, qscale, rss, sigma_fixed, bhat[k], U[k*k] (row-major upper)
end
lik_kind(::GaussRegLogLik) = Int32(1)
function (l::GaussRegLogLik)(p::AbstractVector, data = nothing)   # host evaluation of the same formula
    neq, k, stride, coef, sig = l.iparams; per = 4 + k + k * k; ll = 0.0
    for e in 0:neq-1
        q = l.dparams[per*e+1:per*(e+1)]; T, qs, rss, sfix = q[1:4]; bhat = q[5:4+k]; U = reshape(q[5+k:end], k, k)'
        s = sig >= 0 ? p[sig+e*stride+1] : sfix
        s > 0 || return -Inf
        r = U * (p[coef+e*stride+1:coef+e*stride+k] .- bhat)
        ll += -T * 0.9189385332046727 - T * log(s) - 0.5 * qs * (rss + dot(r, r)) / s^2
    end
    return ll
end
function _suffstats(y::AbstractVector, Z::AbstractMatrix)
    F = qr(Z); k = size(Z, 2); R = Matrix(F.R); s = sign.(diag(R)); s[s .== 0] .= 1.0; R = s .* R
    bhat = R \ ((Matrix(F.Q)' * y)[1:k] .* s)
    return bhat, R, sum(abs2, y - Z * bhat)
end
_pack_eq(T, qs, rss, sfix, bhat, U) = vcat(Float64[T, qs, rss, sfix], bhat, vec(permutedims(U)))
"y = X β + ε, ε ~ N(0, σ2), σ2 known (examples/regression_model/estimate_regression.jl:46-53 with X = [1 x])."
function LinearGaussianLogLik(y::Vector{Float64}, X::Matrix{Float64}; σ2::Float64 = 1.0)
    T, k = size(X); bhat, U, rss = _suffstats(y, X)
    GaussRegLogLik(Int32[1, k, k, 0, -1], _pack_eq(T, 1.0, rss, sqrt(σ2), bhat, U))
end
"test/modelsetup.jl:119-138: y_it = α_i + β_i x_it + e_it, e ~ N(0, σ_i²); parameters ordered (α_i, β_i, σ_i)."
function LinearEquationsLogLik(data::Matrix{Float64}, X::Matrix{Float64})
    neq, T = size(data); eqs = Float64[]
    for i in 1:neq
        bhat, U, rss = _suffstats(data[i, :], hcat(ones(T), X[i, 1:T]))
        append!(eqs, _pack_eq(T, 1.0, rss, 0.0, bhat, U))
    end
    GaussRegLogLik(Int32[neq, 2, 3, 0, 2], eqs)
end
"examples/capm_model/estimate_capm.jl:48-69 (per-period form; `as_written = true` reproduces the example's code literally)."
function CAPMLogLik(lik_data::Matrix{Float64}, market_data::AbstractArray{Float64}; as_written::Bool = false)
    neq, T = size(lik_data); m = vec(market_data)[1:T]
    as_written || return LinearEquationsLogLik(lik_data, repeat(m', neq, 1))
    eqs = Float64[]
    for i in 1:neq
        bhat, U, rss = _suffstats(lik_data[i, :], reshape(1.0 .+ m, T, 1))
        append!(eqs, _pack_eq(T, Float64(T), rss, 0.0, bhat, U))
    end
    GaussRegLogLik(Int32[neq, 1, 3, 0, 2], eqs)
end
# Three-equation An-Schorfheide DSGE model (examples/dsge_models/small_dsge_model.jl:35-50): replaces the closure
#   loglik(p, d) = DSGE.likelihood(m, d; sampler = false, catch_errors = true, use_chand_recursion = true)
# `parameters` must be DSGE.jl's AnSchorfheide ParameterVector (16 entries); data is 3 x T; NaN entries are missing observations (dropped from that period's update).
struct AnSchorfheideLogLik <: DeviceLogLik         # SMCB200_LIK_AS_DSGE
    iparams::Vector{Int32}          # n_periods, n_presample
    dparams::Vector{Float64}        # vec(data): 3 x T column-major
end
AnSchorfheideLogLik(data::Matrix{Float64}; n_presample::Int = 2) =
    AnSchorfheideLogLik(Int32[size(data, 2), n_presample], vec(data))
lik_kind(::AnSchorfheideLogLik) = Int32(2)
(l::AnSchorfheideLogLik)(p, data = nothing) = throw(ArgumentError("AnSchorfheideLogLik is evaluated on the device only"))

# ---- stage structs (include/smcb200.h) ------------------------------------------------------------------
struct StageConfig
    phi_n1::Float64; phi_n::Float64; threshold_ratio::Float64; target::Float64; alpha::Float64
    tempering_target::Float64; prior_weight::Float64; log_prob_old_data::Float64
    n_mh_steps::Int32; n_blocks::Int32; resample_method::Int32; adaptive::Int32; has_old_data::Int32; reserved::Int32
    seed::UInt64; stage::UInt32; reserved2::UInt32
end
mutable struct StageState
    c::Float64; accept::Float64; ess_prev::Float64; phi_prop::Float64; j::Int64
    resampled_last_period::Int32; reserved::Int32
end
struct StageResult
    phi_n::Float64; ess::Float64; sum_weights::Float64; c::Float64; accept::Float64
    resampled::Int32; status::Int32
    ms_correct::Float32; ms_resample::Float32; ms_moments::Float32; ms_mutate::Float32
end

const PRIOR_KIND = Dict(:Normal => 0, :Uniform => 1, :Gamma => 2, :RootInverseGamma => 3, :Beta => 4, :InverseGamma => 5)
resampler_code(m::Symbol) = m == :systematic ? Int32(0) : m in (:multinomial, :polyalgo) ? Int32(1) :   # :polyalgo -> multinomial kernel
    throw("Invalid resampler in SMC. Options are :systematic, :multinomial, or :polyalgo")

# ---- the engine: one context per process / GPU -------------------------------------------------------------
mutable struct Engine
    h::Ptr{Cvoid}
    n_parts::Int; n_para::Int; first::Int; count::Int; rank::Int; world::Int
end
function Engine(device::Int = 0)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:smcb200_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), ctx, device)
    st == 0 || error("smcb200_create(device = $device) failed ($st): a CUDA GPU is required, there is no CPU fallback")
    Engine(ctx[], 0, 0, 0, 0, 0, 1)
end
close!(e::Engine) = (e.h == C_NULL || ccall((:smcb200_destroy, LIB), Int32, (Ptr{Cvoid},), e.h); e.h = C_NULL; nothing)
function unique_id()
    id = zeros(UInt8, 128)
    check(C_NULL, ccall((:smcb200_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id))
    id
end
function comm_init!(e::Engine, rank::Int, world::Int, id::Vector{UInt8})   # replaces the Distributed.jl fan-out (:169-170,471-476)
    check(e.h, ccall((:smcb200_comm_init, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), e.h, rank, world, id))
    e.rank, e.world = rank, world
end
function cloud_create!(e::Engine, n_parts::Int, n_para::Int)
    check(e.h, ccall((:smcb200_cloud_create, LIB), Int32, (Ptr{Cvoid}, Int64, Int32), e.h, n_parts, n_para))
    f = Ref{Int64}(0); c = Ref{Int64}(0)
    check(e.h, ccall((:smcb200_cloud_shard, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), e.h, f, c))
    e.n_parts, e.n_para, e.first, e.count = n_parts, n_para, f[], c[]
end
# `particles` is the GLOBAL n_parts x (n_para + 5) matrix; every rank copies its own rows
upload!(e::Engine, P::Matrix{Float64}) = (size(P) == (e.n_parts, e.n_para + 5) || error("update_cloud!: wrong size");
    GC.@preserve P check(e.h, ccall((:smcb200_cloud_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), e.h, P, size(P, 1), e.first)))
download!(e::Engine, P::Matrix{Float64}) = (size(P) == (e.n_parts, e.n_para + 5) || error("wrong size");
    GC.@preserve P check(e.h, ccall((:smcb200_cloud_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), e.h, P, size(P, 1), e.first)))
function read_column(e::Engine, col::Int)                        # 1-based column of this rank's shard
    v = Vector{Float64}(undef, e.count)
    GC.@preserve v check(e.h, ccall((:smcb200_cloud_read_column, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), e.h, col - 1, v))
    v
end
write_column!(e::Engine, col::Int, v::Vector{Float64}) =
    GC.@preserve v check(e.h, ccall((:smcb200_cloud_write_column, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), e.h, col - 1, v))

function set_model!(e::Engine, parameters, loglikelihood, old_loglikelihood = nothing)
    loglikelihood isa DeviceLogLik ||
        throw(ArgumentError("loglikelihood must be a device likelihood descriptor (e.g. LinearGaussianLogLik); there is no CPU fallback"))
    n_para = length(parameters)
    fixed = Int32[p.fixed for p in parameters]
    lo = Float64[p.valuebounds[1] for p in parameters]; hi = Float64[p.valuebounds[2] for p in parameters]
    prior(p) = p.prior.value                                       # ModelConstructors wraps the prior in a Nullable-like
    kind = Int32[p.fixed ? 0 : PRIOR_KIND[nameof(typeof(prior(p)))] for p in parameters]
    pp(p, i) = (q = prior(p); q isa Distributions.Distribution ? Float64(Distributions.params(q)[i]) : Float64(getfield(q, i)))
    p1 = Float64[p.fixed ? 0. : pp(p, 1) for p in parameters]
    p2 = Float64[p.fixed ? 1. : pp(p, 2) for p in parameters]
    GC.@preserve fixed lo hi kind p1 p2 check(e.h, ccall((:smcb200_set_parameters, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
        e.h, n_para, fixed, lo, hi, kind, p1, p2))
    for (slot, lk) in ((0, loglikelihood), (1, old_loglikelihood))
        lk === nothing && continue
        lk isa DeviceLogLik || throw(ArgumentError("old_loglikelihood must be a device likelihood descriptor"))
        ip, dp = lk.iparams, lk.dparams
        GC.@preserve ip dp check(e.h, ccall((:smcb200_set_likelihood, LIB), Int32,
            (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int64), e.h, slot, lik_kind(lk), ip, length(ip), dp, length(dp)))
    end
end
initial_draw_device!(e::Engine, values::Vector{Float64}, seed::UInt64; max_tries::Int = 1000) =
    GC.@preserve values check(e.h, ccall((:smcb200_initial_draw, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, UInt64, Int32), e.h, values, seed, max_tries))
evaluate!(e::Engine, mode::Int) = check(e.h, ccall((:smcb200_evaluate, LIB), Int32, (Ptr{Cvoid}, Int32), e.h, mode))
resample_cloud!(e::Engine, method::Symbol, seed::UInt64, stage::Int) =
    check(e.h, ccall((:smcb200_resample, LIB), Int32, (Ptr{Cvoid}, Int32, UInt64, UInt32, Float64, Ptr{Int64}),
                     e.h, resampler_code(method), seed, stage, -1.0, C_NULL))
function resample_weights(e::Engine, weights::Vector{Float64}, n_out::Int, method::Symbol, seed::UInt64, stage::Int = 0)
    idx = Vector{Int64}(undef, n_out)
    GC.@preserve weights idx check(e.h, ccall((:smcb200_resample_weights_n, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int32, UInt64, UInt32, Float64, Ptr{Int64}, Ptr{Float64}),
        e.h, weights, length(weights), n_out, resampler_code(method), seed, stage, -1.0, idx, C_NULL))
    idx
end
"Up to `n_stages` consecutive stages in one call (smcb200_run_stages); history columns (shard rows) are optional."
function run_stages!(e::Engine, cfg::StageConfig, state::StageState, schedule::Vector{Float64}, i_first::Int, n_stages::Int,
                     inc::Union{Nothing,Matrix{Float64}}, normw::Union{Nothing,Matrix{Float64}})
    res = Vector{StageResult}(undef, n_stages); n_done = Ref{Int32}(0)
    pinc = inc === nothing ? Ptr{Float64}(C_NULL) : pointer(inc); pnw = normw === nothing ? Ptr{Float64}(C_NULL) : pointer(normw)
    st = GC.@preserve schedule inc normw res ccall((:smcb200_run_stages, LIB), Int32,
        (Ptr{Cvoid}, Ref{StageConfig}, Ref{StageState}, Ptr{Float64}, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int64,
         Ptr{StageResult}, Ref{Int32}), e.h, cfg, state, schedule, length(schedule), i_first, n_stages, pinc, pnw, e.count, res, n_done)
    check(e.h, st)
    res[1:n_done[]]
end

# ---- exported helpers with the reference's signatures --------------------------------------------------------
"""
    resample(weights; n_parts = length(weights), method = :systematic)

src/resample.jl:23 on the device (ancestor indices, 1-based).
"""
function resample(weights::Vector{Float64}; n_parts::Int = length(weights), method::Symbol = :systematic,
                  seed::UInt64 = rand(UInt64), device::Int = 0)
    resampler_code(method)
    e = Engine(device)
    try
        return resample_weights(e, weights, n_parts, method, seed)
    finally
        close!(e)
    end
end

"""
    mvnormal_mixture_draw(θ_old, d_prop; c = 1.0, α = 1.0)

One draw from the mixture proposal of src/helpers.jl:87-100 on the host (a convenience kept for scripts that call it
directly; inside `smc` the draws happen in the mutation kernel): with probability α `θ_old + c L z`, with (1-α)/2 each
`θ_old + c sqrt.(diag(Σ)) .* z` and `θ̄ + c L z`, where Σ = L L' is `d_prop.Σ` and θ̄ = `d_prop.μ`.
"""
function mvnormal_mixture_draw(θ_old::Vector{T}, d_prop::Distribution; c::T = 1.0, α::T = 1.0) where T<:AbstractFloat
    @assert 0 <= α <= 1
    Σ = Matrix(d_prop.Σ); L = cholesky(Symmetric(Σ)).L; z = randn(length(θ_old)); u = rand()
    u < α && return θ_old .+ c .* (L * z)
    u < α + (1 - α) / 2 && return θ_old .+ c .* sqrt.(diag(Σ)) .* z
    return Vector(d_prop.μ) .+ c .* (L * z)
end

"""
    initial_draw!(loglikelihood, parameters, data, c::Cloud; parallel = false)

src/initialization.jl:88-119 on the device: prior draws inside the value bounds, redrawn until the log-likelihood is finite.
"""
function initial_draw!(loglikelihood::Function, parameters, data::Matrix{Float64}, c::Cloud; parallel::Bool = false,
                       regime_switching::Bool = false, toggle::Bool = true, seed::UInt64 = rand(UInt64), device::Int = 0)
    regime_switching && throw(ArgumentError("regime switching is not available in the device engine"))
    e = Engine(device)
    try
        cloud_create!(e, length(c), length(parameters)); set_model!(e, parameters, loglikelihood)
        initial_draw_device!(e, Float64[p.value for p in parameters], seed)
        download!(e, c.particles)
    finally
        close!(e)
    end
    return nothing
end

"""
    mutation(loglikelihood, parameters, data, p, d_μ, d_Σ, n_free_para, blocks_free, blocks_all, ϕ_n, ϕ_n1; c, α, n_mh_steps,
             old_data, old_loglikelihood)

src/mutation.jl:56-138 for ONE particle `p` (a row of the cloud): a one-particle cloud goes through the mutation kernel.
Returns the updated row (parameters, loglh, logprior, old_loglh, accept, weight).
"""
function mutation(loglikelihood::Function, parameters, data::Matrix{S}, p::Vector{S}, d_μ::Vector{S}, d_Σ::Matrix{S},
                  n_free_para::Int, blocks_free::Vector{Vector{Int}}, blocks_all::Vector{Vector{Int}}, ϕ_n::S, ϕ_n1::S;
                  c::S = 1., α::S = 1., n_mh_steps::Int = 1, old_data::AbstractMatrix = Matrix{S}(undef, size(data, 1), 0),
                  old_loglikelihood::Function = loglikelihood, regime_switching::Bool = false, toggle::Bool = true,
                  seed::UInt64 = rand(UInt64), device::Int = 0) where {S<:AbstractFloat}
    regime_switching && throw(ArgumentError("regime switching is not available in the device engine"))
    e = Engine(device)
    try
        n_para = length(parameters); has_old = !isempty(old_data)
        cloud_create!(e, 1, n_para); set_model!(e, parameters, loglikelihood, has_old ? old_loglikelihood : nothing)
        P = reshape(copy(p), 1, :); upload!(e, P)
        sizes = Int32[length(b) for b in blocks_all]
        bf = Int32.(vcat(blocks_free...) .- 1); ba = Int32.(vcat(blocks_all...) .- 1)
        Σ = Matrix{Float64}(d_Σ'); acc = Ref{Float64}(0.0)            # row-major for the C side (symmetric anyway)
        GC.@preserve d_μ Σ sizes bf ba check(e.h, ccall((:smcb200_mutate, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Float64, Float64, Float64, Float64,
             Int32, Int32, UInt64, UInt32, Ref{Float64}),
            e.h, d_μ, Σ, n_free_para, length(sizes), sizes, bf, ba, ϕ_n, ϕ_n1, c, α, n_mh_steps, has_old ? 1 : 0, seed, 0, acc))
        download!(e, P)
        return vec(P)
    finally
        close!(e)
    end
end

# ---- smc ---------------------------------------------------------------------------------------------------------
"""
    smc(loglikelihood, parameters, data; kwargs...)

Same keyword arguments and defaults as the reference (src/smc_main.jl:118-161).  `loglikelihood` must be a device descriptor
(e.g. `LinearGaussianLogLik`); an arbitrary closure raises `ArgumentError`.  `parallel = true` with Distributed workers runs
one rank per worker process, each on its own GPU (worker k uses device k - 1), the cloud sharded across them.
"""
function smc(loglikelihood::Function, parameters, data::Matrix{Float64};
             verbose::Symbol = :low, testing::Bool = false, data_vintage::String = Dates.format(today(), "yymmdd"),
             parallel::Bool = false, n_parts::Int = 5_000, n_blocks::Int = 1, n_mh_steps::Int = 1, λ::Float64 = 2.1, n_Φ::Int = 300,
             resampling_method::Symbol = :systematic, threshold_ratio::Float64 = 0.5,
             c::Float64 = 0.5, α::Float64 = 1.0, target::Float64 = 0.25,
             use_fixed_schedule::Bool = true, tempering_target::Float64 = 0.97,
             old_data::Matrix{Float64} = Matrix{Float64}(undef, size(data, 1), 0), old_cloud::Cloud = Cloud(0, 0),
             old_loglikelihood::Function = loglikelihood, old_vintage::String = "", smc_iteration::Int = 1,
             run_test::Bool = false, filestring_addl::Vector{String} = Vector{String}(), loadpath::String = "",
             savepath::String = "smc_cloud.jld2", particle_store_path::String = "smcsave.h5",
             save_intermediate::Bool = false, intermediate_stage_increment::Int = 10, continue_intermediate::Bool = false,
             intermediate_stage_start::Int = 0, tempered_update_prior_weight::Float64 = 0.0,
             regime_switching::Bool = false, toggle::Bool = true, debug_assertion::Bool = false, log_prob_old_data::Float64 = 0.0,
             seed::UInt64 = UInt64(1793), device::Int = 0)
    regime_switching && throw(ArgumentError("regime switching is not available in the device engine"))
    resampler_code(resampling_method)
    0.0 <= tempered_update_prior_weight <= 1.0 ||                        # src/smc_main.jl:331
        throw(DomainError(tempered_update_prior_weight, "The keyword tempered_update_prior_weight must be within the interval [0, 1]"))
    @assert any(p -> !p.fixed, parameters) "All model parameters are fixed!"     # src/smc_main.jl:237
    kw = (verbose = verbose, testing = testing, n_parts = n_parts, n_blocks = n_blocks, n_mh_steps = n_mh_steps, λ = λ, n_Φ = n_Φ,
          resampling_method = resampling_method, threshold_ratio = threshold_ratio, c = c, α = α, target = target,
          use_fixed_schedule = use_fixed_schedule, tempering_target = tempering_target, old_data = old_data, old_cloud = old_cloud,
          old_loglikelihood = old_loglikelihood, run_test = run_test, loadpath = loadpath, savepath = savepath,
          particle_store_path = particle_store_path, save_intermediate = save_intermediate,
          intermediate_stage_increment = intermediate_stage_increment, continue_intermediate = continue_intermediate,
          tempered_update_prior_weight = tempered_update_prior_weight, log_prob_old_data = log_prob_old_data, seed = seed)
    if parallel && nworkers() > 1 && ispow2(nworkers())
        id = unique_id(); ws = workers(); G = length(ws)
        futs = [remotecall(SMC._smc_rank, w, loglikelihood, parameters, r - 1, G, id, r - 1, kw) for (r, w) in enumerate(ws)]
        outs = fetch.(futs)                                              # (rows, w rows, W rows, cloud meta) per rank, rank order
        cloud = outs[1][4]
        cloud.particles = vcat((o[1] for o in outs)...)
        _write_outputs(cloud, vcat((o[2] for o in outs)...), vcat((o[3] for o in outs)...), length(parameters), kw)
    else
        rows, w, W, cloud = _smc_rank(loglikelihood, parameters, 0, 1, UInt8[], device, kw)
        cloud.particles = rows
        _write_outputs(cloud, w, W, length(parameters), kw)
    end
    nothing
end

function _write_outputs(cloud::Cloud, w_matrix, W_matrix, n_para::Int, kw)
    kw.testing && return                                                     # src/smc_main.jl:513-526
    h5open(kw.particle_store_path, "w") do file
        write(file, "smcparams", cloud.particles[:, 1:n_para])
    end
    jldopen(kw.savepath, true, true, true, IOStream) do file
        write(file, "cloud", cloud); write(file, "w", w_matrix); write(file, "W", W_matrix)
    end
end

# bridge initialisation of a tempered update (src/smc_main.jl:260-329); every rank builds the same global matrix on its own
# GPU (old-cloud resampling and prior draws are keyed on global indices) and uploads its rows
function _bridge!(e::Engine, parameters, lik, old_lik, old_cloud::Cloud, n_parts::Int, pw::Float64, method::Symbol, seed::UInt64, device::Int)
    n_para = length(parameters); n_res = round(Int, (1 - pw) * n_parts); n_prior = n_parts - n_res
    side = e.world == 1 ? e : Engine(device)
    parts = Matrix{Float64}[]
    try
        if n_res > 0
            inds = resample_weights(side, get_weights(old_cloud), n_res, method, seed)
            push!(parts, old_cloud.particles[inds, :])
        end
        if n_prior > 0
            cloud_create!(side, n_prior, n_para); set_model!(side, parameters, old_lik)      # old likelihood on old data, current prior
            initial_draw_device!(side, Float64[p.value for p in parameters], xor(seed, UInt64(0x9E3779B9)))
            Pp = Matrix{Float64}(undef, n_prior, n_para + 5); download!(side, Pp); push!(parts, Pp)
        end
    finally
        side === e || close!(side)
    end
    P = vcat(parts...)
    cloud_create!(e, n_parts, n_para); set_model!(e, parameters, lik, old_lik)
    upload!(e, P); evaluate!(e, 1)                                           # initialize_likelihoods!, :307
    ll = read_column(e, n_para + 1); w = read_column(e, n_para + 5)
    w[ll .== -Inf] .= 0.0                                                    # zero_bad_loglh_weights!, :314
    write_column!(e, n_para + 5, w)
    resample_cloud!(e, method, seed, 1)                                      # normalise + resample + reset, :315-322
end

function _smc_rank(loglikelihood, parameters, rank::Int, world::Int, id::Vector{UInt8}, device::Int, kw)
    n_para = length(parameters); n_parts = kw.n_parts; n_Φ = kw.n_Φ
    tempered_update = !isempty(kw.old_data)
    e = Engine(device)
    try
        world > 1 && comm_init!(e, rank, world, id)
        cloud_create!(e, n_parts, n_para)
        set_model!(e, parameters, loglikelihood, tempered_update ? kw.old_loglikelihood : nothing)
        lo, hi = e.first + 1, e.first + e.count                              # this rank's rows (1-based)
        schedule = ((collect(1:n_Φ) .- 1) / (n_Φ - 1)) .^ kw.λ               # src/smc_main.jl:348-352
        w_matrix = zeros(e.count, 1)
        i = 1; ϕ_n = 0.0; j = 2
        if tempered_update                                                   # src/smc_main.jl:244-333
            old_cloud = cloud_isempty(kw.old_cloud) ? load(kw.loadpath, "cloud") : kw.old_cloud      # :246
            if kw.tempered_update_prior_weight == 0.0 && length(old_cloud) == n_parts
                cloud = Cloud(copy(old_cloud.particles), zeros(1), [old_cloud.ESS[end]], 1, n_Φ, 0, kw.c, kw.target, 0.)
                upload!(e, cloud.particles); evaluate!(e, 1)                 # initialize_likelihoods!
                w0 = get_weights(cloud)
                W_matrix = reshape((sum(w0) <= 1.0 ? w0 .* n_parts : w0)[lo:hi], :, 1)              # :363-366
            else
                _bridge!(e, parameters, loglikelihood, kw.old_loglikelihood, old_cloud, n_parts, kw.tempered_update_prior_weight,
                         kw.resampling_method, kw.seed, device)
                cloud = Cloud(n_para, n_parts); cloud.ESS = [Float64(n_parts)]                        # :325
                W_matrix = ones(e.count, 1)
            end
        elseif kw.continue_intermediate                                      # :334-335,355-361
            cloud = load(kw.loadpath, "cloud")
            size(cloud.particles) == (n_parts, n_para + 5) || error("checkpoint does not match n_parts / the ParameterVector")
            upload!(e, cloud.particles)
            w_matrix = load(kw.loadpath, "w")[lo:hi, :]; W_matrix = load(kw.loadpath, "W")[lo:hi, :]; j = load(kw.loadpath, "j")
            i = cloud.stage_index; ϕ_n = (kw.use_fixed_schedule ? schedule : cloud.tempering_schedule)[i]
        else
            cloud = Cloud(n_para, n_parts)
            initial_draw_device!(e, Float64[p.value for p in parameters], kw.seed)                  # initial_draw!, :341
            cloud.ESS = [Float64(n_parts)]
            W_matrix = ones(e.count, 1)
        end
        if !kw.continue_intermediate                                         # initialize_cloud_settings!, initialization.jl:196-211
            cloud.stage_index = 1; cloud.n_Φ = n_Φ; cloud.resamples = 0; cloud.c = kw.c; cloud.accept = kw.target
            cloud.total_sampling_time = 0.; cloud.tempering_schedule = kw.use_fixed_schedule ? copy(schedule) : zeros(1)
        end
        state = StageState(cloud.c, cloud.accept, cloud.ESS[end], kw.continue_intermediate ? schedule[j] : 0., j, 0, 0)
        while ϕ_n < 1.                                                       # src/smc_main.jl:377
            t0 = time_ns()
            # a batch = the stages the host does not need to look at: all of them on a quiet fixed-schedule run
            n_batch = (kw.use_fixed_schedule && kw.verbose == :none) ? n_Φ - i : 1
            kw.save_intermediate && (n_batch = min(n_batch, kw.intermediate_stage_increment - mod(i, kw.intermediate_stage_increment)))
            kw.run_test && (n_batch = min(n_batch, max(1, 3 - i)))
            ϕ_n1 = kw.use_fixed_schedule ? schedule[i] : cloud.tempering_schedule[i]
            cfg = StageConfig(ϕ_n1, 0., kw.threshold_ratio, kw.target, kw.α, kw.tempering_target, kw.tempered_update_prior_weight,
                              kw.log_prob_old_data, kw.n_mh_steps, kw.n_blocks, resampler_code(kw.resampling_method),
                              kw.use_fixed_schedule ? 0 : 1, tempered_update ? 1 : 0, 0, kw.seed, UInt32(i + 1), 0)
            inc = Matrix{Float64}(undef, e.count, n_batch); normw = Matrix{Float64}(undef, e.count, n_batch)
            results = run_stages!(e, cfg, state, schedule, i + 1, n_batch, inc, normw)
            nd = length(results)
            for res in results
                cloud.stage_index = i += 1
                ϕ_n = res.phi_n
                kw.use_fixed_schedule || push!(cloud.tempering_schedule, ϕ_n)
                push!(cloud.ESS, res.ess); cloud.resamples += res.resampled; cloud.c = res.c; cloud.accept = res.accept
                rank == 0 && VERBOSITY[kw.verbose] >= VERBOSITY[:low] &&
                    @printf(" stage %4d  phi %.6g  c %.4f  accept %.4f  ESS %.1f  (%d resamples)\n", i, ϕ_n, res.c, res.accept, res.ess, cloud.resamples)
            end
            w_matrix = hcat(w_matrix, inc[:, 1:nd]); W_matrix = hcat(W_matrix, normw[:, 1:nd])   # :419-420 (W column = 1 on resample, :445)
            cloud.total_sampling_time += (time_ns() - t0) * 1e-9
            kw.run_test && i >= 3 && break                                   # :495
            if kw.save_intermediate && mod(i, kw.intermediate_stage_increment) == 0 && world == 1     # :499-507
                P = Matrix{Float64}(undef, n_parts, n_para + 5); download!(e, P); cloud.particles = P
                jldopen(replace(kw.savepath, ".jld2" => "_stage=$(i).jld2"), true, true, true, IOStream) do file
                    write(file, "cloud", cloud); write(file, "w", w_matrix); write(file, "W", W_matrix); write(file, "j", state.j)
                end
            end
        end
        P = Matrix{Float64}(undef, n_parts, n_para + 5)
        download!(e, P)                                                      # this rank's rows
        meta = Cloud(Matrix{Float64}(undef, 0, n_para + 5), cloud.tempering_schedule, cloud.ESS, cloud.stage_index, cloud.n_Φ,
                     cloud.resamples, cloud.c, cloud.accept, cloud.total_sampling_time)
        return P[lo:hi, :], w_matrix, W_matrix, meta
    finally
        close!(e)
    end
end

end # module
