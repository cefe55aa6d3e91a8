# SMC.jl shim over libsmcb200 -- keeps the reference's module name, `smc(...)` signature, `Cloud` type and
# exports (src/SMC.jl:14-17, src/smc_main.jl:118-161, src/particle.jl:31-41 of FRBNY-DSGE/SMC.jl v0.1.15) and
# forwards the stage loop (src/smc_main.jl:377-497) to the CUDA engine through `ccall`.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  Every ccall below is mirrored
# one-to-one by the ctypes binding in smc_jl_b200/_lib.py, which IS tested against the same library.
module SMC

using Libdl, Random, LinearAlgebra, Distributions, JLD2, FileIO, HDF5

export smc, Cloud, resample, mutation, LinearGaussianLogLik, GaussRegLogLik, AnSchorfheideLogLik, get_cloud

const LIB = get(ENV, "SMCB200_LIB", "libsmcb200.so")

# ---- status codes -> the exceptions the reference raises (SURVEY 8(b)) ---------------------------------
function check(ctx::Ptr{Cvoid}, st::Int32)
    st == 0 && return
    msg = unsafe_string(ccall((:smcb200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    st == 1 && throw(AssertionError(msg))                       # check_nan_ess, src/helpers.jl:301
    st == 2 && throw("Invalid resampler in SMC. Options are :systematic, :multinomial, or :polyalgo")
    st == 3 && throw(DomainError(msg))
    st == 4 && throw(LinearAlgebra.PosDefException(1))
    st == 7 && throw(ArgumentError(msg))                         # no device kernel; there is no CPU fallback
    error("smcb200 status $st: $msg")
end

# ---- Cloud: identical fields and layout (src/particle.jl:31-53) -----------------------------------------
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

# ---- likelihood descriptors: valid `loglikelihood` arguments that name a device functor ------------------
abstract type DeviceLogLik <: Function end
struct GaussRegLogLik <: DeviceLogLik              # SMCB200_LIK_GAUSSREG
    iparams::Vector{Int32}          # n_eq, k, stride, coef_off, sig_off
    dparams::Vector{Float64}        # per equation: T, qscale, rss, sigma_fixed, bhat[k], U[k*k] (row-major upper)
end
lik_kind(::GaussRegLogLik) = Int32(1)
# Three-equation An-Schorfheide DSGE model (examples/dsge_models/small_dsge_model.jl:35-50): replaces the closure
#   loglik(p, d) = DSGE.likelihood(m, d; sampler = false, catch_errors = true, use_chand_recursion = true)
# `parameters` must be DSGE.jl's AnSchorfheide ParameterVector (16 entries); data is 3 x T.
struct AnSchorfheideLogLik <: DeviceLogLik         # SMCB200_LIK_AS_DSGE
    iparams::Vector{Int32}          # n_periods, n_presample
    dparams::Vector{Float64}        # vec(data): 3 x T column-major
end
AnSchorfheideLogLik(data::Matrix{Float64}; n_presample::Int = 2) =
    AnSchorfheideLogLik(Int32[size(data, 2), n_presample], vec(data))
lik_kind(::AnSchorfheideLogLik) = Int32(2)
function LinearGaussianLogLik(y::Vector{Float64}, X::Matrix{Float64}; σ2::Float64 = 1.0)
    T, k = size(X)
    F = qr(X); R = Matrix(F.R); s = sign.(diag(R)); R = s .* R
    bhat = R \ ((Matrix(F.Q)' * y)[1:k] .* s)
    rss = sum(abs2, y - X * bhat)
    GaussRegLogLik(Int32[1, k, k, 0, -1], vcat(Float64[T, 1.0, rss, sqrt(σ2)], bhat, vec(permutedims(R))))
end

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
mutable struct StageResult
    phi_n::Float64; ess::Float64; sum_weights::Float64; c::Float64; accept::Float64
    resampled::Int32; status::Int32
    ms_correct::Float32; ms_resample::Float32; ms_moments::Float32; ms_mutate::Float32
    StageResult() = new()
end

const PRIOR_KIND = Dict(:Normal => 0, :Uniform => 1, :Gamma => 2, :RootInverseGamma => 3, :Beta => 4, :InverseGamma => 5)

"""
    smc(loglikelihood, parameters, data; kwargs...)

Same keyword arguments and defaults as the reference (src/smc_main.jl:118-161).  `loglikelihood` must be a
device descriptor (e.g. `LinearGaussianLogLik`); an arbitrary closure raises `ArgumentError`.
"""
function smc(loglikelihood::Function, parameters, data::Matrix{Float64};
             verbose::Symbol = :low, testing::Bool = false, parallel::Bool = false,
             n_parts::Int = 5_000, n_blocks::Int = 1, n_mh_steps::Int = 1, λ::Float64 = 2.1, n_Φ::Int = 300,
             resampling_method::Symbol = :systematic, threshold_ratio::Float64 = 0.5,
             c::Float64 = 0.5, α::Float64 = 1.0, target::Float64 = 0.25,
             use_fixed_schedule::Bool = true, tempering_target::Float64 = 0.97,
             old_data::Matrix{Float64} = Matrix{Float64}(undef, size(data, 1), 0), old_cloud::Cloud = Cloud(0, 0),
             old_loglikelihood::Function = loglikelihood, tempered_update_prior_weight::Float64 = 0.0,
             log_prob_old_data::Float64 = 0.0, savepath::String = "smc_cloud.jld2",
             particle_store_path::String = "smcsave.h5", loadpath::String = "", save_intermediate::Bool = false,
             intermediate_stage_increment::Int = 10, continue_intermediate::Bool = false,
             seed::UInt64 = UInt64(1793), device::Int = 0, kwargs...)
    loglikelihood isa DeviceLogLik ||
        throw(ArgumentError("loglikelihood must be a device likelihood descriptor; there is no CPU fallback"))
    resampling_method in (:systematic, :multinomial, :polyalgo) ||    # :polyalgo (i.i.d. categorical draws) -> multinomial kernel
        throw("Invalid resampler in SMC. Options are :systematic, :multinomial, or :polyalgo")
    n_para = length(parameters)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:smcb200_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), ctx, device)
    st == 0 || error("smcb200_create failed ($st): a CUDA GPU is required")
    h = ctx[]
    try
        check(h, ccall((:smcb200_cloud_create, LIB), Int32, (Ptr{Cvoid}, Int64, Int32), h, n_parts, n_para))
        fixed = Int32[p.fixed for p in parameters]
        lo = Float64[p.valuebounds[1] for p in parameters]; hi = Float64[p.valuebounds[2] for p in parameters]
        kind = Int32[p.fixed ? 0 : PRIOR_KIND[nameof(typeof(p.prior.value))] for p in parameters]
        p1 = Float64[p.fixed ? 0. : Distributions.params(p.prior.value)[1] for p in parameters]
        p2 = Float64[p.fixed ? 1. : Distributions.params(p.prior.value)[2] for p in parameters]
        GC.@preserve fixed lo hi kind p1 p2 check(h, ccall((:smcb200_set_parameters, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
            h, n_para, fixed, lo, hi, kind, p1, p2))
        for (slot, lk) in ((0, loglikelihood), (1, isempty(old_data) ? nothing : old_loglikelihood))
            lk === nothing && continue
            GC.@preserve lk check(h, ccall((:smcb200_set_likelihood, LIB), Int32,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int64),
                h, slot, lik_kind(lk), lk.iparams, length(lk.iparams), lk.dparams, length(lk.dparams)))
        end
        # stage 0: initial_draw! on the device (src/initialization.jl:88-119), or the online update's
        # initialize_likelihoods! on the uploaded old cloud (:153-186)
        # (the prior-mixing bridge of src/smc_main.jl:260-329 is composed from smcb200_resample_weights_n,
        #  smcb200_initial_draw with the old likelihood, smcb200_evaluate(1), smcb200_cloud_write_column and
        #  smcb200_resample exactly as smc_jl_b200/driver.py:bridge_cloud does; omitted here for brevity)
        cloud = continue_intermediate ? load(loadpath, "cloud") : (isempty(old_data) ? Cloud(n_para, n_parts) : old_cloud)
        if continue_intermediate                                             # src/smc_main.jl:334-335
            GC.@preserve cloud check(h, ccall((:smcb200_cloud_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64),
                                              h, cloud.particles, n_parts, 0))
        elseif isempty(old_data)
            values = Float64[p.value for p in parameters]
            GC.@preserve values check(h, ccall((:smcb200_initial_draw, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, UInt64, Int32),
                                               h, values, seed, 1000))
        else
            GC.@preserve cloud check(h, ccall((:smcb200_cloud_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64),
                                              h, cloud.particles, n_parts, 0))
            check(h, ccall((:smcb200_evaluate, LIB), Int32, (Ptr{Cvoid}, Int32), h, 1))
        end
        schedule = ((collect(1:n_Φ) .- 1) / (n_Φ - 1)) .^ λ
        if !continue_intermediate                                            # initialize_cloud_settings!, initialization.jl:196-211
            cloud.tempering_schedule = use_fixed_schedule ? schedule : zeros(1)
            cloud.ESS = [isempty(old_data) ? Float64(n_parts) : cloud.ESS[end]]; cloud.n_Φ = n_Φ; cloud.c = c; cloud.accept = target
        end
        state = StageState(c, target, cloud.ESS[end], 0., 2, 0, 0)
        w_matrix = zeros(n_parts, 1)
        W_matrix = isempty(old_data) ? ones(n_parts, 1) :                   # src/smc_main.jl:363-366
                   reshape(sum(cloud.particles[:, end]) <= 1.0 ? cloud.particles[:, end] * n_parts : cloud.particles[:, end], :, 1)
        inc = Vector{Float64}(undef, n_parts); normw = Vector{Float64}(undef, n_parts)
        i = 1; ϕ_n = 0.
        if continue_intermediate                                             # src/smc_main.jl:355-361
            w_matrix = load(loadpath, "w"); W_matrix = load(loadpath, "W"); j = load(loadpath, "j")
            i = cloud.stage_index; ϕ_n = schedule[i]
            state = StageState(cloud.c, cloud.accept, cloud.ESS[end], schedule[j], j, 0, 0)
        end
        while ϕ_n < 1.                                                     # src/smc_main.jl:377
            t0 = time_ns(); cloud.stage_index = i += 1
            ϕ_n1 = use_fixed_schedule ? schedule[i - 1] : cloud.tempering_schedule[i - 1]
            cfg = StageConfig(ϕ_n1, use_fixed_schedule ? schedule[i] : 0., threshold_ratio, target, α, tempering_target,
                              tempered_update_prior_weight, log_prob_old_data, n_mh_steps, n_blocks,
                              resampling_method == :systematic ? 0 : 1, use_fixed_schedule ? 0 : 1,
                              isempty(old_data) ? 0 : 1, 0, seed, UInt32(i), 0)
            res = StageResult()
            GC.@preserve schedule inc normw check(h, ccall((:smcb200_stage, LIB), Int32,
                (Ptr{Cvoid}, Ref{StageConfig}, Ref{StageState}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ref{StageResult}),
                h, cfg, state, schedule, n_Φ, inc, normw, res))
            ϕ_n = res.phi_n
            use_fixed_schedule || push!(cloud.tempering_schedule, ϕ_n)
            push!(cloud.ESS, res.ess); cloud.resamples += res.resampled; cloud.c = res.c; cloud.accept = res.accept
            w_matrix = hcat(w_matrix, inc); W_matrix = hcat(W_matrix, normw)    # :419-420 (normw is reset to 1 on resample, :445)
            cloud.total_sampling_time += (time_ns() - t0) * 1e-9
            if save_intermediate && mod(cloud.stage_index, intermediate_stage_increment) == 0     # :499-507
                GC.@preserve cloud check(h, ccall((:smcb200_cloud_download, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), h, cloud.particles, n_parts, 0))
                jldopen(replace(savepath, ".jld2" => "_stage=$(cloud.stage_index).jld2"), true, true, true, IOStream) do file
                    write(file, "cloud", cloud); write(file, "w", w_matrix); write(file, "W", W_matrix); write(file, "j", state.j)
                end
            end
        end
        GC.@preserve cloud check(h, ccall((:smcb200_cloud_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64),
                                          h, cloud.particles, n_parts, 0))
        if !testing                                                          # :513-526
            jldopen(savepath, true, true, true, IOStream) do file
                write(file, "cloud", cloud); write(file, "w", w_matrix); write(file, "W", W_matrix)
            end
            h5open(particle_store_path, "w") do file                              # :514-520
                write(file, "smcparams", cloud.particles[:, 1:n_para])
            end
        end
    finally
        ccall((:smcb200_destroy, LIB), Int32, (Ptr{Cvoid},), h)
    end
    nothing
end

"""`resample(weights; method)` (src/resample.jl:23) on the device."""
function resample(weights::Vector{Float64}; n_parts::Int = length(weights), method::Symbol = :systematic,
                  seed::UInt64 = rand(UInt64), device::Int = 0)
    method in (:systematic, :multinomial, :polyalgo) || throw("Invalid resampler in SMC. Options are :systematic, :multinomial, or :polyalgo")
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:smcb200_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), ctx, device) == 0 || error("no GPU")
    idx = Vector{Int64}(undef, n_parts)
    try
        GC.@preserve weights idx check(ctx[], ccall((:smcb200_resample_weights_n, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int32, UInt64, UInt32, Float64, Ptr{Int64}, Ptr{Float64}),
            ctx[], weights, length(weights), n_parts, method == :systematic ? 0 : 1, seed, 0, -1.0, idx, C_NULL))
    finally
        ccall((:smcb200_destroy, LIB), Int32, (Ptr{Cvoid},), ctx[])
    end
    idx
end

get_cloud(path::String) = load(path, "cloud")

end # module
