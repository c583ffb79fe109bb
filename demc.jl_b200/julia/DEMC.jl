# DEMC.jl — drop-in Julia host for the B200 generation loop: same module name, option struct, result struct
# and entry points as chrished/DEMC.jl (src/DEMC.jl:10-43, src/demcz.jl:1-3, src/demcz_anneal.jl:14-16), with the
# generation loop delegated to libdemcz_cuda.so through `ccall` (include/demcz.h).  Likelihoods are registered
# GPU descriptors instead of closures; a plain closure raises ArgumentError (there is no CPU fallback).
#
# NOTE: written without a Julia toolchain in the build image (it cannot be executed there); it contains no
# logic beyond argument marshalling — the same call sequence is exercised by the ctypes harness
# (demc.jl_b200/api.py), which the GPU tests drive.
module DEMC

using Libdl
using LinearAlgebra: I

const libdemcz = get(ENV, "DEMCZ_LIB", joinpath(@__DIR__, "..", "libdemcz_cuda.so"))   # const global: a valid ccall library name

# ---- struct MC / DEMCopt: verbatim field names and defaults (src/DEMC.jl:10-43) ---------------------------
struct MC
    chain::Array{Float64,3}
    log_obj::Array{Float64,2}
    Xcurrent::Array{Float64,2}
    log_objcurrent::Array{Float64,1}
end

mutable struct DEMCopt
    N::Int
    K::Int
    Ngeneration::Int
    Nblocks::Int
    blockindex::Array
    eps_scale::Array{Float64,1}
    γ::Float64
    verbose::Bool
    print_step::Int
    T0::Float64
    TN::Float64
    autostop::Symbol
    autostop_every::Int
    autostop_Rhat::Float64
end

function demcopt(Npar; N=4, K=10, Ngeneration=5000, Nblocks=1, blockindex=[1:Npar], eps_scale=1e-4 * ones(Npar), γ=2.38,
                 verbose=true, print_step=100, T0=3, TN=1e-3, autostop=:Rhat, autostop_every=1000, autostop_Rhat=1.05)
    return DEMCopt(N, K, Ngeneration, Nblocks, blockindex, eps_scale, γ, verbose, print_step, T0, TN, autostop, autostop_every, autostop_Rhat)
end

# ---- likelihood descriptors (replace the `logobj` closure, src/demcz.jl:189) -----------------------------
abstract type GPULikelihood end
struct MvNormalLogpdf <: GPULikelihood          # logpdf(MvNormal(μ, Σ), x)   (test/example_normpdf.jl:13-16)
    μ::Vector{Float64}
    Σ::Matrix{Float64}
    drop_constant::Bool
end
MvNormalLogpdf(μ, Σ) = MvNormalLogpdf(Vector{Float64}(μ), Matrix{Float64}(Σ), false)
QuadraticLogobj(μ) = MvNormalLogpdf(Vector{Float64}(μ), Matrix{Float64}(0.5 * I(length(μ))), true)   # -sum((x.-μ).^2), test/test_anneal.jl:10
struct LinRegLoglik <: GPULikelihood            # -0.5*sum((y .- X*b).^2)     (test/example_linreg.jl:32)
    X::Matrix{Float64}
    y::Vector{Float64}
end

# ---- demcz_opts (include/demcz.h), field for field ---------------------------------------------------------
struct COpts
    struct_size::Int32; N::Int32; K::Int32; Ngeneration::Int32; Nblocks::Int32; d::Int32
    block_ptr::Ptr{Int32}; block_idx::Ptr{Int32}; eps_scale::Ptr{Float64}
    gamma::Float64; T0::Float64; TN::Float64
    anneal::Int32; autostop::Int32; autostop_every::Int32; autostop_Rhat::Float64
    adapt::Int32; min_gamma::Float64; max_gamma::Float64; adapt_every::Int32
    seed::UInt64; gen_offset::Int64
    store_chain::Int32; init_mode::Int32; archive_capacity::Int64
    trace::Int32; profile::Int32
    rank::Int32; world::Int32; N_global::Int64; chain_offset::Int64
    sync_mode::Int32; exchange_layout::Int32
    push_engine::Int32; thin::Int32; moments::Int32; persistent::Int32
    snooker_prob::Float64; snooker_gamma::Float64
end

struct CRunInfo
    generations_done::Int64; generations_total::Int64; stopped_by_rhat::Int32; rhat_valid::Int32
    rhat_max::Float64; final_gamma::Float64; M::Int64; proposals::Int64; accepted::Int64
    step_launches::Int64; step_kernel_ms::Float64; device_ms::Float64; exchange_ms::Float64; exchanges::Int64
end

function check(ctx, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:demcz_last_error, libdemcz), Cstring, (Ptr{Cvoid},), ctx))
    error("libdemcz_cuda: $rc: $msg")
end

register(ctx, l::MvNormalLogpdf) = ccall((:demcz_set_likelihood_mvnormal, libdemcz), Cint,
    (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int32), ctx, length(l.μ), l.μ, l.Σ, l.drop_constant)
register(ctx, l::LinRegLoglik) = ccall((:demcz_set_likelihood_linreg, libdemcz), Cint,
    (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Float64}), ctx, size(l.X, 1), size(l.X, 2), l.X, l.y)
struct LogisticLoglik <: GPULikelihood          # sum(y .* (X*b) .- log1p.(exp.(X*b))): Bernoulli-logit GLM (extension)
    X::Matrix{Float64}
    y::Vector{Float64}
end
register(ctx, l::LogisticLoglik) = ccall((:demcz_set_likelihood_logistic, libdemcz), Cint,
    (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Float64}), ctx, size(l.X, 1), size(l.X, 2), l.X, l.y)
register(ctx, f) = throw(ArgumentError("logobj must be a registered GPU likelihood (MvNormalLogpdf, QuadraticLogobj, LinRegLoglik); closures are not supported and there is no CPU fallback"))

tempbaseline(ig, Ng, T0, TN) = T0 * (TN / T0)^(ig / Ng)          # src/demcz_anneal.jl:1-3

function run_gpu(logobj, Zmat, N, K, Ngeneration, Nblocks, blockindex, eps_scale, γ; anneal, prevrun, autostop, autostop_Rhat,
                 autostop_every, T0, TN, temperaturefun, adaptγ, seed=UInt64(31953150), device=0, store_chain=true,
                 thin=1, moments=false, snooker_prob=0.0, snooker_gamma=0.0)
    d = size(Zmat, 2)
    bp = Int32[0; cumsum(Int32[length(b) for b in blockindex])]
    bi = Int32[k - 1 for b in blockindex for k in b]                 # 1-based Julia -> 0-based ABI
    eps = Vector{Float64}(eps_scale)
    gen_offset = prevrun === nothing ? 0 : size(prevrun.chain, 3)
    GC.@preserve bp bi eps begin
        o = COpts(sizeof(COpts), N, K, Ngeneration, Nblocks, d, pointer(bp), pointer(bi), pointer(eps), γ, T0, TN,
                  anneal, (autostop == :Rhat && !anneal), autostop_every, autostop_Rhat,
                  (anneal && get(adaptγ, "adapt", false)), get(adaptγ, "minγ", 0.1), get(adaptγ, "maxγ", 4.0), get(adaptγ, "adapt_every", 500),
                  seed, gen_offset, store_chain, 0, 0, 0, 0, 0, 1, N, 0, 0, 0,
                  0, thin, moments, -1, snooker_prob, snooker_gamma)
        ctxref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:demcz_create, libdemcz), Cint, (Ref{Ptr{Cvoid}}, Ref{COpts}, Cint), ctxref, Ref(o), device)
        rc == 0 || error("demcz_create: " * unsafe_string(ccall((:demcz_last_error, libdemcz), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    ctx = ctxref[]
    try
        check(ctx, register(ctx, logobj))
        Z64 = Matrix{Float64}(Zmat)
        check(ctx, ccall((:demcz_set_archive, libdemcz), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx, Z64, size(Z64, 1)))
        if prevrun !== nothing                                       # src/demcz.jl:20-21
            check(ctx, ccall((:demcz_set_state, libdemcz), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx,
                             Matrix{Float64}(prevrun.Xcurrent), Vector{Float64}(prevrun.log_objcurrent)))
        end
        if anneal && temperaturefun !== tempbaseline
            T = Float64[temperaturefun(ig, Ngeneration, T0, TN) for ig in 1:Ngeneration]
            check(ctx, ccall((:demcz_set_temperature_table, libdemcz), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx, T, length(T)))
        end
        info = Ref(CRunInfo(0, 0, 0, 0, 0.0, 0.0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0))
        check(ctx, ccall((:demcz_run, libdemcz), Cint, (Ptr{Cvoid}, Int64, Ref{CRunInfo}), ctx, 0, info))
        G = Int(info[].generations_total) ÷ thin                         # stored generations (opts.thin)
        chain = zeros(N, d, store_chain ? G : 0); log_obj = zeros(N, store_chain ? G : 0)
        store_chain && G > 0 && check(ctx, ccall((:demcz_get_chain, libdemcz), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx, chain, log_obj))
        X = zeros(N, d); lp = zeros(N)
        check(ctx, ccall((:demcz_get_state, libdemcz), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ctx, X, lp))
        M = Ref{Int64}(0)
        check(ctx, ccall((:demcz_get_archive, libdemcz), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}), ctx, C_NULL, M))
        Zout = zeros(M[], d)
        check(ctx, ccall((:demcz_get_archive, libdemcz), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}), ctx, Zout, M))
        mc = MC(chain, log_obj, X, lp)
        if prevrun !== nothing                                       # src/demcz.jl:59
            mc = MC(cat(prevrun.chain, mc.chain, dims=3), cat(prevrun.log_obj, mc.log_obj, dims=2), mc.Xcurrent, mc.log_objcurrent)
        end
        return mc, Zout
    finally
        ccall((:demcz_destroy, libdemcz), Cvoid, (Ptr{Cvoid},), ctx)
    end
end

# ---- the reference's entry points (src/demcz.jl:1-3, 9; src/demcz_anneal.jl:14-18) ------------------------------
function demcz_sample(logobj, Zmat, opts::DEMCopt; prevrun=nothing)
    demcz_sample(logobj, Zmat, opts.N, opts.K, opts.Ngeneration, opts.Nblocks, opts.blockindex, opts.eps_scale, opts.γ;
                 prevrun=prevrun, verbose=opts.verbose, print_step=opts.print_step, autostop=opts.autostop,
                 autostop_Rhat=opts.autostop_Rhat, autostop_every=opts.autostop_every)
end

# `ext...`: the keyword extensions of run_gpu with no counterpart in the reference (thin, moments, snooker_prob, snooker_gamma,
# seed, device, store_chain), all off by default
function demcz_sample(logobj, Zmat, N=4, K=10, Ngeneration=5000, Nblocks=1, blockindex=[1:size(Zmat, 2)], eps_scale=1e-4 * ones(size(Zmat, 2)), γ=2.38;
                      prevrun=nothing, verbose=true, print_step=100, autostop=:no, autostop_Rhat=1.01, autostop_every=1000, ext...)
    run_gpu(logobj, Zmat, N, K, Ngeneration, Nblocks, blockindex, eps_scale, γ; anneal=false, prevrun=prevrun, autostop=autostop,
            autostop_Rhat=autostop_Rhat, autostop_every=autostop_every, T0=3.0, TN=1e-3, temperaturefun=tempbaseline, adaptγ=Dict(), ext...)
end

function demcz_anneal(logobj, Zmat, opts::DEMCopt; prevrun=nothing, temperaturefun=tempbaseline,
                      adaptγ=Dict("adapt" => true, "minγ" => 0.1, "maxγ" => 4.0, "adapt_every" => 500))
    demcz_anneal(logobj, Zmat, opts.N, opts.K, opts.Ngeneration, opts.Nblocks, opts.blockindex, opts.eps_scale, opts.γ;
                 prevrun=prevrun, verbose=opts.verbose, print_step=opts.print_step, T0=opts.T0, TN=opts.TN,
                 temperaturefun=temperaturefun, adaptγ=adaptγ)
end

function demcz_anneal(logobj, Zmat, N=4, K=10, Ngeneration=5000, Nblocks=1, blockindex=[1:size(Zmat, 2)], eps_scale=1e-4 * ones(size(Zmat, 2)), γ=2.38;
                      prevrun=nothing, verbose=true, print_step=100, T0=3, TN=0,
                      temperaturefun=tempbaseline, adaptγ=Dict("adapt" => true, "minγ" => 0.1, "maxγ" => 4.0, "adapt_every" => 500), ext...)
    run_gpu(logobj, Zmat, N, K, Ngeneration, Nblocks, blockindex, eps_scale, γ; anneal=true, prevrun=prevrun, autostop=:no,
            autostop_Rhat=1.01, autostop_every=1000, T0=Float64(T0), TN=Float64(TN), temperaturefun=temperaturefun, adaptγ=adaptγ, ext...)
end

# ---- the parallel entry points (src/demcz.jl:101-165, src/demcz_anneal.jl:82-140): every chain is a GPU thread, the
# sync block is the R-hat / γ-adaptation window; `workerpool` is accepted and ignored.  This shim drives ONE GPU; the
# multi-GPU runner (one process per GPU, demcz_comm_* of include/demcz.h) is wired in the Python host
# (demc.jl_b200/api.py::_run with sharded=true, dist.py) and needs a per-GPU process launcher on the Julia side ------
function demcz_sample_par(logobj, Zmat, opts::DEMCopt; sync_every=1000, prevrun=nothing, workerpool=nothing)
    demcz_sample(logobj, Zmat, opts.N, opts.K, opts.Ngeneration, opts.Nblocks, opts.blockindex, opts.eps_scale, opts.γ;
                 prevrun=prevrun, verbose=opts.verbose, autostop=opts.autostop, autostop_Rhat=opts.autostop_Rhat, autostop_every=sync_every)
end

function demcz_sample_par(logobj, Zmat, N=4, K=10, Ngeneration=5000, Nblocks=1, blockindex=[1:size(Zmat, 2)], eps_scale=1e-4 * ones(size(Zmat, 2)), γ=2.38;
                          sync_every=1000, prevrun=nothing, autostop=:Rhat, autostop_Rhat=1.1, verbose=true, workerpool=nothing)
    demcz_sample(logobj, Zmat, N, K, Ngeneration, Nblocks, blockindex, eps_scale, γ; prevrun=prevrun, verbose=verbose,
                 autostop=autostop, autostop_Rhat=autostop_Rhat, autostop_every=sync_every)
end

function demcz_anneal_par(logobj, Zmat, opts::DEMCopt; prevrun=nothing, temperaturefun=tempbaseline, sync_every=1000,
                          adaptγ=Dict("adapt" => true, "minγ" => 0.1, "maxγ" => 4.0, "adapt_every" => 500))
    ad = copy(adaptγ); ad["adapt_every"] = sync_every
    demcz_anneal(logobj, Zmat, opts; prevrun=prevrun, temperaturefun=temperaturefun, adaptγ=ad)
end

end # module
