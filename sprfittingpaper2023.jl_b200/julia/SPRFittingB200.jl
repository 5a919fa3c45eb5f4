# SPRFittingB200.jl -- Julia host side of libsprb200.so: a drop-in for the hot path of
# SPRFittingPaper2023.jl (run_spr_sim! and the surrogate builders) that ccalls the sm_100a kernels.
#
# NOTE: Julia is not installed in the build image (no network), so this file has been written
# against the reference sources and the C ABI (include/sprb200.h) but NOT executed there; the same
# entry points are exercised from Python/ctypes in tests/.  See INTEGRATION.md.
#
# Usage:
#     using SPRFittingPaper2023, SPRFittingB200
#     SPRFittingB200.run_spr_sim!(tbo, biopars, simpars, VarianceTerminator())   # same signature
#     SPRFittingB200.enable!()   # optional: make SPRFittingPaper2023.run_spr_sim! / build_surrogate_*
#                                # dispatch here, so fitting.jl (visualisefit/savefit) and the example
#                                # scripts run unchanged
#
# What is replaced (paths in the reference tree):
#   run_spr_sim!                    src/forward_simulator.jl:105-366
#   build_surrogate_serial          src/surrogate.jl:203-242
#   build_surrogate_slice           src/surrogate.jl:273-300
# What is kept from the reference unchanged: parameter structs, outputters, terminators,
# save_surrogate_metadata / save_surrogate / save_surrogate_slice / merge_surrogate_slices (JLD
# writers), Surrogate(lutfile), fitting.jl -- so the surrogate file layout is the reference's own.
module SPRFittingB200

using SPRFittingPaper2023
using SPRFittingPaper2023: BioPhysParams, SimParams, SurrogateParams, TotalBoundOutputter,
    TotalAOutputter, SimNumberTerminator, VarianceTerminator, getdim
import SPRFittingPaper2023: isnotdone, update!, reset!
using OnlineStats: Variance, Mean, nobs      # merge! is Base.merge!, which OnlineStatsBase extends
using StaticArrays: SVector
using Random: RandomDevice

const LIB = get(ENV, "SPRB200_LIB",
                joinpath(@__DIR__, "..", "lib", "libsprb200.so"))

# ---- C structs (include/sprb200.h) ---------------------------------------------------------
struct SprPoint
    kon_eff::Cdouble; koff::Cdouble; konb::Cdouble; reach::Cdouble
end
struct SprSim
    N::Int32; DIM::Int32; L::Cdouble; tstop::Cdouble; tstop_AtoB::Cdouble
    T::Int32; tsave::Ptr{Cdouble}; resample_initlocs::Int32; initlocs::Ptr{Cdouble}
end
struct SprRun
    term_kind::Int32; nsims::Int32; ssetol::Cdouble; minsims::Int32; maxsims::Int32
    observable::Int32; replicate_base::Int32; seed::UInt64; param_index_base::UInt64
end
struct SprOut
    sum::Ptr{Int64}; sumsq::Ptr{UInt64}; n_used::Ptr{Int32}; n_events::Ptr{UInt64}; mean::Ptr{Cdouble}
end
struct SprGrid
    lo::NTuple{4,Cdouble}; hi::NTuple{4,Cdouble}; n::NTuple{4,Int32}
end
const TERM_FIXED, TERM_VARIANCE = Int32(0), Int32(1)
const OBS_TOTAL_BOUND, OBS_TOTAL_A = Int32(0), Int32(1)

# ---- one context per GPU --------------------------------------------------------------------
const HANDLES = Dict{Int,Ptr{Cvoid}}()
const SEED = Ref{Union{Nothing,UInt64}}(nothing)   # set with seed!(); `nothing` = fresh key per call

seed!(s::Integer) = (SEED[] = UInt64(s); nothing)
nextseed() = SEED[] === nothing ? rand(RandomDevice(), UInt64) : (SEED[] += 1; SEED[])

function handle(device::Integer=0)
    get!(HANDLES, device) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:sprb200_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
        rc == 0 || error("sprb200_create failed ($rc): " *
                         unsafe_string(ccall((:sprb200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        h[]
    end
end

function check(h, rc)
    rc == 0 || error("libsprb200 error $rc: " *
                     unsafe_string(ccall((:sprb200_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))
    nothing
end

# Vector{SVector{DIM,Float64}} is DIM contiguous doubles per particle: exactly SprSim.initlocs
function with_sim(f, simpars::SimParams)
    tsave = convert(Vector{Float64}, simpars.tsave)
    locs = simpars.initlocs
    GC.@preserve tsave locs begin
        sim = SprSim(Int32(simpars.N), Int32(getdim(simpars)), simpars.L, simpars.tstop,
                     simpars.tstop_AtoB, Int32(length(tsave)), pointer(tsave),
                     Int32(simpars.resample_initlocs),
                     simpars.resample_initlocs ? Ptr{Cdouble}(C_NULL) : Ptr{Cdouble}(pointer(locs)))
        f(sim)
    end
end

point(b::BioPhysParams) = SprPoint(b.kon * b.antibodyconcen, b.koff, b.konb, b.reach)

# integer moments -> OnlineStats state, so means/vars/sems (callbacks.jl:54-80) keep working
function variance_from_moments(n, S, Q, scale)
    v = Variance()
    n == 0 && return v
    μ = scale * S / n
    σ2 = n > 1 ? scale^2 * max(Q - S * S / n, 0.0) / n : 0.0    # OnlineStats stores the biased value
    v.μ = μ; v.σ2 = σ2; v.n = n
    v
end
mean_from_moments(n, S, scale) = (m = Mean(); m.μ = n == 0 ? 0.0 : scale * S / n; m.n = n; m)

"""
    run_spr_sim!(outputter, biopars, simpars, terminator = SimNumberTerminator(); device=0, seed=nothing)

Same contract as `SPRFittingPaper2023.run_spr_sim!`.  `TotalBoundOutputter`/`TotalAOutputter` with
`SimNumberTerminator`/`VarianceTerminator` are reduced on the GPU; any other callable outputter or
terminator is replayed on the host from raw copynumbers, replicate by replicate.
"""
function run_spr_sim!(outputter, biopars::BioPhysParams, simpars::SimParams,
                      terminator=SimNumberTerminator(); device=0, seed=nothing, param_index=0,
                      replay_chunk=64)
    h = handle(device)
    key = seed === nothing ? nextseed() : UInt64(seed)
    T = length(simpars.tsave)
    builtin_out = outputter isa TotalBoundOutputter || outputter isa TotalAOutputter
    builtin_term = terminator isa SimNumberTerminator || terminator isa VarianceTerminator
    # the on-device VarianceTerminator scans replicates from the first one: an outputter that already holds
    # replicates (the reference keeps accumulating into it, forward_simulator.jl:357) takes the replay path
    fresh = builtin_out && nobs(first(outputter.bindcnt)) == 0
    if builtin_out && builtin_term && (fresh || terminator isa SimNumberTerminator)
        obs = outputter isa TotalAOutputter ? OBS_TOTAL_A : OBS_TOTAL_BOUND
        run = if terminator isa SimNumberTerminator
            todo = simpars.nsims - terminator.num_completed_sims
            todo <= 0 && (outputter(biopars, simpars); return nothing)
            SprRun(TERM_FIXED, Int32(todo), 0.0, 0, 0, obs, Int32(terminator.num_completed_sims), key, UInt64(param_index))
        else
            terminator.notdone || (outputter(biopars, simpars); return nothing)
            SprRun(TERM_VARIANCE, 0, terminator.ssetol, Int32(terminator.minsims), Int32(terminator.maxsims),
                   obs, 0, key, UInt64(param_index))
        end
        S = zeros(Int64, T); Q = zeros(UInt64, T); nused = zeros(Int32, 1)
        pts = [point(biopars)]
        with_sim(simpars) do sim
            GC.@preserve S Q nused pts begin
                out = SprOut(pointer(S), pointer(Q), pointer(nused), Ptr{UInt64}(C_NULL), Ptr{Cdouble}(C_NULL))
                check(h, ccall((:sprb200_simulate, LIB), Cint,
                               (Ptr{Cvoid}, Ref{SprSim}, Ptr{SprPoint}, Int64, Ref{SprRun}, Ref{SprOut}),
                               h, sim, pts, 1, run, out))
            end
        end
        n = Int(nused[1]); scale = biopars.CP / simpars.N
        # accumulate INTO the outputter's running statistics, as the per-replicate fit! of the reference does
        # (forward_simulator.jl:357, callbacks.jl:27-30): OnlineStats merges two Variance/Mean states exactly
        for i in 1:T
            merge!(outputter.bindcnt[i], outputter isa TotalAOutputter ?
                   mean_from_moments(n, Float64(S[i]), scale) :
                   variance_from_moments(n, Float64(S[i]), Float64(Q[i]), scale))
        end
        terminator isa SimNumberTerminator ? (terminator.num_completed_sims += n) : (terminator.notdone = false)
        outputter(biopars, simpars)
        return nothing
    end
    # ---- generic replay from raw copynumbers (rows A, B, C; forward_simulator.jl:189-191)
    done = 0
    copynumbers = zeros(Int, 3, T)
    while isnotdone(terminator, biopars, simpars)
        chunk = terminator isa SimNumberTerminator ?
            min(replay_chunk, simpars.nsims - terminator.num_completed_sims) : replay_chunk
        B = zeros(UInt16, T, chunk); C = zeros(UInt16, T, chunk)
        run = SprRun(TERM_FIXED, Int32(chunk), 0.0, 0, 0, OBS_TOTAL_BOUND, Int32(done), key, UInt64(param_index))
        pts = [point(biopars)]
        with_sim(simpars) do sim
            GC.@preserve B C pts check(h, ccall((:sprb200_simulate_raw, LIB), Cint,
                (Ptr{Cvoid}, Ref{SprSim}, Ptr{SprPoint}, Int64, Ref{SprRun}, Ptr{UInt16}, Ptr{UInt16}, Ptr{UInt64}),
                h, sim, pts, 1, run, B, C, C_NULL))
        end
        for r in 1:chunk
            isnotdone(terminator, biopars, simpars) || break
            @inbounds for s in 1:T
                b = Int(B[s, r]); c = Int(C[s, r])
                copynumbers[1, s] = simpars.N - b - 2c; copynumbers[2, s] = b; copynumbers[3, s] = c
            end
            outputter(simpars.tsave, copynumbers, biopars, simpars)
            update!(terminator, outputter, biopars, simpars)
        end
        done += chunk
    end
    outputter(biopars, simpars)
    nothing
end

# ---- surrogate builders -------------------------------------------------------------------------
grid(sp::SurrogateParams, size4) = SprGrid(
    (sp.logkon_range[1], sp.logkoff_range[1], sp.logkonb_range[1], sp.reach_range[1]),
    (sp.logkon_range[2], sp.logkoff_range[2], sp.logkonb_range[2], sp.reach_range[2]),
    Int32.(Tuple(size4)))

function runspec(terminator, simpars, key)
    if terminator isa VarianceTerminator
        SprRun(TERM_VARIANCE, 0, terminator.ssetol, Int32(terminator.minsims), Int32(terminator.maxsims),
               OBS_TOTAL_BOUND, 0, key, 0)
    elseif terminator isa SimNumberTerminator
        SprRun(TERM_FIXED, Int32(simpars.nsims), 0.0, 0, 0, OBS_TOTAL_BOUND, 0, key, 0)
    else
        error("surrogate builds support SimNumberTerminator and VarianceTerminator")
    end
end

"`build_surrogate_serial` (surrogate.jl:203-242): returns the (nkon,nkoff,nkonb,nreach,ntime) array."
function build_surrogate_serial(surrogate_size::Tuple, surpars::SurrogateParams, simpars::SimParams;
                                terminator=VarianceTerminator(), device=0, devices=nothing, seed=nothing)
    @assert surrogate_size[end] == length(simpars.tsave)
    key = seed === nothing ? nextseed() : UInt64(seed)
    table = zeros(surrogate_size)              # column-major: exactly the library's [T][P] layout
    g = grid(surpars, surrogate_size[1:4])
    if devices !== nothing && length(devices) > 1
        # every listed GPU of this process takes a cost-balanced share; same bytes as one GPU
        devs = Int32.(collect(devices))
        with_sim(simpars) do sim
            GC.@preserve table devs begin
                rc = ccall((:sprb200_build_table_multi, LIB), Cint,
                    (Ptr{Int32}, Int32, Ref{SprGrid}, Ref{SprSim}, Ref{SprRun}, Ptr{Cdouble}, Ptr{Int32}, Ptr{UInt64}),
                    devs, length(devs), g, sim, runspec(terminator, simpars, key), table, C_NULL, C_NULL)
                rc == 0 || error("sprb200 error $rc: " * unsafe_string(ccall((:sprb200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
            end
        end
        return table
    end
    h = handle(devices === nothing ? device : first(devices))
    with_sim(simpars) do sim
        GC.@preserve table check(h, ccall((:sprb200_build_table, LIB), Cint,
            (Ptr{Cvoid}, Ref{SprGrid}, Ref{SprSim}, Ref{SprRun}, Ptr{Cdouble}, Ptr{Int32}, Ptr{UInt64}),
            h, g, sim, runspec(terminator, simpars, key), table, C_NULL, C_NULL))
    end
    table
end

"`build_surrogate_slice` (surrogate.jl:273-300): ntime x nidx matrix, column n = index idxstart+n-1."
function build_surrogate_slice(surmetadata::Dict, idxstart, idxend; terminator=VarianceTerminator(),
                               device=0, seed=get(surmetadata, "seed", 0))
    p = SPRFittingPaper2023.unpack_metadata(surmetadata)   # same quirk: default-geometry SimParams
    h = handle(device)
    T = p.surrogate_size[end]
    out = zeros(T, idxend - idxstart + 1)
    sp = SurrogateParams(; logkon_range=surmetadata["logkon_range"], logkoff_range=surmetadata["logkoff_range"],
                         logkonb_range=surmetadata["logkonb_range"], reach_range=surmetadata["reach_range"])
    g = grid(sp, p.surrogate_size[1:4])
    with_sim(p.simpars) do sim
        GC.@preserve out check(h, ccall((:sprb200_build_slice, LIB), Cint,
            (Ptr{Cvoid}, Ref{SprGrid}, Ref{SprSim}, Ref{SprRun}, Int64, Int64, Ptr{Cdouble}, Ptr{Int32}, Ptr{UInt64}),
            h, g, sim, runspec(terminator, p.simpars, UInt64(seed)), idxstart, idxend, out, C_NULL, C_NULL))
    end
    out
end

# ---------------------------------------------------------------------------------------------
# The consumer of the table: surrogate_sprdata_error (src/fitting.jl:38-91) for a whole population.
struct SprAligned
    nconc::Int32; npts::Ptr{Int32}; times::Ptr{Cdouble}; values::Ptr{Cdouble}; antibodyconcens::Ptr{Cdouble}
end

"""
    load_surrogate!(surrogate::Surrogate; device=0)

Uploads `surrogate`'s FirstMoment table once (`sprb200_surrogate_load`); it stays resident on the GPU.
Needs the raw table: pass the `Array{Float64,5}` the surrogate was built from.
"""
function load_surrogate!(table::Array{Float64,5}, surpars::SurrogateParams, simpars::SimParams; device=0)
    h = handle(device)
    sz = size(table)
    g = SprGrid((surpars.logkon_range[1], surpars.logkoff_range[1], surpars.logkonb_range[1], surpars.reach_range[1]),
                (surpars.logkon_range[2], surpars.logkoff_range[2], surpars.logkonb_range[2], surpars.reach_range[2]),
                (Int32(sz[1]), Int32(sz[2]), Int32(sz[3]), Int32(sz[4])))
    ts = simpars.tsave
    GC.@preserve table check(h, ccall((:sprb200_surrogate_load, LIB), Cint,
        (Ptr{Cvoid}, Ref{SprGrid}, Int32, Cdouble, Cdouble, Ptr{Cdouble}, Int32),
        h, g, Int32(sz[5]), first(ts), last(ts), table, Int32(0)))
    nothing
end

"""
    surrogate_sprdata_errors(population::Matrix{Float64}, aligneddat::AlignedData; device=0)

`surrogate_sprdata_error` for every column `[logkon, logkoff, logkonb, reach, logCP]` of `population` in one
GPU call (after `load_surrogate!`): what a population-based optimiser (BlackBoxOptim's xNES/DE) evaluates per
generation.  `+Inf` where the reference would throw a BoundsError.
"""
function surrogate_sprdata_errors(population::Matrix{Float64}, aligneddat::AlignedData; device=0)
    size(population, 1) == 5 || error("population must be 5 x npop")
    h = handle(device)
    npts = Int32.(length.(aligneddat.times))
    times = reduce(vcat, aligneddat.times); vals = reduce(vcat, aligneddat.refdata)
    abcs = Vector{Float64}(aligneddat.antibodyconcens)
    loss = Vector{Float64}(undef, size(population, 2))
    GC.@preserve npts times vals abcs population loss begin
        al = SprAligned(Int32(length(npts)), pointer(npts), pointer(times), pointer(vals), pointer(abcs))
        check(h, ccall((:sprb200_fit_loss, LIB), Cint, (Ptr{Cvoid}, Ref{SprAligned}, Ptr{Cdouble}, Int64, Ptr{Cdouble}),
                       h, al, population, size(population, 2), loss))
    end
    loss
end

"""
    enable!()

Routes the reference's own entry points through the GPU: after this call
`SPRFittingPaper2023.run_spr_sim!`, `build_surrogate_serial` and `build_surrogate_slice` (and
therefore `save_surrogate`, `save_surrogate_slice`, `visualisefit`, `savefit`, the example scripts)
run on the B200.  The JLD files are still written by the reference's own writers.
"""
function enable!()
    @eval SPRFittingPaper2023 begin
        run_spr_sim!(o, b::BioPhysParams, s::SimParams, t=SimNumberTerminator()) =
            $(run_spr_sim!)(o, b, s, t)
        build_surrogate_serial(sz::Tuple, sp::SurrogateParams, s::SimParams; terminator=VarianceTerminator()) =
            $(build_surrogate_serial)(sz, sp, s; terminator)
        build_surrogate_slice(m::Dict, i0, i1; terminator=VarianceTerminator()) =
            $(build_surrogate_slice)(m, i0, i1; terminator)
    end
    nothing
end

end # module
