# raw_to_jld.jl -- turns the HDF5-free hand-over written by sprb200.api.export_surrogate_raw
# (basename.f64 + basename.toml) into the surrogate .jld file of SPRFittingPaper2023.jl, using the
# reference's own writer (src/surrogate.jl:112-123, 151-158), so that Surrogate(filename) and fitting.jl
# load it unchanged.  Only the TOML standard library is needed besides the package itself.
#
#   julia raw_to_jld.jl basename out.jld
#
# Not executed in the build image (no Julia there); written against the reference sources.
using SPRFittingPaper2023, JLD, TOML

function raw_to_jld(basename::AbstractString, outfile::AbstractString)
    m = TOML.parsefile(basename * ".toml")
    sz = Tuple(Int.(m["surrogate_size"]))
    table = Array{Float64,5}(undef, sz)
    read!(basename * ".f64", table)                     # raw little-endian Float64, column-major
    pair(v) = (Float64(v[1]), Float64(v[2]))
    surpars = SurrogateParams(; logkon_range=pair(m["logkon_range"]), logkoff_range=pair(m["logkoff_range"]),
                              logkonb_range=pair(m["logkonb_range"]), reach_range=pair(m["reach_range"]),
                              antigenconcen=Float64(m["antigenconcen"]), antibodyconcen=Float64(m["antibodyconcen"]),
                              CP=Float64(m["CP"]))
    simpars = SimParams(; antigenconcen=surpars.antigenconcen, N=Int(m["N"]), tstop=Float64(m["tstop"]),
                        tstop_AtoB=Float64(m["tstop_AtoB"]), tsave=Vector{Float64}(m["tsave"]),
                        L=Float64(m["L"]), DIM=Int(m["DIM"]))
    jldopen(outfile, "w") do file
        write(file, "FirstMoment", table)
        SPRFittingPaper2023.save_surrogate_metadata(file, sz, surpars, simpars)
    end
    outfile
end

if abspath(PROGRAM_FILE) == @__FILE__
    length(ARGS) == 2 || error("usage: julia raw_to_jld.jl basename out.jld")
    println(raw_to_jld(ARGS[1], ARGS[2]))
end
