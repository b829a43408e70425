# make_reference_fixtures.jl — run the UNMODIFIED reference (no WTPCuda shim) on the inputs written by
# scripts/export_fixture_inputs.py and store what it returns, so that the CPU oracle and the device path can be pinned to the
# real thing (the reference's own tests hold no neighbour identities and no trajectories: SURVEY.md §8c).
#
#   julia --project=/path/to/WhatsThePoint.jl -t auto scripts/make_reference_fixtures.jl tests/golden/reference
#
# Written against the reference sources (src/topology.jl:79-100, src/repel.jl:202-206, src/repel_forces.jl:88-100,
# src/discretization/spacings.jl:35-39, 93-133); not executed in the build image (no Julia there).
using WhatsThePoint, Meshes, Unitful, JSON        # JSON is not a dependency of the reference: `] add JSON` in the environment used to run this
import WhatsThePoint: _build_knn_neighbors, _build_radius_neighbors, _relax!

root = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "reference")
indir, outdir = joinpath(root, "inputs"), joinpath(root, "outputs")
mkpath(outdir)
manifest = JSON.parsefile(joinpath(indir, "manifest.json"))

eltype_of(s) = s == "float32" ? Float32 : Float64
function read_points(c)
    T, n, d = eltype_of(c["dtype"]), c["n"], c["d"]
    raw = Array{T}(undef, d, n)                      # row-major n × d on disk = column-major d × n here
    read!(joinpath(indir, c["file"]), raw)
    u = u"m"
    return d == 2 ? [Point(raw[1, i] * u, raw[2, i] * u) for i in 1:n] : [Point(raw[1, i] * u, raw[2, i] * u, raw[3, i] * u) for i in 1:n]
end
coords(pts, T) = reduce(hcat, [T.(ustrip.(Meshes.to(p))) for p in pts])           # d × n
write_array(name, a) = open(io -> write(io, a), joinpath(outdir, name), "w")

done = Dict{String, Any}[]
for c in manifest
    T = eltype_of(c["dtype"])
    pts = read_points(c)
    entry = Dict{String, Any}("name" => c["name"], "kind" => c["kind"])
    if c["kind"] == "knn"
        k = c["k"]
        nb = _build_knn_neighbors(pts, k)                                          # Vector{Vector{Int}}, src/topology.jl:79-84
        table = Matrix{Int64}(undef, k, length(pts))
        for (i, row) in enumerate(nb); table[:, i] .= row; end
        write_array(c["name"] * ".idx.bin", table)                                 # k × n column-major = n × k row-major
        entry["idx"] = c["name"] * ".idx.bin"
    elseif c["kind"] == "radius"
        nb = _build_radius_neighbors(pts, T(c["r"]) * u"m")                        # src/topology.jl:91-97 (order = tree traversal: compare as sets)
        offsets = Int64[0]
        for row in nb; push!(offsets, offsets[end] + length(row)); end
        write_array(c["name"] * ".off.bin", offsets)
        write_array(c["name"] * ".ind.bin", Int64.(reduce(vcat, nb; init = Int[])))
        entry["off"] = c["name"] * ".off.bin"; entry["ind"] = c["name"] * ".ind.bin"
    elseif c["kind"] == "repel"
        nf = c["n_fixed"]
        s = c["spacing"]
        bnd = pts[1:nf]
        spacing = s["kind"] == "constant" ? ConstantSpacing(T(s["a"]) * u"m") :
                  BoundaryLayerSpacing(bnd; at_wall = T(s["a"]) * u"m", bulk = T(s["b"]) * u"m", layer_thickness = T(s["c"]) * u"m")
        p = pts[(nf + 1):end]; p_old = copy(p); snap = copy(pts)
        conv = _relax!(p, p_old, snap, spacing, ClippedSpacingForce(T(c["beta"])), (id, xi, xp) -> xp;
                       n_fixed = nf, n_protected = nf, α_lo = T(c["alpha_lo"]), α_max = T(c["alpha_max"]), k = c["k"],
                       max_iters = c["max_iters"], tol = zero(T), rebuild_every = 1, kick_after = 0, trace = nothing,
                       stall_after = 0, cv_target = 0.0)                           # src/repel.jl:202-206
        write_array(c["name"] * ".pos.bin", coords(vcat(bnd, p), T))
        write_array(c["name"] * ".conv.bin", T.(conv))
        entry["pos"] = c["name"] * ".pos.bin"; entry["conv"] = c["name"] * ".conv.bin"; entry["iters"] = length(conv)
    end
    push!(done, entry)
    println("done ", c["name"])
end
open(io -> JSON.print(io, Dict("julia" => string(VERSION), "cases" => done), 1), joinpath(outdir, "manifest.json"), "w")
