# julia_smoke.jl — the first thing to run where Julia and a B200 are available: loads WhatsThePoint with the WTPCuda shim
# (INTEGRATION.md section 2) and runs the calls the shim shadows on a small cloud, comparing with the reference's own
# methods reached through `invoke`.
#
#   WTP_CUDA_LIB=/path/to/libwtp_cuda.so julia --project=/path/to/WhatsThePoint.jl scripts/julia_smoke.jl
using WhatsThePoint, Meshes, Unitful, Random, Statistics, Test
import WhatsThePoint: _build_knn_neighbors, _build_radius_neighbors
@assert isdefined(WhatsThePoint, :WTPCuda) "include(\"WTPCuda.jl\") is missing from src/WhatsThePoint.jl"

Random.seed!(1)
for T in (Float64, Float32)
    pts = [Point(rand(T) * u"m", rand(T) * u"m", rand(T) * u"m") for _ in 1:20_000]
    gpu = _build_knn_neighbors(pts, 21)
    cpu = invoke(_build_knn_neighbors, Tuple{Any, Int}, pts, 21)                   # the reference's method
    @test gpu isa Vector{Vector{Int}} && length(gpu) == length(pts)
    @test mean(gpu[i] == cpu[i] for i in eachindex(pts)) > 0.999                    # equal but for exact distance ties
    r = T(0.05) * u"m"
    g2 = _build_radius_neighbors(pts, r)
    c2 = invoke(_build_radius_neighbors, Tuple{Any, Any}, pts, r)
    @test all(sort(g2[i]) == sort(c2[i]) for i in eachindex(pts))
    bnd = PointBoundary(pts[1:2000], [Meshes.Vec(zero(T), zero(T), one(T)) for _ in 1:2000], fill(T(1.0e-3) * u"m^2", 2000))
    cloud = PointCloud(bnd, PointVolume(pts[2001:end]))
    c3 = set_topology(cloud, KNNTopology, 10)
    @test length(neighbors(c3)) == length(cloud) && all(length(n) == 10 for n in neighbors(c3))
    conv = T[]
    out = repel(cloud, ConstantSpacing(T(0.03) * u"m"); β = T(0.2), max_iters = 5, convergence = conv)
    @test length(conv) <= 5 && all(isfinite, conv)
    m = metrics(cloud; k = 10)
    @test m.mesh_ratio >= 1
end
println("julia smoke ok")
