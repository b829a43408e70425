# WTPCuda.jl — the Julia side of the drop-in boundary.
#
# Julia is not installed in the image this library is built in: this file has been written against the reference
# sources (file:line citations below) and the byte-level calls it makes are the ones tests/ drive through ctypes, but
# it has NOT been executed. scripts/julia_smoke.jl is the first thing to run where Julia is available.
#
# Include this file from src/WhatsThePoint.jl after the last `include` (it needs repel.jl, metrics.jl, isinside.jl):
#
#     include("WTPCuda.jl")          # adds module WTPCuda
#
# Every method below is MORE SPECIFIC than the reference method it shadows (typed `points::AbstractVector{<:Point}`,
# `spacing::BuiltinSpacing`, `cloud::PointCloud{𝔼{N}, C}`), never a redefinition of the same signature: Julia ≥ 1.10
# refuses method overwriting during precompilation, and the reference's own methods stay reachable for everything that
# cannot cross the C ABI (user spacing callables, user force models, other `constrain` closures) — those reach the
# reference's CPU code exactly as before, nothing is silently approximated.
#
#   reference function                            file:line                      C ABI
#   _build_knn_neighbors(points, k)               src/topology.jl:79-84          wtp_knn_{f32,f64}
#   _build_radius_neighbors(points, radius)       src/topology.jl:91-97          wtp_radius_count_{f32,f64} + wtp_radius_fill
#   _relax!(p, p_old, snap, spacing, fm, constrain; ...)   src/repel.jl:202-339  wtp_repel_{f32,f64}
#   repel(cloud, spacing; ...)  (survivor filter) src/repel.jl:56-95             wtp_spacing_eval + wtp_repel + wtp_isinside + wtp_cull_mask
#   _near_duplicate_keep_mask(pts, spacings, r)   src/repel.jl:565-580           wtp_cull_mask_{f32,f64}
#   searchdists(cloud, ::KNearestSearch)          src/neighbors.jl:16-21         wtp_knn_self_{f32,f64}
#   search(cloud, ::KNearestSearch)               src/neighbors.jl:9-14          wtp_knn_self_{f32,f64}
#   metrics(cloud; k)                             src/metrics.jl:19-41           wtp_metrics_{f32,f64}
#   spacing_metrics(cloud, spacing; k)            src/metrics.jl:56-71           wtp_spacing_metrics_{f32,f64}
#   spacing_fidelity_metrics(cloud, spacing; ...) src/metrics.jl:88-129          wtp_spacing_fidelity_{f32,f64}
#   compute_normals(points; k)                    src/normals.jl:9-44            wtp_normals_{f32,f64}
#
# No CUDA.jl, no KernelAbstractions, no CPU fallback inside the library: if libwtp_cuda.so or the GPU is missing the
# calls throw.

module WTPCuda

using Meshes, CoordRefSystems, Unitful, StaticArrays, Random, LinearAlgebra   # CRS comes from CoordRefSystems, as in src/WhatsThePoint.jl:4
import Meshes: search, searchdists
import ..WhatsThePoint
import ..WhatsThePoint: _build_knn_neighbors, _build_radius_neighbors, _relax!, _get_radius, _edge_key,
    _near_duplicate_keep_mask, _cull, repel, metrics, spacing_metrics, spacing_fidelity_metrics, compute_normals,
    RepelForceModel, InverseDistanceForce, SpacingEquilibriumForce, ClippedSpacingForce, StrongSpacingForce,
    ConstantSpacing, LogLike, BoundaryLayerSpacing, PointCloud, PointBoundary, PointSurface, PointVolume, NoTopology,
    points, boundary, volume, surfaces, normal, area

const LIB = get(ENV, "WTP_CUDA_LIB", "libwtp_cuda.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
const BuiltinSpacing = Union{ConstantSpacing, LogLike, BoundaryLayerSpacing}
# above this many points the topology builders return flat storage instead of N heap vectors (src/topology.jl:25: the
# storage type is a parameter of the topology); test/topology.jl:28,61 assert Vector{Vector{Int}} on tiny clouds
const FLAT_THRESHOLD = parse(Int, get(ENV, "WTP_FLAT_THRESHOLD", "1000000"))

# One context per process: a single device (WTP_CUDA_DEVICE, default 0) or, with WTP_CUDA_DEVICES="0,1,2,3", one
# context over several GPUs of the box (wtp_create_multi: one host thread and one stream per device inside the
# library, queries sharded by runs of the sorted order, rows exchanged over NVLink).
function ctx()
    if CTX[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        devs = get(ENV, "WTP_CUDA_DEVICES", "")
        rc = if isempty(devs)
            ccall((:wtp_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), h, parse(Int32, get(ENV, "WTP_CUDA_DEVICE", "0")))
        else
            d = Int32[parse(Int32, strip(s)) for s in split(devs, ',')]
            ccall((:wtp_create_multi, LIB), Int32, (Ref{Ptr{Cvoid}}, Ptr{Int32}, Int32), h, d, length(d))
        end
        rc == 0 || error("libwtp_cuda: context creation failed with status $rc (no usable B200; there is no CPU fallback)")
        CTX[] = h[]
        atexit(() -> ccall((:wtp_destroy, LIB), Cvoid, (Ptr{Cvoid},), CTX[]))
    end
    return CTX[]
end

function check(rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:wtp_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx()))
    rc in (1, 2) ? throw(ArgumentError(msg)) : error("libwtp_cuda: $msg")   # WTP_ERR_BAD_ARG / K_TOO_LARGE -> ArgumentError
end

# Vector{Point{𝔼{D},Cartesian{…,D,Quantity{T}}}} is an isbits array of D contiguous T (units are type-level only):
# reinterpreting it is what _raw_point does element by element (src/repel.jl:350), without a copy.
machine_type(pts) = typeof(ustrip(Meshes.to(first(pts))[1]))
dimension(pts) = length(Meshes.to(first(pts)))
function raw(pts::AbstractVector{<:Point}, ::Type{T}) where {T}
    v = pts isa Vector ? pts : collect(pts)               # SubArrays / lazy vcat results: one dense copy
    isbitstype(eltype(v)) && sizeof(eltype(v)) == dimension(v) * sizeof(T) ||
        error("libwtp_cuda: point type $(eltype(v)) is not D contiguous $T values")
    return reinterpret(T, v)
end
# a dense D × n matrix of T from any vector of small vectors (normals: Vec / SVector / Vector, with or without units)
function dense(vs, ::Type{T}, D::Int) where {T}
    m = Matrix{T}(undef, D, length(vs))
    @inbounds for (j, v) in enumerate(vs), d in 1:D
        m[d, j] = T(ustrip(v[d]))
    end
    return m
end
unbox(x) = x isa Core.Box ? x.contents : x                # a captured variable the compiler boxed

# ------------------------------------------------------------ flat neighbour storage (src/topology.jl:25,45)
"N × k table as `AbstractVector{<:AbstractVector{Int}}`: `rows[i]` is a view of column i, no per-point allocation."
struct FlatRows <: AbstractVector{SubArray{Int, 1, Matrix{Int}, Tuple{Base.Slice{Base.OneTo{Int}}, Int}, true}}
    table::Matrix{Int}                                     # k × N: column i = neighbours of point i (row-major N × k for C)
end
Base.size(r::FlatRows) = (size(r.table, 2),)
Base.@propagate_inbounds Base.getindex(r::FlatRows, i::Int) = view(r.table, :, i)
Base.IndexStyle(::Type{FlatRows}) = IndexLinear()

"Ragged lists over one CSR pair (RadiusTopology storage): `rows[i]` is a view into `indices`."
struct CSRRows <: AbstractVector{SubArray{Int, 1, Vector{Int}, Tuple{UnitRange{Int}}, true}}
    offsets::Vector{Int}                                   # N + 1, 0-based prefix
    indices::Vector{Int}
end
Base.size(r::CSRRows) = (length(r.offsets) - 1,)
Base.@propagate_inbounds Base.getindex(r::CSRRows, i::Int) = view(r.indices, (r.offsets[i] + 1):r.offsets[i + 1])
Base.IndexStyle(::Type{CSRRows}) = IndexLinear()

# ------------------------------------------------------------ C structs (include/wtp_cuda.h)
struct CForce; kind::Int32; beta::Float64; u0::Float64; gamma::Float64; end
struct CSpacing; kind::Int32; a::Float64; b::Float64; c::Float64; bnd::Ptr{Cvoid}; n_bnd::Int64; end
struct CParams
    k::Int32; max_iters::Int32; rebuild_every::Int32; stall_after::Int32; kick_after::Int32; wall::Int32
    want_trace::Int32; reserved::Int32; alpha_lo::Float64; alpha_max::Float64; tol::Float64; cv_target::Float64
    n_protected::Int64; kick_seed::UInt64; deposit_ratio::Float64
end
struct CResult; iters::Int32; stop_reason::Int32; last_cv::Float64; end
struct CTrace; r::Float64; s::Float64; r_over_s::Float64; idx_a::Int64; idx_b::Int64; end
struct CWallMesh
    triangles::Ptr{Cvoid}; feature_normals::Ptr{Cvoid}; n_tri::Int64
    bbox_min::NTuple{3, Float64}; bbox_max::NTuple{3, Float64}; offset_dist::Float64
    is_bnd::Ptr{UInt8}; tri_indices::Ptr{Int64}; escaped::Ptr{UInt8}
end
struct CCloudMetrics; avg::Float64; std::Float64; max::Float64; min::Float64; separation::Float64; fill::Float64; mesh_ratio::Float64; end
struct CSpacingMetrics; max_error::Float64; mean_error::Float64; std_error::Float64; end
struct CSpacingFidelity; mean_dnn_h::Float64; cv::Float64; p05::Float64; p50::Float64; p95::Float64; coordination::Float64; end

cforce(m::InverseDistanceForce) = CForce(0, m.β, 1.0, 3.0)
cforce(m::SpacingEquilibriumForce) = CForce(1, m.β, 1.0, 3.0)
cforce(m::ClippedSpacingForce) = CForce(2, m.β, m.u0, 3.0)
cforce(m::StrongSpacingForce) = CForce(3, m.β, 1.0, m.γ)
const BuiltinForce = Union{InverseDistanceForce, SpacingEquilibriumForce, ClippedSpacingForce, StrongSpacingForce}

# returns (CSpacing, keepalive): the spacing's own boundary set crosses as a dense D × n array in the cloud's T
cspacing(s::ConstantSpacing, ::Type{T}) where {T} = (CSpacing(0, ustrip(s.Δx), 0, 0, C_NULL, 0), nothing)
function cspacing(s::LogLike, ::Type{T}) where {T}
    b = collect(T, raw(s.boundary, machine_type(s.boundary)))
    return (CSpacing(1, ustrip(s.base_size), s.growth_rate, 0, pointer(b), length(s.boundary)), b)
end
function cspacing(s::BoundaryLayerSpacing, ::Type{T}) where {T}
    b = collect(T, raw(s.boundary, machine_type(s.boundary)))
    return (CSpacing(2, ustrip(s.at_wall), ustrip(s.bulk), ustrip(s.layer_thickness), pointer(b), length(s.boundary)), b)
end

# ------------------------------------------------------------ the ccalls, one literal per machine type
for (T, sfx) in ((Float32, "f32"), (Float64, "f64"))
    sym(name) = QuoteNode(Symbol(name, "_", sfx))
    @eval begin
        c_knn(p::AbstractVector{$T}, N, D, k, out, dist) = ccall(($(sym("wtp_knn")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, Int32, Ptr{Int64}, Ptr{Cvoid}), ctx(), p, N, D, k, out, dist)
        c_knn_self(p::AbstractVector{$T}, N, D, k, out, dist) = ccall(($(sym("wtp_knn_self")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, Int32, Ptr{Int64}, Ptr{Cvoid}), ctx(), p, N, D, k, out, dist)
        c_radius_count(p::AbstractVector{$T}, N, D, r, offsets) = ccall(($(sym("wtp_radius_count")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, $T, Ptr{Int64}), ctx(), p, N, D, $T(r), offsets)
        c_repel(s::AbstractVector{$T}, n_fixed, n_move, D, sp, fm, prm, wall, conv::Vector{$T}, tr, res) = ccall(($(sym("wtp_repel")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Int32, Ref{CSpacing}, Ref{CForce}, Ref{CParams}, Ptr{Cvoid}, Ptr{$T}, Ptr{Cvoid}, Ref{CResult}),
            ctx(), s, n_fixed, n_move, D, sp, fm, prm, wall, conv, tr, res)
        c_spacing_eval(sp, p::AbstractVector{$T}, N, D, out::Vector{$T}) = ccall(($(sym("wtp_spacing_eval")), LIB), Int32,
            (Ptr{Cvoid}, Ref{CSpacing}, Ptr{$T}, Int64, Int32, Ptr{$T}), ctx(), sp, p, N, D, out)
        c_isinside(p::AbstractVector{$T}, N, D, bx, bn, ba, M, out) = ccall(($(sym("wtp_isinside")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, Ptr{$T}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{UInt8}, Ptr{Cvoid}), ctx(), p, N, D, bx, bn, ba, M, out, C_NULL)
        c_cull(p::AbstractVector{$T}, N, D, s::Vector{$T}, ratio, keep) = ccall(($(sym("wtp_cull_mask")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, Ptr{$T}, Float64, Ptr{UInt8}), ctx(), p, N, D, s, ratio, keep)
        c_metrics(p::AbstractVector{$T}, N, D, k, out) = ccall(($(sym("wtp_metrics")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, Int32, Ref{CCloudMetrics}), ctx(), p, N, D, k, out)
        c_spacing_metrics(p::AbstractVector{$T}, N, D, k, sp, out) = ccall(($(sym("wtp_spacing_metrics")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, Int32, Ref{CSpacing}, Ref{CSpacingMetrics}), ctx(), p, N, D, k, sp, out)
        c_spacing_fidelity(p::AbstractVector{$T}, N, D, k, cr, sp, out) = ccall(($(sym("wtp_spacing_fidelity")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, Int32, Float64, Ref{CSpacing}, Ref{CSpacingFidelity}), ctx(), p, N, D, k, cr, sp, out)
        c_normals(p::AbstractVector{$T}, N, D, k, out::Matrix{$T}) = ccall(($(sym("wtp_normals")), LIB), Int32,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int32, Int32, Ptr{$T}), ctx(), p, N, D, k, out)
    end
end
c_radius_fill(indices) = ccall((:wtp_radius_fill, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}), ctx(), indices)

# ---------------------------------------------------------------- topology (src/topology.jl:79-100)
function knn_table(points::AbstractVector{<:Point}, k::Int; include_self::Bool = false, dists::Bool = false)
    T, D, N = machine_type(points), dimension(points), length(points)
    out = Matrix{Int}(undef, k, N)
    dist = dists ? Matrix{T}(undef, k, N) : nothing
    p = raw(points, T)
    GC.@preserve p out dist begin
        dp = dists ? Ptr{Cvoid}(pointer(dist)) : C_NULL
        check(include_self ? c_knn_self(p, N, D, k, out, dp) : c_knn(p, N, D, k, out, dp))
    end
    return out, dist
end

function _build_knn_neighbors(points::AbstractVector{<:Point}, k::Int)
    out, _ = knn_table(points, k)
    rows = FlatRows(out)
    return length(points) >= FLAT_THRESHOLD ? rows : [Vector{Int}(r) for r in rows]
end

function _build_radius_neighbors(points::AbstractVector{<:Point}, radius)
    T, D, N = machine_type(points), dimension(points), length(points)
    r = ustrip(_get_radius(radius, points))                # BallSearch uses ustrip(radius) without unit conversion
    offsets = Vector{Int}(undef, N + 1)
    p = raw(points, T)
    GC.@preserve p offsets check(c_radius_count(p, N, D, r, offsets))
    indices = Vector{Int}(undef, offsets[end])
    GC.@preserve indices check(c_radius_fill(indices))
    rows = CSRRows(offsets, indices)
    return N >= FLAT_THRESHOLD ? rows : [Vector{Int}(r) for r in rows]
end

# search / searchdists over cloud types (src/neighbors.jl:9-21): the KNearestSearch object only contributes its k
# (its KD-tree was built by the caller and is not used); per-point results keep the reference's shapes.
# (Euclidean clouds only: strictly more specific than the reference's Union{PointCloud, PointBoundary, PointSurface}
# methods — the same signature would be a method overwrite, which precompilation refuses)
const CloudLike{N} = Union{PointCloud{𝔼{N}}, PointBoundary{𝔼{N}}, PointSurface{𝔼{N}}}
function search(cloud::CloudLike{N}, method::KNearestSearch) where {N}
    out, _ = knn_table(points(cloud), method.k; include_self = true)
    return [Vector{Int}(view(out, :, i)) for i in axes(out, 2)]
end
function searchdists(cloud::CloudLike{N}, method::KNearestSearch) where {N}
    pts = points(cloud)
    out, dist = knn_table(pts, method.k; include_self = true, dists = true)
    u = Unitful.unit(Meshes.to(first(pts))[1])
    return [(Vector{Int}(view(out, :, i)), view(dist, :, i) .* u) for i in axes(out, 2)]
end

# ------------------------------------------------------------------- repel (src/repel.jl)
function spacing_values(spacing::BuiltinSpacing, pts::AbstractVector{<:Point})
    T, D, N = machine_type(pts), dimension(pts), length(pts)
    sp, keep = cspacing(spacing, T)
    out = Vector{T}(undef, N)
    p = raw(pts, T)
    GC.@preserve p keep out check(c_spacing_eval(Ref(sp), p, N, D, out))
    return out
end

"isinside.(pts, Ref(cloud)) on the device: Green's function over the boundary elements in 3-D (src/isinside.jl:86-106), winding number in 2-D (:18-35, :72-74)."
function isinside_batch(pts::AbstractVector{<:Point}, cloud::Union{PointCloud, PointBoundary})
    T, D, N = machine_type(pts), dimension(pts), length(pts)
    N == 0 && return Bool[]
    bnd = boundary(cloud)
    bx = collect(T, raw(points(bnd), machine_type(points(bnd))))
    M = length(points(bnd))
    bn = D == 3 ? dense(normal(bnd), T, D) : nothing
    ba = D == 3 ? T[T(ustrip(a)) for a in area(bnd)] : nothing
    out = Vector{UInt8}(undef, N)
    p = raw(pts, T)
    GC.@preserve p bx bn ba out begin
        check(c_isinside(p, N, D, bx, D == 3 ? Ptr{Cvoid}(pointer(bn)) : C_NULL, D == 3 ? Ptr{Cvoid}(pointer(ba)) : C_NULL, M, out))
    end
    return out .!= 0
end

function _near_duplicate_keep_mask(pts::AbstractVector{<:Point}, spacings::AbstractVector{<:AbstractFloat}, ratio::Real)
    n = length(pts)
    keep = trues(n)
    (ratio <= 0 || n < 2) && return keep
    T, D = machine_type(pts), dimension(pts)
    s = collect(T, spacings)
    flags = Vector{UInt8}(undef, n)
    p = raw(pts, T)
    GC.@preserve p s flags check(c_cull(p, n, D, s, Float64(ratio), flags))
    keep .= flags .!= 0
    return keep
end

# repel(cloud, spacing; ...) for the spacings that cross the ABI: the reference's body (src/repel.jl:56-95) with the
# three per-point host loops batched on the device — the default α (spacing.(to(cloud)), :61), the survivor filter
# (filter(x -> isinside(x, cloud), p), :90) and the cull (:91-93). _relax! dispatches to the override below.
function repel(
        cloud::PointCloud{𝔼{N}, C}, spacing::BuiltinSpacing;
        β = 0.2, force_model::RepelForceModel = ClippedSpacingForce(β),
        α = nothing, α_min = nothing, k = 21, max_iters = 1000, tol = 1.0e-6, rebuild_every::Int = 1,
        cull_ratio::Real = 0.0, kick_after::Int = 0, stall_after::Int = 50, cv_target::Real = 0.0,
        convergence::Union{Nothing, AbstractVector{<:AbstractFloat}} = nothing,
        trace::Union{Nothing, AbstractVector{<:NamedTuple}} = nothing,
    ) where {N, C <: CRS}
    rebuild_every >= 1 || throw(ArgumentError("rebuild_every must be ≥ 1"))
    bnd_p = points(boundary(cloud))
    n_bnd = length(bnd_p)
    p = copy(volume(cloud).points)
    p_old = copy(p)
    snap = vcat(bnd_p, p)
    len_unit = Unitful.unit(Meshes.to(first(snap))[1])
    if isnothing(α)                                        # :61, as a length like the reference's default
        α = minimum(spacing_values(spacing, snap)) * len_unit / 20
    end
    isnothing(α_min) && (α_min = α / 100)
    conv = _relax!(
        p, p_old, snap, spacing, force_model, identity;
        n_fixed = n_bnd, n_protected = n_bnd, α_lo = ustrip(α_min), α_max = ustrip(α),
        k, max_iters, tol, rebuild_every, kick_after, stall_after, cv_target, trace,
    )
    isnothing(convergence) || append!(convergence, conv)
    survivors = p[isinside_batch(p, cloud)]
    if cull_ratio > 0 && !isempty(survivors)
        survivors = survivors[_cull(survivors, spacing, cull_ratio)]
    end
    return PointCloud(boundary(cloud), PointVolume(survivors), NoTopology())
end

# index.vertices / triangles / face / vertex / edge (src/octree/triangle_octree.jl:22-31) -> 9 x n vertex
# coordinates and 21 x n pseudonormals (face, vertex 1..3, edge 12, 13, 23; _feature_pseudonormal :322-337)
function flatten_index(index, ::Type{T}) where {T}
    n = length(index.triangles)
    tri = Matrix{T}(undef, 9, n)
    fn = Matrix{T}(undef, 21, n)
    for i in 1:n
        t = index.triangles[i]
        v1, v2, v3 = index.vertices[t[1]], index.vertices[t[2]], index.vertices[t[3]]
        f = index.face[i]
        tri[1:3, i] .= v1; tri[4:6, i] .= v2; tri[7:9, i] .= v3
        fn[1:3, i] .= f
        fn[4:6, i] .= get(index.vertex, v1, f); fn[7:9, i] .= get(index.vertex, v2, f); fn[10:12, i] .= get(index.vertex, v3, f)
        fn[13:15, i] .= get(index.edge, _edge_key(v1, v2), f)
        fn[16:18, i] .= get(index.edge, _edge_key(v1, v3), f)
        fn[19:21, i] .= get(index.edge, _edge_key(v2, v3), f)
    end
    return tri, fn
end

# The wall rule is recognised by what the closure captured, not by identity: repel(cloud, spacing) passes a fresh
# anonymous (id, xi, x_proposed) -> x_proposed (src/repel.jl:82, no captured variable), repel(cloud, spacing, octree)
# one over (is_bnd, escaped, tri_indices, octree, offset_dist, len_unit) (:158-160).
is_identity_wall(c) = c === identity || fieldcount(typeof(c)) == 0
is_octree_wall(c) = hasproperty(c, :octree) && hasproperty(c, :is_bnd) && hasproperty(c, :tri_indices)

# _relax! for what crosses the ABI: built-in spacing and force law, identity or octree wall. Anything else (a user
# spacing callable, a RepelForceModel subtype, another constrain closure) does not match this signature / falls through
# to the reference's method via invoke, i.e. runs on the CPU exactly as without this file.
function _relax!(
        p::AbstractVector{<:Point}, p_old, snap, spacing::BuiltinSpacing, force_model::BuiltinForce, constrain;
        n_fixed, n_protected, α_lo, α_max, k, max_iters, tol, rebuild_every,
        kick_after, trace, stall_after = 0, cv_target = 0.0, (deposit!) = nothing,
    )
    wall_mode = is_octree_wall(constrain)
    if !(wall_mode || is_identity_wall(constrain))
        return invoke(_relax!, Tuple{Any, Any, Any, Any, Any, Any}, p, p_old, snap, spacing, force_model, constrain;
                      n_fixed, n_protected, α_lo, α_max, k, max_iters, tol, rebuild_every, kick_after, trace, stall_after, cv_target, deposit!)
    end
    # deposit! is the closure of src/repel.jl:161-168 over (k, escaped, is_bnd, tri_indices, octree, spacing,
    # deposit_ratio, ...): only its ratio crosses the ABI, the library runs _deposit_escaped! itself after every sweep
    deposit_ratio = isnothing(deposit!) ? 0.0 : Float64(unbox(deposit!.deposit_ratio))
    T, D = machine_type(snap), dimension(snap)
    n_move = length(p)
    tri = fn = nothing
    is_bnd = UInt8[]; tri_idx = Int64[]; esc = UInt8[]
    wall = Ref{CWallMesh}()
    if wall_mode
        index = unbox(constrain.octree).index
        tri, fn = flatten_index(index, T)
        is_bnd = UInt8.(unbox(constrain.is_bnd)); tri_idx = zeros(Int64, n_move); esc = zeros(UInt8, n_move)
        wall[] = CWallMesh(pointer(tri), pointer(fn), size(tri, 2), Tuple(Float64.(index.bbox_min)), Tuple(Float64.(index.bbox_max)),
                           Float64(unbox(constrain.offset_dist)), pointer(is_bnd), pointer(tri_idx), pointer(esc))
    end
    @views snap[(n_fixed + 1):end] .= p
    sp, keep = cspacing(spacing, T)
    fm = cforce(force_model)
    # the seed of the library's random directions (kicks, coincident points): drawn from a COPY of the task's RNG, so a
    # run is reproducible under Random.seed! without advancing the caller's stream
    seed = rand(copy(Random.default_rng()), UInt64)
    prm = CParams(k, max_iters, rebuild_every, stall_after, kick_after, wall_mode ? 1 : 0, isnothing(trace) ? 0 : 1, 0,
                  Float64(α_lo), Float64(α_max), Float64(tol), Float64(cv_target), Int64(n_protected), seed, deposit_ratio)
    conv = Vector{T}(undef, max(max_iters, 1))
    tr = isnothing(trace) ? CTrace[] : Vector{CTrace}(undef, max(max_iters, 1))
    res = Ref(CResult(0, 0, NaN))
    s = raw(snap, T)
    s === reinterpret(T, snap) || error("libwtp_cuda: the snapshot must be a dense Vector of points")   # updated in place
    GC.@preserve s keep conv tr tri fn is_bnd tri_idx esc wall begin
        check(c_repel(s, n_fixed, n_move, D, Ref(sp), Ref(fm), Ref(prm),
                      wall_mode ? Base.unsafe_convert(Ptr{Cvoid}, wall) : C_NULL, conv,
                      isnothing(trace) ? C_NULL : Ptr{Cvoid}(pointer(tr)), res))
    end
    if wall_mode                                        # side arrays written from inside the sweep (src/repel.jl:462,467)
        unbox(constrain.tri_indices) .= tri_idx
        if deposit_ratio > 0                            # deposition clears the flags it has seen and converts volume points (:493, :509)
            unbox(constrain.escaped) .= (esc .!= 0)
            unbox(constrain.is_bnd) .= (is_bnd .!= 0)
        else
            unbox(constrain.escaped) .|= (esc .!= 0)
        end
    end
    @views p .= snap[(n_fixed + 1):end]                 # final positions (pre-sweep ones on a cv_target stop)
    r = res[]
    if !isnothing(trace)
        for i in 1:r.iters
            t = tr[i]
            push!(trace, (; iteration = i, r = T(t.r), s = T(t.s), r_over_s = T(t.r_over_s), idx_a = Int(t.idx_a), idx_b = Int(t.idx_b)))
        end
    end
    # the reference's log lines (src/repel.jl:315,323,330,336)
    if r.stop_reason == 2
        @info "Node repel stopped in $(r.iters) iterations: spacing CV target reached" cv = r.last_cv cv_target
    elseif r.stop_reason == 3
        @info "Node repel stopped in $(r.iters) iterations: spacing CV stalled for $stall_after iterations" cv = r.last_cv convergence = conv[r.iters]
    elseif r.stop_reason == 1
        @info "Node repel finished in $(r.iters) iterations" convergence = conv[r.iters]
    elseif max_iters > 0
        @warn "Node repel reached maximum iterations" max_iters convergence = conv[r.iters]
    end
    return conv[1:r.iters]
end

# ------------------------------------------------------------------- metrics (src/metrics.jl)
function metrics(cloud::PointCloud{𝔼{N}, C}; k = 20) where {N, C <: CRS}
    pts = points(cloud)
    T, D, n = machine_type(pts), dimension(pts), length(pts)
    u = Unitful.unit(Meshes.to(first(pts))[1])
    m = Ref(CCloudMetrics(0, 0, 0, 0, 0, 0, 0))
    p = raw(pts, T)
    GC.@preserve p check(c_metrics(p, n, D, k, m))
    avg, σ, mx, mn = T(m[].avg) * u, T(m[].std) * u, T(m[].max) * u, T(m[].min) * u
    separation, fill, mesh_ratio = T(m[].separation) * u, T(m[].fill) * u, T(m[].mesh_ratio)
    println("Cloud Metrics")                           # the reference's report, src/metrics.jl:31-39
    println("-------------")
    println("avg. distance to $k nearest neighbors: $avg")
    println("std. distance to $k nearest neighbors: $σ")
    println("max. distance to $k nearest neighbors: $mx")
    println("min. distance to $k nearest neighbors: $mn")
    println("separation (min nearest-neighbor distance): $separation")
    println("fill (max nearest-neighbor distance):       $fill")
    println("mesh ratio (fill / separation, ≥1):         $mesh_ratio")
    return (; avg, std = σ, max = mx, min = mn, separation, fill, mesh_ratio, k)
end

function spacing_metrics(cloud::PointCloud{𝔼{N}, C}, spacing::BuiltinSpacing; k = 20) where {N, C <: CRS}
    pts = points(cloud)
    T, D, n = machine_type(pts), dimension(pts), length(pts)
    sp, keep = cspacing(spacing, T)
    m = Ref(CSpacingMetrics(0, 0, 0))
    p = raw(pts, T)
    GC.@preserve p keep check(c_spacing_metrics(p, n, D, k, Ref(sp), m))
    return (; max_error = T(m[].max_error), mean_error = T(m[].mean_error), std_error = T(m[].std_error), k)
end

function spacing_fidelity_metrics(cloud::PointCloud{𝔼{N}, C}, spacing::BuiltinSpacing; k = 30, coord_radius = 1.4) where {N, C <: CRS}
    pts = points(cloud)
    T, D, n = machine_type(pts), dimension(pts), length(pts)
    k = min(n, k)
    sp, keep = cspacing(spacing, T)
    m = Ref(CSpacingFidelity(0, 0, 0, 0, 0, 0))
    p = raw(pts, T)
    GC.@preserve p keep check(c_spacing_fidelity(p, n, D, k, Float64(coord_radius), Ref(sp), m))
    return (; mean_dnn_h = T(m[].mean_dnn_h), cv = T(m[].cv), p05 = T(m[].p05), p50 = T(m[].p50), p95 = T(m[].p95),
            coordination = m[].coordination, k, coord_radius)
end

# ------------------------------------------------------------------- normals (src/normals.jl:9-44)
# compute_normals(points; k): per point the eigenvector of the smallest eigenvalue of the covariance of its k nearest
# points (self included). Unoriented, like the reference's: orient_normals! (a serial minimum-spanning-tree walk,
# src/normals.jl:75-117) runs on the host over this output.
# (Vector, not AbstractVector: strictly more specific than the reference's method, src/normals.jl:15 — never an overwrite)
function compute_normals(pts::Vector{<:Point{𝔼{D}}}; k::Int = 5) where {D}
    T, n = machine_type(pts), length(pts)
    k = k > n ? n : k
    out = Matrix{T}(undef, D, n)
    p = raw(pts, T)
    GC.@preserve p out check(c_normals(p, n, D, k, out))
    return [SVector{D, T}(view(out, :, i)) for i in 1:n]
end

end # module
