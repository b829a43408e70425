# WTPCuda.jl — the Julia side of the drop-in boundary (cannot be executed in the build image:
# Julia is not installed; the same byte-level calls are exercised through ctypes by tests/).
#
# Include this file from src/WhatsThePoint.jl after `include("repel.jl")`:
#
#     include("WTPCuda.jl")          # adds module WTPCuda and the three method overrides
#
# It overrides the only three call sites of the third-party KD-tree on the hot path
#   _build_knn_neighbors     (src/topology.jl:79-84)
#   _build_radius_neighbors  (src/topology.jl:91-97)
#   _relax!                  (src/repel.jl:202-339)
# with `ccall`s into libwtp_cuda.so (include/wtp_cuda.h). No CUDA.jl, no KernelAbstractions,
# no CPU fallback: if the library or the GPU is missing the calls throw.

module WTPCuda

using Meshes, Unitful, StaticArrays
import ..WhatsThePoint: _build_knn_neighbors, _build_radius_neighbors, _relax!, _get_radius, _edge_key,
    RepelForceModel, InverseDistanceForce, SpacingEquilibriumForce, ClippedSpacingForce, StrongSpacingForce,
    ConstantSpacing, LogLike, BoundaryLayerSpacing

const LIB = get(ENV, "WTP_CUDA_LIB", "libwtp_cuda.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function ctx()
    if CTX[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:wtp_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Int32), h, parse(Int32, get(ENV, "WTP_CUDA_DEVICE", "0")))
        rc == 0 || error("libwtp_cuda: wtp_create failed with status $rc (no usable B200; there is no CPU fallback)")
        CTX[] = h[]
        atexit(() -> ccall((:wtp_destroy, LIB), Cvoid, (Ptr{Cvoid},), CTX[]))
    end
    return CTX[]
end

function check(rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:wtp_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx()))
    rc in (1, 2) ? throw(ArgumentError(msg)) : error("libwtp_cuda: $msg")   # WTP_ERR_BAD_ARG / K_TOO_LARGE -> ArgumentError
end

# Vector{Point{𝔼{D},Cartesian{…,D,Quantity{T}}}} is an isbits AoS of D contiguous T: zero-copy view.
machine_type(pts) = typeof(ustrip(Meshes.to(first(pts))[1]))
dimension(pts) = length(Meshes.to(first(pts)))
raw(pts, ::Type{T}) where {T} = reinterpret(T, pts)

sfx(::Type{Float32}) = "f32"
sfx(::Type{Float64}) = "f64"

# ---------------------------------------------------------------- topology
# Rows of one N×k buffer wrapped as Vector{Vector{Int}} (test/topology.jl:28 asserts that type);
# for 10⁷+ points pass `flat = true` to keep the matrix and use KNNTopology{FlatRows}.
function _build_knn_neighbors(points::AbstractVector{<:Point}, k::Int)
    T, D, N = machine_type(points), dimension(points), length(points)
    out = Matrix{Int}(undef, k, N)                    # column i = neighbours of point i (row-major N×k for C)
    p = raw(points, T)
    GC.@preserve p out begin
        rc = T === Float32 ?
            ccall((:wtp_knn_f32, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64, Int32, Int32, Ptr{Int64}, Ptr{Float32}),
                  ctx(), p, N, D, k, out, C_NULL) :
            ccall((:wtp_knn_f64, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int32, Ptr{Int64}, Ptr{Float64}),
                  ctx(), p, N, D, k, out, C_NULL)
        check(rc)
    end
    return [out[:, i] for i in 1:N]
end

function _build_radius_neighbors(points::AbstractVector{<:Point}, radius)
    T, D, N = machine_type(points), dimension(points), length(points)
    r = T(ustrip(_get_radius(radius, points)))        # BallSearch uses ustrip(radius) without unit conversion
    offsets = Vector{Int}(undef, N + 1)
    p = raw(points, T)
    GC.@preserve p offsets begin
        rc = T === Float32 ?
            ccall((:wtp_radius_count_f32, LIB), Int32, (Ptr{Cvoid}, Ptr{Float32}, Int64, Int32, Float32, Ptr{Int64}), ctx(), p, N, D, r, offsets) :
            ccall((:wtp_radius_count_f64, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Float64, Ptr{Int64}), ctx(), p, N, D, r, offsets)
        check(rc)
    end
    indices = Vector{Int}(undef, offsets[end])
    GC.@preserve indices check(ccall((:wtp_radius_fill, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}), ctx(), indices))
    return [indices[(offsets[i] + 1):offsets[i + 1]] for i in 1:N]
end

# ------------------------------------------------------------------- repel
struct CForce; kind::Int32; beta::Float64; u0::Float64; gamma::Float64; end
struct CSpacing; kind::Int32; a::Float64; b::Float64; c::Float64; bnd::Ptr{Cvoid}; n_bnd::Int64; end
struct CParams
    k::Int32; max_iters::Int32; rebuild_every::Int32; stall_after::Int32; kick_after::Int32; wall::Int32
    want_trace::Int32; reserved::Int32; alpha_lo::Float64; alpha_max::Float64; tol::Float64; cv_target::Float64
    n_protected::Int64; kick_seed::UInt64; deposit_ratio::Float64
end
struct CResult; iters::Int32; stop_reason::Int32; last_cv::Float64; end
struct CTrace; r::Float64; s::Float64; r_over_s::Float64; idx_a::Int64; idx_b::Int64; end

cforce(m::InverseDistanceForce) = CForce(0, m.β, 1.0, 3.0)
cforce(m::SpacingEquilibriumForce) = CForce(1, m.β, 1.0, 3.0)
cforce(m::ClippedSpacingForce) = CForce(2, m.β, m.u0, 3.0)
cforce(m::StrongSpacingForce) = CForce(3, m.β, 1.0, m.γ)
cforce(m::RepelForceModel) = error("libwtp_cuda: user-defined RepelForceModel $(typeof(m)) cannot cross the C ABI (no CPU fallback)")

# returns (CSpacing, keepalive)
cspacing(s::ConstantSpacing, ::Type{T}) where {T} = (CSpacing(0, ustrip(s.Δx), 0, 0, C_NULL, 0), nothing)
function cspacing(s::LogLike, ::Type{T}) where {T}
    b = collect(raw(s.boundary, machine_type(s.boundary)) .|> T)
    return (CSpacing(1, ustrip(s.base_size), s.growth_rate, 0, pointer(b), length(s.boundary)), b)
end
function cspacing(s::BoundaryLayerSpacing, ::Type{T}) where {T}
    b = collect(raw(s.boundary, machine_type(s.boundary)) .|> T)
    return (CSpacing(2, ustrip(s.at_wall), ustrip(s.bulk), ustrip(s.layer_thickness), pointer(b), length(s.boundary)), b)
end
cspacing(s, ::Type) = error("libwtp_cuda: spacing callable $(typeof(s)) cannot cross the C ABI (no CPU fallback)")

const IDENTITY_WALL = 0
const MESH_WALL = 1

# wtp_wall_mesh (include/wtp_cuda.h): the TriangleIndex arrays flattened per triangle, in the cloud's machine type
struct CWallMesh
    triangles::Ptr{Cvoid}; feature_normals::Ptr{Cvoid}; n_tri::Int64
    bbox_min::NTuple{3, Float64}; bbox_max::NTuple{3, Float64}; offset_dist::Float64
    is_bnd::Ptr{UInt8}; tri_indices::Ptr{Int64}; escaped::Ptr{UInt8}
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

function _relax!(
        p, p_old, snap, spacing, force_model, constrain;
        n_fixed, n_protected, α_lo, α_max, k, max_iters, tol, rebuild_every,
        kick_after, trace, stall_after = 0, cv_target = 0.0, (deposit!) = nothing,
    )
    # deposit! is the closure of src/repel.jl:161-168 over (escaped, is_bnd, tri_indices, octree, spacing, deposit_ratio, ...):
    # only its ratio crosses the ABI, the library runs _deposit_escaped! itself after every sweep
    deposit_ratio = isnothing(deposit!) ? 0.0 : Float64(deposit!.deposit_ratio)
    T, D = machine_type(snap), dimension(snap)
    n_move = length(p)
    # repel(cloud, spacing, octree) calls with n_fixed = 0, n_protected = n_boundary (src/repel.jl:172) and a
    # constrain closure over (is_bnd, escaped, tri_indices, octree, offset_dist, len_unit) (:158-160): the octree
    # wall rule. Its captured variables are the closure's fields; the mesh crosses the ABI as flat arrays in T
    # (the reference queries the octree in the octree's own machine type, :453 — identical when the two agree).
    wall_mode = (n_fixed == 0 && n_protected > 0)
    wall_mode || constrain === identity || error("libwtp_cuda: only identity and the octree wall rule can cross the C ABI as `constrain`")
    tri = fn = nothing
    is_bnd = UInt8[]; tri_idx = Int64[]; esc = UInt8[]
    wall = Ref{CWallMesh}()
    if wall_mode
        index = constrain.octree.index
        tri, fn = flatten_index(index, T)
        is_bnd = UInt8.(constrain.is_bnd); tri_idx = zeros(Int64, n_move); esc = zeros(UInt8, n_move)
        wall[] = CWallMesh(pointer(tri), pointer(fn), size(tri, 2), Tuple(Float64.(index.bbox_min)), Tuple(Float64.(index.bbox_max)),
                           Float64(constrain.offset_dist), pointer(is_bnd), pointer(tri_idx), pointer(esc))
    end
    @views snap[(n_fixed + 1):end] .= p
    sp, keep = cspacing(spacing, T)
    fm = cforce(force_model)
    prm = CParams(k, max_iters, rebuild_every, stall_after, kick_after, wall_mode ? MESH_WALL : IDENTITY_WALL, isnothing(trace) ? 0 : 1, 0,
                  Float64(α_lo), Float64(α_max), Float64(tol), Float64(cv_target), Int64(n_protected), rand(UInt64), deposit_ratio)
    conv = Vector{T}(undef, max(max_iters, 1))
    tr = isnothing(trace) ? CTrace[] : Vector{CTrace}(undef, max(max_iters, 1))
    res = Ref(CResult(0, 0, NaN))
    s = raw(snap, T)
    GC.@preserve s keep conv tr tri fn is_bnd tri_idx esc wall begin
        rc = ccall((T === Float32 ? :wtp_repel_f32 : :wtp_repel_f64, LIB), Int32,
                   (Ptr{Cvoid}, Ptr{T}, Int64, Int64, Int32, Ref{CSpacing}, Ref{CForce}, Ref{CParams}, Ptr{Cvoid}, Ptr{T}, Ptr{CTrace}, Ref{CResult}),
                   ctx(), s, n_fixed, n_move, D, Ref(sp), Ref(fm), Ref(prm),
                   wall_mode ? Base.unsafe_convert(Ptr{Cvoid}, wall) : C_NULL, conv, isnothing(trace) ? C_NULL : pointer(tr), res)
        check(rc)
    end
    if wall_mode                                        # side arrays written from inside the sweep (src/repel.jl:462,467)
        constrain.tri_indices .= tri_idx
        if deposit_ratio > 0                            # deposition clears the flags it has seen and converts volume points (:493, :509)
            constrain.escaped .= (esc .!= 0)
            constrain.is_bnd .= (is_bnd .!= 0)
        else
            constrain.escaped .|= (esc .!= 0)
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

end # module
