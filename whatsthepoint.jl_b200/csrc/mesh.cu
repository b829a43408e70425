// mesh.cu — triangle-mesh queries of the 3-argument repel's wall rule.
//
// Replaces, for the hot path only, what the reference answers with its TriangleOctree:
//   _project_to_boundary             src/repel.jl:522-537
//   isinside(::SVector, octree)      src/octree/triangle_octree.jl:71-99
//   _compute_signed_distance_octree  src/octree/triangle_octree.jl:583-607
//   _nearest_element_tree!           src/octree/spatial_octree.jl:283-325
//   closest_point_on_triangle_feature src/octree/geometric_utils.jl:68-136
//   _constrain_octree                src/repel.jl:448-469
//
// The octree itself (host-side construction) is not rebuilt: the
// flattened TriangleIndex arrays arrive through the C ABI and the device gets its own search
// structure — triangles sorted by the Morton code of their centroid, 4-triangle leaves, a
// complete binary tree of boxes in heap layout. One thread answers one query with an explicit
// stack; the nearest triangle is the canonical one, smallest (d2, triangle index), so results
// do not depend on traversal order. Every box bound carries an absolute slack (a few dozen
// ulps of the coordinate magnitude) because the closest-point arithmetic is not monotone.
#include <cfloat>
#include <cmath>

#include "kernels.cuh"
#include "knn_core.cuh"

namespace wtp {

#define LAUNCH_CHECK(ctx)                         \
    do {                                          \
        (ctx)->launches++;                        \
        WTP_CUDA_CHECK(cudaPeekAtLastError());    \
    } while (0)

constexpr int MESH_LEAF = 4;

// sorted triangle: three 16/32-byte vertex records, the first carries the caller's index
template <class T>
struct TriRec { P4<T> v[3]; };

template <class T>
struct MeshView {
    const TriRec<T>* tris;
    const Box<T>* boxes;
    const T* fnorm;        // caller order, 21 per triangle: face, vertex 1..3, edge 12, 13, 23
    int64_t n, leaf_pow2;
    T lo[3], hi[3];        // index.bbox_min / bbox_max
    T offset;              // offset_dist
    T maxabs;              // largest |coordinate| of the mesh box
};

template <class T>
static MeshView<T> mesh_view(const MeshBuffers& mb) {
    MeshView<T> v;
    v.tris = mb.tris.get<TriRec<T>>(); v.boxes = mb.boxes.get<Box<T>>(); v.fnorm = mb.fnorm.get<T>();
    v.n = mb.n; v.leaf_pow2 = mb.leaf_pow2;
    for (int d = 0; d < 3; ++d) { v.lo[d] = (T)mb.lo[d]; v.hi[d] = (T)mb.hi[d]; }
    v.offset = (T)mb.offset; v.maxabs = (T)mb.maxabs;
    return v;
}

// ------------------------------------------------------------------ build
template <class T>
__global__ void __launch_bounds__(256) tri_key_kernel(const T* __restrict__ raw, int64_t n, T lox, T loy, T loz, T sx, T sy, T sz,
                                                      uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* t = raw + i * 9;
    auto quant = [](T c, T lo, T s) {
        int q = (int)((c - lo) * s);
        return (uint32_t)(q < 0 ? 0 : (q > 1023 ? 1023 : q));
    };
    const T cx = (t[0] + t[3] + t[6]) / (T)3, cy = (t[1] + t[4] + t[7]) / (T)3, cz = (t[2] + t[5] + t[8]) / (T)3;
    keys[i] = spread3(quant(cx, lox, sx)) | (spread3(quant(cy, loy, sy)) << 1) | (spread3(quant(cz, loz, sz)) << 2);
    vals[i] = (uint32_t)i;
}

template <class T>
__global__ void __launch_bounds__(256) tri_gather_kernel(const T* __restrict__ raw, const uint32_t* __restrict__ vals, int64_t n,
                                                         TriRec<T>* __restrict__ tris) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = vals[j];
    const T* t = raw + (size_t)i * 9;
    TriRec<T> r;
#pragma unroll
    for (int v = 0; v < 3; ++v) { r.v[v].x = t[3 * v]; r.v[v].y = t[3 * v + 1]; r.v[v].z = t[3 * v + 2]; r.v[v].w = idx_bits((T)0, v == 0 ? i : 0u); }
    tris[j] = r;
}

template <class T>
__global__ void __launch_bounds__(256) tri_leaf_kernel(const TriRec<T>* __restrict__ tris, int64_t n, int64_t leaf_pow2, Box<T>* __restrict__ boxes) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= leaf_pow2) return;
    Box<T> bx;
    for (int d = 0; d < 3; ++d) { bx.lo[d] = t_inf<T>(); bx.hi[d] = -t_inf<T>(); }
    const int64_t j0 = b * MESH_LEAF, j1 = j0 + MESH_LEAF < n ? j0 + MESH_LEAF : n;
    for (int64_t j = j0; j < j1; ++j)
        for (int v = 0; v < 3; ++v) {
            const P4<T> p = tris[j].v[v];
            bx.lo[0] = p.x < bx.lo[0] ? p.x : bx.lo[0]; bx.hi[0] = p.x > bx.hi[0] ? p.x : bx.hi[0];
            bx.lo[1] = p.y < bx.lo[1] ? p.y : bx.lo[1]; bx.hi[1] = p.y > bx.hi[1] ? p.y : bx.hi[1];
            bx.lo[2] = p.z < bx.lo[2] ? p.z : bx.lo[2]; bx.hi[2] = p.z > bx.hi[2] ? p.z : bx.hi[2];
        }
    boxes[leaf_pow2 + b] = bx;
}

template <class T>
void mesh_build(wtp_ctx* ctx, MeshBuffers& mb, const wtp_wall_mesh* wall) {
    WTP_REQUIRE(wall && wall->triangles && wall->feature_normals && wall->n_tri > 0, WTP_ERR_BAD_ARG,
                "mesh wall needs triangles and feature normals (Mesh must contain at least one triangle)");
    const int64_t n = wall->n_tri;
    WTP_REQUIRE(n < (int64_t)0x7ffffff0, WTP_ERR_BAD_ARG, "too many triangles");
    cudaStream_t st = ctx->stream;
    T* raw = mb.raw.as<T>((size_t)n * 9);
    T* fn = mb.fnorm.as<T>((size_t)n * 21);
    WTP_CUDA_CHECK(cudaMemcpyAsync(raw, wall->triangles, (size_t)n * 9 * sizeof(T), cudaMemcpyHostToDevice, st));
    WTP_CUDA_CHECK(cudaMemcpyAsync(fn, wall->feature_normals, (size_t)n * 21 * sizeof(T), cudaMemcpyHostToDevice, st));
    T l[3], s[3];
    double maxabs = 0;
    for (int d = 0; d < 3; ++d) {
        mb.lo[d] = wall->bbox_min[d]; mb.hi[d] = wall->bbox_max[d];
        WTP_REQUIRE(std::isfinite(mb.lo[d]) && std::isfinite(mb.hi[d]) && mb.hi[d] >= mb.lo[d], WTP_ERR_BAD_ARG, "mesh bounding box is not finite");
        const double ext = mb.hi[d] - mb.lo[d];
        l[d] = (T)mb.lo[d];
        s[d] = (T)(ext > 0 ? 1024.0 / ext * (1.0 - 1e-6) : 0.0);
        maxabs = std::max(maxabs, std::max(std::fabs(mb.lo[d]), std::fabs(mb.hi[d])));
    }
    mb.maxabs = maxabs;
    mb.offset = wall->offset_dist;
    uint32_t* keys = mb.ib.keys_a.as<uint32_t>((size_t)n);
    uint32_t* vals = mb.ib.vals_a.as<uint32_t>((size_t)n);
    const unsigned nb = (unsigned)((n + 255) / 256);
    tri_key_kernel<T><<<nb, 256, 0, st>>>(raw, n, l[0], l[1], l[2], s[0], s[1], s[2], keys, vals);
    LAUNCH_CHECK(ctx);
    radix_sort_pairs(ctx, mb.ib, n, 30);
    TriRec<T>* tris = mb.tris.as<TriRec<T>>((size_t)n);
    tri_gather_kernel<T><<<nb, 256, 0, st>>>(raw, mb.ib.vals_a.get<uint32_t>(), n, tris);
    LAUNCH_CHECK(ctx);
    const int64_t nleaf = (n + MESH_LEAF - 1) / MESH_LEAF;
    int64_t P = 1;
    while (P < nleaf) P <<= 1;
    Box<T>* boxes = mb.boxes.as<Box<T>>((size_t)2 * P);
    tri_leaf_kernel<T><<<(unsigned)((P + 255) / 256), 256, 0, st>>>(tris, n, P, boxes);
    LAUNCH_CHECK(ctx);
    bvh_build_levels<T>(ctx, boxes, P);
    mb.n = n;
    mb.leaf_pow2 = P;
}
template void mesh_build<float>(wtp_ctx*, MeshBuffers&, const wtp_wall_mesh*);
template void mesh_build<double>(wtp_ctx*, MeshBuffers&, const wtp_wall_mesh*);

// ----------------------------------------------------------- closest point
// Ericson's region walk (src/octree/geometric_utils.jl:68-136); feature code: 0 face,
// 1..3 vertex, 4 edge12, 5 edge13, 6 edge23. Sums are left-associated, no FMA (-fmad=false).
template <class T>
__device__ __forceinline__ T dot3(T ax, T ay, T az, T bx, T by, T bz) { return (ax * bx + ay * by) + az * bz; }

template <class T>
__device__ __forceinline__ int closest_on_triangle(T px, T py, T pz, const P4<T>& a, const P4<T>& b, const P4<T>& c, T& ox, T& oy, T& oz) {
    const T abx = b.x - a.x, aby = b.y - a.y, abz = b.z - a.z;
    const T acx = c.x - a.x, acy = c.y - a.y, acz = c.z - a.z;
    const T apx = px - a.x, apy = py - a.y, apz = pz - a.z;
    const T d1 = dot3(abx, aby, abz, apx, apy, apz), d2 = dot3(acx, acy, acz, apx, apy, apz);
    if (d1 <= (T)0 && d2 <= (T)0) { ox = a.x; oy = a.y; oz = a.z; return 1; }
    const T bpx = px - b.x, bpy = py - b.y, bpz = pz - b.z;
    const T d3 = dot3(abx, aby, abz, bpx, bpy, bpz), d4 = dot3(acx, acy, acz, bpx, bpy, bpz);
    if (d3 >= (T)0 && d4 <= d3) { ox = b.x; oy = b.y; oz = b.z; return 2; }
    const T vc = d1 * d4 - d3 * d2;
    if (vc <= (T)0 && d1 >= (T)0 && d3 <= (T)0) {
        const T v = d1 / (d1 - d3);
        ox = a.x + v * abx; oy = a.y + v * aby; oz = a.z + v * abz;
        return 4;
    }
    const T cpx = px - c.x, cpy = py - c.y, cpz = pz - c.z;
    const T d5 = dot3(abx, aby, abz, cpx, cpy, cpz), d6 = dot3(acx, acy, acz, cpx, cpy, cpz);
    if (d6 >= (T)0 && d5 <= d6) { ox = c.x; oy = c.y; oz = c.z; return 3; }
    const T vb = d5 * d2 - d1 * d6;
    if (vb <= (T)0 && d2 >= (T)0 && d6 <= (T)0) {
        const T w = d2 / (d2 - d6);
        ox = a.x + w * acx; oy = a.y + w * acy; oz = a.z + w * acz;
        return 5;
    }
    const T va = d3 * d6 - d5 * d4;
    if (va <= (T)0 && (d4 - d3) >= (T)0 && (d5 - d6) >= (T)0) {
        const T w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        ox = b.x + w * (c.x - b.x); oy = b.y + w * (c.y - b.y); oz = b.z + w * (c.z - b.z);
        return 6;
    }
    const T denom = (T)1 / (va + vb + vc);
    const T v = vb * denom, w = vc * denom;
    ox = a.x + abx * v + acx * w; oy = a.y + aby * v + acy * w; oz = a.z + abz * v + acz * w;
    return 0;
}

template <class T>
struct NearestTri {
    T d2, cx, cy, cz;
    uint32_t tri;     // caller's 0-based triangle index (0xffffffff: none)
    uint32_t pos;     // sorted position
    int feature;
};

template <class T>
__device__ __forceinline__ void offer_triangle(const MeshView<T>& m, int64_t j, T px, T py, T pz, NearestTri<T>& best) {
    const P4<T> a = load_p4<T>(&m.tris[j].v[0]), b = load_p4<T>(&m.tris[j].v[1]), c = load_p4<T>(&m.tris[j].v[2]);
    T cx, cy, cz;
    const int f = closest_on_triangle<T>(px, py, pz, a, b, c, cx, cy, cz);
    const T dx = px - cx, dy = py - cy, dz = pz - cz;
    const T d2 = dot3(dx, dy, dz, dx, dy, dz);
    const uint32_t t = idx_of(a);
    if (d2 < best.d2 || (d2 == best.d2 && t < best.tri)) { best.d2 = d2; best.cx = cx; best.cy = cy; best.cz = cz; best.tri = t; best.pos = (uint32_t)j; best.feature = f; }
}

// Exact nearest triangle by (d2, triangle index). `hint`: sorted position to seed the bound with
// (0xffffffff: none); the answer does not depend on it.
template <class T>
__device__ __forceinline__ NearestTri<T> mesh_nearest(const MeshView<T>& m, T px, T py, T pz, uint32_t hint) {
    NearestTri<T> best;
    best.d2 = t_inf<T>(); best.cx = px; best.cy = py; best.cz = pz; best.tri = 0xffffffffu; best.pos = 0xffffffffu; best.feature = 0;
    if (hint < (uint32_t)m.n) offer_triangle<T>(m, hint, px, py, pz, best);
    // closest-point rounding: errors scale with the larger of the mesh and query magnitudes
    const T eps = sizeof(T) == 4 ? (T)1.1920929e-7 : (T)2.220446049250313e-16;
    T mag = m.maxabs;
    mag = fmax(mag, fmax(fabs(px), fmax(fabs(py), fabs(pz))));
    const T slack = (T)64 * eps * mag;
    auto box_lb = [&](int64_t i) -> T {
        const Box<T> b = m.boxes[i];
        if (b.lo[0] > b.hi[0]) return t_inf<T>();   // empty node
        T gx = px < b.lo[0] ? b.lo[0] - px : (px > b.hi[0] ? px - b.hi[0] : (T)0);
        T gy = py < b.lo[1] ? b.lo[1] - py : (py > b.hi[1] ? py - b.hi[1] : (T)0);
        T gz = pz < b.lo[2] ? b.lo[2] - pz : (pz > b.hi[2] ? pz - b.hi[2] : (T)0);
        gx = gx > slack ? gx - slack : (T)0; gy = gy > slack ? gy - slack : (T)0; gz = gz > slack ? gz - slack : (T)0;
        const T s = dot3(gx, gy, gz, gx, gy, gz);
        return s - (T)8 * eps * s;
    };
    int stack[64];
    int sp = 0;
    stack[sp++] = 1;
    while (sp > 0) {
        const int64_t i = stack[--sp];
        if (i >= m.leaf_pow2) {
            const int64_t j0 = (i - m.leaf_pow2) * MESH_LEAF, j1 = j0 + MESH_LEAF < m.n ? j0 + MESH_LEAF : m.n;
            for (int64_t j = j0; j < j1; ++j) offer_triangle<T>(m, j, px, py, pz, best);
            continue;
        }
        const T ll = box_lb(2 * i), lr = box_lb(2 * i + 1);
        // a tie on d2 can still win on the triangle index: prune only strictly farther boxes
        if (ll <= lr) {
            if (lr <= best.d2) stack[sp++] = (int)(2 * i + 1);
            if (ll <= best.d2) stack[sp++] = (int)(2 * i);
        } else {
            if (ll <= best.d2) stack[sp++] = (int)(2 * i);
            if (lr <= best.d2) stack[sp++] = (int)(2 * i + 1);
        }
    }
    return best;
}

// isinside(p, octree): mesh-box test, then sign of dot(p - cp, pseudonormal of the closest feature)
template <class T>
__device__ __forceinline__ bool mesh_inside(const MeshView<T>& m, T px, T py, T pz, uint32_t& hint) {
    if (px < m.lo[0] || px > m.hi[0] || py < m.lo[1] || py > m.hi[1] || pz < m.lo[2] || pz > m.hi[2]) return false;
    const NearestTri<T> nb = mesh_nearest<T>(m, px, py, pz, hint);
    if (nb.tri == 0xffffffffu) return false;
    hint = nb.pos;
    const T* n = m.fnorm + (size_t)nb.tri * 21 + nb.feature * 3;
    const T s = dot3(px - nb.cx, py - nb.cy, pz - nb.cz, n[0], n[1], n[2]);
    return s < (T)0;
}

// _project_to_boundary: closest point nudged inward along the face normal; 1-based triangle, 0: none
template <class T>
__device__ __forceinline__ int64_t mesh_project_point(const MeshView<T>& m, T px, T py, T pz, uint32_t& hint, T& ox, T& oy, T& oz) {
    const NearestTri<T> nb = mesh_nearest<T>(m, px, py, pz, hint);
    if (nb.tri == 0xffffffffu) { ox = px; oy = py; oz = pz; return 0; }
    hint = nb.pos;
    const T* n = m.fnorm + (size_t)nb.tri * 21;
    ox = nb.cx - m.offset * n[0]; oy = nb.cy - m.offset * n[1]; oz = nb.cz - m.offset * n[2];
    return (int64_t)nb.tri + 1;
}

// ----------------------------------------------------------------- kernels
template <class T>
__global__ void __launch_bounds__(128) mesh_isinside_kernel(const MeshView<T> m, const T* __restrict__ pts, int64_t n, uint8_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t hint = 0xffffffffu;
    out[i] = mesh_inside<T>(m, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], hint) ? 1 : 0;
}

template <class T>
__global__ void __launch_bounds__(128) mesh_project_kernel(const MeshView<T> m, const T* __restrict__ pts, int64_t n, T* __restrict__ out,
                                                           int64_t* __restrict__ out_tri) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t hint = 0xffffffffu;
    T ox, oy, oz;
    out_tri[i] = mesh_project_point<T>(m, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], hint, ox, oy, oz);
    out[i * 3] = ox; out[i * 3 + 1] = oy; out[i * 3 + 2] = oz;
}

// _constrain_octree (src/repel.jl:448-469), thread per movable point of this rank's range
template <class T>
__global__ void __launch_bounds__(128) mesh_wall_kernel(const MeshView<T> m, const T* __restrict__ P_old, T* __restrict__ P_new, int64_t id_lo,
                                                        int64_t id_hi, const uint8_t* __restrict__ is_bnd, int64_t* __restrict__ tri_idx,
                                                        uint8_t* __restrict__ escaped, uint32_t* __restrict__ hints) {
    const int64_t id = id_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= id_hi) return;
    const T px = P_new[id * 3], py = P_new[id * 3 + 1], pz = P_new[id * 3 + 2];
    uint32_t hint = hints[id];
    if (is_bnd[id]) {
        T ox, oy, oz;
        int64_t tri = mesh_project_point<T>(m, px, py, pz, hint, ox, oy, oz);
        if (tri == 0) tri = mesh_project_point<T>(m, P_old[id * 3], P_old[id * 3 + 1], P_old[id * 3 + 2], hint, ox, oy, oz);
        tri_idx[id] = tri;
        P_new[id * 3] = ox; P_new[id * 3 + 1] = oy; P_new[id * 3 + 2] = oz;
    } else if (!mesh_inside<T>(m, px, py, pz, hint)) {
        escaped[id] = 1;
        P_new[id * 3] = P_old[id * 3]; P_new[id * 3 + 1] = P_old[id * 3 + 1]; P_new[id * 3 + 2] = P_old[id * 3 + 2];
    }
    hints[id] = hint;
}

template <class T>
void mesh_isinside(wtp_ctx* ctx, const MeshBuffers& mb, const T* d_pts, int64_t n, uint8_t* d_out) {
    if (n <= 0) return;
    mesh_isinside_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(mesh_view<T>(mb), d_pts, n, d_out);
    LAUNCH_CHECK(ctx);
}
template void mesh_isinside<float>(wtp_ctx*, const MeshBuffers&, const float*, int64_t, uint8_t*);
template void mesh_isinside<double>(wtp_ctx*, const MeshBuffers&, const double*, int64_t, uint8_t*);

template <class T>
void mesh_project(wtp_ctx* ctx, const MeshBuffers& mb, const T* d_pts, int64_t n, T* d_out_pts, int64_t* d_out_tri) {
    if (n <= 0) return;
    mesh_project_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(mesh_view<T>(mb), d_pts, n, d_out_pts, d_out_tri);
    LAUNCH_CHECK(ctx);
}
template void mesh_project<float>(wtp_ctx*, const MeshBuffers&, const float*, int64_t, float*, int64_t*);
template void mesh_project<double>(wtp_ctx*, const MeshBuffers&, const double*, int64_t, double*, int64_t*);

template <class T>
void mesh_wall_apply(wtp_ctx* ctx, MeshBuffers& mb, const T* d_P_old, T* d_P_new, int64_t id_lo, int64_t id_hi) {
    if (id_hi <= id_lo) return;
    mesh_wall_kernel<T><<<(unsigned)((id_hi - id_lo + 127) / 128), 128, 0, ctx->stream>>>(
        mesh_view<T>(mb), d_P_old, d_P_new, id_lo, id_hi, mb.is_bnd.get<uint8_t>(), mb.tri_idx.get<int64_t>(), mb.escaped.get<uint8_t>(),
        mb.hint.get<uint32_t>());
    LAUNCH_CHECK(ctx);
}
template void mesh_wall_apply<float>(wtp_ctx*, MeshBuffers&, const float*, float*, int64_t, int64_t);
template void mesh_wall_apply<double>(wtp_ctx*, MeshBuffers&, const double*, double*, int64_t, int64_t);

}  // namespace wtp
