// wtp_oracle.cpp — CPU restatement of the WhatsThePoint.jl hot path.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product path (libwtp_cuda.so, the
// whatsthepoint.jl_b200 package) may call, link or load this file. Allowed users:
// tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.
//
// PARITY STATUS: "parity unpinned" for neighbour identities and repel trajectories.
// The reference cannot run here (no Julia) and its k-NN arithmetic lives in the
// un-vendored NearestNeighbors.jl (compat "0.4.8", Project.toml:42) reached through
// Meshes.jl (compat "0.56, 0.57", Project.toml:41). The reference's own tests pin only
// list shapes, self exclusion, sortedness, compute_force known answers and the cull
// mask (test/topology.jl:28-65, test/neighbors.jl:55,104-106, test/repel.jl:117-183,
// 301-325); those known answers are checked in tests/test_oracle_golden.py. Neighbour
// sets are additionally cross-checked against scipy.spatial.cKDTree.
//
// What is restated, and from where (paths relative to the reference checkout):
//   knn / knn_self      src/topology.jl:79-84, src/neighbors.jl:9-21  (+ published
//                       NearestNeighbors.jl KDTree algorithm: widest-dimension split,
//                       leaf size 25, bounded max-heap, sqrt applied to the k results)
//   radius              src/topology.jl:91-100 (inrange d2 <= r2, self removed by index)
//   compute_force x4    src/repel_forces.jl:37, 57-60, 96-100, 124-127
//   spacings x3         src/discretization/spacings.jl:19-23, 35-39, 67-72, 121-133
//   _relax!             src/repel.jl:202-339 (sweep :256-292, reductions :293, 374-403,
//                       stop logic :305-337)
//   metrics             src/metrics.jl:19-41, spacing_metrics :56-71, spacing_fidelity_metrics :88-129
//   isinside(p, cloud)  src/isinside.jl:18-35, 86-106 (winding number / Green's function over the boundary points)
//   compute_normals     src/normals.jl:9-44, 65-70 (k-NN + smallest eigenvector of the covariance)
//   _gradient_limit_field   src/discretization/algorithms/octree.jl:677-717 (min-plus sweeps over the k-NN graph)
//   closest point / wall rule   src/octree/geometric_utils.jl:68-136, src/repel.jl:448-469,522-537,
//                               src/octree/triangle_octree.jl:71-99,531-549,583-607 (brute force over triangles)
//
// Canonical order: candidates are ordered by (d2, index) ascending with
// d2 = ((dx*dx + dy*dy) + dz*dz) evaluated in T; compile with -ffp-contract=off.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/wtp_cuda.h"

namespace {

// ---------------------------------------------------------------- distances
template <class T, int D>
inline T dist2(const T* a, const T* b) {
    T s = T(0);
    for (int d = 0; d < D; ++d) {
        T t = a[d] - b[d];
        s = s + t * t;  // -ffp-contract=off: no FMA
    }
    return s;
}

template <class T>
struct Cand {
    T d2;
    int64_t idx;  // 0-based
};
template <class T>
inline bool cand_less(const Cand<T>& a, const Cand<T>& b) {
    return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx);
}

// Bounded max-heap on (d2, idx): top = current worst.
template <class T>
struct TopK {
    Cand<T>* h;
    int k, n;
    TopK(Cand<T>* buf, int k_) : h(buf), k(k_), n(0) {}
    inline bool full() const { return n == k; }
    inline const Cand<T>& worst() const { return h[0]; }
    inline void offer(T d2, int64_t idx) {
        Cand<T> c{d2, idx};
        if (n < k) {
            int i = n++;
            h[i] = c;
            while (i > 0) {
                int p = (i - 1) / 2;
                if (cand_less(h[p], h[i])) { std::swap(h[p], h[i]); i = p; } else break;
            }
        } else if (cand_less(c, h[0])) {
            h[0] = c;
            int i = 0;
            for (;;) {
                int l = 2 * i + 1, r = l + 1, m = i;
                if (l < n && cand_less(h[m], h[l])) m = l;
                if (r < n && cand_less(h[m], h[r])) m = r;
                if (m == i) break;
                std::swap(h[m], h[i]);
                i = m;
            }
        }
    }
    inline void sort_ascending() { std::sort(h, h + n, cand_less<T>); }
};

// ------------------------------------------------------------------ KD-tree
// Widest-dimension median split, leaf size 25, points reordered into tree order
// (the published NearestNeighbors.jl KDTree layout choices; see header).
template <class T, int D>
struct KDTree {
    struct Node {
        T lo[D], hi[D];
        int32_t left, right;  // children (internal) or -1
        int64_t begin, end;   // range in perm (leaf and internal)
    };
    static constexpr int LEAF = 25;
    int64_t n = 0;
    std::vector<T> data;        // reordered coordinates
    std::vector<int64_t> perm;  // tree order -> original index
    std::vector<Node> nodes;

    void build(const T* pts, int64_t n_) {
        n = n_;
        perm.resize(n);
        std::iota(perm.begin(), perm.end(), int64_t(0));
        nodes.clear();
        nodes.reserve(size_t(2 * (n / LEAF + 2)));
        if (n > 0) build_rec(pts, 0, n);
        data.resize(size_t(n) * D);
        for (int64_t i = 0; i < n; ++i)
            for (int d = 0; d < D; ++d) data[size_t(i) * D + d] = pts[size_t(perm[i]) * D + d];
    }
    int32_t build_rec(const T* pts, int64_t b, int64_t e) {
        Node nd;
        for (int d = 0; d < D; ++d) { nd.lo[d] = std::numeric_limits<T>::infinity(); nd.hi[d] = -nd.lo[d]; }
        for (int64_t i = b; i < e; ++i)
            for (int d = 0; d < D; ++d) {
                T v = pts[size_t(perm[i]) * D + d];
                nd.lo[d] = std::min(nd.lo[d], v);
                nd.hi[d] = std::max(nd.hi[d], v);
            }
        nd.left = nd.right = -1;
        nd.begin = b; nd.end = e;
        int32_t id = int32_t(nodes.size());
        nodes.push_back(nd);
        if (e - b > LEAF) {
            int sd = 0;
            T best = nd.hi[0] - nd.lo[0];
            for (int d = 1; d < D; ++d) { T w = nd.hi[d] - nd.lo[d]; if (w > best) { best = w; sd = d; } }
            int64_t mid = b + (e - b) / 2;
            std::nth_element(perm.begin() + b, perm.begin() + mid, perm.begin() + e,
                             [&](int64_t x, int64_t y) {
                                 T vx = pts[size_t(x) * D + sd], vy = pts[size_t(y) * D + sd];
                                 return vx < vy || (vx == vy && x < y);
                             });
            int32_t l = build_rec(pts, b, mid);
            int32_t r = build_rec(pts, mid, e);
            nodes[id].left = l; nodes[id].right = r;
        }
        return id;
    }
    // Lower bound of d2 from q to any point of the node's tight box, evaluated with the
    // same operation order as dist2 (monotone rounding => never above a member's d2).
    inline T box_d2(const Node& nd, const T* q) const {
        T s = T(0);
        for (int d = 0; d < D; ++d) {
            T g = T(0);
            if (q[d] < nd.lo[d]) g = nd.lo[d] - q[d];
            else if (q[d] > nd.hi[d]) g = q[d] - nd.hi[d];
            s = s + g * g;
        }
        return s;
    }
    void knn_rec(int32_t id, const T* q, TopK<T>& tk) const {
        const Node& nd = nodes[id];
        if (nd.left < 0) {
            for (int64_t i = nd.begin; i < nd.end; ++i) tk.offer(dist2<T, D>(q, &data[size_t(i) * D]), perm[i]);
            return;
        }
        T dl = box_d2(nodes[nd.left], q), dr = box_d2(nodes[nd.right], q);
        int32_t first = nd.left, second = nd.right;
        if (dr < dl) { std::swap(first, second); std::swap(dl, dr); }
        if (!tk.full() || dl <= tk.worst().d2) knn_rec(first, q, tk);
        if (!tk.full() || dr <= tk.worst().d2) knn_rec(second, q, tk);
    }
    // k nearest (k <= n), ascending (d2, idx).
    void knn(const T* q, int k, Cand<T>* buf) const {
        TopK<T> tk(buf, k);
        if (n > 0) knn_rec(0, q, tk);
        tk.sort_ascending();
    }
    template <class F>
    void inrange_rec(int32_t id, const T* q, T r2, F&& emit) const {
        const Node& nd = nodes[id];
        if (box_d2(nd, q) > r2) return;
        if (nd.left < 0) {
            for (int64_t i = nd.begin; i < nd.end; ++i)
                if (dist2<T, D>(q, &data[size_t(i) * D]) <= r2) emit(perm[i]);
            return;
        }
        inrange_rec(nd.left, q, r2, emit);
        inrange_rec(nd.right, q, r2, emit);
    }
};

// --------------------------------------------------------------- force laws
// src/repel_forces.jl:37, 57-60, 96-100, 124-127. (x)^2 in Julia lowers to x*x.
template <class T>
inline T force_eval(int kind, T beta, T u0, T gamma, T u) {
    T u2 = u * u;
    switch (kind) {
        case WTP_FORCE_INVERSE: { T t = u2 + beta; return T(1) / (t * t); }
        case WTP_FORCE_EQUILIBRIUM: { T t = u2 + beta; return (T(1) - u2) / (t * t); }
        case WTP_FORCE_CLIPPED: { T t = u2 + beta; T F = (u0 * u0 - u2) / (t * t); return F > T(0) ? F : T(0); }
        case WTP_FORCE_STRONG: { return (T(1) - u2) / std::pow(u2 + beta, gamma); }
    }
    return std::numeric_limits<T>::quiet_NaN();
}

// ---------------------------------------------------------------- spacings
// src/discretization/spacings.jl. _min_distance (:19-23) is a 1-NN KD-tree query on
// the spacing's own boundary set, distance in the tree's float type.
template <class T, int D>
struct Spacing {
    int kind;
    T a, b, c;
    KDTree<T, D> tree;
    void init(const wtp_spacing* sp) {
        kind = sp->kind; a = T(sp->a); b = T(sp->b); c = T(sp->c);
        if (kind != WTP_SPACING_CONSTANT) tree.build(static_cast<const T*>(sp->bnd_pts), sp->n_bnd);
    }
    inline T operator()(const T* x) const {
        if (kind == WTP_SPACING_CONSTANT) return a;                         // :35-39
        Cand<T> nn;
        tree.knn(x, 1, &nn);
        T dmin = std::sqrt(nn.d2);
        if (kind == WTP_SPACING_LOGLIKE) {                                  // :67-72
            T inv_growth = T(1) - (b - T(1));
            T aa = a * inv_growth;
            return a * dmin / (aa + dmin);
        }
        T delta = c;                                                         // :121-133
        T center = delta / T(2), width = delta / T(6);
        T sigma = T(1) / (T(1) + std::exp(-(dmin - center) / width));
        return a + (b - a) * sigma;
    }
};

// ------------------------------------------------------ closest point (wall)
// Ericson's closest point on triangle with the feature code of the region
// (src/octree/geometric_utils.jl:68-136): 0 face, 1..3 vertex, 4 edge12, 5 edge13, 6 edge23.
template <class T>
inline int closest_point_on_triangle(const T* p, const T* a, const T* b, const T* c, T* out) {
    T ab[3], ac[3], ap[3];
    for (int i = 0; i < 3; ++i) { ab[i] = b[i] - a[i]; ac[i] = c[i] - a[i]; ap[i] = p[i] - a[i]; }
    auto dot = [](const T* x, const T* y) { return (x[0] * y[0] + x[1] * y[1]) + x[2] * y[2]; };
    T d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) { for (int i = 0; i < 3; ++i) out[i] = a[i]; return 1; }
    T bp[3]; for (int i = 0; i < 3; ++i) bp[i] = p[i] - b[i];
    T d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) { for (int i = 0; i < 3; ++i) out[i] = b[i]; return 2; }
    T vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { T v = d1 / (d1 - d3); for (int i = 0; i < 3; ++i) out[i] = a[i] + v * ab[i]; return 4; }
    T cp[3]; for (int i = 0; i < 3; ++i) cp[i] = p[i] - c[i];
    T d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) { for (int i = 0; i < 3; ++i) out[i] = c[i]; return 3; }
    T vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { T w = d2 / (d2 - d6); for (int i = 0; i < 3; ++i) out[i] = a[i] + w * ac[i]; return 5; }
    T va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        T w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int i = 0; i < 3; ++i) out[i] = b[i] + w * (c[i] - b[i]);
        return 6;
    }
    T denom = T(1) / (va + vb + vc);
    T v = vb * denom, w = vc * denom;
    for (int i = 0; i < 3; ++i) out[i] = a[i] + ab[i] * v + ac[i] * w;
    return 0;
}

// Brute-force nearest triangle (src/octree/triangle_octree.jl:531-549 without the tree):
// canonical tie-break (d2, triangle index). Returns the 0-based triangle, -1 if none.
template <class T>
struct Nearest { T d2; int64_t tri; T cp[3]; int feature; };
template <class T>
inline Nearest<T> nearest_triangle(const T* tris, int64_t n_tri, const T* p) {
    Nearest<T> best{std::numeric_limits<T>::max(), -1, {p[0], p[1], p[2]}, 0};
    for (int64_t t = 0; t < n_tri; ++t) {
        const T* v = tris + size_t(t) * 9;
        T cp[3];
        int f = closest_point_on_triangle<T>(p, v, v + 3, v + 6, cp);
        T dv[3] = {p[0] - cp[0], p[1] - cp[1], p[2] - cp[2]};
        T d2 = (dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2];
        if (d2 < best.d2) { best.d2 = d2; best.tri = t; best.feature = f; for (int i = 0; i < 3; ++i) best.cp[i] = cp[i]; }
    }
    return best;
}
// isinside(p, octree) (src/octree/triangle_octree.jl:71-99, 583-607): mesh-bbox test, then the
// sign of dot(p - cp, pseudonormal of the closest feature) < 0.
template <class T>
inline bool mesh_isinside(const wtp_wall_mesh* m, const T* p) {
    for (int d = 0; d < 3; ++d) if (p[d] < T(m->bbox_min[d]) || p[d] > T(m->bbox_max[d])) return false;
    const T* tris = static_cast<const T*>(m->triangles);
    Nearest<T> nb = nearest_triangle<T>(tris, m->n_tri, p);
    if (nb.tri < 0) return false;
    const T* n = static_cast<const T*>(m->feature_normals) + size_t(nb.tri) * 21 + size_t(nb.feature) * 3;
    T s = ((p[0] - nb.cp[0]) * n[0] + (p[1] - nb.cp[1]) * n[1]) + (p[2] - nb.cp[2]) * n[2];
    return s < T(0);
}
// _project_to_boundary (src/repel.jl:522-537): closest point nudged inward along the face normal.
template <class T>
inline int64_t mesh_project(const wtp_wall_mesh* m, const T* p, T* out) {
    const T* tris = static_cast<const T*>(m->triangles);
    Nearest<T> nb = nearest_triangle<T>(tris, m->n_tri, p);
    if (nb.tri < 0) { for (int i = 0; i < 3; ++i) out[i] = p[i]; return 0; }
    const T* n = static_cast<const T*>(m->feature_normals) + size_t(nb.tri) * 21;
    const T off = T(m->offset_dist);
    for (int i = 0; i < 3; ++i) out[i] = nb.cp[i] - off * n[i];
    return nb.tri + 1;
}

// ------------------------------------------------- the random direction of coincident pairs
// _safe_direction (src/repel.jl:358-364) returns randn(...)/norm for r == 0, drawn from Julia's global RNG: a stream no
// other implementation can reproduce. The library defines its own counter-based stream (include/wtp_cuda.h, "random
// directions"); this is its restatement: the first point of a hashed sequence in [-1, 1)^D inside the unit ball and at
// least 2^-5 from the origin, normalised. Exactly rounded operations only, so the bits agree with the device.
inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
inline uint64_t sweep_key(uint64_t seed, uint64_t iteration) { return mix64(seed ^ mix64(iteration)); }
template <class T, int D>
inline void random_unit(uint64_t key, uint32_t i, uint32_t j, T* out) {
    const uint64_t h1 = mix64(key ^ ((uint64_t(i) << 32) | uint64_t(j)));
    for (uint64_t c = 0;; ++c) {
        const uint64_t h = mix64(h1 + c);
        const T scale = T(1.0 / 1048576.0);
        T v[3] = {T(int(h & 0x1fffffu) - 1048576) * scale, T(int((h >> 21) & 0x1fffffu) - 1048576) * scale,
                  D == 3 ? T(int((h >> 42) & 0x1fffffu) - 1048576) * scale : T(0)};
        T n2 = v[0] * v[0] + v[1] * v[1];
        if (D == 3) n2 = n2 + v[2] * v[2];
        if (n2 <= T(1) && n2 >= T(1.0 / 1024.0)) {
            const T n = std::sqrt(n2);
            for (int d = 0; d < D; ++d) out[d] = v[d] / n;
            return;
        }
    }
}

// ------------------------------------------------------------------- relax
template <class T, int D>
int32_t relax(T* snap, int64_t n_fixed, int64_t n_move, const wtp_spacing* sp_in, const wtp_force* fm,
              const wtp_repel_params* prm, const wtp_wall_mesh* wall, T* conv, wtp_trace_entry* trace, wtp_repel_result* res,
              int threads) {
    const int64_t n_all = n_fixed + n_move;
    if (prm->rebuild_every < 1) return WTP_ERR_BAD_ARG;                      // src/repel.jl:74
    if (prm->kick_after > 0) return WTP_ERR_UNSUPPORTED;                     // randn, :430
    if ((prm->wall == WTP_WALL_MESH) != (wall != nullptr)) return WTP_ERR_BAD_ARG;
    if (wall && D != 3) return WTP_ERR_BAD_ARG;
    if (prm->deposit_ratio < 0 || (prm->deposit_ratio > 0 && (!wall || n_fixed != 0))) return WTP_ERR_BAD_ARG;   // :143, :172
    const int kk = int(std::min<int64_t>(prm->k, n_all));                   // :208
    Spacing<T, D> spacing;
    spacing.init(sp_in);
    const T beta = T(fm->beta), u0 = T(fm->u0), gamma = T(fm->gamma);
    std::vector<T> spacings(size_t(n_all) > 0 ? size_t(n_all) : 1);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t i = 0; i < n_all; ++i) spacings[i] = spacing(&snap[size_t(i) * D]);   // :209
    std::vector<T> p(snap + size_t(n_fixed) * D, snap + size_t(n_all) * D);  // movable, current
    std::vector<T> p_old(p);
    std::vector<T> coords(snap, snap + size_t(n_all) * D);                   // :216
    KDTree<T, D> tree;
    tree.build(coords.data(), n_all);                                         // :218
    const T a_lo = T(prm->alpha_lo), a_max = T(prm->alpha_max);              // :226
    std::vector<T> forces(size_t(n_move), T(0)), nn_dist(size_t(n_move), std::numeric_limits<T>::max());
    std::vector<int64_t> nn_id(size_t(n_move), 0);
    T best_cv = std::numeric_limits<T>::max();
    int64_t last_impr = 0;
    int it = 1, n_conv = 0;
    int stop = WTP_STOP_MAX_ITERS;
    double last_cv = std::numeric_limits<double>::quiet_NaN();
    (void)threads;
    while (it <= prm->max_iters) {                                            // :243
        p_old = p;                                                            // :244
        if ((it - 1) % prm->rebuild_every == 0) {                             // :245-253
            std::copy(p.begin(), p.end(), coords.begin() + size_t(n_fixed) * D);
            // (serial in the reference, :251; elementwise, so the threads change nothing but the time)
#pragma omp parallel for num_threads(threads) schedule(static)
            for (int64_t i = 0; i < n_move; ++i) spacings[size_t(n_fixed + i)] = spacing(&p[size_t(i) * D]);
            tree.build(coords.data(), n_all);
        }
#pragma omp parallel num_threads(threads)
        {
            std::vector<Cand<T>> buf(size_t(kk) > 0 ? size_t(kk) : 1);
#pragma omp for schedule(dynamic, 256)
            for (int64_t id = 0; id < n_move; ++id) {                         // :256-292
                const T* xi = &p_old[size_t(id) * D];
                tree.knn(xi, kk, buf.data());                                 // :259
                const T s = spacing(xi);                                      // :260
                const int64_t self = id + n_fixed;                            // :266
                int64_t nid = 0; T nd = std::numeric_limits<T>::max();
                T F[D]; for (int d = 0; d < D; ++d) F[d] = T(0);
                for (int j = 0; j < kk; ++j) {                                // :270-280
                    if (buf[j].idx == self) continue;
                    const T r = std::sqrt(buf[j].d2);
                    if (nid == 0) { nid = buf[j].idx + 1; nd = r; }
                    const T* xj = &coords[size_t(buf[j].idx) * D];
                    const T f = force_eval<T>(fm->kind, beta, u0, gamma, r / s);
                    if (r > T(0)) {                                           // _safe_direction :358-364
                        for (int d = 0; d < D; ++d) F[d] = F[d] + f * ((xi[d] - xj[d]) / r);
                    } else {                                                  // coincident: the library's random stream
                        T dir[3];
                        random_unit<T, D>(sweep_key(prm->kick_seed, uint64_t(it)), uint32_t(self), uint32_t(buf[j].idx), dir);
                        for (int d = 0; d < D; ++d) F[d] = F[d] + f * dir[d];
                    }
                }
                T n2 = T(0); for (int d = 0; d < D; ++d) n2 = n2 + F[d] * F[d];
                const T Fn = std::sqrt(n2);                                   // :282
                forces[size_t(id)] = Fn * s;                                  // :283
                T ai = T(1) / (Fn + T(1.0e-30));                              // :285
                ai = ai > a_max ? a_max : (ai < a_lo ? a_lo : ai);
                T disp[D]; const T sa = s * ai;
                T dn2 = T(0);
                for (int d = 0; d < D; ++d) { disp[d] = sa * F[d]; dn2 = dn2 + disp[d] * disp[d]; }
                const T dn = std::sqrt(dn2);
                if (dn > s) { const T sc = s / dn; for (int d = 0; d < D; ++d) disp[d] = disp[d] * sc; }  // :288-290
                T prop[3] = {T(0), T(0), T(0)};
                for (int d = 0; d < D; ++d) prop[d] = xi[d] + disp[d];
                if (wall && D == 3) {                                         // _constrain_octree :448-469
                    T xi3[3] = {xi[0], xi[1], xi[D - 1]};
                    if (wall->is_bnd[id]) {
                        T sv[3];
                        int64_t tri = mesh_project<T>(wall, prop, sv);
                        if (tri == 0) tri = mesh_project<T>(wall, xi3, sv);
                        wall->tri_indices[id] = tri;
                        for (int d = 0; d < 3; ++d) prop[d] = sv[d];
                    } else if (!mesh_isinside<T>(wall, prop)) {
                        wall->escaped[id] = 1;
                        for (int d = 0; d < 3; ++d) prop[d] = xi3[d];
                    }
                }
                for (int d = 0; d < D; ++d) p[size_t(id) * D + d] = prop[d];  // :291
                nn_id[size_t(id)] = nid; nn_dist[size_t(id)] = nd;
            }
        }
        T mx = T(0); for (int64_t i = 0; i < n_move; ++i) mx = std::max(mx, forces[size_t(i)]);
        conv[n_conv++] = mx;                                                  // :293
        if (n_move > 0 && trace) {                                            // :294-296, 396-403
            int64_t i = 0; for (int64_t j = 1; j < n_move; ++j) if (nn_dist[size_t(j)] < nn_dist[size_t(i)]) i = j;
            int64_t j = nn_id[size_t(i)], ig = i + n_fixed + 1;
            T r = nn_dist[size_t(i)];
            T s = j > 0 ? (spacings[size_t(ig - 1)] + spacings[size_t(j - 1)]) / T(2) : spacings[size_t(ig - 1)];
            wtp_trace_entry& te = trace[n_conv - 1];
            te.r = double(r); te.s = double(s); te.r_over_s = double(r / s);
            te.idx_a = std::min(ig, j); te.idx_b = std::max(ig, j);
        }
        bool stopped = false;
        if ((prm->stall_after > 0 || prm->cv_target > 0) && n_move > 0) {     // :305-327
            T s1 = T(0), s2 = T(0);                                           // _dnn_cv :374-386 (serial, in T)
            if (prm->reserved & 1) {
                // NOT the reference: the same terms accumulated in double, the device's documented deviation
                // (DESIGN.md section 2). Test-only switch that isolates what the accumulation precision does to a
                // Float32 stall_after run; the reference-faithful path is the else branch.
                double a1 = 0, a2 = 0;
                for (int64_t i = 0; i < n_move; ++i) {
                    T u = nn_dist[size_t(i)] / spacings[size_t(i + n_fixed)];
                    a1 += double(u); a2 += double(u * u);
                }
                s1 = T(a1); s2 = T(a2);
                T mu = T(a1 / double(n_move));
                T var = T(a2 / double(n_move)) - mu * mu;
                T cv = std::sqrt(var > T(0) ? var : T(0)) / mu;
                last_cv = double(cv);
                if (prm->cv_target > 0 && double(cv) <= prm->cv_target) { p = p_old; stop = WTP_STOP_CV_TARGET; stopped = true; }
                else if (prm->stall_after > 0) {
                    if (double(cv) < double(best_cv) * (1 - 1.0e-3)) { best_cv = cv; last_impr = it; }
                    else if (it - last_impr >= prm->stall_after) { stop = WTP_STOP_STALL; stopped = true; }
                }
            } else {
            for (int64_t i = 0; i < n_move; ++i) {
                T u = nn_dist[size_t(i)] / spacings[size_t(i + n_fixed)];
                s1 = s1 + u; s2 = s2 + u * u;
            }
            T mu = s1 / T(n_move);
            T var = s2 / T(n_move) - mu * mu;
            T cv = std::sqrt(var > T(0) ? var : T(0)) / mu;
            last_cv = double(cv);
            if (prm->cv_target > 0 && double(cv) <= prm->cv_target) {
                p = p_old; stop = WTP_STOP_CV_TARGET; stopped = true;         // :307-317
            } else if (prm->stall_after > 0) {
                // :319 — Julia evaluates best_cv * (1 - 1.0e-3) with a Float64 literal.
                if (double(cv) < double(best_cv) * (1 - 1.0e-3)) { best_cv = cv; last_impr = it; }
                else if (it - last_impr >= prm->stall_after) { stop = WTP_STOP_STALL; stopped = true; }
            }
            }
        }
        if (stopped) break;
        if (wall && D == 3 && prm->deposit_ratio > 0) {                        // deposit!(p, tree, i) :328 -> _deposit_escaped! :483-514
            uint8_t* is_bnd = const_cast<uint8_t*>(wall->is_bnd);
            const int kq = int(std::min<int64_t>(prm->k, n_move));            // min(k, length(p_cur)) :163
            std::vector<Cand<T>> near(size_t(kq) > 0 ? size_t(kq) : 1);
            for (int64_t id = 0; id < n_move; ++id) {                          // serial on purpose
                if (!wall->escaped[id]) continue;
                wall->escaped[id] = 0;
                if (is_bnd[id]) continue;
                T here[3] = {p[size_t(id) * D], p[size_t(id) * D + 1], p[size_t(id) * D + (D - 1)]}, site[3];
                const int64_t tri = mesh_project<T>(wall, here, site);        // :495
                if (tri == 0) continue;
                const double thr = prm->deposit_ratio * double(spacing(site));   // :501 (Float64 ratio times T spacing)
                tree.knn(site, kq, near.data());                              // the sweep's snapshot tree :502
                bool occupied = false;
                for (int j = 0; j < kq && !occupied; ++j) {                   // :503-506 (n_fixed = 0: snapshot index = movable id)
                    const int64_t jj = near[size_t(j)].idx;
                    if (jj == id || !is_bnd[jj]) continue;
                    const T* pj = &p[size_t(jj) * D];
                    const T dx = pj[0] - site[0], dy = pj[1] - site[1], dz = pj[D - 1] - site[2];
                    occupied = double(std::sqrt((dx * dx + dy * dy) + dz * dz)) < thr;
                }
                if (occupied) continue;
                for (int d = 0; d < 3; ++d) p[size_t(id) * D + d] = site[d];  // :508-510
                is_bnd[id] = 1;
                wall->tri_indices[id] = tri;
            }
        }
        if (double(conv[n_conv - 1]) < prm->tol) { stop = WTP_STOP_TOL; break; }   // :329-332
        ++it;
    }
    std::copy(p.begin(), p.end(), snap + size_t(n_fixed) * D);
    res->iters = n_conv; res->stop_reason = stop; res->last_cv = last_cv;
    return WTP_OK;
}

template <class T, int D>
int32_t knn_impl(const T* pts, int64_t N, int k, int drop_first, int algo, int threads, int64_t* out_idx, T* out_dist) {
    const int K1 = k + (drop_first ? 1 : 0);
    if (K1 > N) return WTP_ERR_K_TOO_LARGE;
    if (algo == 0) {  // brute force: the independent check of the KD-tree
#pragma omp parallel num_threads(threads)
        {
            std::vector<Cand<T>> all(static_cast<size_t>(N));
#pragma omp for schedule(static)
            for (int64_t i = 0; i < N; ++i) {
                for (int64_t j = 0; j < N; ++j) all[size_t(j)] = Cand<T>{dist2<T, D>(&pts[size_t(i) * D], &pts[size_t(j) * D]), j};
                std::partial_sort(all.begin(), all.begin() + K1, all.end(), cand_less<T>);
                for (int j = 0; j < k; ++j) {
                    const Cand<T>& c = all[size_t(j + (drop_first ? 1 : 0))];
                    out_idx[size_t(i) * k + j] = c.idx + 1;
                    if (out_dist) out_dist[size_t(i) * k + j] = std::sqrt(c.d2);
                }
            }
        }
        return WTP_OK;
    }
    KDTree<T, D> tree;
    tree.build(pts, N);
#pragma omp parallel num_threads(threads)
    {
        std::vector<Cand<T>> buf(static_cast<size_t>(K1));
#pragma omp for schedule(dynamic, 1024)
        for (int64_t i = 0; i < N; ++i) {
            tree.knn(&pts[size_t(i) * D], K1, buf.data());
            for (int j = 0; j < k; ++j) {
                const Cand<T>& c = buf[size_t(j + (drop_first ? 1 : 0))];
                out_idx[size_t(i) * k + j] = c.idx + 1;
                if (out_dist) out_dist[size_t(i) * k + j] = std::sqrt(c.d2);
            }
        }
    }
    return WTP_OK;
}

template <class T, int D>
int32_t radius_impl(const T* pts, int64_t N, T r, int threads, int64_t* offsets, int64_t* indices) {
    KDTree<T, D> tree;
    tree.build(pts, N);
    const T r2 = r * r;
    if (!indices) {
        offsets[0] = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1024)
        for (int64_t i = 0; i < N; ++i) {
            int64_t c = 0;
            tree.inrange_rec(0, &pts[size_t(i) * D], r2, [&](int64_t j) { if (j != i) ++c; });
            offsets[i + 1] = c;
        }
        for (int64_t i = 0; i < N; ++i) offsets[i + 1] += offsets[i];
        return WTP_OK;
    }
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1024)
    for (int64_t i = 0; i < N; ++i) {
        int64_t* row = indices + offsets[i];
        int64_t c = 0;
        tree.inrange_rec(0, &pts[size_t(i) * D], r2, [&](int64_t j) { if (j != i) row[c++] = j + 1; });
        std::sort(row, row + c);
    }
    return WTP_OK;
}

template <class T, int D>
int32_t metrics_impl(const T* pts, int64_t N, int k, int threads, wtp_cloud_metrics* out) {
    // src/metrics.jl:19-41: k nearest including self, drop the first, per-point
    // mean/std/max/min of the remaining k-1 distances, then means over points.
    if (k > N || k < 2) return WTP_ERR_K_TOO_LARGE;
    KDTree<T, D> tree;
    tree.build(pts, N);
    const int m = k - 1;
    const size_t NN = static_cast<size_t>(N);
    std::vector<double> a(NN), sd(NN), mx(NN), mn(NN), nn(NN);
#pragma omp parallel num_threads(threads)
    {
        std::vector<Cand<T>> buf(static_cast<size_t>(k));
#pragma omp for schedule(dynamic, 1024)
        for (int64_t i = 0; i < N; ++i) {
            tree.knn(&pts[size_t(i) * D], k, buf.data());
            T s = T(0);
            for (int j = 1; j < k; ++j) s = s + std::sqrt(buf[size_t(j)].d2);
            T mean = s / T(m);
            T v = T(0);
            for (int j = 1; j < k; ++j) { T e = std::sqrt(buf[size_t(j)].d2) - mean; v = v + e * e; }
            a[size_t(i)] = double(mean);
            sd[size_t(i)] = m > 1 ? double(std::sqrt(v / T(m - 1))) : std::numeric_limits<double>::quiet_NaN();
            mx[size_t(i)] = double(std::sqrt(buf[size_t(k - 1)].d2));
            mn[size_t(i)] = double(std::sqrt(buf[1].d2));
            nn[size_t(i)] = mn[size_t(i)];
        }
    }
    auto mean_of = [&](const std::vector<double>& v) { double s = 0; for (double x : v) s += x; return s / double(N); };
    out->avg = mean_of(a); out->std = mean_of(sd); out->max = mean_of(mx); out->min = mean_of(mn);
    out->separation = *std::min_element(nn.begin(), nn.end());
    out->fill = *std::max_element(nn.begin(), nn.end());
    out->mesh_ratio = out->fill / out->separation;
    return WTP_OK;
}

// _near_duplicate_keep_mask (src/repel.jl:565-580): ball search at ratio * max spacing, greedy sweep in index order
template <class T, int D>
int32_t cull_mask_impl(const T* pts, int64_t N, const T* spacings, double ratio, uint8_t* keep) {
    std::fill(keep, keep + N, uint8_t(1));
    if (!(ratio > 0) || N < 2) return WTP_OK;
    KDTree<T, D> tree;
    tree.build(pts, N);
    const T r = T(ratio) * *std::max_element(spacings, spacings + N);
    std::vector<int64_t> hits;
    for (int64_t i = 0; i < N; ++i) {
        if (!keep[i]) continue;
        const T thr = T(ratio) * spacings[i];
        hits.clear();
        if (tree.n > 0) tree.inrange_rec(0, &pts[size_t(i) * D], r * r, [&](int64_t j) { hits.push_back(j); });
        for (int64_t j : hits) {
            if (j == i || !keep[j]) continue;
            T d2 = T(0);
            for (int d = 0; d < D; ++d) { const T v = pts[size_t(j) * D + d] - pts[size_t(i) * D + d]; d2 = d2 + v * v; }
            if (std::sqrt(d2) < thr) keep[j] = 0;
        }
    }
    return WTP_OK;
}

// spacing_metrics (src/metrics.jl:56-71) and spacing_fidelity_metrics (:88-129)
template <class T, int D>
int32_t spacing_metrics_impl(const T* pts, int64_t N, int k, const wtp_spacing* sp_in, int threads, wtp_spacing_metrics_t* out) {
    if (k > N || k < 2) return WTP_ERR_K_TOO_LARGE;
    KDTree<T, D> tree;
    tree.build(pts, N);
    Spacing<T, D> spacing;
    spacing.init(sp_in);
    std::vector<T> err(static_cast<size_t>(N));
#pragma omp parallel num_threads(threads)
    {
        std::vector<Cand<T>> buf(static_cast<size_t>(k));
#pragma omp for schedule(dynamic, 1024)
        for (int64_t i = 0; i < N; ++i) {
            tree.knn(&pts[size_t(i) * D], k, buf.data());
            T s = T(0);
            for (int j = 1; j < k; ++j) s = s + std::sqrt(buf[size_t(j)].d2);
            const T actual = s / T(k - 1), target = spacing(&pts[size_t(i) * D]);
            err[size_t(i)] = std::fabs(actual - target) / target;
        }
    }
    double mx = 0, sum = 0;
    for (T e : err) { mx = std::max(mx, double(e)); sum += double(e); }
    const double mu = sum / double(N);
    double ss = 0;
    for (T e : err) ss += (double(e) - mu) * (double(e) - mu);
    out->max_error = mx; out->mean_error = mu; out->std_error = N > 1 ? std::sqrt(ss / double(N - 1)) : std::numeric_limits<double>::quiet_NaN();
    return WTP_OK;
}

template <class T, int D>
int32_t spacing_fidelity_impl(const T* pts, int64_t N, int k, double coord_radius, const wtp_spacing* sp_in, int threads, wtp_spacing_fidelity_t* out) {
    k = int(std::min<int64_t>(k, N));                                              // :93
    if (k < 1) return WTP_ERR_BAD_ARG;
    KDTree<T, D> tree;
    tree.build(pts, N);
    Spacing<T, D> spacing;
    spacing.init(sp_in);
    std::vector<T> u(static_cast<size_t>(N));
    std::vector<int> coord(static_cast<size_t>(N));
    const T cr = T(coord_radius);
#pragma omp parallel num_threads(threads)
    {
        std::vector<Cand<T>> buf(static_cast<size_t>(k));
#pragma omp for schedule(dynamic, 1024)
        for (int64_t i = 0; i < N; ++i) {
            tree.knn(&pts[size_t(i) * D], k, buf.data());
            const T h = spacing(&pts[size_t(i) * D]);
            T dmin = std::numeric_limits<T>::max();
            int c = 0;
            for (int j = 0; j < k; ++j) {
                if (buf[size_t(j)].idx == i) continue;                              // skip self BY INDEX (:106)
                const T d = std::sqrt(buf[size_t(j)].d2);
                dmin = std::min(dmin, d);
                if (d <= cr * h) ++c;
            }
            u[size_t(i)] = dmin / h;
            coord[size_t(i)] = c;
        }
    }
    double sum = 0, csum = 0;
    for (int64_t i = 0; i < N; ++i) { sum += double(u[size_t(i)]); csum += coord[size_t(i)]; }
    const double mu = sum / double(N);
    double ss = 0;
    for (T x : u) ss += (double(x) - mu) * (double(x) - mu);
    std::sort(u.begin(), u.end());
    auto quant = [&](double p) {                                                    // Julia quantile, type 7
        const double hh = double(N - 1) * p;
        const int64_t lo = int64_t(std::floor(hh));
        const int64_t hi = std::min<int64_t>(lo + 1, N - 1);
        return double(u[size_t(lo)]) + (hh - double(lo)) * (double(u[size_t(hi)]) - double(u[size_t(lo)]));
    };
    out->mean_dnn_h = mu;
    out->cv = (N > 1 ? std::sqrt(ss / double(N - 1)) : std::numeric_limits<double>::quiet_NaN()) / mu;
    out->p05 = quant(0.05); out->p50 = quant(0.5); out->p95 = quant(0.95);
    out->coordination = csum / double(N);
    return WTP_OK;
}

// ------------------------------------------------------------------ normals
// compute_normals (src/normals.jl:9-44) + _compute_normal (:65-70): neighbors = search.(points, Ref(KNearestSearch(points, k)))
// (self included), then per point eigen(Symmetric(cov(v))) of the gathered coordinates, Q[:, 1] = the eigenvector of the
// smallest eigenvalue. cov in T (Statistics.cov: mean, centred second moments over k - 1); the symmetric eigenproblem by
// cyclic Jacobi rotations in double (LAPACK in the reference: same eigenvector up to sign and rounding). The sign is
// not defined by the reference; the first nonzero component is made positive.
template <int D>
inline void smallest_eigenvector(double C[3][3], double* out) {
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0;
        for (int p = 0; p < D; ++p) for (int q = p + 1; q < D; ++q) off += C[p][q] * C[p][q];
        if (off == 0.0) break;
        for (int p = 0; p < D; ++p)
            for (int q = p + 1; q < D; ++q) {
                if (C[p][q] == 0.0) continue;
                const double theta = (C[q][q] - C[p][p]) / (2.0 * C[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int r = 0; r < D; ++r) { const double a = C[r][p], b = C[r][q]; C[r][p] = c * a - s * b; C[r][q] = s * a + c * b; }
                for (int r = 0; r < D; ++r) { const double a = C[p][r], b = C[q][r]; C[p][r] = c * a - s * b; C[q][r] = s * a + c * b; }
                for (int r = 0; r < D; ++r) { const double a = V[r][p], b = V[r][q]; V[r][p] = c * a - s * b; V[r][q] = s * a + c * b; }
            }
    }
    int m = 0;
    for (int d = 1; d < D; ++d) if (C[d][d] < C[m][m]) m = d;
    double n2 = 0;
    for (int d = 0; d < D; ++d) n2 += V[d][m] * V[d][m];
    const double inv = 1.0 / std::sqrt(n2);
    double sign = 1.0;
    for (int d = 0; d < D; ++d) if (V[d][m] != 0.0) { sign = V[d][m] > 0 ? 1.0 : -1.0; break; }
    for (int d = 0; d < D; ++d) out[d] = sign * V[d][m] * inv;
}

template <class T, int D>
int32_t normals_impl(const T* pts, int64_t N, int k, int threads, T* out) {
    if (k < 1 || N < 1) return WTP_ERR_BAD_ARG;
    if (int64_t(k) > N) k = int(N);                                           // :16
    KDTree<T, D> tree;
    tree.build(pts, N);
#pragma omp parallel num_threads(threads)
    {
        std::vector<Cand<T>> buf((size_t)k);
#pragma omp for schedule(dynamic, 512)
        for (int64_t i = 0; i < N; ++i) {
            tree.knn(&pts[size_t(i) * D], k, buf.data());
            T mean[3] = {T(0), T(0), T(0)};
            for (int j = 0; j < k; ++j) for (int d = 0; d < D; ++d) mean[d] = mean[d] + pts[size_t(buf[j].idx) * D + d];
            for (int d = 0; d < D; ++d) mean[d] = mean[d] / T(k);
            T S[3][3] = {};
            for (int j = 0; j < k; ++j) {
                T c[3];
                for (int d = 0; d < D; ++d) c[d] = pts[size_t(buf[j].idx) * D + d] - mean[d];
                for (int a = 0; a < D; ++a) for (int b = a; b < D; ++b) S[a][b] = S[a][b] + c[a] * c[b];
            }
            double C[3][3] = {};
            const T denom = T(k > 1 ? k - 1 : 1);
            for (int a = 0; a < D; ++a) for (int b = a; b < D; ++b) { C[a][b] = double(S[a][b] / denom); C[b][a] = C[a][b]; }
            double v[3];
            smallest_eigenvector<D>(C, v);
            for (int d = 0; d < D; ++d) out[size_t(i) * D + d] = T(v[d]);
        }
    }
    return WTP_OK;
}

// --------------------------------------------------------- gradient-limit field
// _gradient_limit_field (src/discretization/algorithms/octree.jl:677-717) on the leaf centres.
template <class T, int D>
int32_t gradient_limit_impl(const T* centers, int64_t n, const T* h0, T g, int k, double tol, int max_sweeps, int threads, T* out, int32_t* sweeps_out) {
    if (n < 1 || k < 1) return WTP_ERR_BAD_ARG;
    const int kk = int(std::min<int64_t>(k, n));                              // :684
    KDTree<T, D> tree;
    tree.build(centers, n);                                                   // :685
    std::vector<int64_t> idx(size_t(n) * kk);
    std::vector<T> dist(size_t(n) * kk);
#pragma omp parallel num_threads(threads)
    {
        std::vector<Cand<T>> buf((size_t)kk);
#pragma omp for schedule(dynamic, 512)
        for (int64_t a = 0; a < n; ++a) {                                     // knn(tree, centers, kk, true) :686
            tree.knn(&centers[size_t(a) * D], kk, buf.data());
            for (int t = 0; t < kk; ++t) { idx[size_t(a) * kk + t] = buf[t].idx; dist[size_t(a) * kk + t] = std::sqrt(buf[t].d2); }
        }
    }
    std::vector<T> h(h0, h0 + n), hnew((size_t)n);
    int sweeps = 0;
    for (int s = 0; s < max_sweeps; ++s) {                                    // :693
#pragma omp parallel for num_threads(threads) schedule(static)
        for (int64_t a = 0; a < n; ++a) {                                     // :694-702
            T hi = h[size_t(a)];
            for (int t = 0; t < kk; ++t) {
                const T cand = h[size_t(idx[size_t(a) * kk + t])] + g * dist[size_t(a) * kk + t];
                if (cand < hi) hi = cand;
            }
            hnew[size_t(a)] = hi;
        }
        T maxrel = T(0);                                                      // :703-707
        for (int64_t a = 0; a < n; ++a) {
            const T rel = std::fabs(hnew[size_t(a)] - h[size_t(a)]) / h[size_t(a)];
            if (rel > maxrel) maxrel = rel;
        }
        h = hnew;                                                             // :708
        ++sweeps;
        if (sizeof(T) == 4 ? float(maxrel) < float(tol) : double(maxrel) < tol) break;   // :709
    }
    std::copy(h.begin(), h.end(), out);
    if (sweeps_out) *sweeps_out = sweeps;
    return WTP_OK;
}

int default_threads(int threads) {
#ifdef _OPENMP
    return threads > 0 ? threads : omp_get_max_threads();
#else
    (void)threads; return 1;
#endif
}

}  // namespace

#define DISPATCH_D(T, D, call2, call3) ((D) == 2 ? (call2) : (D) == 3 ? (call3) : int32_t(WTP_ERR_BAD_ARG))

// isinside(p, cloud) against the boundary point cloud (src/isinside.jl): 3-D Green's-function sum (:86-106),
// 2-D winding number (:18-35, the signed angle A-p-B is atan(u x v, u . v)). Serial sums in element order.
template <class T>
static void isinside_points(const T* pts, int64_t N, int D, const T* bx, const T* bn, const T* ba, int64_t M, int threads, uint8_t* out, T* sums) {
    const T eps = std::numeric_limits<T>::epsilon();
    #pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        T g = T(0);
        bool in;
        if (D == 3) {
            const T* p = pts + size_t(i) * 3;
            for (int64_t j = 0; j < M; ++j) {
                const T dx = p[0] - bx[j * 3], dy = p[1] - bx[j * 3 + 1], dz = p[2] - bx[j * 3 + 2];
                const T dn = std::sqrt((dx * dx + dy * dy) + dz * dz);
                const T dot = (dx * bn[j * 3] + dy * bn[j * 3 + 1]) + dz * bn[j * 3 + 2];
                g = g + ba[j] * dot / (dn * dn * dn);
            }
            in = g < T(-6.283185307179586);
        } else {
            const T* p = pts + size_t(i) * 2;
            bool on = false;
            for (int64_t j = 0; j < M; ++j) {
                const int64_t k = j + 1 == M ? 0 : j + 1;
                const T ux = bx[j * 2] - p[0], uy = bx[j * 2 + 1] - p[1], vx = bx[k * 2] - p[0], vy = bx[k * 2 + 1] - p[1];
                on = on || std::sqrt(ux * ux + uy * uy) < T(1.0e2) * eps;
                g = g + std::atan2(ux * vy - uy * vx, ux * vx + uy * vy);
            }
            in = on || !(std::fabs(g) < T(1.0e3) * eps);
        }
        out[i] = in ? 1 : 0;
        if (sums) sums[i] = g;
    }
}

extern "C" {

int32_t wtpo_max_threads(void) { return default_threads(0); }

// algo: 0 brute force, 1 KD-tree. drop_first: 1 = _build_knn_neighbors, 0 = search/searchdists.
int32_t wtpo_knn_f32(const float* pts, int64_t N, int32_t D, int32_t k, int32_t drop_first, int32_t algo,
                     int32_t threads, int64_t* out_idx, float* out_dist) {
    threads = default_threads(threads);
    return DISPATCH_D(float, D, (knn_impl<float, 2>(pts, N, k, drop_first, algo, threads, out_idx, out_dist)),
                      (knn_impl<float, 3>(pts, N, k, drop_first, algo, threads, out_idx, out_dist)));
}
int32_t wtpo_knn_f64(const double* pts, int64_t N, int32_t D, int32_t k, int32_t drop_first, int32_t algo,
                     int32_t threads, int64_t* out_idx, double* out_dist) {
    threads = default_threads(threads);
    return DISPATCH_D(double, D, (knn_impl<double, 2>(pts, N, k, drop_first, algo, threads, out_idx, out_dist)),
                      (knn_impl<double, 3>(pts, N, k, drop_first, algo, threads, out_idx, out_dist)));
}

// indices == NULL: fill offsets[N+1]; else fill indices using offsets.
int32_t wtpo_radius_f32(const float* pts, int64_t N, int32_t D, float r, int32_t threads, int64_t* offsets, int64_t* indices) {
    threads = default_threads(threads);
    return DISPATCH_D(float, D, (radius_impl<float, 2>(pts, N, r, threads, offsets, indices)),
                      (radius_impl<float, 3>(pts, N, r, threads, offsets, indices)));
}
int32_t wtpo_radius_f64(const double* pts, int64_t N, int32_t D, double r, int32_t threads, int64_t* offsets, int64_t* indices) {
    threads = default_threads(threads);
    return DISPATCH_D(double, D, (radius_impl<double, 2>(pts, N, r, threads, offsets, indices)),
                      (radius_impl<double, 3>(pts, N, r, threads, offsets, indices)));
}

void wtpo_force_f32(const wtp_force* f, const float* u, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = force_eval<float>(f->kind, float(f->beta), float(f->u0), float(f->gamma), u[i]);
}
void wtpo_force_f64(const wtp_force* f, const double* u, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = force_eval<double>(f->kind, f->beta, f->u0, f->gamma, u[i]);
}

#define SPACING_EVAL(T, DD)                                                   \
    {                                                                         \
        Spacing<T, DD> s; s.init(sp);                                         \
        for (int64_t i = 0; i < N; ++i) out[i] = s(&pts[size_t(i) * DD]);     \
        return WTP_OK;                                                        \
    }
int32_t wtpo_spacing_f32(const wtp_spacing* sp, const float* pts, int64_t N, int32_t D, float* out) {
    if (D == 2) SPACING_EVAL(float, 2) else if (D == 3) SPACING_EVAL(float, 3)
    return WTP_ERR_BAD_ARG;
}
int32_t wtpo_spacing_f64(const wtp_spacing* sp, const double* pts, int64_t N, int32_t D, double* out) {
    if (D == 2) SPACING_EVAL(double, 2) else if (D == 3) SPACING_EVAL(double, 3)
    return WTP_ERR_BAD_ARG;
}

int32_t wtpo_repel_f32(float* snap, int64_t n_fixed, int64_t n_move, int32_t D, const wtp_spacing* sp,
                       const wtp_force* fm, const wtp_repel_params* prm, const wtp_wall_mesh* wall, float* conv,
                       wtp_trace_entry* trace, wtp_repel_result* res, int32_t threads) {
    threads = default_threads(threads);
    return DISPATCH_D(float, D, (relax<float, 2>(snap, n_fixed, n_move, sp, fm, prm, wall, conv, trace, res, threads)),
                      (relax<float, 3>(snap, n_fixed, n_move, sp, fm, prm, wall, conv, trace, res, threads)));
}
int32_t wtpo_repel_f64(double* snap, int64_t n_fixed, int64_t n_move, int32_t D, const wtp_spacing* sp,
                       const wtp_force* fm, const wtp_repel_params* prm, const wtp_wall_mesh* wall, double* conv,
                       wtp_trace_entry* trace, wtp_repel_result* res, int32_t threads) {
    threads = default_threads(threads);
    return DISPATCH_D(double, D, (relax<double, 2>(snap, n_fixed, n_move, sp, fm, prm, wall, conv, trace, res, threads)),
                      (relax<double, 3>(snap, n_fixed, n_move, sp, fm, prm, wall, conv, trace, res, threads)));
}

#define MESH_QUERIES(T, sfx)                                                                                        \
    void wtpo_mesh_isinside_##sfx(const wtp_wall_mesh* m, const T* pts, int64_t N, int32_t threads, uint8_t* out) { \
        threads = default_threads(threads);                                                                         \
        _Pragma("omp parallel for num_threads(threads) schedule(dynamic, 64)")                                      \
        for (int64_t i = 0; i < N; ++i) out[i] = mesh_isinside<T>(m, pts + size_t(i) * 3) ? 1 : 0;                  \
    }                                                                                                               \
    void wtpo_mesh_project_##sfx(const wtp_wall_mesh* m, const T* pts, int64_t N, int32_t threads, T* out_pts, int64_t* out_tri) { \
        threads = default_threads(threads);                                                                         \
        _Pragma("omp parallel for num_threads(threads) schedule(dynamic, 64)")                                      \
        for (int64_t i = 0; i < N; ++i) out_tri[i] = mesh_project<T>(m, pts + size_t(i) * 3, out_pts + size_t(i) * 3); \
    }
void wtpo_isinside_f32(const float* pts, int64_t N, int32_t D, const float* bx, const float* bn, const float* ba, int64_t M, int32_t threads, uint8_t* out, float* sums) {
    isinside_points<float>(pts, N, D, bx, bn, ba, M, default_threads(threads), out, sums);
}
void wtpo_isinside_f64(const double* pts, int64_t N, int32_t D, const double* bx, const double* bn, const double* ba, int64_t M, int32_t threads, uint8_t* out, double* sums) {
    isinside_points<double>(pts, N, D, bx, bn, ba, M, default_threads(threads), out, sums);
}
MESH_QUERIES(float, f32)
MESH_QUERIES(double, f64)

int32_t wtpo_metrics_f32(const float* pts, int64_t N, int32_t D, int32_t k, int32_t threads, wtp_cloud_metrics* out) {
    threads = default_threads(threads);
    return DISPATCH_D(float, D, (metrics_impl<float, 2>(pts, N, k, threads, out)), (metrics_impl<float, 3>(pts, N, k, threads, out)));
}
int32_t wtpo_metrics_f64(const double* pts, int64_t N, int32_t D, int32_t k, int32_t threads, wtp_cloud_metrics* out) {
    threads = default_threads(threads);
    return DISPATCH_D(double, D, (metrics_impl<double, 2>(pts, N, k, threads, out)), (metrics_impl<double, 3>(pts, N, k, threads, out)));
}

int32_t wtpo_cull_mask_f32(const float* pts, int64_t N, int32_t D, const float* s, double ratio, uint8_t* keep) {
    return DISPATCH_D(float, D, (cull_mask_impl<float, 2>(pts, N, s, ratio, keep)), (cull_mask_impl<float, 3>(pts, N, s, ratio, keep)));
}
int32_t wtpo_cull_mask_f64(const double* pts, int64_t N, int32_t D, const double* s, double ratio, uint8_t* keep) {
    return DISPATCH_D(double, D, (cull_mask_impl<double, 2>(pts, N, s, ratio, keep)), (cull_mask_impl<double, 3>(pts, N, s, ratio, keep)));
}
int32_t wtpo_spacing_metrics_f32(const float* pts, int64_t N, int32_t D, int32_t k, const wtp_spacing* sp, int32_t threads, wtp_spacing_metrics_t* out) {
    threads = default_threads(threads);
    return DISPATCH_D(float, D, (spacing_metrics_impl<float, 2>(pts, N, k, sp, threads, out)), (spacing_metrics_impl<float, 3>(pts, N, k, sp, threads, out)));
}
int32_t wtpo_spacing_metrics_f64(const double* pts, int64_t N, int32_t D, int32_t k, const wtp_spacing* sp, int32_t threads, wtp_spacing_metrics_t* out) {
    threads = default_threads(threads);
    return DISPATCH_D(double, D, (spacing_metrics_impl<double, 2>(pts, N, k, sp, threads, out)), (spacing_metrics_impl<double, 3>(pts, N, k, sp, threads, out)));
}
int32_t wtpo_spacing_fidelity_f32(const float* pts, int64_t N, int32_t D, int32_t k, double cr, const wtp_spacing* sp, int32_t threads, wtp_spacing_fidelity_t* out) {
    threads = default_threads(threads);
    return DISPATCH_D(float, D, (spacing_fidelity_impl<float, 2>(pts, N, k, cr, sp, threads, out)), (spacing_fidelity_impl<float, 3>(pts, N, k, cr, sp, threads, out)));
}
int32_t wtpo_spacing_fidelity_f64(const double* pts, int64_t N, int32_t D, int32_t k, double cr, const wtp_spacing* sp, int32_t threads, wtp_spacing_fidelity_t* out) {
    threads = default_threads(threads);
    return DISPATCH_D(double, D, (spacing_fidelity_impl<double, 2>(pts, N, k, cr, sp, threads, out)), (spacing_fidelity_impl<double, 3>(pts, N, k, cr, sp, threads, out)));
}

int32_t wtpo_normals_f32(const float* pts, int64_t N, int32_t D, int32_t k, int32_t threads, float* out) {
    threads = default_threads(threads);
    return DISPATCH_D(float, D, (normals_impl<float, 2>(pts, N, k, threads, out)), (normals_impl<float, 3>(pts, N, k, threads, out)));
}
int32_t wtpo_normals_f64(const double* pts, int64_t N, int32_t D, int32_t k, int32_t threads, double* out) {
    threads = default_threads(threads);
    return DISPATCH_D(double, D, (normals_impl<double, 2>(pts, N, k, threads, out)), (normals_impl<double, 3>(pts, N, k, threads, out)));
}
int32_t wtpo_gradient_limit_f32(const float* c, int64_t n, int32_t D, const float* h0, float g, int32_t k, double tol, int32_t max_sweeps, int32_t threads,
                                float* out, int32_t* sweeps) {
    threads = default_threads(threads);
    return DISPATCH_D(float, D, (gradient_limit_impl<float, 2>(c, n, h0, g, k, tol, max_sweeps, threads, out, sweeps)),
                      (gradient_limit_impl<float, 3>(c, n, h0, g, k, tol, max_sweeps, threads, out, sweeps)));
}
int32_t wtpo_gradient_limit_f64(const double* c, int64_t n, int32_t D, const double* h0, double g, int32_t k, double tol, int32_t max_sweeps, int32_t threads,
                                double* out, int32_t* sweeps) {
    threads = default_threads(threads);
    return DISPATCH_D(double, D, (gradient_limit_impl<double, 2>(c, n, h0, g, k, tol, max_sweeps, threads, out, sweeps)),
                      (gradient_limit_impl<double, 3>(c, n, h0, g, k, tol, max_sweeps, threads, out, sweeps)));
}

int32_t wtpo_closest_point_on_triangle_f64(const double* p, const double* a, const double* b, const double* c, double* out) {
    return closest_point_on_triangle<double>(p, a, b, c, out);
}
int32_t wtpo_closest_point_on_triangle_f32(const float* p, const float* a, const float* b, const float* c, float* out) {
    return closest_point_on_triangle<float>(p, a, b, c, out);
}

}  // extern "C"
