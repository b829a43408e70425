// kernels.cuh — host-side entry points of the CUDA translation units.
#pragma once
#include "common.cuh"

namespace wtp {

// ---- grid.cu -------------------------------------------------------------
// Bounding box of N x D points on the device (one small D2H + stream sync).
template <class T>
void compute_bbox(wtp_ctx* ctx, IndexBuffers& ib, const T* d_pts, int64_t N, int D, double lo[3], double hi[3]);

// Grid parameters from a bounding box: cell size from the target occupancy, never
// below min_cell (radius search), cell count capped.
template <class T>
Grid<T> make_grid(int64_t N, int D, const double lo[3], const double hi[3], double occupancy, double min_cell, int K);

// cell keys -> radix sort -> sorted P4 tiles + cell starts. Returns radix passes run.
template <class T>
int build_index(wtp_ctx* ctx, IndexBuffers& ib, const T* d_pts, int64_t N, int D, const Grid<T>& g);

// The same for the part of the grid a sharded k-NN rank needs: the layers of the sorted positions [sb, se) plus `halo`
// layers on either side (see grid.cu). On success g.w_lo / g.w_hi and ib.cs_rebase describe the window and the
// sorted records are the run [win->P0, win->P0 + win->M) of the whole sorted order. False: not worthwhile, nothing built.
template <class T>
bool build_index_window(wtp_ctx* ctx, IndexBuffers& ib, const T* d_pts, int64_t N, int D, Grid<T>& g, int64_t sb, int64_t se, int halo,
                        IndexWindow* win, int* passes_out);

// Stable LSD radix sort (8-bit digits) of the pairs in ib.keys_a / ib.vals_a on the low `bits` key bits.
int radix_sort_pairs(wtp_ctx* ctx, IndexBuffers& ib, int64_t N, int bits);

// Exclusive scans (count -> offsets). out has n+1 entries; out[n] = total.
void exclusive_scan_u32(wtp_ctx* ctx, DevBuf& tmp, const uint32_t* d_in, uint32_t* d_out, int64_t n);
void exclusive_scan_u32_to_i64(wtp_ctx* ctx, DevBuf& tmp, const uint32_t* d_in, int64_t* d_out, int64_t n);

// ---- knn.cu --------------------------------------------------------------
// Warp-per-query exact k-NN on a built index. Queries are the index's own points,
// restricted to caller-order indices [q_begin, q_end) (d_qlist: their sorted positions,
// or null = all). K1 = list length; drop_first drops rank 0 from the output.
// Rows are written where `rows` says (RowMap, common.cuh).
template <class T>
void knn_query(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, int K1, int drop_first,
               const uint32_t* d_qlist, int64_t n_queries, const RowMap& rows, void* d_out_idx, T* d_out_dist,
               unsigned long long* d_expanded_counter, bool out32 = false);   // out32: uint32 rows instead of int64; counter[1]: window misses

// CTA-tiled front end + general kernel for the leftovers, over the sorted positions [s_begin, s_end)
// (K1 <= 32). Rows are written at (orig - q_begin) * (K1 - drop_first) like knn_query. Fully asynchronous.
template <class T>
void knn_query_tiled(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, int K1, int drop_first,
                     int64_t s_begin, int64_t s_end, const RowMap& rows, void* d_out_idx, T* d_out_dist,
                     unsigned long long* d_expanded_counter, bool out32 = false);
// K nearest index points (1-based caller indices, uint32 rows of K, ascending (d2, index)) of n_q arbitrary query points
template <class T>
void knn_points(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int D, int K, const T* d_q, int64_t n_q, uint32_t* d_out_idx32);
// caller indices (1-based) of the sorted positions [s_begin, s_end)
void owned_ids(wtp_ctx* ctx, const IndexBuffers& ib, bool f64, int64_t s_begin, int64_t s_end, int64_t* d_ids);
// the n rows (uint32, k wide; optionally their distances) of the run starting at sorted position s_begin, reordered by
// ascending caller index (uses the index's radix sort scratch); d_ids_out: the caller indices + 1 in that order
template <class T>
void rows_by_caller_index(wtp_ctx* ctx, IndexBuffers& ib, int64_t N, int64_t s_begin, int64_t n, int k, const uint32_t* d_rows, const T* d_dist,
                          uint32_t* d_rows_out, T* d_dist_out, uint32_t* d_ids_out);
void owned_ids32(wtp_ctx* ctx, const IndexBuffers& ib, bool f64, int64_t s_begin, int64_t s_end, uint32_t* d_ids);   // the same as 4-byte values

// Sinks of a tiled pass inside ctx->d_fail: 16 counters (zeroed here), then the list of up to n sorted positions.
// slot / slots: one of several sinks that are alive at the same time (the density classes of a graded repel).
TileFails tile_fails(wtp_ctx* ctx, int64_t n, int slot = 0, int slots = 1);

// Compact list of sorted positions whose original index is in [q_begin, q_end).
void build_query_list(wtp_ctx* ctx, const IndexBuffers& ib, int64_t N, int64_t q_begin, int64_t q_end,
                      bool f64, DevBuf& flags, DevBuf& scan, DevBuf& qlist);

// ---- radius.cu -----------------------------------------------------------
template <class T>
void radius_count(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, T r,
                  const uint32_t* d_qlist, int64_t n_queries, int64_t q_begin, uint32_t* d_counts);
template <class T>
void radius_fill(wtp_ctx* ctx, const IndexBuffers& ib, const Grid<T>& g, int64_t N, int D, T r,
                 const uint32_t* d_qlist, int64_t n_queries, int64_t q_begin, const int64_t* d_offsets,
                 int64_t* d_indices);

// ---- bvh.cu --------------------------------------------------------------
template <class T> struct ForceP { int kind; T beta, u0, gamma; };
template <class T> struct SpacingP { int kind; T a, b, c; };
template <class T> struct Box { T lo[3], hi[3]; };
template <class T> struct BvhView { const P4<T>* pts; const Box<T>* boxes; int64_t n, leaf_pow2; };

// Morton sort + bottom-up boxes over the spacing's boundary set (device pointer, n x D).
template <class T>
void bvh_build(wtp_ctx* ctx, BvhBuffers& bv, const T* d_bnd, int64_t n, int D);
template <class T>
BvhView<T> bvh_view(const BvhBuffers& bv);
// out[i] = spacing(pts[i]) for i < n (thread per point; BVH 1-NN for the variable kinds).
template <class T>
void spacing_eval(wtp_ctx* ctx, const SpacingP<T>& sp, const BvhBuffers& bv, const T* d_pts, int64_t n, int D, T* d_out,
                  uint32_t* d_nn_cache = nullptr, bool use_cache = false);
// the same for the movable points [n_fixed, n_fixed + n) visited in the order of a sorted copy of the snapshot
template <class T>
void spacing_eval_ordered(wtp_ctx* ctx, const SpacingP<T>& sp, const BvhBuffers& bv, const T* d_pts, int64_t n, int D, T* d_out,
                          uint32_t* d_nn_cache, bool use_cache, const P4<T>* d_order, int64_t n_order, int64_t n_fixed);
template <class T>
void force_eval(wtp_ctx* ctx, const ForceP<T>& f, const T* d_u, int64_t n, T* d_out);
// inner boxes of a complete binary tree in heap layout whose P leaves (boxes[P..2P)) are set
template <class T>
void bvh_build_levels(wtp_ctx* ctx, Box<T>* boxes, int64_t P);

// ---- mesh.cu -------------------------------------------------------------
// Triangle-mesh queries of the wall rule (_constrain_octree, src/repel.jl:448-469): Morton BVH
// over the triangles, exact nearest triangle by (d2, triangle index), Ericson closest point,
// pseudonormal-signed inside test.
// H2D of the TriangleIndex arrays (host pointers in `wall`) + BVH build on the device.
template <class T>
void mesh_build(wtp_ctx* ctx, MeshBuffers& mb, const wtp_wall_mesh* wall);
template <class T>
void mesh_isinside(wtp_ctx* ctx, const MeshBuffers& mb, const T* d_pts, int64_t n, uint8_t* d_out);
template <class T>
void mesh_project(wtp_ctx* ctx, const MeshBuffers& mb, const T* d_pts, int64_t n, T* d_out_pts, int64_t* d_out_tri);
// Wall rule on the movable ids [id_lo, id_hi): P_new holds the proposals on entry and the
// constrained positions on return; mb.tri_idx / mb.escaped (n_move entries) are updated.
template <class T>
void mesh_wall_apply(wtp_ctx* ctx, MeshBuffers& mb, const T* d_P_old, T* d_P_new, int64_t id_lo, int64_t id_hi);

// ---- inside.cu -----------------------------------------------------------
// isinside(points, cloud): 3-D Green's-function sum over the boundary elements (positions, unit normals, areas);
// 2-D winding number over the ordered boundary polygon (src/isinside.jl). d_g / d_w (nullable): the raw sums.
template <class T>
void greens_isinside(wtp_ctx* ctx, const T* d_pts, int64_t n, const T* d_bx, const T* d_bn, const T* d_ba, int64_t m, T* d_g, uint8_t* d_inside);
template <class T>
void winding_isinside(wtp_ctx* ctx, const T* d_pts, int64_t n, const T* d_poly, int64_t m, T* d_w, uint8_t* d_inside);

}  // namespace wtp
