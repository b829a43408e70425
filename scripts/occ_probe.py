"""Graded clouds: k-NN / repel time against the grid's target occupancy (points per cell)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import __graft_entry__ as g
import synth
pkg = g.load_package()
ctx = pkg.Context(0)
ctx.set_timing(True)
pts, nw, hw = synth.graded_cube(2_000_000, np.float32)
q2, hm = synth.graded_square(4_000_000, np.float32)
idx = np.empty((len(pts), 21), dtype=np.int64)
idx2 = np.empty((len(q2), 21), dtype=np.int64)
for occ in (0.0, 5.0, 3.0, 2.0, 1.2, 0.7):
    ctx.set_cell_occupancy(occ)
    for _ in range(2):
        ctx.knn(pts, 21, out_idx=idx)
    t = ctx.timing()
    for _ in range(2):
        ctx.knn(q2, 21, out_idx=idx2)
    t2 = ctx.timing()
    sp, keep = ctx.make_spacing("boundary_layer", hw, 4 * hw, 0.2, pts[:nw])
    o, conv, res, _ = ctx.repel(pts, nw, sp, ctx.make_force("clipped", 0.2), max_iters=6, tol=0.0, stall_after=0, alpha_lo=hw / 2000, alpha_max=hw / 20)
    tr = ctx.timing()
    print(json.dumps({"occ": occ, "knn3d_query_ms": round(t["ms_query"], 3), "knn3d_index_ms": round(t["ms_sort"] + t["ms_reorder"] + t["ms_cellkey"], 3), "cells3d": t["n_cells"],
                      "left3d": [t["n_leftover_sparse"], t["n_leftover_dense"], t["n_leftover_other"]], "expanded3d": t["n_ring_expanded"],
                      "knn2d_query_ms": round(t2["ms_query"], 3), "left2d": [t2["n_leftover_sparse"], t2["n_leftover_dense"], t2["n_leftover_other"]],
                      "repel_ms_iter": round(tr["ms_total"] / 6, 3), "repel_sweep": round(tr["ms_query"] / 6, 3), "repel_spacing": round(tr["ms_scan"] / 6, 3),
                      "repel_left": [tr["n_leftover_sparse"], tr["n_leftover_dense"], tr["n_leftover_other"]], "conv": float(conv[-1])}), flush=True)
