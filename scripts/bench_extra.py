"""Secondary measurements on the GPU box: the other BASELINE configs (graded repel, 2-D radius, f64)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import __graft_entry__ as g
import synth
pkg = g.load_package()
ctx = pkg.Context(0)
ctx.set_timing(True)
out = {}

def show(name, **kw):
    out[name] = kw
    print(name, json.dumps(kw), flush=True)

def t_knn(name, pts, k=21, reps=3):
    idx = np.empty((len(pts), k), dtype=np.int64)
    for _ in range(reps):
        t = time.perf_counter(); ctx.knn(pts, k, out_idx=idx); wall = time.perf_counter() - t
    tm = ctx.timing()
    show(name, n=len(pts), query_ms=tm["ms_query"], index_ms=tm["ms_bbox"] + tm["ms_cellkey"] + tm["ms_sort"] + tm["ms_reorder"],
         mq_s=len(pts) / tm["ms_query"] / 1e3, expanded=tm["n_ring_expanded"], cells=tm["n_cells"], wall_ms=wall * 1e3)

which = sys.argv[1:] or ["knn", "graded", "repel", "radius"]
if "knn" in which:
    for dt in (np.float32, np.float64):
        t_knn(f"knn_uniform3d_1M_{dt.__name__}", synth.uniform_cube(1_000_000, dt))
        t_knn(f"knn_uniform3d_10M_{dt.__name__}", synth.uniform_cube(10_000_000, dt))
    p2 = np.random.default_rng(0).random((4_000_000, 2)).astype(np.float32)
    t_knn("knn_uniform2d_4M_float32", p2)
if "graded" in which:
    pts, nw, hw = synth.graded_cube(2_000_000, np.float32)
    t_knn("knn_graded3d_2M_float32", pts)
    q2, hm = synth.graded_square(4_000_000, np.float32)
    t_knn("knn_graded2d_4M_float32", q2)
if "repel" in which:
    for dt in (np.float64, np.float32):
        pts, nw, hw = synth.graded_cube(2_000_000, dt)
        sp, keep = ctx.make_spacing("boundary_layer", hw, 4 * hw, 0.2, pts[:nw])
        smin = hw
        iters = 20
        for rep in range(2):
            t = time.perf_counter()
            o, conv, res, _ = ctx.repel(pts, nw, sp, ctx.make_force("clipped", 0.2), max_iters=iters, tol=0.0, stall_after=0,
                                        alpha_lo=smin / 2000, alpha_max=smin / 20)
            wall = time.perf_counter() - t
        tm = ctx.timing()
        show(f"repel_graded3d_2M_{dt.__name__}", n=len(pts), n_wall=nw, h_wall=hw, iters=res["iters"], ms_per_iter=tm["ms_total"] / iters,
             sweep_ms=tm["ms_query"] / iters, sort_ms=tm["ms_sort"] / iters, wall_ms_per_iter=wall * 1e3 / iters, conv0=float(conv[0]), conv_last=float(conv[-1]))
        u = synth.uniform_cube(2_000_000, dt, stream=5)
        h = 2_000_000 ** (-1 / 3)
        spc, _ = ctx.make_spacing("constant", h)
        for rep in range(2):
            o, conv, res, _ = ctx.repel(u, 0, spc, ctx.make_force("clipped", 0.2), max_iters=iters, tol=0.0, stall_after=0, alpha_lo=h / 2000, alpha_max=h / 20)
        tm = ctx.timing()
        show(f"repel_uniform3d_2M_{dt.__name__}", iters=res["iters"], ms_per_iter=tm["ms_total"] / iters, sweep_ms=tm["ms_query"] / iters)
if "radius" in which:
    q2, hm = synth.graded_square(10_000_000, np.float64)
    for rep in range(2):
        t = time.perf_counter(); off, ind = ctx.radius(q2, 2.5 * hm); wall = time.perf_counter() - t
    show("radius_graded2d_10M_float64", n=len(q2), nnz=int(off[-1]), wall_ms=wall * 1e3, mpts_s=len(q2) / wall / 1e6, max_row=int(np.diff(off).max()))
    u2 = np.random.default_rng(1).random((10_000_000, 2))
    h = 10_000_000 ** -0.5
    for rep in range(2):
        t = time.perf_counter(); off, ind = ctx.radius(u2, 2.5 * h); wall = time.perf_counter() - t
    show("radius_uniform2d_10M_float64", nnz=int(off[-1]), wall_ms=wall * 1e3, mpts_s=len(u2) / wall / 1e6)
json.dump(out, open(os.path.join("gpurun_out", "bench_extra.json"), "w"), indent=1)
