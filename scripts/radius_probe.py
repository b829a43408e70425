"""Radius-topology probe: device phases of count / fill on 10M 2-D points (uniform and graded), pinned host buffers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import __graft_entry__ as g
import synth
pkg = g.load_package()
ctx = pkg.Context(0)
ctx.set_timing(True)
dev = torch.device("cuda", 0)
def run(name, pts, r):
    n, d = pts.shape
    dp = torch.from_numpy(pts).to(dev)
    off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.radius_dev(dp.data_ptr(), n, d, r, pts.dtype, off.data_ptr())
        tc = ctx.timing()
        nnz = int(off[-1].item())
        ind = torch.empty(nnz, dtype=torch.int64, device=dev)
        ctx.radius_fill_dev(ind.data_ptr())
        tf = ctx.timing()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    idx_ms = tc["ms_bbox"] + tc["ms_cellkey"] + tc["ms_sort"] + tc["ms_reorder"]
    print(f"{name}: n={n} nnz={nnz} index={idx_ms:.2f} ms count={tc['ms_query']:.2f} ms scan={tc['ms_scan']:.2f} ms fill={tf['ms_query']:.2f} ms "
          f"wall(dev api)={wall*1e3:.1f} ms -> {n/ (idx_ms+tc['ms_query']+tc['ms_scan']+tf['ms_query'])/1e3:.1f} Mpts/s device", flush=True)
for dt in (np.float32, np.float64):
    u2 = np.random.default_rng(1).random((10_000_000, 2)).astype(dt)
    run(f"radius_uniform2d_10M_{dt.__name__}", u2, 2.5 * 10_000_000 ** -0.5)
    q2, hm = synth.graded_square(10_000_000, dt)
    run(f"radius_graded2d_10M_{dt.__name__}", q2, 2.5 * hm)
u3 = np.random.default_rng(2).random((2_000_000, 3)).astype(np.float32)
run("radius_uniform3d_2M_float32", u3, 1.6 * 2_000_000 ** (-1 / 3))
