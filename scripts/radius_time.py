"""Device time of the radius CSR step (index + count + scan + fill) on config #4 (10 M quadtree-graded 2-D points, Float64)
and on a uniform 2-D cloud of the same size (Float32): CUDA events on the launching stream, 5 steps after 2 warm-ups."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np, torch
import __graft_entry__ as g
import synth

pkg = g.load_package()
ctx = pkg.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
ctx.set_stream(stream.cuda_stream)
ctx.set_timing(True)
cases = []
q2, hm = synth.graded_square(10_000_000, np.float64)
cases.append(("config4 graded f64", q2, 2.5 * hm))
rng = np.random.default_rng(5)
u2 = rng.random((10_000_000, 2), dtype=np.float32)
cases.append(("uniform 2-D f32", u2, 2.5 / np.sqrt(10_000_000)))
for name, pts, r in cases:
    d = torch.from_numpy(pts).to(dev)
    off = torch.empty(len(pts) + 1, dtype=torch.int64, device=dev)
    nnz = ctx.radius_dev(d.data_ptr(), len(pts), 2, float(r), pts.dtype.type, off.data_ptr())
    ind = torch.empty(max(nnz, 1), dtype=torch.int64, device=dev)
    def step():
        ctx.radius_dev(d.data_ptr(), len(pts), 2, float(r), pts.dtype.type, off.data_ptr())
        ctx.radius_fill_dev(ind.data_ptr())
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): step()
    e1.record(stream)
    torch.cuda.synchronize()
    t = ctx.timing()
    print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms per step, nnz {nnz}, fill phase {t['ms_query']:.3f} ms, leftovers "
          f"{int(t['n_leftover_sparse'])}/{int(t['n_leftover_dense'])}/{int(t['n_leftover_other'])}", flush=True)
    del d, off, ind
ctx.close()
