"""One RadiusTopology build (count + fill) on a 2-D cloud, for ncu. usage: radius_profile.py [uniform|graded] [f32|f64] [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import __graft_entry__ as g
import synth
pkg = g.load_package()
kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
dt = np.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else np.float32
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
ctx = pkg.Context(0)
ctx.set_timing(True)
dev = torch.device("cuda", 0)
if kind == "uniform":
    pts = np.random.default_rng(1).random((n, 2)).astype(dt); r = 2.5 * n ** -0.5
else:
    pts, hm = synth.graded_square(n, dt); r = 2.5 * hm
dp = torch.from_numpy(pts).to(dev)
off = torch.empty(n + 1, dtype=torch.int64, device=dev)
for rep in range(2):
    ctx.radius_dev(dp.data_ptr(), n, 2, r, pts.dtype, off.data_ptr())
    tc = ctx.timing()
    nnz = int(off[-1].item())
    ind = torch.empty(nnz, dtype=torch.int64, device=dev)
    ctx.radius_fill_dev(ind.data_ptr())
    tf = ctx.timing()
    torch.cuda.synchronize()
print(f"{kind} {dt.__name__} n={n} nnz={nnz} count={tc['ms_query']:.3f} ms fill={tf['ms_query']:.3f} ms leftovers(count)={tc['n_leftover_sparse'] + tc['n_leftover_dense'] + tc['n_leftover_other']}")
