"""Config #3 (graded cube 2 M, BoundaryLayerSpacing), a few repel iterations: run under
ncu --metrics gpu__time_duration.sum to get the per-kernel share of an iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import __graft_entry__ as g
import synth
pkg = g.load_package()
dt = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = pkg.Context(0)
gp, nw, hw = synth.graded_cube(2_000_000, dt)
dev = torch.device("cuda", 0)
d_g = torch.from_numpy(gp).to(dev)
d_b = d_g[:nw].clone()
sp, _ = ctx.make_spacing("boundary_layer", hw, 4 * hw, 0.2, bnd_ptr=d_b.data_ptr(), n_bnd=nw)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
ctx.set_timing(True)
conv, res = ctx.repel_dev(d_g.data_ptr(), nw, len(gp) - nw, 3, dt, sp, ctx.make_force("clipped", 0.2), k=21, max_iters=iters, tol=0.0,
                          stall_after=0, alpha_lo=hw / 2000, alpha_max=hw / 20)
torch.cuda.synchronize()
t = ctx.timing()
print({k: round(v, 3) if isinstance(v, float) else v for k, v in t.items()})
