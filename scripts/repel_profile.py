"""Small repel run for ncu: graded 2M cloud, BoundaryLayerSpacing, a few iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import __graft_entry__ as g
import synth
pkg = g.load_package()
ctx = pkg.Context(0)
ctx.set_timing(True)
dt = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pts, nw, hw = synth.graded_cube(2_000_000, dt)
sp, keep = ctx.make_spacing("boundary_layer", hw, 4 * hw, 0.2, pts[:nw])
o, conv, res, _ = ctx.repel(pts, nw, sp, ctx.make_force("clipped", 0.2), max_iters=iters, tol=0.0, stall_after=0, alpha_lo=hw / 2000, alpha_max=hw / 20)
print({k: round(v, 3) if isinstance(v, float) else v for k, v in ctx.timing().items()})
