"""python scripts/multidevice_check.py G — ONE process, one context over G GPUs (wtp_create_multi): the host entry points
that shard must fill the caller's arrays exactly like a single-device context and like the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg, oracle = g.load_package(), g.load_oracle()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = pkg.Context(devices=list(range(G)))
one = pkg.Context(0)
ok = True
def check(name, cond):
    global ok
    print(f"{'PASS' if cond else 'FAIL'} {name}", flush=True)
    ok = ok and bool(cond)
rng = np.random.default_rng(11)
for dt in (np.float32, np.float64):
    for D in (2, 3):
        pts = rng.random((120001, D)).astype(dt)
        ref = oracle.knn(pts, 21)
        check(f"knn {dt.__name__} D={D}", np.array_equal(ctx.knn(pts, 21), ref))
        idx, dist = ctx.knn(pts, 9, dists=True, include_self=True)
        ridx, rdist = oracle.knn(pts, 9, drop_first=False, dists=True)
        check(f"searchdists {dt.__name__} D={D}", np.array_equal(idx, ridx) and np.array_equal(dist, rdist))
        r = 0.012 if D == 2 else 0.05
        off, ind = ctx.radius(pts, r)
        roff, rind = oracle.radius(pts, r)
        check(f"radius CSR {dt.__name__} D={D} nnz={roff[-1]}", np.array_equal(off, roff) and np.array_equal(ind, rind))
        nf = 5001
        h = len(pts) ** (-1.0 / D)
        for skind in ("constant", "boundary_layer"):
            args = ("constant", h, 0, 0, None) if skind == "constant" else ("boundary_layer", 0.7 * h, 1.3 * h, 0.2, pts[:nf])
            sp, k1 = ctx.make_spacing(*args); osp, k2 = oracle.make_spacing(*args)
            kw = dict(max_iters=6, tol=0.0, stall_after=0, alpha_lo=0.7 * h / 2000, alpha_max=0.7 * h / 20, trace=True)
            out, conv, res, tr = ctx.repel(pts, nf, sp, ctx.make_force("clipped", 0.2), **kw)
            oout, oconv, ores, otr = oracle.repel(pts, nf, osp, oracle.make_force("clipped", 0.2), **kw)
            tol = (1e-6 if dt == np.float64 else 1e-3) * 0.7 * h
            check(f"repel {dt.__name__} D={D} {skind} err={np.abs(out - oout).max() / h:.2e}",
                  np.abs(out - oout).max() <= tol and res["iters"] == 6 and np.allclose(conv, oconv, rtol=1e-5)
                  and [(t["idx_a"], t["idx_b"]) for t in tr] == [(t["idx_a"], t["idx_b"]) for t in otr])
# small sets and the entry points that do not shard run on the first device
small = rng.random((500, 3))
check("small knn (single-device path)", np.array_equal(ctx.knn(small, 5), oracle.knn(small, 5)))
m, om = ctx.metrics(pts, 10), oracle.metrics(pts, 10)
check("metrics (first device)", all(abs(m[k] - om[k]) <= 1e-9 * abs(om[k]) for k in om))
try:
    ctx.knn_dev(0, 10, 3, 2, np.float32, 0)
    check("device-pointer entry points are refused", False)
except pkg.WtpError:
    check("device-pointer entry points are refused", True)
# the bench size: all devices together vs one
big = np.random.default_rng(3).random((10_000_000, 3)).astype(np.float32)
outm, outs = np.empty((len(big), 21), dtype=np.int64), np.empty((len(big), 21), dtype=np.int64)
for c, o, name in ((one, outs, "1 device"), (ctx, outm, f"{G} devices, one process")):
    for _ in range(2):
        c.knn(big, 21, out_idx=o)
    t = time.perf_counter(); c.knn(big, 21, out_idx=o); dt_ = time.perf_counter() - t
    print(f"wtp_knn_f32 10M k=21, pageable int64 table out, {name}: {dt_ * 1e3:.1f} ms ({len(big) / dt_ / 1e6:.0f} Mq/s)", flush=True)
check("10M rows: multi-device == single-device", np.array_equal(outm, outs))
print("MULTIDEVICE_CHECK", "OK" if ok else "FAILED", flush=True)
ctx.close(); one.close()
sys.exit(0 if ok else 1)
