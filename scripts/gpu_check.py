"""First-contact GPU battery: parity of every entry point against the oracle + timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); oracle = g.load_oracle()
ctx = pkg.Context(0)
rng = np.random.default_rng(1)
fails = 0
def check(name, ok, extra=""):
    global fails
    print(("PASS " if ok else "FAIL ") + name, extra, flush=True)
    if not ok: fails += 1

for dt in (np.float32, np.float64):
    for D in (2, 3):
        for (N, k) in ((50, 5), (1000, 21), (30000, 21), (5000, 40), (3000, 100)):
            pts = rng.random((N, D)).astype(dt)
            try:
                a, ad = ctx.knn(pts, k, dists=True)
                b, bd = oracle.knn(pts, k, dists=True)
                ok = np.array_equal(a, b) and np.array_equal(ad, bd)
                check(f"knn {dt.__name__} D={D} N={N} k={k}", ok, f"mismatch rows={(a!=b).any(1).sum()}" if not ok else "")
            except Exception as e:
                check(f"knn {dt.__name__} D={D} N={N} k={k}", False, repr(e))
        pts = rng.random((20000, D)).astype(dt)
        r = 0.03 if D == 2 else 0.08
        try:
            off, ind = ctx.radius(pts, r); roff, rind = oracle.radius(pts, r)
            ok = np.array_equal(off, roff) and np.array_equal(ind, rind)
            check(f"radius {dt.__name__} D={D} nnz={roff[-1]} maxrow={np.diff(roff).max()}", ok)
        except Exception as e:
            check(f"radius {dt.__name__} D={D}", False, repr(e))
# clustered / graded / duplicates
pts = np.concatenate([rng.normal(0, 0.01, (5000, 3)), rng.random((5000, 3)) * 10]).astype(np.float32)
a = ctx.knn(pts, 21); b = oracle.knn(pts, 21); check("knn clustered f32", np.array_equal(a, b), str(ctx.timing()) )
pts = np.repeat(rng.random((500, 3)), 4, axis=0).astype(np.float64)
a = ctx.knn(pts, 9); b = oracle.knn(pts, 9); check("knn duplicates f64", np.array_equal(a, b))
grid = np.array([[i * 0.1, j * 0.1] for i in range(5) for j in range(5)])
off, ind = ctx.radius(grid, 0.15); roff, rind = oracle.radius(grid, 0.15); check("radius 5x5 grid", np.array_equal(off, roff) and np.array_equal(ind, rind), str(np.diff(off)))
a2 = ctx.knn(rng.random((40, 3)), 39); b2 = oracle.knn(rng.random((40,3)),39)
# self-including search
pts = rng.random((4000, 3)).astype(np.float32)
a, ad = ctx.knn(pts, 8, include_self=True, dists=True); b, bd = oracle.knn(pts, 8, drop_first=False, dists=True)
check("search incl self", np.array_equal(a, b) and np.array_equal(ad, bd))
# forces / spacing
u = np.linspace(0, 2.5, 101)
for kind in ("inverse", "equilibrium", "clipped", "strong"):
    for dt in (np.float32, np.float64):
        f = ctx.make_force(kind, 0.2, 1.0, 3.0); of = oracle.make_force(kind, 0.2, 1.0, 3.0)
        x = ctx.force_eval(f, u.astype(dt)); y = oracle.force(of, u.astype(dt))
        check(f"force {kind} {dt.__name__}", np.allclose(x, y, rtol=1e-6 if dt == np.float32 else 1e-14, atol=0), f"maxrel={np.max(np.abs(x-y)/np.maximum(np.abs(y),1e-300)):.2e}")
bnd = rng.random((3000, 3)); bnd[:, 0] = 0
q = rng.random((20000, 3))
for kind, a_, b_, c_ in (("loglike", 0.1, 1.5, 0), ("boundary_layer", 0.01, 0.04, 0.2)):
    for dt in (np.float32, np.float64):
        sp, k1 = ctx.make_spacing(kind, a_, b_, c_, bnd.astype(dt)); osp, k2 = oracle.make_spacing(kind, a_, b_, c_, bnd.astype(dt))
        x = ctx.spacing_eval(sp, q.astype(dt)); y = oracle.spacing_eval(osp, q.astype(dt))
        check(f"spacing {kind} {dt.__name__}", np.allclose(x, y, rtol=1e-5 if dt == np.float32 else 1e-13), f"maxrel={np.max(np.abs(x-y)/np.abs(y)):.2e} exact={np.array_equal(x,y)}")
# repel
for dt, tolf in ((np.float64, 1e-6), (np.float32, 1e-3)):
    for D in (2, 3):
        N = 6000; nf = 800
        snap = rng.random((N, D)).astype(dt)
        h = N ** (-1.0 / D)
        for skind in ("constant", "boundary_layer"):
            if skind == "constant":
                sp, k1 = ctx.make_spacing("constant", h); osp, k2 = oracle.make_spacing("constant", h)
            else:
                sp, k1 = ctx.make_spacing("boundary_layer", 0.7 * h, 1.3 * h, 0.2, snap[:nf]); osp, k2 = oracle.make_spacing("boundary_layer", 0.7 * h, 1.3 * h, 0.2, snap[:nf])
            kw = dict(k=21, max_iters=10, tol=0.0, stall_after=0, alpha_lo=h / 2000, alpha_max=h / 20, trace=True)
            out, conv, res, tr = ctx.repel(snap, nf, sp, ctx.make_force("clipped", 0.2), **kw)
            oout, oconv, ores, otr = oracle.repel(snap, nf, osp, oracle.make_force("clipped", 0.2), **kw)
            err = np.abs(out - oout).max() / h
            check(f"repel {dt.__name__} D={D} {skind}", err <= tolf and res["iters"] == ores["iters"],
                  f"max|dx|/h={err:.2e} conv_rel={np.max(np.abs(conv-oconv)/np.abs(oconv)):.2e} trace_eq={[ (t['idx_a'],t['idx_b']) for t in tr]==[(t['idx_a'],t['idx_b']) for t in otr]}")
# stop logic
snap = rng.random((4000, 3)); h = 4000 ** (-1 / 3)
sp, _ = ctx.make_spacing("constant", h); osp, _ = oracle.make_spacing("constant", h)
for kw in (dict(stall_after=5, tol=1e-12, max_iters=200), dict(cv_target=10.0, tol=1e-12, max_iters=50, stall_after=0), dict(tol=1e6, max_iters=50, stall_after=0), dict(rebuild_every=3, max_iters=12, tol=0.0, stall_after=0)):
    out, conv, res, _ = ctx.repel(snap, 500, sp, ctx.make_force("clipped", 0.2), alpha_lo=h/2000, alpha_max=h/20, **kw)
    oout, oconv, ores, _ = oracle.repel(snap, 500, osp, oracle.make_force("clipped", 0.2), alpha_lo=h/2000, alpha_max=h/20, **kw)
    check(f"repel stop {kw}", res["iters"] == ores["iters"] and res["stop_reason"] == ores["stop_reason"] and np.abs(out-oout).max() <= 1e-6*h, f"{res} vs {ores} err={np.abs(out-oout).max()/h:.2e}")
# metrics
m = ctx.metrics(snap, 20); om = oracle.metrics(snap, 20)
check("metrics", all(abs(m[k] - om[k]) <= 1e-9 * abs(om[k]) for k in om), str(m))
# timings
ctx.set_timing(True)
for N in (1_000_000, 10_000_000):
    pts = rng.random((N, 3)).astype(np.float32)
    idx = np.empty((N, 21), dtype=np.int64)
    for rep in range(2):
        t = time.time(); ctx.knn(pts, 21, out_idx=idx); dt_ = time.time() - t
    tm = ctx.timing()
    print(f"knn N={N}: wall {dt_*1e3:.1f} ms  timing={ {k: round(v,3) if isinstance(v,float) else v for k,v in tm.items()} }", flush=True)
    if N == 1_000_000:
        b = oracle.knn(pts, 21); check("knn 1M f32 parity", np.array_equal(idx, b))
N = 2_000_000
snap = rng.random((N, 3)); h = N ** (-1/3)
sp, _ = ctx.make_spacing("constant", h)
t = time.time(); out, conv, res, _ = ctx.repel(snap, 100000, sp, ctx.make_force("clipped", 0.2), max_iters=20, tol=0.0, stall_after=0, alpha_lo=h/2000, alpha_max=h/20); dt_ = time.time() - t
tm = ctx.timing()
print(f"repel f64 N={N} 20 iters: wall {dt_*1e3:.1f} ms timing={ {k: round(v,3) if isinstance(v,float) else v for k,v in tm.items()} }")
print("FAILS", fails)
sys.exit(1 if fails else 0)
