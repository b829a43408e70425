"""torchrun --nproc-per-node G scripts/multigpu_check.py — parity of the sharded device paths:
every rank's k-NN / radius rows and the NCCL-all-gathered repel against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as g
pkg, oracle = g.load_package(), g.load_oracle()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = pkg.Context(local)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(pkg.Context.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
ok = True
def check(name, cond):
    global ok
    print(f"[rank {rank}] {'PASS' if cond else 'FAIL'} {name}", flush=True)
    ok = ok and bool(cond)
rng = np.random.default_rng(5)
for dt in (np.float32, np.float64):
    for D in (2, 3):
        pts = rng.random((50001, D)).astype(dt)
        b, e = ctx.shard(len(pts))
        idx = np.zeros((len(pts), 21), dtype=np.int64)
        ctx.knn(pts, 21, out_idx=idx)
        ref = oracle.knn(pts, 21)
        own = ctx.owned() - 1                                   # this rank's points: a contiguous run of the sorted order
        rest = np.ones(len(pts), dtype=bool); rest[own] = False
        counts = torch.zeros(len(pts), dtype=torch.int32, device=dev); counts[torch.from_numpy(own).to(dev)] += 1
        dist.all_reduce(counts)                                 # every point is owned by exactly one rank
        check(f"knn shard rows {dt.__name__} D={D}", len(own) == e - b and np.array_equal(idx[own], ref[own]) and (idx[rest] == 0).all()
              and bool((counts == 1).all().item()))
        off, ind = ctx.radius(pts, 0.02 if D == 2 else 0.06)
        roff, rind = oracle.radius(pts, 0.02 if D == 2 else 0.06)
        check(f"radius shard CSR {dt.__name__} D={D}", np.array_equal(off, roff[b:e + 1] - roff[b]) and np.array_equal(ind, rind[roff[b]:roff[e]]))
        nf = 3001
        h = len(pts) ** (-1.0 / D)
        for skind in ("constant", "boundary_layer"):
            args = ("constant", h, 0, 0, None) if skind == "constant" else ("boundary_layer", 0.7 * h, 1.3 * h, 0.2, pts[:nf])
            sp, k1 = ctx.make_spacing(*args); osp, k2 = oracle.make_spacing(*args)
            kw = dict(max_iters=6, tol=0.0, stall_after=0, alpha_lo=0.7 * h / 2000, alpha_max=0.7 * h / 20, trace=True)
            out, conv, res, tr = ctx.repel(pts, nf, sp, ctx.make_force("clipped", 0.2), **kw)
            oout, oconv, ores, otr = oracle.repel(pts, nf, osp, oracle.make_force("clipped", 0.2), **kw)
            tol = (1e-6 if dt == np.float64 else 1e-3) * 0.7 * h
            check(f"repel sharded {dt.__name__} D={D} {skind} err={np.abs(out - oout).max() / h:.2e}",
                  np.abs(out - oout).max() <= tol and res["iters"] == 6 and np.allclose(conv, oconv, rtol=1e-5)
                  and [(t["idx_a"], t["idx_b"]) for t in tr] == [(t["idx_a"], t["idx_b"]) for t in otr])
        # stop logic must agree on every rank
        sp, _ = ctx.make_spacing("constant", h); osp, _ = oracle.make_spacing("constant", h)
        out, conv, res, _ = ctx.repel(pts, nf, sp, ctx.make_force("clipped", 0.2), max_iters=50, tol=1e-12, cv_target=10.0, stall_after=0, alpha_lo=h / 2000, alpha_max=h / 20)
        check(f"repel cv_target stop {dt.__name__} D={D}", res["iters"] == 1 and res["stop_reason"] == "cv_target" and np.array_equal(out, pts))
# the opt-in path of the row exchange: a pinned table, int64 written by the copy engines (WTP_D2H_DIRECT=1)
pts = rng.random((200003, 3)).astype(np.float32)
ref = oracle.knn(pts, 21)
pinned = torch.zeros((len(pts), 21), dtype=torch.int64).pin_memory()
os.environ["WTP_D2H_DIRECT"] = "1"
ctx.knn(pts, 21, out_idx=pinned.numpy())
os.environ.pop("WTP_D2H_DIRECT")
own = ctx.owned() - 1
b, e = ctx.shard(len(pts))
tm = ctx.timing()
check(f"knn row exchange, direct int64 DMA (bytes_d2h={tm['bytes_d2h']})", np.array_equal(own, np.arange(b, e)) and np.array_equal(pinned.numpy()[own], ref[own])
      and tm["n_peer_ranks"] == world and tm["bytes_d2h"] == (e - b) * 21 * 8)
staged = np.zeros((len(pts), 21), dtype=np.int64)
ctx.knn(pts, 21, out_idx=staged)
tm = ctx.timing()
check(f"knn row exchange, staged 3-byte indices (bytes_d2h={tm['bytes_d2h']}, bytes_h2d={tm['bytes_h2d']})", np.array_equal(staged[own], ref[own]) and (staged[:b] == 0).all() and (staged[e:] == 0).all()
      and tm["bytes_d2h"] < (e - b) * 21 * 4 and tm["bytes_h2d"] == (e - b) * 3 * 4)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_CHECK", "OK" if int(t.item()) == 1 else "FAILED", flush=True)
ctx.close()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
