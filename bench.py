#!/usr/bin/env python
"""bench.py — headline benchmark of the WhatsThePoint hot path on B200.

Workload (BASELINE.json `metric`): KNNTopology k=21 on a synthetic uniform 3-D cloud of
10 M float32 points ("U3(10M)"). One step = one full set_topology pass:
bounding box -> cell keys -> radix sort -> gather -> tiled k-NN, N x 21 int64 out.

  value   : Mqueries/s with the points already resident in HBM (wtp_knn_dev_f32), CUDA
            events on the launching stream, max over ranks.
  e2e     : the same metric through the host C-ABI call a Julia user makes (wtp_knn_f32):
            pinned host points in, N x 21 int64 table in host memory out, copies inside the timed region
            (the rows cross PCIe as 3- or 4-byte indices and are widened on the host by the library).
  roofline: the k-NN query kernels, 96 algorithmic bytes per query, timed by
            CUDA events inside the library on the same stream, against MEASURED_PEAKS.json.
            `traffic` is the ncu dram__bytes of the same launch at the same N when a capture of
            this round is committed (profiles/r02_knn_traffic_n<N>.json), else null.
  cpu_baseline: the CPU oracle (KD-tree port of the reference's path) on the same cloud.
  repel   : extra object — iterations/s of the fused repel sweep on the same cloud size.
  extras  : the other BASELINE configurations, each with its own roofline object: k-NN on the
            10 M cloud in Float64, config #3 (repel on the 2 M graded cube, Float32 and Float64),
            config #4 (radius CSR on the 10 M quadtree-graded square), the int32 device table, and config #5's size
            (k-NN + 3 repel iterations on 100 M points generated on the device; --points5 0 skips it).
  parity_check: outside the timed regions every rank brute-forces 256 of the rows it
            answered and 64 positions of one sharded repel sweep in numpy.

`--impl reference` times the CPU port (the reference itself is Julia and cannot run here) on
the SAME cloud (all 10 M points per step) with every core of the process's affinity mask.

Multi-GPU (torchrun, one rank per GPU): the point set is replicated, every rank indexes the
window of the grid around its contiguous 1/N run of the sorted order and answers that run (no
data-path collective); repel sweeps the same kind of run and its kernels store the moved points
into every rank's buffer over NVLink peer memory (NCCL all-gather as the fallback). Total work
is fixed: "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

K = 21
ALGO_BYTES_PER_QUERY = {"f32": 96.0, "f64": 108.0}    # read D*T coords + write 21 x 4 B indices
ALGO_BYTES_REPEL = {"f32": 136.0, "f64": 196.0}       # per point per iteration incl. index rebuild


def synth_uniform(n: int, seed: int = 0x57545031) -> np.ndarray:
    """U3(N): iid uniform in [0,1)^3, float64 draw rounded once to float32 (scripts/synth.uniform_cube, stream 0)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    return rng.random((n, 3)).astype(np.float32)


def workload_config(n: int) -> dict:
    """The `config` object of BOTH arms (the driver compares them): the workload only, nothing about how it is run."""
    return {"workload": f"U3({n}) uniform 3-D unit cube, KNNTopology k=21, float32, N x 21 int64 out", "points": n, "k": K}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(name: str, world: int):
    """dram__bytes_read + dram__bytes_write per launch of the kernel from an ncu --set full capture of THIS round at
    THIS GPU count (profiles/r02_<name>_traffic_n<world>.json, written from the .ncu-rep by scripts/ncu_traffic.py), or
    None: a number from another configuration is not evidence for this line."""
    p = os.path.join(ROOT, "profiles", f"r02_{name}_traffic_n{world}.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


def roofline(kernel, algo_bytes, kernel_ms, traffic=None, note=None, **more):
    hbm, how = peaks()
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms and kernel_ms > 0 else None
    r = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": hbm, "unit": "GB/s",
         "frac": achieved / hbm if achieved else None, "traffic": traffic, "peak_source": how,
         "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms}
    if note:
        r["note"] = note
    r.update(more)
    return r


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML every 5 ms during the timed region. NVML is initialised
    when the sampler is made (outside the timed region), so that the first sample is taken as the region opens."""

    def __init__(self, index: int):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)
        self.max_mhz, self.nv, self.h, self.error = None, None, None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover - NVML missing
            self.error = repr(e)

    def _sample(self):
        nv, h = self.nv, self.h
        reasons = (nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                   else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons)))

    def _run(self):
        if self.nv is None:
            return
        try:
            while not self.stop.is_set():
                self._sample()
                self.stop.wait(0.005)
            self._sample()
        except Exception as e:  # pragma: no cover
            self.error = repr(e)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [s[0] for s in self.samples]
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted({name for s in self.samples for bit, name in bits.items() if s[1] & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------ numpy checkers (parity_check, tests)
def _slab_order(pts):
    """Sort once by x so that the candidates of a sample are a contiguous slice (brute force inside the slab only)."""
    order = np.argsort(pts[:, 0], kind="stable")
    return order, np.ascontiguousarray(pts[order, 0])


def _candidates(pts, order, xs, x, half):
    lo, hi = np.searchsorted(xs, x - half), np.searchsorted(xs, x + half, side="right")
    return order[lo:hi]


def sample_rows_brute_force(pts: np.ndarray, qi, k: int, half: float | None = None) -> np.ndarray:
    """Canonical (d2, index) k nearest OTHER points of the sampled queries, in the input precision without FMA (numpy
    evaluates ((dx*dx + dy*dy) + dz*dz) operation by operation), 1-based. Candidates come from an x-slab of half-width
    `half` (default: wide enough for ~40 k points per unit density); the result is checked to lie inside it."""
    n, d = pts.shape
    if half is None:
        half = min(1.0, 1.3 * (float(k + 1) / n) ** (1.0 / d)) * float(np.ptp(pts[:, 0]) or 1.0)   # ~2x the k-th neighbour distance of a uniform cloud
    order, xs = _slab_order(pts)
    out = np.empty((len(qi), k), dtype=np.int64)
    for a, i in enumerate(qi):
        while True:
            cand = _candidates(pts, order, xs, pts[i, 0], pts.dtype.type(half))
            diff = pts[cand] - pts[i]
            d2 = diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]
            if d == 3:
                d2 = d2 + diff[:, 2] * diff[:, 2]
            if len(cand) > k:
                sel = np.lexsort((cand, d2))[:k + 1]
                if float(d2[sel[-1]]) < float(half) ** 2 or len(cand) == n:
                    break
            half *= 2.0
        out[a] = cand[sel[1:]] + 1        # position 0 is the query itself (d2 = 0, or a lower-indexed twin: n[2:end] drops it either way)
    return out


def sample_repel_sweep(snap: np.ndarray, n_fixed: int, ids, s: float, beta: float, alpha_lo: float, alpha_max: float, k: int = K):
    """One Jacobi sweep of _relax! (src/repel.jl:256-292) for the sampled movable ids, constant spacing s, clipped force
    (u0 = 1), identity wall — in float64 numpy on the given coordinates. Returns the new positions of those points."""
    pts = snap.astype(np.float64)
    n, d = pts.shape
    half = min(1.0, 1.3 * (float(k) / n) ** (1.0 / d)) * float(np.ptp(pts[:, 0]) or 1.0)
    order, xs = _slab_order(pts)
    out = np.empty((len(ids), d))
    for a, mid in enumerate(ids):
        i = int(mid) + n_fixed
        h = half
        while True:
            cand = _candidates(pts, order, xs, pts[i, 0], h)
            diff = pts[i] - pts[cand]
            d2 = (diff * diff).sum(1)
            sel = np.lexsort((cand, d2))[:k]
            if len(cand) >= k and (d2[sel[-1]] < h * h or len(cand) == n):
                break
            h *= 2.0
        F = np.zeros(d)
        for j in sel:
            if cand[j] == i:
                continue
            r = np.sqrt(d2[j])
            u = r / s
            f = max((1.0 - u * u) / (u * u + beta) ** 2, 0.0)
            if r > 0:
                F += f * diff[j] / r
        fn = np.linalg.norm(F)
        ai = min(max(1.0 / (fn + 1e-30), alpha_lo), alpha_max)
        disp = s * ai * F
        dn = np.linalg.norm(disp)
        if dn > s:
            disp *= s / dn
        out[a] = pts[i] + disp
    return out


# ------------------------------------------------------------------------------ the CPU arm
def run_reference(args, n_points):
    """CPU arm: the oracle's KD-tree k-NN (port of the reference's NearestNeighbors path) on the same cloud, every
    step the whole problem (tree build + N queries), with all the cores of this process's affinity mask."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    pts = synth_uniform(n_points)
    threads = oracle.host_threads()                      # not OpenMP's env: torchrun exports OMP_NUM_THREADS=1
    for _ in range(args.warmup):
        oracle.knn(pts[: max(n_points // 10, 1000)], K, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.knn(pts, K, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = n_points / dt / 1e6
    s1 = max(n_points // 80, 1000)
    t0 = time.perf_counter()
    oracle.knn(pts[:s1], K, threads=1)
    one = s1 / (time.perf_counter() - t0) / 1e6
    line = {
        "impl": "reference", "metric": "knn_k21_Mqueries_per_s", "value": val, "unit": "Mqueries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_points),
        "cpu_baseline": {"value": val, "unit": "Mqueries/s", "cores": threads, "kind": "port",
                         "sample": f"the whole cloud per step: KD-tree build + {n_points} queries, {threads} threads (sched_getaffinity); warm-up steps "
                                   f"on a tenth of it; reference-faithful single thread (set_topology is serial, src/topology.jl:81) on {s1} points: "
                                   f"{one:.3f} Mqueries/s",
                         "single_thread_value": one},
        "e2e": {"value": val, "unit": "Mqueries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is 100% Julia (not runnable in this image); this arm is oracle/wtp_oracle.cpp, the CPU port of its path, "
                "on all host threads (the reference's own set_topology is single-threaded, src/topology.jl:81)",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--repel-points", type=int, default=10_000_000)
    ap.add_argument("--repel-iters", type=int, default=20)
    ap.add_argument("--points5", type=int, default=100_000_000, help="size of the config-#5 extra (k-NN + a few repel iterations, points generated on the device); 0 skips it")
    ap.add_argument("--no-repel", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations (f64 k-NN, config #3, config #4)")
    ap.add_argument("--no-parity", action="store_true", help="skip the numpy self-check of the sharded paths (world > 1)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (very large clouds: N x 21 int64 pinned per rank)")
    args = ap.parse_args()
    n = args.points
    if args.impl == "reference":
        run_reference(args, n)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = pkg.Context(local)   # raises if libwtp_cuda.so or the GPU is missing: no fallback
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(pkg.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks_ok(flag: bool) -> bool:
        if world == 1:
            return flag
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def timed_dev(step, steps, warmup):
        """W untimed + K timed calls of `step`, CUDA events on the launching stream, max over ranks -> ms per call."""
        for _ in range(warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    pts_h = synth_uniform(n)
    qb, qe = ctx.shard(n)
    nq = qe - qb
    # ---------------------------------------------------------------- device-resident arm
    d_pts = torch.from_numpy(pts_h).to(dev)
    d_idx = torch.empty((nq, K), dtype=torch.int64, device=dev)
    ctx.set_timing(True)

    def step_dev():
        ctx.knn_dev(d_pts.data_ptr(), n, 3, K, np.float32, d_idx.data_ptr())

    for _ in range(args.warmup):
        step_dev()
    barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    with sampler as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step_dev()
        ev1.record(stream)
        barrier()
    # per-phase CUDA-event times of the last timed step (the library records them on the launching stream during the
    # call; reading them after every step would put a host round trip between the steps)
    t = ctx.timing()
    launches = ctx.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = n / (ms_step * 1e-3) / 1e6
    q_ms = float(t["ms_query"])
    # quick self-check of the measured output (self excluded). Sharded: row t of the compact table belongs to owned()[t].
    chk = d_idx[:1000].cpu().numpy()
    own_all = ctx.owned() if world > 1 else None
    own = own_all[:1000] if world > 1 else np.arange(1, 1001)
    assert (chk >= 1).all() and (chk <= n).all() and not (chk == own[:, None]).any()

    # ---------------------------------------------------------------- parity of the sharded rows (world > 1)
    parity = None
    if not args.no_parity:
        rng = np.random.default_rng(1000 + rank)
        rows_t = np.sort(rng.choice(nq, size=min(256, nq), replace=False))
        got = d_idx[torch.from_numpy(rows_t).to(dev)].cpu().numpy()
        want = sample_rows_brute_force(pts_h, (own_all[rows_t] - 1) if world > 1 else rows_t, K)
        parity = {"knn_rows_checked_per_rank": int(len(rows_t)), "knn_ok": all_ranks_ok(bool(np.array_equal(got, want)))}

    # ------------------------------------------------------------------ end-to-end arm
    ctx.set_timing(False)
    e2e_ms, e2e_val, e2e_phases, e2e_direct, e2e_bytes, nq_e2e = None, None, None, False, (int(pts_h.nbytes), None), nq
    e2e_per_rank = None
    if not args.no_e2e:
        h_pts = torch.from_numpy(pts_h).pin_memory()
        h_idx = torch.empty((n, K), dtype=torch.int64).pin_memory()
        h_pts_np, h_idx_np = h_pts.numpy(), h_idx.numpy()

        def step_e2e():
            ctx.knn(h_pts_np, K, out_idx=h_idx_np)

        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        mine_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        e2e_ms = max_over_ranks(mine_ms)
        e2e_val = n / (e2e_ms * 1e-3) / 1e6
        e2e_per_rank = [mine_ms]
        if world > 1:
            tm = torch.tensor([mine_ms], dtype=torch.float64, device=dev)
            allm = [torch.zeros_like(tm) for _ in range(world)]
            dist.all_gather(allm, tm)
            e2e_per_rank = [float(x.item()) for x in allm]
        if world == 1:
            assert np.array_equal(h_idx_np[own - 1], chk), "host and device entry points disagree"
        if True:   # the host call fills the rows wtp_shard_owned reports (a contiguous caller range with the row exchange)
            own_h = ctx.owned() if world > 1 else np.arange(1, n + 1)
            qs = own_h[:: max(len(own_h) // 64, 1)][:64] - 1
            ok_rows = bool(np.array_equal(h_idx_np[qs], sample_rows_brute_force(pts_h, qs, K)))
            if parity is not None:
                parity["e2e_rows_checked_per_rank"] = int(len(qs))
                parity["e2e_ok"] = all_ranks_ok(ok_rows)
            else:
                assert ok_rows, "host entry point rows disagree with brute force"
        # where the time of one host call goes (one more, untimed call with the library's phase events on)
        ctx.set_timing(True)
        t0 = time.perf_counter()
        step_e2e()
        wall = (time.perf_counter() - t0) * 1e3
        te = ctx.timing()
        ctx.set_timing(False)
        dev_ms = te["ms_bbox"] + te["ms_cellkey"] + te["ms_sort"] + te["ms_reorder"] + te["ms_query"]
        e2e_direct = bool(te["bytes_d2h"] == nq_e2e * K * 8)
        e2e_bytes = (int(te["bytes_h2d"]), int(te["bytes_d2h"]))
        e2e_phases = {"ms_h2d": float(te["ms_h2d"]), "ms_device_compute": float(dev_ms), "ms_exchange_barrier": float(te["ms_comm"]),
                      "ms_d2h_and_widen": float(max(wall - te["ms_h2d"] - dev_ms, 0.0)), "ms_wall_this_call": float(wall)}
        if world > 1:   # every rank's wall time of that one call (the ranks leave the call together only as far as its barrier goes)
            tw = torch.tensor([wall], dtype=torch.float64, device=dev)
            allw = [torch.zeros_like(tw) for _ in range(world)]
            dist.all_gather(allw, tw)
            e2e_phases["ms_wall_this_call_per_rank"] = [round(float(x.item()), 3) for x in allw]
        del h_pts, h_idx

    # --------------------------------------------------------------------- repel extra
    repel = None
    if not args.no_repel:
        nr = args.repel_points
        ctx.set_timing(True)
        snap_h = synth_uniform(nr, seed=0x57545032)
        snap = torch.from_numpy(snap_h).to(dev)
        h = nr ** (-1.0 / 3.0)
        sp, _ = ctx.make_spacing("constant", a=h)
        fm = ctx.make_force("clipped", 0.2)
        kw = dict(k=K, tol=0.0, stall_after=0, alpha_lo=h / 2000, alpha_max=h / 20)
        if not args.no_parity:
            # one (sharded) sweep from the known snapshot, every rank checks 64 points anywhere in the cloud (so also
            # points another rank swept and sent over NVLink) against the numpy restatement of the sweep
            one = snap.clone()
            ctx.repel_dev(one.data_ptr(), 0, nr, 3, np.float32, sp, fm, max_iters=1, **kw)
            ids = np.sort(np.random.default_rng(2000 + rank).choice(nr, size=64, replace=False))
            got = one[torch.from_numpy(ids).to(dev)].cpu().numpy().astype(np.float64)
            want = sample_repel_sweep(snap_h, 0, ids, h, 0.2, h / 2000, h / 20)
            err = float(np.abs(got - want).max() / h)
            parity.update({"repel_points_checked_per_rank": 64, "repel_max_err_over_spacing": max_over_ranks(err),
                           "repel_ok": all_ranks_ok(err <= 1e-3)})
            del one
        ctx.repel_dev(snap.data_ptr(), 0, nr, 3, np.float32, sp, fm, max_iters=3, **kw)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        conv, res = ctx.repel_dev(snap.data_ptr(), 0, nr, 3, np.float32, sp, fm, max_iters=args.repel_iters, **kw)
        r1.record(stream)
        barrier()
        rt = ctx.timing()
        it = max(res["iters"], 1)
        rms = max_over_ranks(r0.elapsed_time(r1)) / it
        repel = {"metric": "repel_iters_per_s", "value": 1e3 / rms, "unit": "iters/s", "points": nr, "dtype": "f32",
                 "iters": res["iters"], "ms_per_iter": rms, "sweep_ms_per_iter": rt["ms_query"] / it, "comm_ms_per_iter": rt["ms_comm"] / it,
                 "exchange": ("sweep kernels store their runs into every rank's buffer over NVLink peer memory" if rt["n_peer_ranks"] > 0
                              else ("ncclAllGather of the runs after the sweep" if world > 1 else None)),
                 "conv_last": float(conv[-1]),
                 "roofline": roofline("repel_tile_kernel<float,3,clipped> + index rebuild (whole iteration)", ALGO_BYTES_REPEL["f32"] * nr / world, rms,
                                      traffic=measured_traffic("repel", world))}
        del snap

    # ------------------------------------------------------------------ the other BASELINE configurations
    extras = None
    if not args.no_extras:
        import synth
        extras = {}
        ctx.set_timing(True)
        steps_x, warm_x = max(3, min(args.steps, 5)), 3
        # k-NN on the same cloud in Float64 (Julia's default coordinate type)
        d_idx_chk = d_idx[:4096].clone()           # rows of the timed int64 table, for the int32-table extra below
        d64 = d_pts.double()
        d_idx.zero_()
        ms = timed_dev(lambda: ctx.knn_dev(d64.data_ptr(), n, 3, K, np.float64, d_idx.data_ptr()), steps_x, warm_x)
        tq = ctx.timing()
        extras["knn_f64_10M"] = {"metric": "knn_k21_Mqueries_per_s", "value": n / (ms * 1e-3) / 1e6, "unit": "Mqueries/s", "dtype": "f64", "ms_per_step": ms,
                                 "config": {"workload": f"U3({n}) uniform 3-D, KNNTopology k=21, float64", "points": n},
                                 "roofline": roofline("knn_tile_kernel<double,3> (+ leftovers)", ALGO_BYTES_PER_QUERY["f64"] * nq, float(tq["ms_query"]),
                                                      traffic=measured_traffic("knn_f64", world))}
        del d64
        # the same k-NN step with the int32 device table (wtp_knn_dev_i32_f32): what the int64 rows cost the kernels
        d_idx32 = torch.empty((nq, K), dtype=torch.int32, device=dev)
        ms32 = timed_dev(lambda: ctx.knn_dev(d_pts.data_ptr(), n, 3, K, np.float32, d_idx32.data_ptr(), idx32=True), steps_x, warm_x)
        t32 = ctx.timing()
        same32 = bool(torch.equal(d_idx32[:4096].to(torch.int64), d_idx_chk)) if d_idx_chk is not None else None
        extras["knn_f32_10M_int32_table"] = {"metric": "knn_k21_Mqueries_per_s", "value": n / (ms32 * 1e-3) / 1e6, "unit": "Mqueries/s", "dtype": "f32", "ms_per_step": ms32,
                                            "rows_equal_int64_table": same32,
                                            "config": {"workload": f"U3({n}) uniform 3-D, KNNTopology k=21, float32, N x 21 int32 device table", "points": n},
                                            "roofline": roofline("knn_tile_kernel<float,3> (+ leftovers), int32 rows", ALGO_BYTES_PER_QUERY["f32"] * nq, float(t32["ms_query"]),
                                                                 traffic=None)}
        del d_idx32
        # config #3: repel on the 2 M graded cube, BoundaryLayerSpacing, Float32 and Float64
        for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
            gp, nw, hw = synth.graded_cube(2_000_000, dt)
            d_g = torch.from_numpy(gp).to(dev)
            d_b = d_g[:nw].clone()
            spg, _ = ctx.make_spacing("boundary_layer", hw, 4 * hw, 0.2, bnd_ptr=d_b.data_ptr(), n_bnd=nw)
            kwg = dict(k=K, tol=0.0, stall_after=0, alpha_lo=hw / 2000, alpha_max=hw / 20)
            ctx.repel_dev(d_g.data_ptr(), nw, len(gp) - nw, 3, dt, spg, ctx.make_force("clipped", 0.2), max_iters=3, **kwg)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            convg, resg = ctx.repel_dev(d_g.data_ptr(), nw, len(gp) - nw, 3, dt, spg, ctx.make_force("clipped", 0.2), max_iters=args.repel_iters, **kwg)
            g1.record(stream)
            barrier()
            tg = ctx.timing()
            itg = max(resg["iters"], 1)
            msg = max_over_ranks(g0.elapsed_time(g1)) / itg
            extras[f"repel_config3_graded_2M_{tag}"] = {
                "metric": "repel_iters_per_s", "value": 1e3 / msg, "unit": "iters/s", "dtype": tag, "ms_per_iter": msg, "iters": resg["iters"],
                "sweep_ms_per_iter": tg["ms_query"] / itg, "spacing_ms_per_iter": tg["ms_scan"] / itg,
                "leftovers_last_iter": {k2: int(tg[k2]) for k2 in ("n_leftover_sparse", "n_leftover_dense", "n_leftover_other")},
                "config": {"workload": f"G3({len(gp)}) graded unit cube (wall lattice {nw} fixed points, h_bulk/h_wall = 4, delta = 0.2), repel beta=0.2 "
                                       f"k=21 BoundaryLayerSpacing, {tag}", "points": int(len(gp)), "seconds_per_1000_iters": msg},
                "roofline": roofline(f"repel_tile_kernel<{'float' if tag == 'f32' else 'double'},3,clipped> + spacing evaluation + index rebuild (whole iteration)",
                                     ALGO_BYTES_REPEL[tag] * len(gp) / world, msg, traffic=measured_traffic(f"repel_config3_{tag}", world))}
            del d_g, d_b
        # config #4: radius CSR on the 10 M quadtree-graded square (2-D, Float64), r = 2.5 h_mid
        q2, hm = synth.graded_square(10_000_000, np.float64)
        d_q = torch.from_numpy(q2).to(dev)
        b4, e4 = ctx.shard(len(q2))
        d_off = torch.empty(e4 - b4 + 1, dtype=torch.int64, device=dev)
        nnz = ctx.radius_dev(d_q.data_ptr(), len(q2), 2, 2.5 * hm, np.float64, d_off.data_ptr())
        d_ind = torch.empty(max(nnz, 1), dtype=torch.int64, device=dev)
        ctx.radius_fill_dev(d_ind.data_ptr())

        def step_radius():
            ctx.radius_dev(d_q.data_ptr(), len(q2), 2, 2.5 * hm, np.float64, d_off.data_ptr())
            ctx.radius_fill_dev(d_ind.data_ptr())

        msr = timed_dev(step_radius, steps_x, warm_x)
        nnz_all = nnz
        if world > 1:
            tt = torch.tensor([nnz], dtype=torch.int64, device=dev)
            dist.all_reduce(tt)
            nnz_all = int(tt.item())
        extras["radius_config4_graded2d_10M_f64"] = {
            "metric": "radius_Mpoints_per_s", "value": len(q2) / (msr * 1e-3) / 1e6, "unit": "Mpoints/s", "dtype": "f64", "ms_per_step": msr, "nnz": int(nnz_all),
            "config": {"workload": f"Q2({len(q2)}) quadtree-graded unit square (h, h/2, h/4), RadiusTopology r = 2.5 h_mid = {2.5 * hm:.3e}, CSR int64 out, float64",
                       "points": int(len(q2))},
            "roofline": roofline("radius_tile_count_kernel + scan + radius_tile_fill_kernel (whole step incl. index build)",
                                 (40.0 * (e4 - b4) + 4.0 * nnz), msr, traffic=measured_traffic("radius_config4", world),
                                 note="algorithmic bytes = 2*D*T read + 4 count + 4 offset per point + 4 per entry (SURVEY.md §8d; the device writes int64 entries)")}
        del d_q, d_off, d_ind
        # config #5's size: U3(100 M) float32 generated on the device (same stream of torch's Philox on every rank), k-NN into
        # the int32 device table, then a few repel iterations; 8 rows per rank against brute force on the device. Last of the
        # extras and fenced: whatever happens here, the line above it is printed.
        # (one GPU by default: the sharded run of it is opt-in, WTP_BENCH_CONFIG5_SHARDED=1 — a rank failing inside this block
        # would leave its peers waiting in a collective)
        if args.points5 > 0 and (world == 1 or os.environ.get("WTP_BENCH_CONFIG5_SHARDED") == "1"):
            try:
                n5 = args.points5
                gen5 = torch.Generator(device=dev)
                gen5.manual_seed(0x57545035)
                p5 = torch.rand((n5, 3), generator=gen5, device=dev, dtype=torch.float32)
                b5, e5 = ctx.shard(n5)
                t5 = torch.empty((e5 - b5, K), dtype=torch.int32, device=dev)
                ms5 = timed_dev(lambda: ctx.knn_dev(p5.data_ptr(), n5, 3, K, np.float32, t5.data_ptr(), idx32=True), 3, 1)
                tq5 = ctx.timing()
                own5 = ctx.owned() if world > 1 else None
                ok5 = True
                for tt in np.random.default_rng(5000 + rank).choice(e5 - b5, size=8, replace=False):
                    i5 = int(own5[tt] - 1) if own5 is not None else int(tt)
                    q5 = p5[i5]
                    dx5 = p5[:, 0] - q5[0]
                    d25 = dx5 * dx5
                    dx5 = p5[:, 1] - q5[1]
                    d25 = d25 + dx5 * dx5
                    dx5 = p5[:, 2] - q5[2]
                    d25 = d25 + dx5 * dx5                                    # the operation order of the library's d2 (eager mode: no FMA)
                    d25[i5] = float("inf")
                    want5 = torch.topk(d25, K, largest=False, sorted=True).values
                    ok5 = ok5 and bool(torch.equal(d25[t5[int(tt)].to(torch.int64) - 1], want5))
                    del d25, dx5
                extras["knn_config5_100M_f32"] = {
                    "metric": "knn_k21_Mqueries_per_s", "value": n5 / (ms5 * 1e-3) / 1e6, "unit": "Mqueries/s", "dtype": "f32", "ms_per_step": ms5,
                    "rows_checked_per_rank": 8, "rows_ok": all_ranks_ok(ok5),
                    "phases_ms": {k2: float(tq5[k2]) for k2 in ("ms_bbox", "ms_cellkey", "ms_sort", "ms_reorder", "ms_query")},
                    "config": {"workload": f"U3({n5}) uniform 3-D generated on the device, KNNTopology k=21, float32, int32 device table", "points": n5},
                    "roofline": roofline("knn_tile_kernel<float,3> (+ leftovers)", ALGO_BYTES_PER_QUERY["f32"] * (e5 - b5), float(tq5["ms_query"]), traffic=None)}
                del t5
                h5 = n5 ** (-1.0 / 3.0)
                sp5, _ = ctx.make_spacing("constant", a=h5)
                kw5 = dict(k=K, tol=0.0, stall_after=0, alpha_lo=h5 / 2000, alpha_max=h5 / 20)
                ctx.repel_dev(p5.data_ptr(), 0, n5, 3, np.float32, sp5, ctx.make_force("clipped", 0.2), max_iters=1, **kw5)
                barrier()
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                v0.record(stream)
                _, res5 = ctx.repel_dev(p5.data_ptr(), 0, n5, 3, np.float32, sp5, ctx.make_force("clipped", 0.2), max_iters=3, **kw5)
                v1.record(stream)
                barrier()
                it5 = max(res5["iters"], 1)
                msr5 = max_over_ranks(v0.elapsed_time(v1)) / it5
                extras["repel_config5_100M_f32"] = {
                    "metric": "repel_iters_per_s", "value": 1e3 / msr5, "unit": "iters/s", "dtype": "f32", "ms_per_iter": msr5, "iters": res5["iters"],
                    "config": {"workload": f"U3({n5}) uniform 3-D, repel beta=0.2 k=21 constant spacing, float32", "points": n5, "seconds_per_100_iters": msr5 / 10.0},
                    "roofline": roofline("repel_tile_kernel<float,3,clipped> + index rebuild (whole iteration)", ALGO_BYTES_REPEL["f32"] * n5 / world, msr5, traffic=None)}
                del p5
            except Exception as ex:                                          # noqa: BLE001 — reported in the line, never fatal to it
                extras["config5_100M_error"] = str(ex)[:400]

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            import oracle
            threads = oracle.host_threads()
            t0 = time.perf_counter()
            oracle.knn(pts_h, K, threads=threads)
            dt_all = time.perf_counter() - t0
            s1 = max(n // 80, 1000)
            t0 = time.perf_counter()
            oracle.knn(pts_h[:s1], K, threads=1)
            dt_one = time.perf_counter() - t0
            cpu = {"value": n / dt_all / 1e6, "unit": "Mqueries/s", "cores": threads, "kind": "port",
                   "sample": f"the whole cloud once: KD-tree build + {n} queries, {threads} threads (sched_getaffinity); reference-faithful single thread "
                             f"(set_topology is serial, src/topology.jl:81) on {s1} points: {s1 / dt_one / 1e6:.3f} Mqueries/s",
                   "single_thread_value": s1 / dt_one / 1e6}
        line = {
            "metric": "knn_k21_Mqueries_per_s", "value": value, "unit": "Mqueries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(n),
            "run": {"sharding": f"queries split in {world} contiguous runs of the spatially sorted order; every GPU holds the point set and indexes "
                                f"the window of the grid around its run (the whole grid at 1 GPU); no collective",
                    "l2": "working set (120 MB points + 160 MB sorted tiles + 1.68 GB output per step) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_val, "unit": "Mqueries/s", "ms_per_step": e2e_ms, "ms_per_step_per_rank": e2e_per_rank, "h2d_bytes_per_step": e2e_bytes[0], "d2h_bytes_per_step": e2e_bytes[1],   # per rank, as counted by the library
                    "phases_ms": e2e_phases,
                    "api": "wtp_knn_f32: pinned host points in, N x 21 int64 table in host memory out. " +
                           ("Sharded: the kernels hand every row to the rank owning its caller range over NVLink (peer stores); each rank's contiguous "
                            "part of the table goes back as int64 written by the DMA engine (the table is pinned)" if e2e_direct else
                            (("The rows cross PCIe packed to 3 bytes per index (every index is below 2^24)" if e2e_bytes[1] is not None and e2e_bytes[1] < nq_e2e * K * 4
                              else "The rows cross PCIe as 4-byte indices") + " through a ring of pinned staging slots and are widened to int64 by the "
                             "library's host threads while later chunks are on the wire" +
                             ("; sharded: the kernels hand every row to the rank owning its caller range over NVLink (peer stores), every rank uploads its "
                              "slice of the points (all-gathered over NVLink) and brings back one contiguous part of the table" if world > 1 else "")))},
            "gpu_launches": int(launches),
            "roofline": roofline("knn_tile_kernel<float,3> (+ knn_kernel<float,3,1> for its leftovers)", ALGO_BYTES_PER_QUERY["f32"] * nq, q_ms,
                                 traffic=measured_traffic("knn", world),
                                 note="k-NN is bound by instruction issue and the shared-memory pipe, not by DRAM: 96 B/query is compulsory traffic only "
                                      "(SURVEY.md §8d); kernel_ms = tiled pass + leftover pass",
                                 kernel_ms_source="CUDA events around the query launches of the last timed step, recorded by the library on the launching stream",
                                 kernel_share_of_step=q_ms / ms_step),
            "phases_ms": {k2: float(t[k2]) for k2 in ("ms_bbox", "ms_cellkey", "ms_sort", "ms_reorder", "ms_query")},
            "ring_expanded_queries": int(t["n_ring_expanded"]),
            "tiled_pass_leftovers": {k2: int(t[k2]) for k2 in ("n_leftover_sparse", "n_leftover_dense", "n_leftover_other")},
            "index_window": {"points_indexed_rank0": int(t["n_window_points"]) or n, "missed": int(t["n_window_missed"])},
            "cpu_baseline": cpu,
            "repel": repel,
            "extras": extras,
            "parity_check": (dict(parity, ok=bool(parity.get("knn_ok", True) and parity.get("repel_ok", True) and parity.get("e2e_ok", True)))
                             if parity is not None else None),
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
