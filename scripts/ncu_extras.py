"""One of bench.py's extra configurations, a fixed number of steps and nothing else — to be run under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file X.csv
scripts/ncu_extras.py <which> <steps>; then `scripts/ncu_extras.py sum X.csv <name> <steps>` writes
profiles/r02_<name>_traffic_n1.json (DRAM bytes of ALL launches of the run divided by the steps: these rooflines are
quoted per whole step / iteration, set-up launches included)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))

if sys.argv[1] == "sum":
    path, name, steps = sys.argv[2], sys.argv[3], int(sys.argv[4])
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    kernels = {}
    for r in rows:
        m, unit, val = r[12], r[13], float(r[14].replace(",", ""))
        if m in tot:
            tot[m] += val * scale.get(unit, 1)
            if m != "gpu__time_duration.sum":
                k = r[4].split("(")[0][:60]
                kernels[k] = kernels.get(k, 0.0) + val * scale.get(unit, 1)
    out = {"dram_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / steps, "steps": steps,
           "gpu_time_us_per_step": tot["gpu__time_duration.sum"] / steps, "launches": len(rows) // 3,
           "top_kernels_bytes_per_step": {k: v / steps for k, v in sorted(kernels.items(), key=lambda kv: -kv[1])[:6]},
           "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, {os.path.basename(path)}: "
                     f"all launches of {steps} steps (set-up included), per step"}
    dst = os.path.join(ROOT, "profiles", f"r02_{name}_traffic_n1.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(dst, json.dumps(out)[:400])
    sys.exit(0)

import numpy as np, torch
import __graft_entry__ as g
import synth
which, steps = sys.argv[1], int(sys.argv[2])
pkg = g.load_package()
ctx = pkg.Context(0)
dev = torch.device("cuda", 0)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
if which == "knn_f64":
    n = 10_000_000
    d = torch.from_numpy(synth.uniform_cube(n, np.float64)).to(dev)
    out = torch.empty((n, 21), dtype=torch.int64, device=dev)
    for _ in range(steps):
        ctx.knn_dev(d.data_ptr(), n, 3, 21, np.float64, out.data_ptr())
elif which.startswith("cfg3"):
    dt = np.float64 if which.endswith("f64") else np.float32
    gp, nw, hw = synth.graded_cube(2_000_000, dt)
    d_g = torch.from_numpy(gp).to(dev)
    d_b = d_g[:nw].clone()
    sp, _ = ctx.make_spacing("boundary_layer", hw, 4 * hw, 0.2, bnd_ptr=d_b.data_ptr(), n_bnd=nw)
    ctx.repel_dev(d_g.data_ptr(), nw, len(gp) - nw, 3, dt, sp, ctx.make_force("clipped", 0.2), k=21, max_iters=steps, tol=0.0, stall_after=0,
                  alpha_lo=hw / 2000, alpha_max=hw / 20)
elif which == "cfg4":
    q2, hm = synth.graded_square(10_000_000, np.float64)
    d_q = torch.from_numpy(q2).to(dev)
    d_off = torch.empty(len(q2) + 1, dtype=torch.int64, device=dev)
    d_ind = None
    for _ in range(steps):
        nnz = ctx.radius_dev(d_q.data_ptr(), len(q2), 2, 2.5 * hm, np.float64, d_off.data_ptr())
        if d_ind is None:
            d_ind = torch.empty(nnz, dtype=torch.int64, device=dev)
        ctx.radius_fill_dev(d_ind.data_ptr())
torch.cuda.synchronize()
ctx.close()
