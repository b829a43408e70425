"""profiles/r02_<name>_traffic_n<world>.json from an ncu --set full capture: dram__bytes_read.sum + dram__bytes_write.sum
of the launches whose kernel name matches, summed per step (e.g. tiled pass + leftover pass) and averaged over the
captured steps. usage: ncu_traffic.py <report.ncu-rep> <kernel-regex> <name> <world> [launches_per_step]"""
import csv, io, json, os, re, subprocess, sys

rep, rx, name, world = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3], int(sys.argv[4])
per_step = int(sys.argv[5]) if len(sys.argv) > 5 else 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
units = rows[1]
kn, rd, wr, du = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


sel = [r for r in rows[2:] if len(r) == len(h) and rx.search(r[kn])]
tot = sum(to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]) for r in sel)
steps = max(len(sel) // per_step, 1)
out = {"dram_bytes_per_launch": tot / steps, "launches_matched": len(sel), "launches_per_step": per_step,
       "kernels": sorted({r[kn][:120] for r in sel}), "gpu_time_us_per_step": sum(float(r[du].replace(",", "")) for r in sel) / steps * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[du], 1.0),
       "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}: dram__bytes_read.sum + dram__bytes_write.sum"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"r02_{name}_traffic_n{world}.json")
json.dump(out, open(path, "w"), indent=1)
print(path, json.dumps(out))
