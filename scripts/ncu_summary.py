"""Summarise an .ncu-rep: per-kernel headline metrics, SASS opcode mix and stall reasons."""
import csv, sys, subprocess, collections, io
rep = sys.argv[1]
per_unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0   # e.g. queries per launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
for k in keys:
    if k in h:
        i = h.index(k)
        print(f"{k:75s}", [r[i][:60] for r in rows[1:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
names = [rows[i - 1][1][:90] for i in hdrs]
for n, i in enumerate(hdrs):
    h = rows[i]
    end = hdrs[n + 1] - 1 if n + 1 < len(hdrs) else len(rows)
    data = [r for r in rows[i + 1:end] if len(r) == len(h)]
    ie = h.index('Instructions Executed'); sc = h.index('Source')
    stall_cols = [j for j, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    tot = sum(int(r[ie]) for r in data)
    print(f"\n== {names[n]}\n   SASS instrs {len(data)}, warp-instr executed {tot:.3e} ({tot / per_unit:.1f} per unit)")
    ops = collections.Counter()
    for r in data:
        t = r[sc].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += int(r[ie])
    print("   " + "  ".join(f"{k}:{v / per_unit:.1f}" for k, v in ops.most_common(22)))
    st = collections.Counter()
    for r in data:
        for j in stall_cols:
            st[h[j]] += int(r[j] or 0)
    ts = sum(st.values()) or 1
    print("   " + "  ".join(f"{k}:{100 * v / ts:.1f}%" for k, v in st.most_common(9)))
