"""Join an ncu SASS source page with nvdisasm line info: warp-instructions executed per source line.
usage: ncu_lines.py <report.ncu-rep> <disassembly from `nvdisasm -g -c`> <mangled-kernel-substring> [units]"""
import csv, io, re, subprocess, sys, collections
rep, sass, kern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
# nvdisasm: track //## File "...", line N ; instruction lines: /*0010*/ OP ... ;
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
addr2line, cur = {}, ("?", 0)
inl = ""
for l in lines[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
ie = h.index("Instructions Executed")
isamp = h.index("# Samples") if "# Samples" in h else None
stall_cols = [j for j, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[hdr + 1:]:
    if len(r) != len(h) or r[0] == "Address":
        break
    data.append(r)
base = int(data[0][0], 16)
per = collections.Counter()
samp = collections.Counter()
why = collections.defaultdict(collections.Counter)
tot = 0
for r in data:
    off = int(r[0], 16) - base
    n = int(r[ie])
    key = addr2line.get(off, ("?", 0))
    per[key] += n
    tot += n
    if isamp is not None:
        samp[key] += int(r[isamp] or 0)
        for j in stall_cols:
            why[key][h[j]] += int(r[j] or 0)
print(f"total warp-instr {tot:.3e} = {tot / units:.1f} per unit")
ts = sum(samp.values()) or 1
print("instr/unit  instr%  samples%  top stall reasons   source line")
for (f, ln), n in sorted(per.items(), key=lambda kv: -max(kv[1] / tot, samp[kv[0]] / ts))[:45]:
    top = " ".join(f"{k[6:]}:{100 * v / max(sum(why[(f, ln)].values()), 1):.0f}" for k, v in why[(f, ln)].most_common(3))
    print(f"{n / units:8.1f}  {100 * n / tot:5.1f}%  {100 * samp[(f, ln)] / ts:5.1f}%  {top:42s} {f}:{ln}")
