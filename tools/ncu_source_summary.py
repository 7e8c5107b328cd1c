"""Summarise an `ncu --page source --csv` dump: per-opcode executed instructions / samples and
the hottest contiguous SASS regions.  Usage: ncu -i X.ncu-rep --page source --csv > src.csv;
python tools/ncu_source_summary.py src.csv [n_regions]"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
end = next((i for i, r in enumerate(data) if not r or r[0] == "Kernel Name"), len(data))  # first launch only
data = data[:end]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print(f"SASS instructions: {len(data)}, executed warp-instr: {tot_inst}, samples: {tot_samp}")
by_op = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    by_op[op][0] += int(r[ix["Instructions Executed"]]); by_op[op][1] += int(r[ix["# Samples"]]); by_op[op][2] += 1
print("\nopcode         static   executed  share   samples share")
for op, (e, s, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:12s} {n:7d} {e:11d} {100*e/tot_inst:5.1f}% {s:8d} {100*s/max(tot_samp,1):5.1f}%")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# regions: split on big changes of executed count
print("\nregions (contiguous SASS with equal execution count):")
regions = []
cur = None
for i, r in enumerate(data):
    e = int(r[ix["Instructions Executed"]])
    if cur is None or e != cur["e"]:
        cur = {"e": e, "start": i, "n": 0, "samp": 0, "stalls": collections.Counter(), "ops": collections.Counter()}
        regions.append(cur)
    cur["n"] += 1; cur["samp"] += int(r[ix["# Samples"]])
    for c in stall_cols:
        cur["stalls"][c] += int(r[ix[c]] or 0)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]]); cur["ops"][m.group(2) if m else "?"] += 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for g in sorted(regions, key=lambda g: -g["e"] * g["n"])[:N]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in g["stalls"].most_common(4))
    ops = ", ".join(f"{k}:{v}" for k, v in g["ops"].most_common(6))
    print(f"  sass[{g['start']:5d}+{g['n']:4d}] exec/inst={g['e']:9d} warp-instr={g['e']*g['n']:11d} ({100*g['e']*g['n']/tot_inst:4.1f}%) samples={g['samp']:6d} ({100*g['samp']/max(tot_samp,1):4.1f}%) | {top} | {ops}")
