"""Hottest instructions (pc samples with their top stall reasons) of the first kernel in an ncu report.
    python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
first, seen = [], set()
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        if first:
            break
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    first.append(r)
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0
tot = sum(num(r[idx["# Samples"]]) for r in first) or 1
print("%s: %d instructions, %d samples" % (rows[0][1][:80], len(first), tot))
pos = {r[0]: i for i, r in enumerate(first)}
for r in sorted(first, key=lambda r: -num(r[idx["# Samples"]]))[:N]:
    n = num(r[idx["# Samples"]])
    st = {k: num(r[idx[k]]) for k in hdr if k.startswith("stall_") and "Not" not in k}
    main = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
    print("%5.1f%%  #%-5d %-64s %s" % (100.0 * n / tot, pos[r[0]], r[1].strip()[:64], main))
