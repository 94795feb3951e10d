"""Text summary of an ncu --set full report: one block per kernel launch (or per kernel name, --merge).

    python tools/ncu_summary.py report.ncu-rep [--merge] > profiles/xyz.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
]


def main():
    rep = sys.argv[1]
    merge = "--merge" in sys.argv
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    print("# ncu --set full --clock-control none summary of `%s`\n" % rep.split("/")[-1])
    groups = collections.OrderedDict()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        key = name.split("(")[0] if merge else "%s  [id %s]" % (name.split("(")[0], r[idx["ID"]])
        groups.setdefault(key, []).append(r)
    for key, rs in groups.items():
        r = max(rs, key=lambda q: float(q[idx["gpu__time_duration.sum"]].replace(",", "") or 0))
        print("## `%s`%s" % (key, "  (%d launches; the longest shown)" % len(rs) if merge and len(rs) > 1 else ""))
        for k, label in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                print("- %s: %s %s" % (label, r[idx[k]], units[idx[k]]))
        tot_t = sum(float(q[idx["gpu__time_duration.sum"]].replace(",", "") or 0) for q in rs)
        if merge and len(rs) > 1:
            print("- total over the %d launches: %.1f %s" % (len(rs), tot_t, units[idx["gpu__time_duration.sum"]]))
        st = {}
        for h in stall:
            try:
                st[h] = float(r[idx[h]].replace(",", ""))
            except ValueError:
                pass
        tot = sum(st.values()) or 1.0
        top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
        print("- top stall reasons (pc sampling): " + ", ".join(
            "%s %.0f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot) for k, v in top))
        print()


if __name__ == "__main__":
    main()
