"""profiles/ table from one `ncu --set full` pass over tools/kernel_zoo.py: joins the raw-page CSV with the case names the zoo
prints after each case (ncu's "Profiling" lines and the zoo's lines interleave in the log in launch order).
    python tools/zoo_summary.py gpurun_out/ncu_zoo.log gpurun_out/ncu_zoo.csv > profiles/r02_ncu_kernels.md
"""
import csv
import re
import sys

log, raw = sys.argv[1], sys.argv[2]
case_of = {}  # profiled-launch ordinal -> case name
pending, ordinal = [], 0
for line in open(log, errors="replace"):
    m = re.match(r'==PROF== Profiling "', line)
    if m:
        pending.append(ordinal)
        ordinal += 1
        continue
    m = re.match(r"\s*\d+ libxv2 launches\s+(.*)", line)
    if m:
        for o in pending:
            case_of[o] = m.group(1).strip()
        pending = []
rows = list(csv.reader(open(raw)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (ValueError, KeyError):
        return float("nan")


print("| # | kernel | zoo case (C2 shape) | grid | regs | time us | DRAM rd MB | DRAM wr MB | DRAM % of peak | tensor pipe % | SM % |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
unit_rd = rows[1][ix["dram__bytes_read.sum"]]
unit_wr = rows[1][ix["dram__bytes_write.sum"]]
scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}
ours_tagged = any("xv2::" in r[ix["Kernel Name"]] for r in rows[2:])  # a capture filtered with -k regex:xv2:: prints bare names
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if ours_tagged and "xv2::" not in name:
        continue
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("xv2::", "")
    o = int(r[ix["ID"]])
    print(f"| {o} | `{short}` | {case_of.get(o, case_of.get(o - 1, '?'))} | {r[ix['launch__grid_size']]} | {r[ix['launch__registers_per_thread']]} | "
          f"{num(r, 'gpu__time_duration.sum'):.1f} | {num(r, 'dram__bytes_read.sum') * scale.get(unit_rd, 1):.1f} | "
          f"{num(r, 'dram__bytes_write.sum') * scale.get(unit_wr, 1):.1f} | {num(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {num(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} |")
