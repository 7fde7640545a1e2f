"""Kernel shares of one training step from an ncu launch list (tools/gpu_r2.sh launches).
    python tools/launch_shares.py gpurun_out/launches.csv "header comment" > profiles/r02_ncu_launch_shares.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    if name.startswith("at::"):
        name = re.sub(r"<\d+, at::", "<", name)
        name = re.sub(r", std::array.*", ">", name)
    ms = float(r[ix["Metric Value"]].replace(",", "")) / 1e6
    agg[name][0] += 1
    agg[name][1] += ms
    tot += ms
print("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off on `python tools/layer_profile.py --ncu`")
print("# (ONE eager training step of BASELINE config 2 after two warm-up steps; serialised + cold cache -> compare SHARES)")
for line in sys.argv[2:]:
    print("# " + line)
print("kernel,launches,total_ms,share")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"\"{k}\",{a[0]},{a[1]:.3f},{a[1] / tot:.4f}")
print(f"TOTAL,{sum(a[0] for a in agg.values())},{tot:.3f},1.0")
