"""Aggregates a tools/layer_profile.py table by kernel kind (entry point + conv role).   python tools/layer_sum.py layers.txt [-n 20]"""
import collections
import re
import sys

path = sys.argv[1]
top = int(sys.argv[sys.argv.index("-n") + 1]) if "-n" in sys.argv else 40
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for line in open(path):
    p = line.split()
    if len(p) < 3 or not p[0].isdigit():
        if line.startswith("#"):
            print(line.rstrip())
        continue
    name, ms, tag = p[1], float(p[2]), " ".join(p[3:])
    kind, tf = name, 0.0
    m = re.search(r"(convT )?(fwd|dgrad|wgrad) n\d+ (\d+)x\d+ c(\d+)(\+\d+)? k(\d+)( (\dx\d) g(\d))?", tag)
    if m and name in ("xv2_conv_tc", "xv2_wgrad_tc"):
        kind = f"{name} {'convT' if m.group(1) else m.group(8)} {m.group(2)}"
        try:
            tf = float(p[3]) * ms
        except ValueError:
            pass
    a = agg[kind]
    a[0] += 1
    a[1] += ms
    a[2] += tf
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{k:36s} n={a[0]:3d} ms={a[1]:7.3f} {a[1] / tot * 100:5.1f}%  TF/s={a[2] / a[1] if a[1] else 0:7.1f}")
print(f"{'sum of libxv2 launches':36s}       ms={tot:7.3f}")
