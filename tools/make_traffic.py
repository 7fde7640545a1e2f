"""profiles/traffic.json from an ncu launch list with DRAM byte counters (tools/gpu_round.sh traffic):
per libxv2 entry point, the average dram__bytes_read + dram__bytes_write per launch of its kernels in one training step.

    python tools/make_traffic.py gpurun_out/traffic.csv profiles/traffic.json
"""
import csv
import json
import sys

ENTRY_OF = {"conv_tc_kernel": "xv2_conv_tc", "conv_strip_kernel": "xv2_conv_tc", "wgrad_tc_kernel": "xv2_wgrad_tc",
            "wgrad_strip_kernel": "xv2_wgrad_tc", "bn_stream_kernel<0>": "xv2_bn_stats", "bn_stream_kernel<1>": "xv2_bn_train_apply",
            "bn_stream_kernel<2>": "xv2_bn_bwd_reduce", "bn_stream_kernel<3>": "xv2_bn_bwd_apply"}  # kernels not listed here are reported under their own name


def main(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    iname, imetric, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    per = {}
    for r in rows[1:]:
        name = r[iname]
        entry = next((e for k, e in ENTRY_OF.items() if k.replace("<", "<(int)") in name or k in name), None)
        if entry is None:  # every other kernel of the step under its own (demangled, argument-free) name
            import re
            entry = re.sub(r"\(.*", "", name).replace("void ", "")
            entry = re.sub(r"<.*", "", entry) if entry.startswith("at::") else entry
        v = float(r[ival].replace(",", ""))
        unit = r[iunit].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9,
                 "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(unit, 1)
        d = per.setdefault(entry, {"launches": set(), "bytes": 0.0, "seconds": 0.0})
        d["launches"].add(r[iid])
        if r[imetric].startswith("dram__bytes"):
            d["bytes"] += v * scale
        elif r[imetric].startswith("gpu__time"):
            d["seconds"] += v * scale
    out = {e: {"launches": len(d["launches"]), "dram_bytes_per_launch": d["bytes"] / max(1, len(d["launches"])),
               "dram_bytes_per_step": d["bytes"], "ncu_ms_per_step": d["seconds"] * 1e3} for e, d in per.items()}
    out = dict(sorted(out.items(), key=lambda kv: -kv[1]["dram_bytes_per_step"]))
    out["TOTAL"] = {"launches": sum(v["launches"] for v in out.values()),
                    "dram_bytes_per_step": sum(v["dram_bytes_per_step"] for v in out.values()),
                    "ncu_ms_per_step": sum(v["ncu_ms_per_step"] for v in out.values())}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
