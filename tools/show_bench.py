"""Prints the headline fields of a bench.py JSON line (gpurun tail output is short)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        print(path, "unreadable:", exc)
        continue
    lib = d.get("library_baseline") or {}
    roof = d.get("roofline") or {}
    print(f"{path}: {d.get('value'):.2f} {d.get('unit')} ({d.get('ms_per_step'):.2f} ms/step, N={d.get('n_gpus')}), e2e {(d.get('e2e') or {}).get('value')}, "
          f"conv_roofline_frac {d.get('conv_roofline_frac')}, library {lib.get('value')} ({lib.get('ms_per_step')} ms; e2e {(lib.get('e2e') or {}).get('value')}) {lib.get('unavailable', '')}")
    print("  roofline:", {k: roof.get(k) for k in ("kernel", "achieved", "frac", "hbm_frac", "share_of_step")}, "clocks:", d.get("clocks"))
    for k, v in list((d.get("kernels") or {}).items())[:12]:
        print(f"    {k:28s} {v}")
    st = roof.get("stages") or {}
    for k, v in st.items():
        print(f"    stage {k:16s} {v}")
    for k in ("cpu_baseline", "cpu_baseline_c1"):
        if d.get(k):
            print(f"  {k}: {d[k].get('value')} {d[k].get('unit')} cores {d[k].get('cores')} :: {str(d[k].get('sample'))[:160]}")
