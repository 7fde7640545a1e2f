"""Runs every tensor-core case in its own process (a faulting kernel poisons the CUDA context) and prints a table.

    python tools/tc_probe.py [case ...]       # under gpurun; writes gpurun_out/tc_probe.txt
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name):
    import torch
    from tests.tc_cases import run_case
    try:
        print("RESULT", name, run_case(name))
    except Exception as e:  # noqa: BLE001
        print("RESULT", name, "EXC", repr(e)[:300])


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    from tests.tc_cases import CASES
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    lines = []
    for name in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--child", name], capture_output=True, text=True, timeout=180)
            res = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            line = res[0] if res else f"RESULT {name} CRASH rc={r.returncode} {r.stderr.strip().splitlines()[-1:]}"
        except subprocess.TimeoutExpired:
            line = f"RESULT {name} TIMEOUT"
        print(line, flush=True)
        lines.append(line)
    with open(os.path.join(ROOT, "gpurun_out", "tc_probe.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
