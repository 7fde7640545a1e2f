"""The oracle is test infrastructure: the product package must never import, call or execute anything under oracle/, and
bench.py may do so only in its CPU-baseline / reference / library-baseline legs (never inside run_ours' timed region).

Static check over the source (AST), plus a dynamic check that importing the whole product package leaves `oracle` out of
sys.modules."""
import ast
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "xview2_b200")


def _imports(path):
    tree = ast.parse(open(path).read(), path)
    names = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            names += [(a.name, node.lineno) for a in node.names]
        elif isinstance(node, ast.ImportFrom) and node.module:
            names.append((node.module, node.lineno))
    return names


def test_product_package_never_imports_the_oracle():
    bad = []
    for dirpath, _dirs, files in os.walk(PKG):
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(dirpath, f)
            for mod, line in _imports(path):
                if mod == "oracle" or mod.startswith("oracle."):
                    bad.append(f"{os.path.relpath(path, ROOT)}:{line} imports {mod}")
            src = open(path).read()
            if "oracle" in src and ("import_module(\"oracle" in src or "__import__(\"oracle" in src):
                bad.append(f"{os.path.relpath(path, ROOT)} imports the oracle dynamically")
    for f in ("main.py",):
        for mod, line in _imports(os.path.join(ROOT, f)):
            if mod == "oracle" or mod.startswith("oracle."):
                bad.append(f"{f}:{line} imports {mod}")
    assert not bad, bad


def test_native_sources_do_not_reference_the_oracle():
    for dirpath, _dirs, files in os.walk(os.path.join(PKG, "csrc")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                assert "oracle/" not in open(os.path.join(dirpath, f)).read(), f
    assert "oracle" not in open(os.path.join(ROOT, "include", "xv2.h")).read()


def test_bench_uses_the_oracle_only_in_baseline_legs():
    """bench.py's own arm (run_ours) may reach the oracle only through the baseline helpers it calls AFTER its timed regions."""
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"oracle_state", "oracle_step_fn"}  # helpers of the cpu_baseline / --impl reference / --impl library legs
    for node in tree.body:
        if not isinstance(node, ast.FunctionDef):
            continue
        uses = [n for n in ast.walk(node) if isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")]
        if uses:
            assert node.name in allowed, f"bench.py:{node.name} imports the oracle"
    ours = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    calls = [n for n in ast.walk(ours) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name)]
    # the line where e2e_value is ASSIGNED: both timed regions (device-resident and end-to-end) have ended there
    timed_end = min(n.lineno for n in ast.walk(ours) if isinstance(n, ast.Name) and n.id == "e2e_value" and isinstance(n.ctx, ast.Store))
    for c in calls:
        if c.func.id in ("time_cpu_reference", "time_cpu_c1", "time_library"):
            assert c.lineno > timed_end, f"baseline leg {c.func.id} is called before the timed regions end (line {c.lineno})"


def test_importing_the_product_does_not_load_the_oracle():
    code = ("import sys; sys.path.insert(0, %r); import xview2_b200, xview2_b200.ops, xview2_b200.model.plt, "
            "xview2_b200.model.unet, xview2_b200.trainer, xview2_b200.data_loading.data_module, xview2_b200.utils.post_process; "
            "import main; bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; "
            "print('BAD' if bad else 'CLEAN', bad)" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "CLEAN" in out.stdout, out.stdout
