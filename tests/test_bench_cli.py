"""bench.py contract checks that need no GPU: every helper `main()` calls exists, and the reference arm
(`--impl reference`: the unmodified reference on the host cores — or the oracle port where its tree is absent) prints
one JSON line with the contract's keys."""
import ast
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_calls_only_names_that_exist():
    """A helper deleted by accident (it happened) must fail here, not on the GPU box."""
    import builtins
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    bound = set()
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            bound.add(n.name)
            if isinstance(n, ast.FunctionDef):
                bound |= {a.arg for a in n.args.args + n.args.kwonlyargs}
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            bound |= {(a.asname or a.name).split(".")[0] for a in n.names}
        elif isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
            bound.add(n.id)
        elif isinstance(n, ast.arg):
            bound.add(n.arg)
    called = {c.func.id for c in ast.walk(tree) if isinstance(c, ast.Call) and isinstance(c.func, ast.Name)}
    missing = {c for c in called if c not in bound and not hasattr(builtins, c)}
    assert not missing, missing


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "1", "--batch", "2", "--src-lo", "10", "--src-hi", "14"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mel_frames_per_sec" and line["unit"] == "mel-frames/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    from oracle import ref_shim
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_shim.reference_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert len(r.stdout.strip().splitlines()) == 1          # the reference's own prints must not reach stdout
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]
