"""`bench.py --impl reference` is the reference's own CPU path and must not touch the product: run on the tiny workload in a
subprocess whose import machinery refuses `sparsefactormixedmodel_b200` (the arm mapped the product library)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GUARD = r'''
import builtins, runpy, sys
real = builtins.__import__
def guard(name, *a, **k):
    if name.startswith("sparsefactormixedmodel_b200"):
        raise ImportError("product imported by the reference arm: " + name)
    return real(name, *a, **k)
builtins.__import__ = guard
sys.argv = ["bench.py", "--impl", "reference", "--workload", "tiny", "--steps", "3", "--warmup", "3"]
runpy.run_path("bench.py", run_name="__main__")
assert not [m for m in sys.modules if m.startswith("sparsefactormixedmodel_b200")]
'''


def test_reference_arm_runs_without_the_product():
    out = subprocess.run([sys.executable, "-c", GUARD], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "iter/s"
    assert line["cpu_baseline"]["extrapolated"] is False and line["steps_timed"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("tiny: n=96 individuals, p=160 traits, k=8 factors")
