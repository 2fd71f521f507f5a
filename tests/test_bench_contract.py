"""bench.py contract pieces that run without a GPU: the reference arm (`--impl reference`) prints one JSON line with the
agreed keys, keeps every host thread busy, and never loads the CUDA library."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(ref):
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--codec', 'bc1', '--size', '256', '--steps', '2', "
            "'--warmup', '1', '--cpu-budget', '0.2']\n"
            "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "print('PRODUCT_LIB_LOADED' if 'libgfx_imagecompress_b200' in maps else 'PRODUCT_LIB_NOT_LOADED')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert lines[-1] == "PRODUCT_LIB_NOT_LOADED"
    d = json.loads(lines[-2])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"] and cb["threads_used"] == cb["cores"] >= 1
    assert d["config"]["workload"] and "model" not in d["config"]
