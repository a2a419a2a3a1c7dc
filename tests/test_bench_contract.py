"""bench.py's command-line contract on a box without a GPU: the reference arm prints ONE JSON line with the keys the driver reads and
loads nothing of the product library; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

from util import ROOT


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--frames", "1", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "poses/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic"
    assert d["config"]["frames_per_gpu_per_step"] == 1 and d["config"]["crops_per_gpu_per_step"] == 8
    assert d["config"]["num_points"] == 500 and d["config"]["num_obj"] == 21 and d["config"]["refine_iterations"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "libdensefusion_b200" not in r.stderr                    # (the arm must not touch the product library)


def test_reference_arm_does_not_load_the_product_library():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--frames', '1', '--steps', '1', '--warmup', '0'];\n"
            "runpy.run_path(%r, run_name='__main__');\n"
            "maps = open('/proc/self/maps').read();\n"
            "assert 'libdensefusion_b200' not in maps, 'product library mapped by the reference arm'\n" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-2000:])


def test_product_arm_refuses_to_run_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]      # and prints no result line
