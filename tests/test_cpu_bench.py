"""bench.py contract on a box without a GPU: the reference arm prints exactly ONE JSON line on stdout with the keys
the driver reads; the multi-GPU-only workload refuses to run on one GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Gvoxels/s seg+CC" and d["unit"] == "Gvoxels/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    # "reference": the reference's own files were found (/root/reference here, baseline/_ref/reference on the GPU box)
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    staged = os.path.exists("/root/reference/inference/inference.py") or os.path.exists(os.path.join(ROOT, "baseline", "_ref", "reference", "inference", "inference.py"))
    assert (d["cpu_baseline"]["kind"] == "reference") == (staged and os.path.exists(os.path.join(ROOT, "baseline", "_ref", "inference_weights.tar")))
    assert d["e2e"] == {"value": d["value"], "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_whole_volume_workload_needs_several_gpus():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "cfg4"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode != 0 and "multi-GPU" in (out.stderr + out.stdout)
