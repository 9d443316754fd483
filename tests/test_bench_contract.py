"""CPU: the driver-facing contract of bench.py that can be checked without a GPU.

  * `bench.py --impl reference` (the reference arm: the C oracle on the host cores) prints exactly ONE JSON line on stdout with
    the keys the driver reads, the same metric / unit / config.workload wording as the GPU arm, and zero-byte e2e copies;
  * the GPU arm fails loudly without a CUDA device (no CPU fallback) and prints no JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT,
                          env=dict(os.environ, **(env or {})), timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run(["--impl", "reference", "--tiles", "4096", "--steps", "2", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "covt_tile_batch_decode_compressed_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("config5: ONE batch of 4096 synthetic") and "model" not in d["config"]
    # the keys of `config` are the ones the GPU arm prints too (the driver compares the two dicts)
    assert sorted(d["config"]) == ["container", "flags", "l2", "partition", "tiles", "workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == "GB/s" and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["dtype"] == "int32"


def test_gpu_arm_fails_loudly_without_a_device():
    p = _run(["--tiles", "1024", "--steps", "1", "--warmup", "0"], env={"CUDA_VISIBLE_DEVICES": ""})
    assert p.returncode != 0
    assert not [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")]
    assert "CUDA" in p.stderr or "cuda" in p.stderr
