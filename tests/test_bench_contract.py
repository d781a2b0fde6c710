"""bench.py contract checks that need no GPU: the reference arm's JSON line, the frame cache, and that the product arm
refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, tmp_path, timeout=300):
    env = dict(os.environ, EF_BENCH_CACHE=str(tmp_path))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line(tmp_path):
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"], tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    for k in ("metric", "value", "unit", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert "640x480" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_cuda(tmp_path):
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline"], tmp_path)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]  # no bench line from a machine without CUDA


def test_frame_cache_serves_prefixes(tmp_path, small_K, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    from elasticfusion_b200 import synth

    monkeypatch.setenv("EF_BENCH_CACHE", str(tmp_path))
    a = bench.make_frames(small_K, 3, 5)
    b = bench.make_frames(small_K, 5, 5)  # extends the cached run
    c = bench.make_frames(small_K, 2, 5)  # served from it
    ref = list(synth.sequence(5, small_K, seed=5, noise=True))
    for i in range(5):
        assert np.array_equal(b[0][i], ref[i][0]) and np.array_equal(b[1][i], ref[i][1])
    assert np.array_equal(a[0], b[0][:3]) and np.array_equal(c[1], b[1][:2])
