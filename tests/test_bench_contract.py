"""bench.py contract checks that need no GPU: the reference arm (CPU port of the reference path) prints
ONE JSON line with the keys the driver reads, on the small 5-way 5-shot shape so that it runs in
seconds; the product arm must refuse to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _run(*args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--shape", "5w5s", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference"
    assert rec["metric"] == "gnn_head_episodes_per_sec_fwd_bwd" and rec["unit"] == "episodes/s"
    assert rec["higher_is_better"] is True and rec["n_gpus"] == 1 and rec["steps"] == 1 and rec["warmup"] == 1
    assert rec["value"] > 0 and rec["ms_per_step"] > 0
    from oracle import ref_loader
    assert rec["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert rec["cpu_baseline"]["cores"] >= 1
    assert rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": rec["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert rec["config"]["workload"].startswith("GnnNet head fwd+bwd, 5w5s")
    assert rec["vs_baseline"] is None


def test_product_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a machine without CUDA")
    r = _run("--steps", "1", "--warmup", "3", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0
    assert r.stdout.strip() == ""                      # no JSON line from a fallback
    assert "CUDA" in r.stderr or "cuda" in r.stderr


def test_stdout_carries_only_the_json_line():
    """Whatever native code writes to file descriptor 1 during the run (NCCL's version banner under
    torchrun) must land on stderr; stdout gets the one JSON line."""
    code = (
        "import os, sys, json; sys.path.insert(0, %r); import bench\n"
        "with bench._OnlyJsonOnStdout() as out:\n"
        "    os.write(1, b'NCCL version x.y.z\\n'); print('python noise')\n"
        "    out.emit(json.dumps({'metric': 'm', 'value': 1}))\n"
        "print('after')\n" % ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.splitlines()
    assert lines[0] == json.dumps({"metric": "m", "value": 1}) and lines[1:] == ["after"]
    assert "NCCL version x.y.z" in res.stderr and "python noise" in res.stderr
