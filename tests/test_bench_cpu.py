"""CPU-side checks of bench.py's contract: the reference arm (the CPU restatement timed on the host cores) prints one
well-formed JSON line, under torchrun only rank 0 prints it, the roofline's algorithmic-bytes figure is SURVEY §8(d)'s
formula, and our own arm refuses to run without a CUDA device instead of measuring anything on the CPU."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_per_update_is_the_survey_formula():
    b = _bench_module()
    # B(G_f) = 8d + (8d + 8)/G_f + 4d/K   (two archive rows, state + log-density in and out, append)
    assert b.algorithmic_bytes_per_update(20, 10, 1) == pytest.approx(336.0)      # C4, one generation per launch
    assert b.algorithmic_bytes_per_update(20, 10, 10) == pytest.approx(184.8)     # C4, K = 10 generations fused
    assert b.algorithmic_bytes_per_update(5, 10, 1) == pytest.approx(90.0)        # C1
    assert b.algorithmic_bytes_per_update(10, 10, 1) == pytest.approx(172.0)      # C3


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "chain-updates/s" and d["unit"] == "chain-updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("C2 linreg")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_runs_on_rank_0_only():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--workload", "c1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_our_arm_fails_loudly_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "1", "--warmup", "0",
                        "--no-cpu-baseline", "--no-e2e"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not _json_lines(r.stdout)                     # nothing that could be mistaken for a measurement
