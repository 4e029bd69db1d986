"""CPU-side checks of bench.py: the reference arm (oracle port on host cores) runs and reports a positive
throughput for every workload kind, and the GPU arm refuses to run without a CUDA device."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("workload", ["encoder_base", "pretrain_base"])
def test_cpu_oracle_leg_runs(workload):
    sys.path.insert(0, str(ROOT))
    import bench
    v, cores, desc = bench.cpu_oracle_throughput(workload, budget_s=0.1, batch=1)
    assert v > 0 and cores >= 1 and workload in desc


def test_workload_table_matches_baseline_configs():
    sys.path.insert(0, str(ROOT))
    import bench
    cfg = json.loads((ROOT / "BASELINE.json").read_text())["configs"]
    assert "batch 256" in cfg[1] and bench.WORKLOADS["encoder_large"][2] == 256
    assert bench.WORKLOADS["encoder_large"][1] == ["bscan", "slo"]
    assert bench.WORKLOADS["pretrain_large"][1] == ["bscan", "slo", "bscanlayermap"]
    assert "batch 64" in cfg[4] and bench.WORKLOADS["cls_large"][2] == 64
    for name in bench.WORKLOADS:
        assert name in bench.GFLOP_FWD


def test_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
