"""GPU (-m gpu), needs >= 2 devices: the point-partitioned multi-rank path (NCCL all-reduce of camera-sized buffers)
reproduces the single-GPU solve.  Skipped on single-GPU boxes; the partition arithmetic itself is covered on CPU by
tests/test_partition_gloo.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_solve_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "multi_gpu_check.py"), "ladybug", "12"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["max_rel_cost_diff_first8"] < 1e-9 and res["max_rel_cost_diff"] < 1e-7 and res["accept_equal"]
