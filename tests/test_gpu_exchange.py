"""Fused gradient exchange (csrc/exchange.cu) against the NCCL path, on >= 2 GPUs of one box.

Opt-in like the path it covers (``NGP_B200_TEST_EXCHANGE=1``): the kernel spins on its peers, so it only runs where two
ranks are known to come up together; single-GPU boxes skip it."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("NGP_B200_TEST_EXCHANGE") != "1", reason="opt-in: NGP_B200_TEST_EXCHANGE=1"),
    pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one NVLink domain"),
]


@pytest.mark.parametrize("mode", ["peer", "peer-p2p"])
def test_peer_exchange_matches_nccl_arm(mode):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools_exchange_check.py"), "--mode", mode, "--time", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
