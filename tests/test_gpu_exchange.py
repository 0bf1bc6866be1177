"""Fused gradient exchange (csrc/exchange.cu) against the NCCL path, on >= 2 GPUs of one box: two ranks under torchrun,
both flavours (multimem through the NVSwitch, per-peer loads), replicas bit-identical and equal to the NCCL arm
(tools/exchange_check.py).  Single-GPU boxes skip it; tests/test_exchange_protocol.py model-checks the handshake on the CPU."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one NVLink domain"),
]


def test_peer_exchange_matches_nccl_arm():
    for mode in ("peer", "peer-p2p"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.join(ROOT, "tools", "exchange_check.py"), "--mode", mode, "--time", "0"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode == 0, mode + r.stdout[-2000:] + r.stderr[-2000:]
