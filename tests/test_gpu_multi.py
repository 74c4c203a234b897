"""GPU, 2 ranks over NCCL (skipped with fewer than 2 devices): data-parallel train step of the agent --
rank-averaged gradients equal the single-process mean of both ranks' gradients (agent_seg.py:459-495 run
per rank + all-reduce), parameters stay bit-identical across ranks through eager steps, capture and replays."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradient_mean_and_parameter_agreement():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "ddp_check.py")]
    env = dict(os.environ, STEPS="7", GRAD_MEAN="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    print(r.stdout[-4000:])
    print(r.stderr[-4000:])
    assert r.returncode == 0
    assert "GRAD MEAN OK" in r.stdout and "DDP CHECK OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_exact_data_parallel_dice():
    """[training] exact_dp_dice: the loss sums are all-reduced inside the (graph-captured) step; every rank reports the
    same GLOBAL loss and the parameters stay bit-identical across ranks."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "ddp_check.py")]
    env = dict(os.environ, STEPS="6", EXACT_DP="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0 and "DDP CHECK OK" in r.stdout
