"""GPU, needs >= 2 devices: fused NVLink scatter of the m-major product (tools/check_peer_scatter.py
under torchrun).  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_peer_scatter_two_ranks():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_peer_scatter.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("PASS") == 2


def _torchrun(script, port, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))


def test_generate_two_ranks_sharing_one_gpu():
    """The multi-rank product path on a single-GPU box: two processes on cuda:0 (gloo for the host
    plumbing, since NCCL refuses two ranks on one device), the m-blocks of each owner mapped by the
    other through CUDA IPC and filled by its pack kernel -- BeamTransfer.generate() must write exactly
    the single-process product (fp32x3 with complex64 on the wire, fp64, and a multi-chunk run)."""
    res = _torchrun("check_generate_multi.py", 29541, env={"DSB_CHECK_BACKEND": "gloo"})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("PASS") == 3 and "peer-scatter" in res.stdout


def test_generate_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    res = _torchrun("check_generate_multi.py", 29543)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("PASS") == 3 and "peer-scatter" in res.stdout
