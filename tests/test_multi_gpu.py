"""x-slab sharding on real GPUs: the 2-rank run must equal the single-GPU run bit for bit, on the
peer-memory transport (in-kernel ordering over NVLink) and on the NCCL plane exchange, and the
peer -> NCCL fallback must engage when a rank cannot export its arrays.  Needs >= 2 visible GPUs
(skipped otherwise); the host-side logic is covered on CPU by tests/test_dist_gloo.py."""

import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(n, script, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", script)]
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)


def _need_gpus(n):
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.gpu
def test_two_gpu_slabs_equal_single_gpu():
    _need_gpus(2)
    r = _torchrun(2, "check_slabs.py")
    assert r.returncode == 0 and "SLAB CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "halo=peer" in r.stdout


@pytest.mark.gpu
def test_peer_export_failure_falls_back_to_nccl_on_every_rank():
    _need_gpus(2)
    r = _torchrun(2, "check_slabs.py", {"FDTDX_B200_PEER_FAIL": "1"})
    assert r.returncode == 0 and "SLAB CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "peer-memory halo unavailable" in (r.stdout + r.stderr)
