"""N>1 path on CPU: world_size 2, gloo backend (the GPU tier uses the same helpers over NCCL in bench.py)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sample_split_partition():
    from turner_b200 import dist as tdist
    for pps in (1, 2, 5, 8, 128):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                b, s = tdist.sample_split(r, world)
                idx = list(range(b, pps, s))
                assert len(idx) == tdist.local_sample_count(pps, r, world)
                seen += idx
            assert sorted(seen) == list(range(pps))  # every sample exactly once
    with pytest.raises(ValueError):
        tdist.sample_split(2, 2)


def test_two_rank_gloo_reduce(ob):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_OK world=2" in out.stdout
