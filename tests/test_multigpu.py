"""Row-slab sharding on real GPUs (needs >= 2 devices; skipped on a 1-GPU box). The CPU-side protocol is covered by
tests/test_sharding_cpu.py with gloo."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_sharded_solve_matches_oracle():
    from iterative_solvers_b200 import capi

    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port),
                           os.path.join(ROOT, "tests", "run_multigpu.py")], capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and "MULTIGPU_OK" in proc.stdout, proc.stdout[-3000:] + proc.stderr[-3000:]
