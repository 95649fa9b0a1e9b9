"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise): launches
tools/mgpu_check.py under torchrun, one rank per GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return len([l for l in out.splitlines() if l.startswith("GPU ")])
    except Exception:
        return 0


@pytest.mark.parametrize("fused", ["1", "0"])      # single applications: peer-store exchange | NCCL all-gather
def test_row_sharded_operator_and_fused_loops_two_ranks(fused):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, SDFS_FUSED_EXCHANGE=fused, SDFS_KRON_SHARD_MIN="0"))   # shard the small test grids too
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ALL PASS" in r.stdout and "FAIL" not in r.stdout, r.stdout[-3000:]
