"""Row-sharded path on real GPUs: one process per GPU over NCCL (skipped on a box with a single GPU; the
host-side logic of the same path runs on CPU with gloo in tests/test_distributed.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_two_gpus():
    import pyci_b200
    ngpu = pyci_b200.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % ngpu)
    for attempt in range(2):  # (a rendezvous right after another multi-process job on the box has failed once: retry)
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        if out.returncode == 0:
            break
        sys.stderr.write("[test_gpu_multi] attempt %d failed:\n%s\n" % (attempt, (out.stdout[-2000:] + out.stderr[-4000:])))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0 of 2 ok" in out.stdout and "rank 1 of 2 ok" in out.stdout
