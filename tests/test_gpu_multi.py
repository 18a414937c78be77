"""One process per GPU (torchrun, NCCL bootstrap, CUDA IPC windows): runs tools/multi_gpu_check.py on every GPU of the box,
at most 4.  Skipped on a single-GPU box -- tests/test_gpu_comm.py covers the same kernels with the ranks on one device."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ranks_on_separate_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "multi_gpu_check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
