"""GPU, two ranks: the data-parallel training step (SUM all-reduce of the gradients over NCCL inside the library,
include/cvb200.h cvb_allreduce_init) equals the single-GPU step on the same global batch -- train.py's model.train
(clairvoyante_v3.py:183-205) sharded over GPUs.  Skipped on a one-GPU box; `gpurun --gpus 2` runs it."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("variant", ["v3", "v3_slim"])
def test_two_rank_dp_step_equals_single_gpu_step(variant):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dp_check.py"), variant]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert "dp_check ok: %s, 2 ranks" % variant in r.stdout
