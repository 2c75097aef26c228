"""N-rank data-parallel step == single-rank step on the concatenated batch, through the CUDA kernels and NCCL
(NCCL form: needs >= 2 GPUs, `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`; the gloo form runs on one GPU)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(backend):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "dp_equivalence.py")], capture_output=True, text=True, timeout=900,
                       env={**os.environ, "NSD_DP_BACKEND": backend})
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "DP_EQUIVALENCE_OK" in r.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_step_equals_single_rank():
    _run("nccl")


def test_two_rank_step_on_one_gpu_equals_single_rank():
    """The same equivalence on a ONE-GPU box: two ranks time-slice cuda:0 (both run the real kernels, the bucketed
    all-reduce is fired from inside the backward) and exchange gradients through gloo, since NCCL refuses two ranks on one
    device.  Keeps the data-parallel path under test where only one GPU is visible."""
    _run("gloo")
