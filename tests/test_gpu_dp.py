"""Multi-GPU parity of the CUDA path inside the driver-run suite: launches tests/dp_check.py under torchrun on
two GPUs (NCCL) and keeps its log.  dp_check compares the sharded, all-reduced closure (loss, all 110 gradients,
the centre numerator / denominator sums that ride in the same exchange buffer), the graphed data-parallel loop and
the sharded loader statistics with the unsharded CPU oracle."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests -m gpu -k dp)")
def test_data_parallel_closure_matches_unsharded_oracle_nccl():
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    log = r.stdout + "\n---- stderr ----\n" + r.stderr
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, f"dp_check_world{world}.log"), "w") as fh:
            fh.write(log)
    except OSError:
        pass
    assert r.returncode == 0, log[-4000:]
    assert r.stdout.count("PASS") >= 4 and "FAIL" not in r.stdout, log[-4000:]
