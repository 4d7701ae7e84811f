"""`-m gpu`, needs >= 2 GPUs (skipped otherwise): the fused data-parallel optimiser step over peer memory
(`pka_dp_adam_step`: gradient reduce-scatter + Adam on the shard + parameter all-gather in one kernel) must leave every
rank with the parameters of {all-reduce SUM, Adam}.  Runs tools/check_peer_adam.py under torchrun on two ranks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_peer_adam_step_equals_allreduce_plus_adam():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tools", "check_peer_adam.py"), "--numel", "200003",
           "--steps", "3"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert '"ok": true' in res.stdout
