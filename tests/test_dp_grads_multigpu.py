"""GPU (needs >= 2 devices; skipped on a one-GPU box): data-parallel gradient parity.  Two ranks render halves of one ray batch,
all-reduce through the hand-written switch kernel (csrc/collective.cu) and must end up with the gradients of a single-GPU backward
over the whole batch; the visibility votes must match exactly.  The work is in tools/dp_grad_parity.py (launched under torchrun)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("reducer", ["switch", "nccl"])
def test_two_rank_gradients_match_single_gpu(reducer):
    env = dict(os.environ, OCCNERF_REDUCER=reducer, ENGINE="tf32")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29513" if reducer == "switch" else "29514", os.path.join(ROOT, "tools", "dp_grad_parity.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(r.stdout[r.stdout.index("{"):])
    assert res["all_ranks_ok"] and res["votes_equal"] and res["worst"] < 2e-2, res
