"""Data parallelism on the path that is timed: needs >= 2 GPUs, skipped otherwise.
Runs tools/dp_check.py under torchrun: the graphed DP step (NCCL buckets captured in the step graph, dynamic tile
schedule, side-stream weight gradients) leaves the SUM of the ranks' single-GPU gradients in every rank's buffer and
the ranks bit-identical after clip + AdamW."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_graphed_data_parallel_step_reduces_exactly_and_keeps_ranks_identical():
    env = dict(os.environ, DP_CHECK_BATCH="4")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "dp_check.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, (p.stdout[-1500:], p.stderr[-3000:])
    d = json.loads(lines[-1])
    assert d["ok"] and d["world"] == 2 and d["ranks_bit_identical_after_adamw"] and d["optimizer_steps"] == 3
    assert min(d["grad_cos"]) >= 0.99999 and d["buckets"] >= 6
