"""Root-parallel search with the merge fused into the search kernel (peer stores over NVLink) against the NCCL
all-reduce of the same statistics.  Needs two GPUs on the box; skipped otherwise (the 1-GPU tiers cover the kernel's
single-rank path, tests/test_dist_cpu.py the host-side merge on gloo)."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_fused_merge_equals_nccl_all_reduce():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "fused_merge_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=500)
    assert res.returncode == 0 and "FUSED-MERGE-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
