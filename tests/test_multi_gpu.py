"""N > 1 on real GPUs (skipped on single-GPU boxes; the host logic is covered on CPU by test_sharding_gloo.py):
runs scripts/check_multi_gpu.py under torchrun — fused peer-store all-gather and NCCL all-gather against the
single-GPU matrix, bit for bit, plus first-error propagation."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_column_sharded_psi_equals_single_gpu():
    n = min(_ngpu(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "scripts", "check_multi_gpu.py")], capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
