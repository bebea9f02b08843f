"""Row-sharded bank over >= 2 B200s: ncclAllGather + merge kernel vs the oracle on the whole bank.
Skipped on a single-GPU box (the driver's gpu tier); run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`."""
import subprocess
import sys
from pathlib import Path

import pytest

from speaker_diarization_toolkit_b200 import _native

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_sharded_bank_two_gpus():
    n = _native.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "tools" / "sharded_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
