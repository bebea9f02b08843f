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


def test_assign_batch_two_gpus_dp_and_sharded_bank(tmp_path):
    """The product CLI on 2 GPUs: `assign-batch --gpus 2` (recordings split, no collective) and
    `--gpus 2 --shard-bank` (bank rows split, NCCL all-gather + merge inside the library) both give the 1-GPU answer."""
    if _native.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import json
    from test_gpu_cli import build_batch_store, cli, strip_time
    manifest, mpath, env = build_batch_store(tmp_path, n_rec=7)
    base = ["speaker-assign", "-q", "assign-batch", str(mpath), "--threshold", "0.2", "--format", "json", "--dry-run"]
    rc, out, err = cli(*base, env=env)
    assert rc == 0, err
    one = strip_time(json.loads(out))
    rc, out, err = cli(*base, "--gpus", "2", env=env)
    assert rc == 0, err
    assert strip_time(json.loads(out)) == one
    rc, out, err = cli(*base, "--gpus", "2", "--shard-bank", env=env)
    assert rc == 0, err
    assert strip_time(json.loads(out)) == one
