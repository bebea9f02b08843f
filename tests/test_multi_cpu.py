"""N>1 host logic on CPU: world-size-2 gloo.  Each rank works on its shard with the NumPy oracle (test
infrastructure), the candidates are all-gathered over gloo and merged with the host mirror of the device merge
kernel; the result must equal the single-process answer.  Covers both multi-GPU modes of SURVEY 8e."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import matching_np as mnp
    from speaker_diarization_toolkit_b200 import sharding, synth
    try:
        rng = np.random.default_rng(3)
        rps = rng.choice([1, 2, 3], size=200)
        case = synth.make_case(404, synth.zipf_counts(rng, 300, 6), 200, 64, rows_per_speaker=rps, neighbours=4, impostor_frac=0.0)
        k = 6
        full = mnp.identify(case.seg, case.goff, case.bank, case.row_speaker, threshold=-1.0, k=k)
        # ---- mode 1: bank row-sharded, one all-gather of the local top-k, merge ----
        p0, p1 = sharding.shard_bank_rows(case.row_speaker, world)[rank]
        assert p0 == 0 or case.row_speaker[p0] != case.row_speaker[p0 - 1], "shards must be cut on speaker boundaries"
        loc = mnp.identify(case.seg, case.goff, case.bank[p0:p1], case.row_speaker[p0:p1], threshold=-1.0, k=k, row_offset=p0)
        rows = [torch.empty((case.G, k), dtype=torch.int64) for _ in range(world)]
        scores = [torch.empty((case.G, k), dtype=torch.float32) for _ in range(world)]
        counts = [torch.empty((case.G,), dtype=torch.int32) for _ in range(world)]
        dist.all_gather(rows, torch.from_numpy(loc[0]))
        dist.all_gather(scores, torch.from_numpy(loc[1]))
        dist.all_gather(counts, torch.from_numpy(loc[2]))
        mr, ms, mc, _ = sharding.merge_topk_lists(torch.stack(rows).numpy(), torch.stack(scores).numpy(), torch.stack(counts).numpy(), k)
        assert np.array_equal(mr, full[0]) and np.array_equal(mc, full[2])
        np.testing.assert_allclose(ms, full[1], rtol=1e-6)
        # ---- mode 2: data-parallel over recordings, no collective on the data path ----
        rec_counts = np.asarray([case.goff[2] - case.goff[0], case.goff[4] - case.goff[2], case.goff[6] - case.goff[4]])
        r0, r1 = sharding.partition_recordings(rec_counts, world)[rank]
        g0, g1 = 2 * r0, 2 * r1
        s0, s1 = int(case.goff[g0]), int(case.goff[g1])
        mine = mnp.identify(case.seg[s0:s1], case.goff[g0:g1 + 1] - s0, case.bank, case.row_speaker, threshold=-1.0, k=k)
        assert np.array_equal(mine[0], full[0][g0:g1]) and np.array_equal(mine[2], full[2][g0:g1])
        covered = [None] * world
        dist.all_gather_object(covered, (r0, r1))
        assert covered[0][0] == 0 and covered[-1][1] == len(rec_counts) and all(covered[i][1] == covered[i + 1][0] for i in range(world - 1))
        q.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        q.put((rank, f"{type(exc).__name__}: {exc}"))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_and_dp():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_partition_and_shard_edges():
    sys.path.insert(0, str(ROOT))
    from speaker_diarization_toolkit_b200 import sharding
    assert sharding.partition_recordings([5, 5, 5, 5], 2) == [(0, 2), (2, 4)]
    parts = sharding.partition_recordings([1, 100, 1, 1, 1], 4)
    assert parts[0][0] == 0 and parts[-1][1] == 5 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert sharding.partition_recordings([7], 3)[-1][1] == 1
    spk = np.asarray([0, 0, 0, 1, 2, 2, 3, 3, 3, 3])
    sh = sharding.shard_bank_rows(spk, 2)
    assert sh == [(0, 6), (6, 10)] or sh == [(0, 4), (4, 10)]
    for p0, p1 in sharding.shard_bank_rows(spk, 4):
        assert p0 == p1 or p0 == 0 or spk[p0] != spk[p0 - 1]
    with pytest.raises(ValueError):
        sharding.shard_bank_rows(np.asarray([0, 1, 0]), 2)
    r = np.asarray([[[5, 9, -1]], [[2, 7, 8]]], np.int64)
    s = np.asarray([[[0.9, 0.5, 0]], [[0.9, 0.6, 0.1]]], np.float32)
    c = np.asarray([[2], [3]], np.int32)
    mr, ms, mc, _ = sharding.merge_topk_lists(r, s, c, 3)
    assert mr[0].tolist() == [2, 5, 7] and mc[0] == 3          # 0.9 tie -> lower global row first
