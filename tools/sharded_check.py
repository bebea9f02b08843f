#!/usr/bin/env python3
"""Multi-GPU parity check of the row-sharded bank path (config 4): run under torchrun, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/sharded_check.py

Every rank loads its shard of the bank (cut on speaker boundaries), runs sdk_identify -- which ends with the single
ncclAllGather + merge kernel -- and compares the GLOBAL result with the C oracle run on the whole bank.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    from oracle import canonical
    from speaker_diarization_toolkit_b200 import _native, sharding, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [_native.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = _native.Context(local, world, rank, box[0])
    ok = True
    for name, path, dtype, pool, thr, k in [("tc-top10", 2, 1, 0, -1.0, 10), ("tc-thr", 2, 1, 1, 0.354, 10),
                                            ("exact-f32", 1, 0, 0, 0.3, 5), ("small-query", 2, 1, 0, -1.0, 10)]:
        rng = np.random.default_rng(11)
        rps = rng.choice([1, 2, 3], size=900)
        # "small-query": 7 segments of 4 labels -> the bank-stream kernels (last_path 4) on every shard, settled before the gather
        counts = [2, 1, 3, 1] if name == "small-query" else synth.zipf_counts(rng, 900, 8)
        case = synth.make_case(404, counts, 900, 128 if path == 1 else 512, rows_per_speaker=rps,
                               neighbours=5, impostor_frac=0.1)
        shards = sharding.shard_bank_rows(case.row_speaker, world)
        p0, p1 = shards[rank]
        ctx.set_option("path", path)
        ctx.bank_load(case.bank[p0:p1], case.row_speaker[p0:p1], case.row_trust[p0:p1], dtype=dtype, global_row_offset=p0)
        rows, scores, counts = ctx.identify(case.seg, case.seg_label, case.G, pool=pool, threshold=thr, k=k)
        ctx.assign(0.2, "low")
        out = ctx.fetch(with_assign=True)
        ref = canonical.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=dtype, pool=pool,
                                 threshold=thr, k=k)
        same = (np.array_equal(rows, ref[0]) and np.array_equal(counts, ref[2])
                and np.array_equal(scores.view(np.uint32), ref[1].view(np.uint32)))
        tr_ref = np.where(ref[0] >= 0, case.row_trust[np.clip(ref[0], 0, None)], 4)
        same = same and np.array_equal(out["trust"], tr_ref.astype(np.uint8))
        a = canonical.assign(ref[0], ref[1], tr_ref.astype(np.uint8), ref[2], 0.2, 2)
        same = same and np.array_equal(out["assign_idx"], a[0]) and np.array_equal(out["assign_score"], a[1])
        if name == "small-query":
            same = same and ctx.last_path()[0] == 4
        print(f"[rank {rank}/{world}] {name}: shard rows [{p0},{p1}) path={ctx.last_path()} -> {'OK' if same else 'MISMATCH'}", flush=True)
        ok = ok and same
    # ---- empty shards: more ranks than speaker runs: the ranks without rows still join the all-gather ----
    case = synth.make_case(7, [30, 12, 25], 1, 128, rows_per_speaker=3, impostor_frac=0.0)
    shards = sharding.shard_bank_rows(case.row_speaker, world)
    p0, p1 = shards[rank]
    assert sum(1 for a, b in shards if b > a) == 1
    ctx.set_option("path", 0)
    ctx.bank_load(case.bank[p0:p1], case.row_speaker[p0:p1], case.row_trust[p0:p1], dtype=1, global_row_offset=p0)
    rows, scores, counts = ctx.identify(case.seg, case.seg_label, case.G, pool=0, threshold=-1.0, k=3)
    ref = canonical.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=1, pool=0, threshold=-1.0, k=3)
    same = np.array_equal(rows, ref[0]) and np.array_equal(counts, ref[2]) and np.array_equal(scores.view(np.uint32), ref[1].view(np.uint32))
    print(f"[rank {rank}/{world}] empty-shard: shard rows [{p0},{p1}) -> {'OK' if same else 'MISMATCH'}", flush=True)
    ok = ok and same

    # ---- a rank that fails locally must not hang its peers: it joins the collective with a status word, returns its
    #      own error, and the others get SDK_EPEER at fetch time ----
    rng = np.random.default_rng(3)
    case = synth.make_case(9, synth.zipf_counts(rng, 300, 4), 64 * world, 64, impostor_frac=0.0)
    shards = sharding.shard_bank_rows(case.row_speaker, world)
    p0, p1 = shards[rank]
    ctx.bank_load(case.bank[p0:p1], case.row_speaker[p0:p1], case.row_trust[p0:p1], dtype=1, global_row_offset=p0)
    bad = world - 1
    ctx.set_option("path", 1)
    err = None
    try:
        if rank == bad:
            ctx.set_option("inject_fail", 1)       # test knob: the next local pass fails (SDK_EINVAL) after its first kernels
        ctx.identify(case.seg, case.seg_label, case.G, pool=0, threshold=-1.0, k=3)
    except _native.NativeError as e:
        err = e
    want = (err is not None) and (err.code == (-22 if rank == bad else -70))
    print(f"[rank {rank}/{world}] failing-peer: error {err.code if err else None} ({'expected' if want else 'UNEXPECTED'})", flush=True)
    ok = ok and want
    # the communicator is still usable afterwards
    rows, scores, counts = ctx.identify(case.seg, case.seg_label, case.G, pool=0, threshold=-1.0, k=3)
    ref = canonical.identify(case.seg, case.goff, case.bank, case.row_speaker, case.n_speakers, mode=1, pool=0, threshold=-1.0, k=3)
    same = np.array_equal(rows, ref[0]) and np.array_equal(scores.view(np.uint32), ref[1].view(np.uint32))
    print(f"[rank {rank}/{world}] after-failure: -> {'OK' if same else 'MISMATCH'}", flush=True)
    ok = ok and same

    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    ctx.close()
    dist.destroy_process_group()
    if flag.item() != 0:
        sys.exit(1)
    if rank == 0:
        print("SHARDED_CHECK_OK")


if __name__ == "__main__":
    main()
