"""N>1 path on CPU: two gloo ranks shard a batch, decode their shards (here with the CPU emulation of the CTA
program standing in for the GPU) and the source rank must get exactly the unsharded result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ctcdecode_b200.sharding import decode_sharded, shard_bounds, shard_by_length


def test_shard_bounds_partition():
    for B in (0, 1, 7, 8, 256, 2049):
        for W in (1, 2, 3, 8):
            got = [shard_bounds(B, W, r) for r in range(W)]
            assert got[0][0] == 0 and got[-1][1] == B
            assert all(got[i][1] == got[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_length_deals_longest_first():
    lens = [40, 3, 0, 25, 40, 17, 25]
    parts = shard_by_length(lens, 3)
    assert sorted(i for p in parts for i in p) == list(range(7))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert parts == [[0, 6, 2], [4, 5], [3, 1]]
    frames = [sum(lens[i] for i in p) for p in parts]
    assert max(frames) - min(frames) <= max(lens)           # contiguous thirds would give 43 / 82 / 25
    assert shard_by_length([], 2) == [[], []]


def _emul_decode(probs, seq_lens):
    from tests import emul
    r = emul.decode(probs.numpy(), None if seq_lens is None else seq_lens.numpy(), beam=12)
    return (torch.from_numpy(r["tokens"]), torch.from_numpy(r["scores"]), torch.from_numpy(r["timesteps"]),
            torch.from_numpy(r["lens"]))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ctcdecode_b200.synth import ctc_like_probs
    probs = ctc_like_probs(5, 40, 9, seed=3) if rank == 0 else None
    lens = torch.tensor([40, 3, 0, 25, 40], dtype=torch.int32) if rank == 0 else None
    out = decode_sharded(_emul_decode, probs, lens)
    bal = decode_sharded(_emul_decode, probs, lens, balance="length")   # utterances dealt out by length
    if rank == 0:
        q.put([t.numpy() for t in out] + [t.numpy() for t in bal])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single():
    from tests import emul
    emul.build()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from ctcdecode_b200.synth import ctc_like_probs
    full = _emul_decode(ctc_like_probs(5, 40, 9, seed=3), torch.tensor([40, 3, 0, 25, 40], dtype=torch.int32))
    lens = full[3].numpy()
    assert np.array_equal(got[3], lens) and np.array_equal(got[1].view(np.int32), full[1].numpy().view(np.int32))
    for b in range(5):
        for p in range(12):
            L = lens[b, p]
            assert np.array_equal(got[0][b, p, :L], full[0].numpy()[b, p, :L])
            assert np.array_equal(got[2][b, p, :L], full[2].numpy()[b, p, :L])
    # the length-balanced partition returns the same rows in the same (original) order
    assert np.array_equal(got[7], lens) and np.array_equal(got[5].view(np.int32), got[1].view(np.int32))
    for b in range(5):
        for p in range(12):
            L = lens[b, p]
            assert np.array_equal(got[4][b, p, :L], got[0][b, p, :L]) and np.array_equal(got[6][b, p, :L], got[2][b, p, :L])
