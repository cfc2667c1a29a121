"""Batch sharding across the GPUs of one box: one process per GPU, utterances are independent.

The reference's only parallelism is a thread pool over utterances (reference
ctc_beam_search_decoder.cpp:259-284); the multi-GPU equivalent is a plain partition of the batch.  There is
no exchange step inside the algorithm, so the only communication is the optional scatter of inputs from /
gather of results to one rank (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world_size, rank):
    """Contiguous shard [lo, hi) of `batch` items for `rank`: sizes differ by at most one."""
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_length(seq_lens, world_size):
    """Length-balanced partition: utterances sorted by length, longest first (ties: lower index
    first), dealt round-robin -- rank r gets the r-th, (r + W)-th, ... of that order.  An utterance is as many serial
    frames as it is long, so this evens out both the frames per GPU and the longest utterance per GPU.  Returns one
    index list per rank (sizes differ by at most one)."""
    lens = [int(x) for x in seq_lens]
    order = sorted(range(len(lens)), key=lambda i: (-lens[i], i))
    return [order[r::world_size] for r in range(world_size)]


def decode_sharded(decode_fn, probs, seq_lens=None, group=None, src=0, device=None, balance=None):
    """Rank `src` holds probs [B, T, V] (+ seq_lens [B]); every rank decodes its shard with
    decode_fn(probs_shard, seq_lens_shard) -> (tokens [b, K, T], scores [b, K], timesteps [b, K, T],
    lens [b, K]) and rank `src` gets the result in the original order (others get None).  Shards are contiguous
    (shard_bounds) unless balance="length" and seq_lens is given: then shard_by_length deals the utterances out.

    Tensors travel on `device` (the rank's GPU for NCCL; CPU for gloo)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if balance not in (None, "length"):
        raise ValueError("balance must be None or 'length'")
    meta = [None]
    if rank == src:
        assign = shard_by_length(seq_lens.tolist(), world) if (balance == "length" and seq_lens is not None) else None
        meta = [(tuple(probs.shape), seq_lens is not None, assign)]
    dist.broadcast_object_list(meta, src=src, group=group)
    (B, T, V), has_lens, assign = meta[0]
    if assign is not None:
        return _decode_assigned(decode_fn, probs, seq_lens, group, src, device, assign, (B, T, V), has_lens)
    lo, hi = shard_bounds(B, world, rank)
    dev = device if device is not None else (probs.device if rank == src else torch.device("cpu"))
    my_probs = torch.empty(hi - lo, T, V, dtype=torch.float32, device=dev)
    my_lens = torch.empty(hi - lo, dtype=torch.int32, device=dev) if has_lens else None
    # scatter (variable sizes -> point-to-point)
    if rank == src:
        reqs = []
        for r in range(world):
            a, b = shard_bounds(B, world, r)
            if r == src:
                my_probs.copy_(probs[a:b])
                if has_lens:
                    my_lens.copy_(seq_lens[a:b])
            elif b > a:
                reqs.append(dist.isend(probs[a:b].to(dev).contiguous(), r, group=group))
                if has_lens:
                    reqs.append(dist.isend(seq_lens[a:b].to(dev, torch.int32).contiguous(), r, group=group))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(my_probs, src, group=group)
        if has_lens:
            dist.recv(my_lens, src, group=group)
    out = decode_fn(my_probs, my_lens) if hi > lo else None
    # gather
    if rank != src:
        if hi > lo:
            for t in out:
                dist.send(t.to(dev).contiguous(), src, group=group)
        return None
    parts = [None] * world
    parts[src] = out
    for r in range(world):
        a, b = shard_bounds(B, world, r)
        if r == src or b == a:
            continue
        K = out[1].shape[1] if out is not None else None
        if K is None:
            raise RuntimeError("source rank must own a non-empty shard")
        bufs = [torch.empty(b - a, K, T, dtype=torch.int32, device=dev), torch.empty(b - a, K, dtype=torch.float32, device=dev),
                torch.empty(b - a, K, T, dtype=torch.int32, device=dev), torch.empty(b - a, K, dtype=torch.int32, device=dev)]
        for t in bufs:
            dist.recv(t, r, group=group)
        parts[r] = tuple(bufs)
    parts = [p for p in parts if p is not None]
    return tuple(torch.cat([p[i].to(dev) for p in parts], dim=0) for i in range(4))


def _decode_assigned(decode_fn, probs, seq_lens, group, src, device, assign, shape, has_lens):
    """decode_sharded for an arbitrary assignment of utterances to ranks (index lists): rows travel gathered by index
    and come back scattered to where they belong."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, T, V = shape
    mine = assign[rank]
    dev = device if device is not None else (probs.device if rank == src else torch.device("cpu"))
    my_probs = torch.empty(len(mine), T, V, dtype=torch.float32, device=dev)
    my_lens = torch.empty(len(mine), dtype=torch.int32, device=dev) if has_lens else None
    if rank == src:
        reqs = []
        for r in range(world):
            if not assign[r]:
                continue
            idx = torch.tensor(assign[r], dtype=torch.long, device=probs.device)
            p_r = probs.index_select(0, idx)
            l_r = seq_lens.to(probs.device).index_select(0, idx).to(torch.int32) if has_lens else None
            if r == src:
                my_probs.copy_(p_r)
                if has_lens:
                    my_lens.copy_(l_r)
            else:
                reqs.append(dist.isend(p_r.to(dev).contiguous(), r, group=group))
                if has_lens:
                    reqs.append(dist.isend(l_r.to(dev).contiguous(), r, group=group))
        for q in reqs:
            q.wait()
    elif mine:
        dist.recv(my_probs, src, group=group)
        if has_lens:
            dist.recv(my_lens, src, group=group)
    out = decode_fn(my_probs, my_lens) if mine else None
    if rank != src:
        if mine:
            for t in out:
                dist.send(t.to(dev).contiguous(), src, group=group)
        return None
    if out is None:
        raise RuntimeError("source rank must own a non-empty shard")
    K = out[1].shape[1]
    res = [torch.empty(B, K, T, dtype=torch.int32, device=dev), torch.empty(B, K, dtype=torch.float32, device=dev),
           torch.empty(B, K, T, dtype=torch.int32, device=dev), torch.empty(B, K, dtype=torch.int32, device=dev)]
    for r in range(world):
        if not assign[r]:
            continue
        if r == src:
            part = [t.to(dev) for t in out]
        else:
            n = len(assign[r])
            part = [torch.empty(n, K, T, dtype=torch.int32, device=dev), torch.empty(n, K, dtype=torch.float32, device=dev),
                    torch.empty(n, K, T, dtype=torch.int32, device=dev), torch.empty(n, K, dtype=torch.int32, device=dev)]
            for t in part:
                dist.recv(t, r, group=group)
        idx = torch.tensor(assign[r], dtype=torch.long, device=dev)
        for dst, t in zip(res, part):
            dst.index_copy_(0, idx, t)
    return tuple(res)
