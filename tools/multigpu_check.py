#!/usr/bin/env python
"""Multi-GPU checks on hardware (gpurun --gpus N), two modes:

  torchrun --nproc-per-node N tools/multigpu_check.py nccl [--batch 512]
      ctcdecode_b200.sharding.decode_sharded over NCCL: rank 0 holds the batch, every rank decodes its contiguous shard
      on its own GPU, rank 0 gets the gathered result and compares it with decoding everything itself.

  python tools/multigpu_check.py host [--batch 2048]
      ONE process, ONE call: ctcdec_decode_batch_host_multi shards a host batch over all visible GPUs (a worker thread
      per device); compares with the single-GPU host call and times both end to end (host buffers in, host buffers out).

Each prints one JSON line (kept under profiles/)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctcdecode_b200 import CTCBeamDecoder  # noqa: E402
from ctcdecode_b200.synth import ctc_like_probs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("mode", choices=["nccl", "host"])
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
T, V, K = 1000, 29, 100
labels = [str(i) for i in range(V)]

if a.mode == "nccl":
    import torch.distributed as dist
    from ctcdecode_b200.sharding import decode_sharded
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = a.batch or 512
    dec = CTCBeamDecoder(labels, beam_width=K, device_outputs=True)
    probs = ctc_like_probs(B, T, V, seed=5).to(dev) if rank == 0 else None
    fn = lambda p, sl: dec.decode(p, sl)  # noqa: E731
    out = decode_sharded(fn, probs, None, device=dev)      # warm-up (communicator, kernels)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out = decode_sharded(fn, probs, None, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / a.steps
    if rank == 0:
        ref = dec.decode(probs)
        lens = ref[3]
        col = torch.arange(T, device=dev)[None, None, :] < lens[:, :, None]
        same = bool(torch.equal(out[1], ref[1]) and torch.equal(out[3], lens)
                    and torch.equal(torch.where(col, out[0], 0), torch.where(col, ref[0], 0))
                    and torch.equal(torch.where(col, out[2], 0), torch.where(col, ref[2], 0)))
        print(json.dumps({"check": "decode_sharded over NCCL", "world_size": world, "batch": B, "T": T, "beam": K,
                          "equals_single_gpu": same, "ms_per_call": dt * 1e3, "utterances_per_s": B / dt,
                          "note": "scatter from / gather to rank 0 over NCCL point-to-point, device tensors"}))
        assert same
    dist.barrier()
    dist.destroy_process_group()
else:
    n = torch.cuda.device_count()
    B = a.batch or 2048
    probs = ctc_like_probs(B, T, V, seed=6).pin_memory()
    one = CTCBeamDecoder(labels, beam_width=K, device="cuda:0")
    many = CTCBeamDecoder(labels, beam_width=K, device="all")

    def timed(dec):
        dec.decode(probs)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            r = dec.decode(probs)
        return (time.perf_counter() - t0) / a.steps, r

    t1, r1 = timed(one)
    tn, rn = timed(many)
    lens = r1[3]
    col = torch.arange(T)[None, None, :] < lens[:, :, None]
    same = bool(torch.equal(r1[1], rn[1]) and torch.equal(lens, rn[3]) and torch.equal(torch.where(col, r1[0], 0), torch.where(col, rn[0], 0)))
    # the entry point itself, the way a serving host calls it: page-locked result arrays allocated once and reused (the
    # Python wrapper above allocates fresh [B, beam, T] tensors per call like the reference's __init__.py does -- at this
    # batch 1.6 GB of pages to fault in, which is most of its time)
    import ctypes
    from ctcdecode_b200 import _native
    lib = _native.load()
    cfg = _native.Config(V, K, 0, 0, 40, 1.0)
    tok, ts = torch.empty(B, K, T, dtype=torch.int32).pin_memory(), torch.empty(B, K, T, dtype=torch.int32).pin_memory()
    sc, ln = torch.empty(B, K, dtype=torch.float32).pin_memory(), torch.zeros(B, K, dtype=torch.int32).pin_memory()
    nres, fl = torch.zeros(B, dtype=torch.int32), torch.zeros(B, dtype=torch.int32)

    def timed_c(devs):
        arr = (ctypes.c_int * len(devs))(*devs)
        def call():
            _native.check(lib.ctcdec_decode_batch_host_multi(
                ctypes.byref(cfg), probs.data_ptr(), None, B, T, tok.data_ptr(), ts.data_ptr(), sc.data_ptr(),
                ln.data_ptr(), nres.data_ptr(), fl.data_ptr(), arr, len(devs)))
        call()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            call()
        return (time.perf_counter() - t0) / a.steps

    c1 = timed_c([0])
    cn = timed_c(list(range(n)))
    same_c = bool(torch.equal(sc, rn[1]) and torch.equal(ln, rn[3]) and torch.equal(torch.where(col, tok, 0), torch.where(col, rn[0], 0)))
    print(json.dumps({"check": "ctcdec_decode_batch_host_multi (one process, one call)", "gpus": n, "batch": B, "T": T,
                      "beam": K, "equals_single_gpu": same and same_c,
                      "c_abi_page_locked_buffers": {"one_gpu_ms": c1 * 1e3, "all_gpus_ms": cn * 1e3,
                                                    "one_gpu_utt_per_s": B / c1, "all_gpus_utt_per_s": B / cn,
                                                    "speedup": c1 / cn},
                      "python_wrapper_fresh_tensors": {"one_gpu_ms": t1 * 1e3, "all_gpus_ms": tn * 1e3,
                                                       "one_gpu_utt_per_s": B / t1, "all_gpus_utt_per_s": B / tn,
                                                       "speedup": t1 / tn},
                      "note": "end to end, host probs in, host results out; the wrapper allocates [B, beam, T] result "
                              "tensors per call (reference semantics), the C-ABI leg reuses page-locked ones"}))
    assert same and same_c
