"""N > 1 path on CPU: two gloo ranks each take their contiguous range of LZ4 blocks (balanced by compressed bytes), decode it
through the C ABI (emulation build: no GPU here), all-gather the decoded shards and compare with the oracle.  Mirrors what
bench.py / a multi-GPU caller does with one process per GPU over NCCL."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_partition_covers_and_balances():
    shard = importlib.import_module("rust-compress_b200.shard")
    rs = np.random.RandomState(0)
    for n, world in [(0, 4), (1, 8), (7, 2), (256, 8), (1000, 3)]:
        w = rs.randint(1, 1000, size=n)
        parts = shard.partition(w, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        if n >= 64 * world:
            loads = [w[a:b].sum() for a, b in parts]
            assert max(loads) <= 1.25 * (w.sum() / world)
    assert shard.partition([5, 5, 5, 5], 2) == [(0, 2), (2, 4)]
    assert shard.partition(np.ones(256), 8) == [(32 * r, 32 * r + 32) for r in range(8)]


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        rcz = importlib.import_module("rust-compress_b200")
        shard = importlib.import_module("rust-compress_b200.shard")
        from oracle import oracle
        from tools import gen
        unit, count = 30000, 13
        raw = gen.units("lzsyn", gen.unit_seed(2, 0), unit, count, nthreads=1)
        packed, off, lens = gen.lz4_compress_units(raw, unit, count, nthreads=1)
        lo, hi = shard.my_range(lens, world, rank)
        ctx = rcz.Context(emu=True)
        out = np.zeros(unit * (hi - lo) + 64, dtype=np.uint8)
        out_off = np.arange(hi - lo, dtype=np.uint64) * unit
        out_len, st = ctx.lz4_decode_blocks(packed, off[lo:hi], lens[lo:hi], out, out_off, np.full(hi - lo, unit, np.uint64))
        assert (st == 0).all() and (out_len == unit).all()
        parts = shard.gather_shards(torch.from_numpy(out), unit * (hi - lo), dist)
        whole = b"".join(bytes(p.numpy()) for p in parts)
        ref = np.zeros(unit * count + 64, dtype=np.uint8)
        oracle.lz4_decode_blocks_mt(packed, off, lens, ref, np.arange(count, dtype=np.uint64) * unit, np.full(count, unit, np.uint64), 1)
        ok = whole == raw.tobytes() == bytes(ref[: unit * count])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok, (lo, hi)))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc() + repr(e)))


def test_two_rank_gloo_decode_and_gather():
    import torch.multiprocessing as mp
    importlib.import_module("rust-compress_b200.build").build_emu()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in res:
        assert ok is True, info
    ranges = sorted(info for _, _, info in res)
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == 13
