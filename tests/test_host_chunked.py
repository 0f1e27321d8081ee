"""HOST-buffer batches in pipelined chunks (csrc/rcz_internal.h host_chunked): bwt decode / encode, flate and zlib cut a big host batch
into chunks of consecutive units, upload chunk k+1 and download chunk k-1 while chunk k runs.  The results must not depend on where
the cuts fall: every op is run with a tiny chunk size (RCZ_HOST_CHUNK_BYTES) and with chunking off, on ragged units with gaps, bad
units included, against the oracle.  On the emulator (CPU) and on the GPU with page-locked and pageable buffers."""
import zlib as pyzlib

import numpy as np
import pytest

from util import pack, out_layout


def _host(arr, pinned):
    if not pinned:
        return arr, arr
    import torch
    t = torch.from_numpy(arr).pin_memory()
    return t, t.numpy()


def _run(ctx, oracle, gen, pinned):
    sizes = [3000 + 977 * i for i in range(23)]
    raw = [gen.one("hextext" if i % 2 else "lzsyn", 900 + i, n) for i, n in enumerate(sizes)]
    # ---- bwt encode / decode
    inb, in_off, in_len = pack(raw, pad_front=3, gap=5, align=1)
    out_off, out_cap, total = out_layout(sizes, gap=7)
    n_arr = in_len.copy()
    n_arr[4] = 0                                                            # empty block: MALFORMED on encode (bwt/mod.rs:186-188)
    keep_in, h_in = _host(inb, pinned)
    keep_out, h_out = _host(np.full(total, 0xAA, dtype=np.uint8), pinned)
    origin, status = ctx.bwt_encode_blocks(h_in, in_off, n_arr, h_out, out_off)
    ls = []
    for i, r in enumerate(raw):
        if i == 4:
            assert status[i] != 0
            ls.append(b"")
            continue
        st, l, org = oracle.bwt_encode(r)
        assert status[i] == 0 and int(origin[i]) == org and h_out[int(out_off[i]): int(out_off[i]) + len(r)].tobytes() == l, i
        ls.append(l)
    lb, l_off, l_len = pack(ls, pad_front=1, gap=2, align=1)
    org2 = np.array([int(o) for o in origin], dtype=np.uint32)
    org2[7] = 0xFFFFFF                                                      # origin beyond the block: an error in the reference
    keep_l, h_l = _host(lb, pinned)
    keep_o2, h_o2 = _host(np.full(total, 0xAA, dtype=np.uint8), pinned)
    out_len, status = ctx.bwt_decode_blocks(h_l, l_off, l_len, org2, h_o2, out_off)
    for i, r in enumerate(raw):
        if i == 4:
            continue
        if i == 7:
            assert status[i] != 0
            continue
        assert status[i] == 0 and int(out_len[i]) == len(r) and h_o2[int(out_off[i]): int(out_off[i]) + len(r)].tobytes() == r, i
    # ---- flate / zlib
    z = [pyzlib.compress(r, 6) for r in raw]
    z[9] = z[9][: len(z[9]) // 2]                                           # truncated stream
    caps = list(sizes)
    caps[11] -= 50                                                          # too small
    for wrapper in (False, True):
        units = z if wrapper else [u[2:-4] for u in z]
        zb, z_off, z_len = pack(units, pad_front=2, gap=3, align=1)
        o_off, o_cap, tot = out_layout(caps, gap=5)
        keep_z, h_z = _host(zb, pinned)
        keep_zo, h_zo = _host(np.full(tot, 0xAA, dtype=np.uint8), pinned)
        res = (ctx.zlib_decode_streams if wrapper else ctx.flate_decode_streams)(h_z, z_off, z_len, h_zo, o_off, o_cap)
        out_len, status = res[0], res[1]
        for i, r in enumerate(raw):
            ref = oracle.zlib_decode(units[i], caps[i]) if wrapper else oracle.flate_decode(units[i], caps[i])
            assert int(status[i]) == ref[0], (wrapper, i, int(status[i]), ref[0])
            if ref[0] == 0:
                assert int(out_len[i]) == len(r) and h_zo[int(o_off[i]): int(o_off[i]) + len(r)].tobytes() == r, (wrapper, i)
        assert status[9] != 0 and status[11] != 0


@pytest.mark.parametrize("chunk", ["20000", "0"])
def test_host_chunked_emu(emu_ctx, oracle, gen, chunk, monkeypatch):
    monkeypatch.setenv("RCZ_HOST_CHUNK_BYTES", chunk)
    _run(emu_ctx, oracle, gen, False)


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("chunk", ["20000", "0"])
def test_host_chunked_gpu(gpu_ctx, oracle, gen, chunk, pinned, monkeypatch):
    monkeypatch.setenv("RCZ_HOST_CHUNK_BYTES", chunk)
    _run(gpu_ctx, oracle, gen, pinned)


def test_host_chunked_layout_edge_cases_emu(emu_ctx, oracle, gen, monkeypatch):
    """Layouts the chunk cutter must cope with: empty units between full ones, units stored in REVERSE order (a chunk's span would drag
    in the other chunks' bytes: the batch must fall back to the single-shot path), a batch of one unit."""
    monkeypatch.setenv("RCZ_HOST_CHUNK_BYTES", "8000")
    raw = [gen.one("hextext", 40 + i, 2500 + 300 * i) for i in range(12)]
    z = [pyzlib.compress(r, 6)[2:-4] for r in raw]
    # (a) empty streams in between: status of an empty DEFLATE stream is the oracle's, the neighbours decode
    units = []
    for i, u in enumerate(z):
        units.append(u)
        if i % 3 == 0:
            units.append(b"")
    caps = [4000 + 300 * 12] * len(units)
    zb, z_off, z_len = pack(units, pad_front=1, gap=2, align=1)
    o_off, o_cap, tot = out_layout(caps, gap=3)
    out = np.full(tot, 0xAA, dtype=np.uint8)
    out_len, status, used, detail = emu_ctx.flate_decode_streams(zb, z_off, z_len, out, o_off, o_cap)
    k = 0
    for i, u in enumerate(units):
        ref = oracle.flate_decode(u, caps[i])
        assert int(status[i]) == ref[0], i
        if u:
            assert out[int(o_off[i]): int(o_off[i]) + int(out_len[i])].tobytes() == raw[k]
            k += 1
    # (b) the same streams, stored back to front
    order = list(range(len(z)))[::-1]
    zb, z_off, z_len = pack([z[i] for i in order], pad_front=0, gap=1, align=1)
    z_off, z_len = z_off[::-1].copy(), z_len[::-1].copy()         # unit i is z[i] again, at descending offsets
    caps = [len(r) for r in raw]
    o_off, o_cap, tot = out_layout(caps, gap=3)
    out = np.full(tot, 0xAA, dtype=np.uint8)
    out_len, status, used, detail = emu_ctx.flate_decode_streams(zb, z_off, z_len, out, o_off, o_cap)
    for i, r in enumerate(raw):
        assert status[i] == 0 and out[int(o_off[i]): int(o_off[i]) + len(r)].tobytes() == r, i
    # (c) one unit
    zb, z_off, z_len = pack([z[0]], pad_front=3, gap=0, align=1)
    o_off, o_cap, tot = out_layout([len(raw[0])], gap=0)
    out = np.zeros(tot, dtype=np.uint8)
    out_len, status, used, detail = emu_ctx.flate_decode_streams(zb, z_off, z_len, out, o_off, o_cap)
    assert status[0] == 0 and out[: len(raw[0])].tobytes() == raw[0]
