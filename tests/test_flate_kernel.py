"""Inflate kernel (csrc/flate_decode.cu) against the oracle (flate.rs:69-147, 195-450): bytes, status, flate detail code,
bytes consumed.  Known-answer fixtures of flate.rs:528-548 (test.z.0-9 minus zlib framing, raw test.z.go), all three
BTYPEs from Python zlib (stored / Z_FIXED / dynamic, multi-block, sync-flush empty blocks), truncations and byte fuzz."""
import random
import zlib

import numpy as np
import pytest

from conftest import golden
from util import run_batch

TXT = golden("ref_test.txt")


def raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=0):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    if not flush_every:
        return c.compress(data) + c.flush()
    out = b""
    for i in range(0, len(data), flush_every):
        out += c.compress(data[i: i + flush_every]) + c.flush(zlib.Z_SYNC_FLUSH)
    return out + c.flush()


def _streams(gen, big):
    hx = gen.one("hextext", 21, big)
    lz = gen.one("lzsyn", 22, big)
    rnd = gen.one("random", 23, big // 4)
    s = {
        "fixtures": [golden("ref_test.z.%d" % i)[2:-4] for i in range(10)] + [golden("ref_test.z.go")],
        "dynamic": [raw_deflate(TXT), raw_deflate(hx), raw_deflate(lz, 9), raw_deflate(bytes(100000)), raw_deflate(rnd)],
        "fixed": [raw_deflate(TXT, 6, zlib.Z_FIXED), raw_deflate(hx[:50000], 6, zlib.Z_FIXED), raw_deflate(b"a", 6, zlib.Z_FIXED),
                  raw_deflate(b"abcabcabcabcabcabc" * 50, 6, zlib.Z_FIXED)],
        "stored": [raw_deflate(TXT, 0), raw_deflate(rnd, 0), raw_deflate(b"", 0), raw_deflate(b"")],
        "multiblock": [raw_deflate(hx, 6, flush_every=7000), raw_deflate(lz, 1, flush_every=30011), raw_deflate(TXT * 40, 6, zlib.Z_FIXED, 999),
                       raw_deflate(hx, 0, flush_every=5000)],
        "overlap": [raw_deflate(bytes((i * 7) & 255 for i in range(p)) * (60000 // p + 1)) for p in (1, 2, 3, 5, 31, 32, 33, 258, 300)],
        "huffman_only": [raw_deflate(hx[:80000], 6, zlib.Z_HUFFMAN_ONLY), raw_deflate(rnd[:30000], 6, zlib.Z_RLE)],
    }
    return s


def _expected(oracle, stream, cap):
    st, out, used, detail = oracle.flate_decode(stream, cap)
    return st, out, used, detail


def _check(ctx, oracle, streams, caps, device=False, check_used=True):
    got, (in_used, detail) = run_batch(ctx, "flate_decode_streams", streams, caps, device=device, pad_front=3, gap=1)
    for i, (s, cap) in enumerate(zip(streams, caps)):
        st, out, used, det = _expected(oracle, s, cap)
        assert got[i][0] == st, (i, got[i][0], st, int(detail[i]), det)
        assert int(detail[i]) == (det if st == oracle.E_INVALID_INPUT else 0), (i, int(detail[i]), det)
        if st == 0:
            assert got[i][1] == out, "stream %d differs" % i
            if check_used:
                assert int(in_used[i]) == used, (i, int(in_used[i]), used)
        elif st != oracle.E_OUTPUT_FULL:
            assert got[i][1] == out, "stream %d: bytes before the error differ" % i


GROUPS = ["fixtures", "dynamic", "fixed", "stored", "multiblock", "overlap", "huffman_only"]


@pytest.mark.parametrize("group", GROUPS)
def test_inflate_emu(emu_ctx, oracle, gen, group):
    streams = _streams(gen, 120000)[group]
    caps = [len(zlib.decompress(s, -15)) + 5 for s in streams]
    _check(emu_ctx, oracle, streams, caps)
    if group == "fixtures":
        got, _ = run_batch(emu_ctx, "flate_decode_streams", streams, caps)
        assert all(g == (0, TXT) for g in got)


def _bad_streams(gen):
    good = [raw_deflate(TXT), raw_deflate(TXT, 6, zlib.Z_FIXED), raw_deflate(TXT, 0), raw_deflate(gen.one("hextext", 5, 30000), 6, flush_every=4000)]
    bad = []
    for g in good:
        for cut in (1, 2, 3, 7, len(g) // 3, len(g) // 2, len(g) - 5, len(g) - 1):
            bad.append(g[:cut])
    rnd = random.Random(7)
    for g in good:
        for _ in range(24):
            b = bytearray(g)
            for _ in range(rnd.randint(1, 3)):
                b[rnd.randrange(min(len(b), 200))] = rnd.randrange(256)
            bad.append(bytes(b))
    bad += [b"", b"\x07", b"\x06\x00", bytes([0x01, 0x05, 0x00, 0xfa, 0xfe]), bytes([0x01, 0x05, 0x00, 0xfa, 0xff, 1, 2, 3])]
    return bad


def test_inflate_emu_errors(emu_ctx, oracle, gen):
    bad = _bad_streams(gen)
    _check(emu_ctx, oracle, bad, [40000] * len(bad), check_used=False)
    # caller buffer too small
    g = raw_deflate(TXT)
    got, _ = run_batch(emu_ctx, "flate_decode_streams", [g, g], [3049, 100])
    assert got[0][0] == got[1][0] == oracle.E_OUTPUT_FULL


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_inflate_gpu(gpu_ctx, oracle, gen, device):
    groups = _streams(gen, 1 << 20)
    for name in GROUPS:
        streams = groups[name]
        caps = [len(zlib.decompress(s, -15)) + 5 for s in streams]
        _check(gpu_ctx, oracle, streams, caps, device=device)
    bad = _bad_streams(gen)
    _check(gpu_ctx, oracle, bad, [40000] * len(bad), device=device, check_used=False)


@pytest.mark.gpu
def test_inflate_gpu_64k_streams(gpu_ctx, oracle, gen):
    """BASELINE config 4 shape at reduced count: 2048 independent 64 KiB hexdump-text streams (zlib level 6, raw DEFLATE),
    against the oracle and the generator's bytes."""
    import torch
    unit, count = 65536, 2048
    raw = gen.units("hextext", gen.unit_seed(4, 0), unit, count)
    comp = [raw_deflate(raw[i * unit: (i + 1) * unit].tobytes()) for i in range(count)]
    lens = np.array([len(c) for c in comp], dtype=np.uint64)
    off = np.zeros(count, dtype=np.uint64)
    off[1:] = np.cumsum(lens)[:-1]
    packed = np.frombuffer(b"".join(comp) + bytes(64), dtype=np.uint8).copy()
    d_out = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    out_off = np.arange(count, dtype=np.uint64) * unit
    caps = np.full(count, unit, dtype=np.uint64)
    out_len, status, in_used, detail = gpu_ctx.flate_decode_streams(torch.from_numpy(packed).cuda(), off, lens, d_out, out_off, caps)
    assert (status == 0).all() and (out_len == unit).all() and (in_used == lens).all()
    assert bytes(d_out.cpu().numpy()) == raw.tobytes()
    ref = np.zeros(unit * count + 64, dtype=np.uint8)
    rl, rs = oracle.flate_decode_streams_mt(packed, off, lens, ref, out_off, caps, 8)
    assert (rs == 0).all() and bytes(ref[: unit * count]) == raw.tobytes()
