"""zlib wrapper + Adler-32 (csrc/flate_decode.cu: zlib_finish_kernel, adler32_kernel) against the oracle
restatement of zlib.rs / checksum/adler.rs, the reference's own fixtures (zlib.rs:152-165) and Python's zlib."""
import random
import zlib

import numpy as np
import pytest

from conftest import golden
from util import pack, out_layout

TXT = golden("ref_test.txt")


def _streams(gen):
    rl = random.Random(3)
    units, caps, names = [], [], []
    for i in range(10):                                        # the reference fixtures: zlib.rs:152-165
        units.append(golden("ref_test.z.%d" % i)); caps.append(4096); names.append("fixture.%d" % i)
    for level in (0, 1, 6, 9):
        for kind, n in (("hextext", 70000), ("random", 3000), ("lzsyn", 40000), ("runs", 20000)):
            d = gen.one(kind, 50 + level, n)
            units.append(zlib.compress(d, level)); caps.append(n); names.append("%s.l%d" % (kind, level))
    units.append(zlib.compress(b"")); caps.append(16); names.append("empty")
    # streams that end on / contain a block of zero bytes: the only place where the reference reads the trailer (zlib.rs:106-109)
    co = zlib.compressobj(6)
    sync = co.compress(gen.one("hextext", 77, 3000)) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(b"tail") + co.flush()
    units.append(sync); caps.append(4000); names.append("sync_flush_mid_stream")
    co = zlib.compressobj(6)
    body = co.compress(gen.one("hextext", 78, 3000)) + co.flush(zlib.Z_SYNC_FLUSH)
    d = gen.one("hextext", 78, 3000)
    units.append(body + zlib.adler32(d).to_bytes(4, "big")); caps.append(4000); names.append("sync_then_valid_trailer")
    good = zlib.compress(gen.one("hextext", 9, 5000), 6)
    bad = [("short0", b""), ("short1", good[:1]), ("bad_method", bytes([0x77, good[1]]) + good[2:]),
           ("bad_window", bytes([0x68, 0x81]) + good[2:]),        # CINFO 6, FCHECK valid: only the window test fails
           ("preset_dict", bytes([0x78, 0xBB]) + good[2:]),      # FDICT set, FCHECK valid
           ("bad_header_check", bytes([0x78, 0x9D]) + good[2:]),
           ("bad_checksum", good[:-1] + bytes([good[-1] ^ 1])), ("no_trailer", good[:-4]), ("half_trailer", good[:-2]),
           ("truncated_deflate", good[: len(good) // 2]), ("output_full", good)]
    for k in range(6):
        bb = bytearray(good)
        bb[rl.randrange(2, len(bb) - 4)] ^= 1 << rl.randrange(8)
        bad.append(("fuzz%d" % k, bytes(bb)))
    for nm, u in bad:
        units.append(u); caps.append(3000 if nm == "output_full" else 6000); names.append(nm)
    return units, caps, names


def _check(ctx, oracle, gen, device=False):
    units, caps, names = _streams(gen)
    inb, in_off, in_len = pack(units, pad_front=3, gap=5)
    out_off, out_cap, total = out_layout(caps, gap=3)
    outb = np.zeros(total, dtype=np.uint8)
    if device:
        import torch
        d_in, d_out = torch.from_numpy(inb).cuda(), torch.zeros(total, dtype=torch.uint8, device="cuda")
        out_len, status, in_used, detail, adler = ctx.zlib_decode_streams(d_in, in_off, in_len, d_out, out_off, out_cap)
        outb = d_out.cpu().numpy()
    else:
        out_len, status, in_used, detail, adler = ctx.zlib_decode_streams(inb, in_off, in_len, outb, out_off, out_cap)
    seen = set()
    for i, (u, cap, nm) in enumerate(zip(units, caps, names)):
        st, ref, used, det, ad = oracle.zlib_decode(u, cap)
        assert int(status[i]) == st, (nm, int(status[i]), st)
        seen.add(st)
        if st == 0:
            got = outb[int(out_off[i]): int(out_off[i]) + int(out_len[i])].tobytes()
            assert got == ref, nm
            assert int(adler[i]) == ad == zlib.adler32(ref), nm
            assert int(in_used[i]) == used, nm
            # the reference reads the 4 trailer bytes only after a block of zero bytes (zlib.rs:104-109)
            assert used in (len(u), len(u) - 4, len(u) - 2) or nm.startswith(("fuzz", "sync")), (nm, used, len(u))
        elif st == -1:
            assert int(detail[i]) == det, (nm, int(detail[i]), det)
    assert {0, -1, -2, -5} <= seen                                # ok, InvalidInput, UnexpectedEof, output full all occur
    for i in range(10):
        assert int(status[i]) == 0 and outb[int(out_off[i]): int(out_off[i]) + int(out_len[i])].tobytes() == TXT


def test_oracle_zlib_on_reference_fixtures(oracle):
    """Pins the restatement: zlib.rs:152-165 (test.z.0-9 -> test.txt); the fixtures' trailer == Adler-32 of the text, and the
    reference never reads it for them (their final block is not empty: zlib.rs:104-105 answers first)."""
    for i in range(10):
        z = golden("ref_test.z.%d" % i)
        st, out, used, det, ad = oracle.zlib_decode(z, 4096)
        assert (st, out, used) == (0, TXT, len(z) - 4) and ad == zlib.adler32(TXT) == int.from_bytes(z[-4:], "big")


def test_oracle_zlib_trailer_rule(oracle, gen):
    """zlib.rs:99-124: trailer compared only after a block of zero bytes."""
    d = gen.one("hextext", 9, 5000)
    good = zlib.compress(d, 6)
    for mangled in (good[:-1] + bytes([good[-1] ^ 1]), good[:-4], good[:-2]):      # wrong / missing trailer after a non-empty final block
        st, out, used, det, ad = oracle.zlib_decode(mangled, 6000)
        assert (st, out, used, ad) == (0, d, len(good) - 4, zlib.adler32(d))
    empty = zlib.compress(b"")                                                      # one empty final block: trailer is read and checked
    assert oracle.zlib_decode(empty, 16)[:3] == (0, b"", len(empty))
    assert oracle.zlib_decode(empty[:-1] + b"\x02", 16)[0] == -1
    assert oracle.zlib_decode(empty[:-2], 16)[0] == -2
    co = zlib.compressobj(6)
    sync = co.compress(d) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(b"tail") + co.flush()
    st, out, used, det, ad = oracle.zlib_decode(sync, 6000)                         # the 4 bytes after the empty stored block are not the checksum
    assert st == -1 and det == 20 and out == d


def test_zlib_emu(emu_ctx, oracle, gen):
    _check(emu_ctx, oracle, gen)


def test_adler32_emu(emu_ctx, oracle, gen):
    units = [b"", b"a", b"abracadabra", TXT, gen.one("random", 1, 4096), gen.one("random", 2, 4097), gen.one("hextext", 3, 300000), bytes([255]) * 70000]
    inb, off, ln = pack(units, pad_front=1, gap=2)
    got = emu_ctx.adler32_streams(inb, off, ln)
    for i, u in enumerate(units):
        assert int(got[i]) == zlib.adler32(u) == oracle.adler32(u)


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_zlib_gpu(gpu_ctx, oracle, gen, device):
    _check(gpu_ctx, oracle, gen, device=device)


@pytest.mark.gpu
def test_adler32_gpu(gpu_ctx, gen):
    import torch
    units = [b"", b"abracadabra", TXT, gen.one("random", 2, 4097), gen.one("hextext", 3, 3000000), bytes([255]) * 5000000]
    inb, off, ln = pack(units, pad_front=1, gap=2)
    got = gpu_ctx.adler32_streams(torch.from_numpy(inb).cuda(), off, ln)
    for i, u in enumerate(units):
        assert int(got[i]) == zlib.adler32(u)
