"""LZ4 block-decode kernel (csrc/lz4_decode.cu) against the oracle: bit-exact output, identical status.

The cases run twice: on the CPU SIMT emulation build (no GPU needed; catches logic errors here) and, marked
`gpu`, on the real sm_100a library through the C ABI with device-resident and host-resident buffers."""
import random

import numpy as np
import pytest

from conftest import golden
from util import out_layout, pack, run_batch

TXT = golden("ref_test.txt")


PERIODS = (1, 2, 3, 4, 5, 7, 13, 31, 32, 33, 255, 256, 300, 32767, 32768, 32769, 50000, 65534, 65535)


def _lz4_raw(seqs, tail=b""):
    """Hand-assembled LZ4 block: [(literals, offset, match_len >= 4)], then a final literal run."""
    out = bytearray()
    for lit, off, ml in seqs + [(tail, 0, 0)]:
        l, m = len(lit), (ml - 4 if ml else 0)
        out.append((min(l, 15) << 4) | (min(m, 15) if ml else 0))
        if l >= 15:
            r = l - 15
            while r >= 255:
                out.append(255); r -= 255
            out.append(r)
        out += lit
        if ml:
            out += bytes([off & 255, off >> 8])
            if m >= 15:
                r = m - 15
                while r >= 255:
                    out.append(255); r -= 255
                out.append(r)
    return bytes(out)


def _cases(oracle, gen):
    c = gen.lz4_compress(TXT)
    cases = {
        "empty": ([b""], [16]),
        "tiny": ([bytes([0x10, 0x41])], [16]),
        "txt_ref_encoder": ([oracle.lz4_encode_block(TXT)], [4096]),
        "txt_liblz4": ([c], [3050]),
        "txt_liblz4_hc": ([gen.lz4_compress(TXT, hc=True)], [3050]),
        "zeros_long_match": ([gen.lz4_compress(bytes(300000))], [300000]),
        "zeros_ref_encoder": ([oracle.lz4_encode_block(bytes(300000))], [300000]),
        "incompressible_long_literals": ([gen.lz4_compress(gen.one("random", 7, 100000))], [100000]),
        "lzsyn_512k": ([gen.lz4_compress(gen.one("lzsyn", gen.unit_seed(2, 0), 1 << 19))], [1 << 19]),
        "lzsyn_ref_encoder": ([oracle.lz4_encode_block(gen.one("lzsyn", gen.unit_seed(2, 1), 1 << 18))], [1 << 18]),
        "hextext": ([gen.lz4_compress(gen.one("hextext", 3, 200000))], [200000]),
        "runs": ([gen.lz4_compress(gen.one("runs", 4, 150000))], [150000]),
        "batch_ragged": ([gen.lz4_compress(gen.one("lzsyn", 100 + i, 20000 + i * 777)) for i in range(7)] + [b"", c],
                         [20000 + i * 777 for i in range(7)] + [0, 3050]),
        "output_full": ([c], [3000]),
        "truncated_half": ([c[: len(c) // 2]], [4096]),
        "truncated_1": ([c[:-1]], [4096]),
        "offset_zero": ([bytes([0x10, 0x41, 0, 0, 0x00])], [64]),
        "offset_before_start": ([bytes([0x10, 0x41, 5, 0, 0x00])], [64]),
        "literal_overrun": ([bytes([0xF0, 0xFF, 0xFF])], [100000]),
        "ends_after_match": ([bytes([0x14, 0x41, 1, 0])], [64]),
    }
    for per in PERIODS:     # offsets 1..3 exercise the DECR path (lz4.rs:100-102); large ones the 64 KiB ring reach
        d = (bytes((i * 37 + 11) & 255 for i in range(per)) * (140000 // per + 2))[:140001]
        cases["period_%d" % per] = ([gen.lz4_compress(d)], [len(d)])
    # matches whose seed is (partly) the sequence's own literals, offsets <, == and > the literal run, overlapping and not
    rl = random.Random(5)
    seqs = []
    for _ in range(600):
        l = rl.choice([1, 2, 3, 5, 17, 40])
        off = rl.choice([1, 2, l, max(1, l - 1), l + 1, l + 7]) if seqs else rl.choice([1, l])
        seqs.append((bytes(rl.randrange(256) for _ in range(l)), off, rl.choice([4, 5, 19, 33, 70, 300])))
    blk = _lz4_raw(seqs, b"tail!")
    cases["own_literal_seeds"] = ([blk], [200000])
    blk = _lz4_raw([(b"ab", 2, 4)] + [(b"", rl.choice([1, 2, 3, 4, 6]), 4) for _ in range(5000)], b"z")
    cases["many_tiny_sequences"] = ([blk], [40000])
    rnd = random.Random(1)
    for k in range(8):
        bb = bytearray(c)
        for _ in range(3):
            bb[rnd.randrange(len(bb))] = rnd.randrange(256)
        cases["fuzz_%d" % k] = ([bytes(bb)], [8192])
    return cases


CASE_NAMES = ["empty", "tiny", "txt_ref_encoder", "txt_liblz4", "txt_liblz4_hc", "zeros_long_match", "zeros_ref_encoder",
              "incompressible_long_literals", "lzsyn_512k", "lzsyn_ref_encoder", "hextext", "runs", "batch_ragged", "output_full",
              "truncated_half", "truncated_1", "offset_zero", "offset_before_start", "literal_overrun", "ends_after_match"] + \
             ["period_%d" % p for p in PERIODS] + ["fuzz_%d" % k for k in range(8)] + ["own_literal_seeds", "many_tiny_sequences"]


def _check(ctx, oracle, units, caps, **kw):
    got, _ = run_batch(ctx, "lz4_decode_blocks", units, caps, **kw)
    for i, (u, cap) in enumerate(zip(units, caps)):
        st, ref = oracle.lz4_decode_block(u, cap)
        assert got[i][0] == st, (i, got[i][0], st)
        if st == 0:
            assert got[i][1] == ref, "block %d differs" % i
        else:
            assert got[i][1] == b"", "block %d failed: out_len must be 0 (rcz.h)" % i


@pytest.mark.parametrize("name", CASE_NAMES)
def test_lz4_emu(emu_ctx, oracle, gen, name):
    units, caps = _cases(oracle, gen)[name]
    _check(emu_ctx, oracle, units, caps, pad_front=5, gap=3)


def test_lz4_emu_frame_fixtures(emu_ctx, oracle):
    """The nine reference frames (lz4.rs:647-659): one compressed block each, FLG 0x64 / BD 0x70."""
    for i in range(1, 10):
        f = golden("ref_test.lz4.%d" % i)
        size = int.from_bytes(f[7:11], "little")
        got, _ = run_batch(emu_ctx, "lz4_decode_blocks", [f[11: 11 + size]], [4096])
        assert got[0] == (0, TXT)


def test_lz4_emu_pipelined_host_path(emu_ctx, oracle, gen, monkeypatch):
    """Host buffers with >= 8 blocks take the chunked H2D / decode / D2H pipeline; small chunks force many of them."""
    monkeypatch.setenv("RCZ_LZ4_CHUNK_BYTES", "50000")
    units = [gen.lz4_compress(gen.one("lzsyn", 200 + i, 9000 + 1311 * i)) for i in range(23)] + [b"", bytes([0x10, 0x41, 0, 0, 0])]
    caps = [9000 + 1311 * i for i in range(23)] + [0, 64]
    _check(emu_ctx, oracle, units, caps, pad_front=5, gap=3)


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_lz4_gpu_cases(gpu_ctx, oracle, gen, device):
    cases = _cases(oracle, gen)
    for name in CASE_NAMES:
        units, caps = cases[name]
        _check(gpu_ctx, oracle, units, caps, device=device, pad_front=5, gap=3)
    for i in range(1, 10):
        f = golden("ref_test.lz4.%d" % i)
        size = int.from_bytes(f[7:11], "little")
        got, _ = run_batch(gpu_ctx, "lz4_decode_blocks", [f[11: 11 + size]], [4096], device=device)
        assert got[0] == (0, TXT)


@pytest.mark.gpu
def test_lz4_gpu_4mib_blocks(gpu_ctx, oracle, gen):
    """BASELINE config 2 shape at reduced count: 16 x 4 MiB lzsyn blocks, liblz4-compressed, vs the oracle."""
    import torch
    unit, count = 4 << 20, 16
    raw = gen.units("lzsyn", gen.unit_seed(2, 0), unit, count)
    packed, off, lens = gen.lz4_compress_units(raw, unit, count)
    d_in = torch.from_numpy(packed).cuda()
    d_out = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    out_off = np.arange(count, dtype=np.uint64) * unit
    out_len, status = gpu_ctx.lz4_decode_blocks(d_in, off, lens, d_out, out_off, np.full(count, unit, dtype=np.uint64))
    assert (status == 0).all() and (out_len == unit).all()
    assert bytes(d_out.cpu().numpy()) == raw.tobytes()
    ref_len, ref_st = oracle.lz4_decode_blocks_mt(packed, off, lens, np.zeros(unit * count, np.uint8), out_off, np.full(count, unit, np.uint64), 4)
    assert (ref_st == 0).all() and (ref_len == out_len).all()


@pytest.mark.gpu
def test_lz4_gpu_pipelined_host_path(gpu_ctx, oracle, gen):
    """64 x 4 MiB blocks from pinned host buffers: 4 pipeline chunks of 64 MiB; bytes equal the generator's."""
    import torch
    unit, count = 4 << 20, 64
    raw = gen.units("lzsyn", gen.unit_seed(2, 0), unit, count)
    packed, off, lens = gen.lz4_compress_units(raw, unit, count)
    h_in = torch.from_numpy(packed).pin_memory()
    h_out = torch.zeros(unit * count + 64, dtype=torch.uint8).pin_memory()
    out_off = np.arange(count, dtype=np.uint64) * unit
    out_len, status = gpu_ctx.lz4_decode_blocks(h_in.numpy(), off, lens, h_out.numpy(), out_off, np.full(count, unit, dtype=np.uint64))
    assert (status == 0).all() and (out_len == unit).all()
    assert bytes(h_out.numpy()[: unit * count]) == raw.tobytes()


@pytest.mark.gpu
def test_lz4_gpu_full_config_roundtrip(gpu_ctx, gen):
    """BASELINE config 2 at full size (256 x 4 MiB) via a size-independent property: decode(compress(x)) == x,
    checked on the device by comparing against the uploaded original."""
    import torch
    unit, count = 4 << 20, 256
    raw = gen.units("lzsyn", gen.unit_seed(2, 0), unit, count)
    packed, off, lens = gen.lz4_compress_units(raw, unit, count)
    d_in = torch.from_numpy(packed).cuda()
    d_raw = torch.from_numpy(raw).cuda()
    d_out = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    out_off = np.arange(count, dtype=np.uint64) * unit
    out_len, status = gpu_ctx.lz4_decode_blocks(d_in, off, lens, d_out, out_off, np.full(count, unit, dtype=np.uint64))
    assert (status == 0).all() and (out_len == unit).all()
    assert torch.equal(d_out, d_raw)


# ---------------------------------------------------------------------------------------------------------------------
# Cases aimed at the window-parallel parse (8 KiB windows of compressed bytes, 320-byte halo, look-back chain) and at the
# ring-buffer materialise kernel (8 KiB tiles, 64 KiB of history): fields that straddle window ends, 0xFF bytes that the
# speculative parse may take for length extensions, long matches / literal runs that span windows and tiles.
# ---------------------------------------------------------------------------------------------------------------------
def _window_cases(oracle, gen):
    rl = random.Random(11)
    cases = {}
    # literal runs of many lengths (also > halo, > window) between compressible stretches: tokens whose fields cross window ends
    parts = []
    for i in range(60):
        parts.append(gen.one("lzsyn", 300 + i, rl.choice([700, 3000, 9000])))
        parts.append(gen.one("random", 400 + i, rl.choice([1, 14, 15, 16, 270, 300, 330, 500, 8000, 8200, 20000])))
    d = b"".join(parts)
    cases["literal_runs_across_windows"] = ([gen.lz4_compress(d)], [len(d)])
    # a long match whose length extension bytes straddle the end of window 0 at every alignment near it
    units, caps = [], []
    for pre in list(range(8150, 8200, 3)) + [16360, 16383, 16384, 16385]:
        d = gen.one("random", 900 + pre, pre) + bytes(700000 + pre)
        units.append(gen.lz4_compress(d)); caps.append(len(d))
    cases["match_extension_across_windows"] = (units, caps)
    # hand-assembled: literals full of 0xFF (every byte looks like a token with both lengths extended), long and short, with matches
    # of all sizes in between; several windows long
    seqs = []
    for i in range(900):
        lit = bytes([0xFF]) * rl.choice([0, 1, 2, 15, 16, 33, 254, 255, 256, 300, 700]) + bytes(rl.randrange(256) for _ in range(rl.choice([0, 1, 3])))
        if not seqs and not lit:
            lit = b"\xff"
        have = sum(len(l) + m for l, _, m in seqs) + len(lit)
        off = rl.choice([1, 2, 3, 15, 16, 17, 255, 4097]) if have > 4097 else 1
        seqs.append((lit, off, rl.choice([4, 5, 18, 19, 20, 273, 274, 275, 529, 2000])))
    blk = _lz4_raw(seqs, b"\xff" * 40)
    cases["ff_literals"] = ([blk], [sum(len(l) + m for l, _, m in seqs) + 40])
    # many small blocks (also empty ones) in one batch
    units = [gen.lz4_compress(gen.one("lzsyn", 700 + i, (i * 37) % 2100)) if (i % 9) else b"" for i in range(300)]
    caps = [(i * 37) % 2100 if (i % 9) else 0 for i in range(300)]
    cases["many_small_blocks"] = (units, caps)
    # hexdump text: 9-byte sequences, every match reaches into the current tile (long in-tile dependency chains)
    d = gen.one("hextext", 5, 700000)
    cases["hextext_700k"] = ([gen.lz4_compress(d)], [len(d)])
    # matches of every length class at offsets around the 64 KiB reach of the ring and around the tile size
    seqs = [(bytes(rl.randrange(256) for _ in range(70000)), 1, 4)]
    for i in range(3000):
        seqs.append((bytes(rl.randrange(256) for _ in range(rl.choice([0, 0, 1, 5, 40]))),
                     rl.choice([65535, 65534, 65520, 65000, 57344, 49152, 40000, 32769, 32768, 32767, 16384, 8193, 8192, 8191, 4096, 100, 17, 16, 15, 8, 4, 1]),
                     rl.choice([4, 7, 16, 17, 31, 32, 33, 64, 100, 1000, 9000])))
    blk = _lz4_raw(seqs, b"end")
    cases["ring_reach"] = ([blk], [sum(len(l) + m for l, _, m in seqs) + 3])
    # corruptions of a block that spans many windows: status (and bytes when it still decodes) must match the oracle
    c = gen.lz4_compress(gen.one("lzsyn", 77, 300000))
    for k in range(12):
        bb = bytearray(c)
        for _ in range(2):
            bb[rl.randrange(len(bb))] = rl.randrange(256)
        cases["fuzz_multiwindow_%d" % k] = ([bytes(bb), bytes(bb[: rl.randrange(len(bb))])], [300000, 300000])
    cases["output_full_multiwindow"] = ([c, c, c], [299999, 150000, 8000])
    return cases


WINDOW_CASE_NAMES = ["literal_runs_across_windows", "match_extension_across_windows", "ff_literals", "many_small_blocks", "hextext_700k",
                     "ring_reach", "output_full_multiwindow"] + ["fuzz_multiwindow_%d" % k for k in range(12)]


@pytest.mark.parametrize("name", WINDOW_CASE_NAMES)
def test_lz4_emu_window_cases(emu_ctx, oracle, gen, name):
    units, caps = _window_cases(oracle, gen)[name]
    _check(emu_ctx, oracle, units, caps, pad_front=7, gap=1)


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_lz4_gpu_window_cases(gpu_ctx, oracle, gen, device):
    cases = _window_cases(oracle, gen)
    for name in WINDOW_CASE_NAMES:
        units, caps = cases[name]
        _check(gpu_ctx, oracle, units, caps, device=device, pad_front=7, gap=1)


@pytest.mark.gpu
def test_lz4_gpu_stage_timings(gpu_ctx, gen):
    """rcz_last_stage_ms: the three kernels of a device-resident call (parse, scan, materialise) are timed separately."""
    import torch
    unit, count = 1 << 20, 32
    raw = gen.units("lzsyn", gen.unit_seed(2, 0), unit, count)
    packed, off, lens = gen.lz4_compress_units(raw, unit, count)
    d_in = torch.from_numpy(packed).cuda()
    d_out = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    out_off = np.arange(count, dtype=np.uint64) * unit
    out_len, status = gpu_ctx.lz4_decode_blocks(d_in, off, lens, d_out, out_off, np.full(count, unit, dtype=np.uint64))
    assert (status == 0).all() and bytes(d_out.cpu().numpy()) == raw.tobytes()
    st = gpu_ctx.last_stage_ms()
    assert len(st) == 3 and all(t > 0 for t in st)
    assert abs(sum(st) - gpu_ctx.last_kernel_ms()) < 0.25 * gpu_ctx.last_kernel_ms() + 0.05


def test_lz4_emu_gather_peers(emu_ctx, oracle, gen):
    """rcz_lz4_decode_blocks_gather: every output byte also lands at the same offset of the peers' buffers (the multi-GPU fused
    gather; on the emulator the 'peers' are two more host buffers)."""
    raw = [gen.one("lzsyn", 5, 50000), gen.one("hextext", 6, 33333), b"x" * 70000]
    units = [gen.lz4_compress(r) for r in raw]
    inb, in_off, in_len = pack(units, pad_front=5, gap=3, align=16)
    caps = [len(r) for r in raw]
    out_off, out_cap, total = out_layout(caps, gap=7)
    mine, p1, p2 = (np.zeros(total, dtype=np.uint8) for _ in range(3))
    out_len, status = emu_ctx.lz4_decode_blocks_gather(inb, in_off, in_len, mine, out_off, out_cap, [p1.ctypes.data, p2.ctypes.data], async_="emu-device")
    assert (status == 0).all()
    for o, r in zip(out_off, raw):
        for buf in (mine, p1, p2):
            assert buf[int(o): int(o) + len(r)].tobytes() == r
    # nothing outside the blocks' regions is touched (the peers' whole-chunk bulk stores stop at the ragged edges)
    assert mine.tobytes() == p1.tobytes() == p2.tobytes()
    used = np.zeros(total, dtype=bool)
    for o, r in zip(out_off, raw):
        used[int(o): int(o) + len(r)] = True
    assert not mine[~used].any()


def test_lz4_emu_gather_many_blocks(emu_ctx, gen):
    """the gather call on more blocks than one CTA wave of the emulator, ragged sizes and unaligned offsets"""
    raw = [gen.one("lzsyn" if i % 3 else "hextext", 100 + i, 3000 + 517 * i) for i in range(37)]
    units = [gen.lz4_compress(r) for r in raw]
    inb, in_off, in_len = pack(units, pad_front=3, gap=1, align=1)
    out_off, out_cap, total = out_layout([len(r) for r in raw], gap=5)
    mine, p1 = (np.zeros(total, dtype=np.uint8) for _ in range(2))
    out_len, status = emu_ctx.lz4_decode_blocks_gather(inb, in_off, in_len, mine, out_off, out_cap, [p1.ctypes.data], async_="emu-device")
    assert (status == 0).all() and [int(x) for x in out_len] == [len(r) for r in raw]
    for o, r in zip(out_off, raw):
        assert mine[int(o): int(o) + len(r)].tobytes() == r
    assert mine.tobytes() == p1.tobytes()


def _host_pipeline_case(ctx, oracle, gen, sizes, pinned):
    """The HOST-buffer pipeline (>= 8 blocks, several chunks): bytes, lengths and statuses equal the oracle's, a block that does not
    fit included, and nothing outside the blocks' capacity regions is written to the caller's buffer."""
    raw = [gen.one("lzsyn" if i % 2 else "hextext", 300 + i, n) for i, n in enumerate(sizes)]
    units = [gen.lz4_compress(r) for r in raw]
    inb, in_off, in_len = pack(units, pad_front=7, gap=2, align=1)
    caps = [len(r) for r in raw]
    caps[5] -= 100                                               # block 5 does not fit: OUTPUT_FULL, nothing of it is materialised
    out_off, out_cap, total = out_layout(caps, gap=9)
    if pinned:
        import torch
        keep = (torch.from_numpy(inb).pin_memory(), torch.full((total + 3,), 0xAA, dtype=torch.uint8).pin_memory())
        inb_, out = keep[0].numpy(), keep[1].numpy()[3:]
    else:
        inb_, out = inb, np.full(total + 3, 0xAA, dtype=np.uint8)[3:]           # odd alignment of the host base
    out_len, status = ctx.lz4_decode_blocks(inb_, in_off, in_len, out, out_off, out_cap)
    ref = np.zeros(total, dtype=np.uint8)
    ref_len, ref_st = oracle.lz4_decode_blocks_mt(inb, in_off, in_len, ref, out_off, out_cap, 2)
    assert [int(x) for x in status] == [int(x) for x in ref_st] and status[5] != 0
    used = np.zeros(total, dtype=bool)
    for i, (o, r) in enumerate(zip(out_off, raw)):
        if status[i] == 0:
            assert int(out_len[i]) == len(r) and out[int(o): int(o) + len(r)].tobytes() == r
            used[int(o): int(o) + len(r)] = True
        else:
            assert int(out_len[i]) == 0
            used[int(o): int(o) + int(out_cap[i])] = True       # a failed block's region is unspecified (rcz.h)
    assert (out[~used] == 0xAA).all()


def test_lz4_emu_host_pipeline(emu_ctx, oracle, gen, monkeypatch):
    monkeypatch.setenv("RCZ_LZ4_CHUNK_BYTES", "60000")
    _host_pipeline_case(emu_ctx, oracle, gen, [9000 + 1301 * i for i in range(13)], False)


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [True, False])
def test_lz4_gpu_host_pipeline_ragged(gpu_ctx, oracle, gen, pinned, monkeypatch):
    """blocks of 0.2 - 3 MiB in 4 MiB chunks, page-locked and pageable host buffers"""
    monkeypatch.setenv("RCZ_LZ4_CHUNK_BYTES", str(4 << 20))
    _host_pipeline_case(gpu_ctx, oracle, gen, [200000 + 120001 * i for i in range(24)], pinned)


@pytest.mark.gpu
def test_lz4_gpu_gather_bulk_stores_one_gpu(gpu_ctx, gen):
    """The fused gather on ONE GPU: the 'peers' are two more buffers of the same device, so the TMA bulk stores from the output ring
    (and the byte stores of the ragged chunks) run exactly as they do over NVLink.  Ragged block sizes, unaligned offsets: all three
    buffers must hold every block's bytes and nothing else."""
    import torch
    raw = [gen.one("lzsyn" if i % 3 else "hextext", 700 + i, 150000 + 91003 * i) for i in range(40)]
    units = [gen.lz4_compress(r) for r in raw]
    inb, in_off, in_len = pack(units, pad_front=5, gap=3, align=1)
    out_off, out_cap, total = out_layout([len(r) for r in raw], gap=7)
    d_in = torch.from_numpy(inb).cuda()
    bufs = [torch.full((total,), 0xAA, dtype=torch.uint8, device="cuda") for _ in range(3)]
    out_len, status = gpu_ctx.lz4_decode_blocks_gather(d_in, in_off, in_len, bufs[0], out_off, out_cap, [bufs[1].data_ptr(), bufs[2].data_ptr()])
    torch.cuda.synchronize()
    status = status.cpu().numpy() if hasattr(status, "cpu") else status
    assert (status == 0).all()
    want = np.full(total, 0xAA, dtype=np.uint8)
    for o, r in zip(out_off, raw):
        want[int(o): int(o) + len(r)] = np.frombuffer(r, dtype=np.uint8)
    for b in bufs:
        assert bytes(b.cpu().numpy()) == want.tobytes()
