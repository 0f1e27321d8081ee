"""Range-coder (csrc/ari.cu) and RLE (csrc/rle.cu) kernels against the oracle: bit-exact bytes, lengths, statuses.

ari: entropy/ari/table.rs:185-273 ByteEncoder/ByteDecoder over ari/mod.rs:117-150 RangeEncoder::process.
rle: rle.rs:40-123 Encoder, rle.rs:176-281 Decoder (inline KATs rle.rs:320-352).
Every case runs on the CPU SIMT emulation build here and, marked `gpu`, on librcz.so through the C ABI."""
import numpy as np
import pytest

from conftest import golden
from test_oracle_golden import RLE_KATS
from util import run_batch

TXT = golden("ref_test.txt")


# ------------------------------------------------------------------------------------------------ ari
def _ari_inputs(gen, big):
    return [b"", b"a", b"abracadabra", TXT, bytes(5000), gen.one("random", 9, big), gen.one("hextext", 10, big),
            gen.one("runs", 11, big // 2), bytes([255]) * 3000, bytes(range(256)) * 40]


def _check_ari(ctx, oracle, gen, big, device=False):
    raws = _ari_inputs(gen, big)
    caps = [2 * len(r) + 64 for r in raws]
    enc, _ = run_batch(ctx, "ari_encode_streams", raws, caps, device=device, pad_front=3, gap=2)
    refs = [oracle.ari_encode(r) for r in raws]
    for i, (st, e) in enumerate(enc):
        assert st == 0 and e == refs[i], "ari encode stream %d differs" % i
    # decode: the oracle's streams plus garbage appended (the decoder must stop at the terminator)
    streams = [r + b"\x5a" * 7 for r in refs]
    dec, (in_used,) = run_batch(ctx, "ari_decode_streams", streams, [len(r) + 8 for r in raws], device=device, pad_front=1, gap=4)
    for i, (st, d) in enumerate(dec):
        ost, od, used_read, used_finish = oracle.ari_decode(streams[i], len(raws[i]) + 8)
        assert st == ost == 0 and d == od == raws[i], "ari decode stream %d differs" % i
        assert int(in_used[i]) == used_finish == len(refs[i])
    # error parity: truncated stream (the reference panics in feed().unwrap(), ari/mod.rs:282), output too small
    bad = [refs[3][: len(refs[3]) // 2], refs[2][:3], refs[3]]
    caps = [4096, 64, 100]
    got, _ = run_batch(ctx, "ari_decode_streams", bad, caps, device=device)
    for i, (st, d) in enumerate(got):
        ost, od, _, _ = oracle.ari_decode(bad[i], caps[i])
        assert st == ost, (i, st, ost)
        assert st != 0


def test_ari_emu(emu_ctx, oracle, gen):
    _check_ari(emu_ctx, oracle, gen, 30000)


def test_ari_emu_golden_vectors(emu_ctx):
    """SURVEY Appendix C vectors (restatement-derived; the reference pins ARI by roundtrip only)."""
    enc, _ = run_batch(emu_ctx, "ari_encode_streams", [b"abracadabra", b""], [64, 64])
    assert enc[0] == (0, bytes.fromhex("6101aba17aa9d5cc68d39733f600"))
    assert enc[1] == (0, bytes.fromhex("ff00ff0000"))


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_ari_gpu(gpu_ctx, oracle, gen, device):
    _check_ari(gpu_ctx, oracle, gen, 200000, device=device)


@pytest.mark.gpu
def test_ari_gpu_many_streams_roundtrip(gpu_ctx, gen):
    """2048 x 16 KiB streams: decode(encode(x)) == x on the device (size-independent property)."""
    import torch
    unit, count = 16384, 2048
    raw = gen.units("hextext", gen.unit_seed(5, 0), unit, count)
    d_raw = torch.from_numpy(raw).cuda()
    off = np.arange(count, dtype=np.uint64) * unit
    cap = 2 * unit + 64
    coff = np.arange(count, dtype=np.uint64) * cap
    d_enc = torch.zeros(cap * count, dtype=torch.uint8, device="cuda")
    elen, st = gpu_ctx.ari_encode_streams(d_raw, off, np.full(count, unit, np.uint64), d_enc, coff, np.full(count, cap, np.uint64))
    assert (st == 0).all()
    d_dec = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    dlen, st, used = gpu_ctx.ari_decode_streams(d_enc, coff, elen, d_dec, off, np.full(count, unit, np.uint64))
    assert (st == 0).all() and (dlen == unit).all() and (used == elen).all()
    assert torch.equal(d_dec, d_raw)


# ------------------------------------------------------------------------------------------------ rle
def _rle_cases(oracle, gen, big):
    raws = [r for r, _ in RLE_KATS] + [TXT, gen.one("runs", 3, big), gen.one("random", 4, 13579), bytes(70000), b"ab" * 500,
                                      bytes([9]) * 2 + bytes([8]) * 3 + bytes([7]) * 130 + bytes([6]) * 16386]
    return raws


def _check_rle(ctx, oracle, gen, big, device=False):
    raws = _rle_cases(oracle, gen, big)
    encs = [oracle.rle_encode(r) for r in raws]
    got, _ = run_batch(ctx, "rle_encode_streams", raws, [2 * len(r) + 16 for r in raws], device=device, pad_front=2, gap=1)
    for i, (st, e) in enumerate(got):
        assert st == 0 and e == encs[i], "rle encode %d differs" % i
    got, _ = run_batch(ctx, "rle_decode_streams", encs, [len(r) + 4 for r in raws], device=device, pad_front=7, gap=1)
    for i, (st, d) in enumerate(got):
        assert (st, d) == (0, raws[i]), "rle decode %d differs" % i
    # inline KAT encodings decode to the KAT inputs (rle.rs:336-352)
    got, _ = run_batch(ctx, "rle_decode_streams", [e for _, e in RLE_KATS], [len(r) + 4 for r, _ in RLE_KATS], device=device)
    assert [g for g in got] == [(0, r) for r, _ in RLE_KATS]
    # error / edge parity (rle.rs:151-154 overly long run; :247-256 partial flush; maximal 9-group run clipped by capacity)
    odd = [bytes([7, 7] + [1] * 10), bytes([7, 7, 3]), bytes([7, 7]), bytes([7, 7] + [1] * 9), bytes([1, 2, 2, 0x85, 3, 3]),
           bytes([4, 4, 0x80, 4, 4, 0x80, 4]), bytes([5, 5, 0x7f, 0x7f, 0x81])]
    caps = [64, 64, 64, 64, 64, 64, 100000]
    got, _ = run_batch(ctx, "rle_decode_streams", odd, caps, device=device)
    for i, (st, d) in enumerate(got):
        ost, od = oracle.rle_decode(odd[i], caps[i])
        assert st == ost, (i, st, ost)
        if st == 0:
            assert d == od, i


def test_rle_emu(emu_ctx, oracle, gen):
    _check_rle(emu_ctx, oracle, gen, 60000)


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_rle_gpu(gpu_ctx, oracle, gen, device):
    _check_rle(gpu_ctx, oracle, gen, 1 << 20, device=device)
