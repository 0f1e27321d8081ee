"""LZ4 block ENCODE kernel (csrc/lz4_encode.cu) against the oracle restatement of `BlockEncoder::encode`
(lz4.rs:226-310): the compressed bytes must be identical (same probe sequence, skip acceleration, rewind, tail rule), and they
must decode back with liblz4 and with our own decoder.  The oracle's encoder is pinned by SURVEY Appendix C
(`encode_block(test.txt)` -> 2,724 bytes) and by liblz4 decoding its output."""
import hashlib

import numpy as np
import pytest

from conftest import golden
from util import run_batch

TXT = golden("ref_test.txt")


def _inputs(gen, big):
    return [b"", b"a", b"abcdefghijk", b"abcdefghijkl", b"abcabcabcabcabcabcabc", TXT, bytes(5000), bytes(70000),
            gen.one("lzsyn", 31, big), gen.one("hextext", 32, big), gen.one("random", 33, big // 2), gen.one("runs", 34, big),
            (gen.one("random", 35, 300) + bytes(200)) * 40, gen.one("lzsyn", 36, 65536 + 777), bytes(range(256)) * 300,
            gen.one("random", 37, 20000) + gen.one("random", 37, 20000)]      # a far repeat after a long incompressible stretch (rewind path)


def _check(ctx, oracle, gen, big, device=False):
    units = _inputs(gen, big)
    caps = [oracle.lz4_compression_bound(len(u)) for u in units]
    got, _ = run_batch(ctx, "lz4_encode_blocks", units, caps, device=device, pad_front=5, gap=3)
    for i, u in enumerate(units):
        ref = oracle.lz4_encode_block(u)
        assert got[i][0] == 0, (i, got[i][0])
        assert got[i][1] == ref, "block %d: compressed bytes differ from BlockEncoder::encode (%d vs %d bytes)" % (i, len(got[i][1]), len(ref))
    return units, [g[1] for g in got]


def test_oracle_encoder_pinned(oracle, gen):
    enc = oracle.lz4_encode_block(TXT)
    assert len(enc) == 2724 and hashlib.sha256(enc).hexdigest().startswith("92921c4321ae45b3")
    n, back = gen.lz4_decompress(enc, len(TXT))
    assert n == len(TXT) and back == TXT


def test_lz4_encode_emu(emu_ctx, oracle, gen):
    units, encs = _check(emu_ctx, oracle, gen, 60000)
    for u, e in zip(units, encs):
        if len(u):
            n, back = gen.lz4_decompress(e, len(u))
            assert n == len(u) and back == u
    # our own decoder reads what our encoder wrote
    got, _ = run_batch(emu_ctx, "lz4_decode_blocks", encs, [len(u) for u in units])
    assert all(g == (0, u) for g, u in zip(got, units))


def test_lz4_encode_emu_small_cap(emu_ctx, oracle):
    got, _ = run_batch(emu_ctx, "lz4_encode_blocks", [TXT, TXT], [oracle.lz4_compression_bound(len(TXT)), 100])
    assert got[0][0] == 0 and got[1] == (oracle.E_OUTPUT_FULL, b"")


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_lz4_encode_gpu(gpu_ctx, oracle, gen, device):
    units, encs = _check(gpu_ctx, oracle, gen, 1 << 20, device=device)
    got, _ = run_batch(gpu_ctx, "lz4_decode_blocks", encs, [len(u) for u in units], device=device)
    assert all(g == (0, u) for g, u in zip(got, units))


@pytest.mark.gpu
def test_lz4_encode_gpu_4mib_blocks(gpu_ctx, oracle, gen):
    """BASELINE configs[1] made self-contained: 32 lzsyn blocks of 4 MiB compressed by the kernel == the oracle's bytes, and the decode
    kernel returns the input."""
    import torch
    unit, nb = 4 << 20, 32
    raw = gen.units("lzsyn", gen.unit_seed(2, 0), unit, nb)
    bound = oracle.lz4_compression_bound(unit)
    off = np.arange(nb, dtype=np.uint64) * unit
    coff = np.arange(nb, dtype=np.uint64) * ((bound + 15) // 16 * 16)
    d_raw = torch.from_numpy(raw).cuda()
    d_enc = torch.zeros(int(coff[-1]) + bound + 64, dtype=torch.uint8, device="cuda")
    clen, st = gpu_ctx.lz4_encode_blocks(d_raw, off, np.full(nb, unit, np.uint64), d_enc, coff, np.full(nb, bound, np.uint64))
    assert (st == 0).all()
    enc = d_enc.cpu().numpy()
    for i in range(0, nb, 4):
        ref = oracle.lz4_encode_block(raw[i * unit: (i + 1) * unit].tobytes())
        assert enc[int(coff[i]): int(coff[i]) + int(clen[i])].tobytes() == ref, "block %d" % i
    d_back = torch.zeros(unit * nb, dtype=torch.uint8, device="cuda")
    olen, st = gpu_ctx.lz4_decode_blocks(d_enc, coff, clen, d_back, off, np.full(nb, unit, np.uint64))
    assert (st == 0).all() and torch.equal(d_back, d_raw)
