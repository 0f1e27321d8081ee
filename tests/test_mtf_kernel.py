"""Move-to-front stream coder (csrc/mtf.cu) against the oracle restatement of bwt/mtf.rs: bit-exact ranks,
bit-exact inverse, the reference's own round trips (mtf.rs:188-192)."""
import numpy as np
import pytest

from conftest import golden
from util import run_batch

TXT = golden("ref_test.txt")


def _units(gen, oracle):
    units = [b"teeesst_mtf", b"", TXT, bytes(range(256)) * 3, bytes(reversed(range(256))), bytes([255]) * 1000 + bytes([0]) * 33,
             gen.one("random", 5, 70001), gen.one("hextext", 6, 50000), gen.one("runs", 7, 30011), b"a", b"ab" * 17]
    # what the coder sees in the bzip-style chain: the L column of a BWT block
    units.append(oracle.bwt_encode(gen.one("hextext", 8, 20000))[1])
    return units


def _check(ctx, oracle, gen, **kw):
    units = _units(gen, oracle)
    caps = [len(u) for u in units]
    enc, _ = run_batch(ctx, "mtf_encode_streams", units, caps, pad_front=3, gap=2, **kw)
    ranks = []
    for i, u in enumerate(units):
        ref = oracle.mtf_encode(u)
        assert enc[i] == (0, bytes(ref)), "ranks of unit %d differ" % i
        ranks.append(bytes(ref))
    dec, _ = run_batch(ctx, "mtf_decode_streams", ranks, caps, pad_front=1, gap=4, **kw)
    for i, u in enumerate(units):
        assert dec[i] == (0, u) and bytes(oracle.mtf_decode(ranks[i])) == u
    # output buffer too small: the first cap bytes, status OUTPUT_FULL
    short, _ = run_batch(ctx, "mtf_encode_streams", [TXT], [100], **kw)
    assert short[0] == (-5, bytes(oracle.mtf_encode(TXT))[:100])


def test_oracle_mtf_known_answers(oracle):
    """mtf.rs:63-91 by hand: alphabetical list, 'a' = 97."""
    assert bytes(oracle.mtf_encode(b"aaa")) == bytes([97, 0, 0])
    assert bytes(oracle.mtf_encode(b"abab")) == bytes([97, 98, 1, 1])
    assert bytes(oracle.mtf_encode(bytes([0, 1, 2, 2, 0]))) == bytes([0, 1, 2, 0, 2])
    for u in (b"teeesst_mtf", b"", TXT):                       # mtf.rs:188-192 some_roundtrips
        assert bytes(oracle.mtf_decode(oracle.mtf_encode(u))) == u


def test_mtf_emu(emu_ctx, oracle, gen):
    _check(emu_ctx, oracle, gen)


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_mtf_gpu(gpu_ctx, oracle, gen, device):
    _check(gpu_ctx, oracle, gen, device=device)


@pytest.mark.gpu
def test_mtf_gpu_4mib_blocks(gpu_ctx, oracle, gen):
    """C5 shape: 32 x 4 MiB blocks, round trip on the device + ranks of one block against the oracle."""
    import torch
    unit, count = 4 << 20, 32
    raw = gen.units("hextext", gen.unit_seed(5, 0), unit, count)
    d_raw = torch.from_numpy(raw).cuda()
    d_rk = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    off = np.arange(count, dtype=np.uint64) * unit
    n = np.full(count, unit, dtype=np.uint64)
    ol, st = gpu_ctx.mtf_encode_streams(d_raw, off, n, d_rk, off, n)
    assert (st == 0).all() and (ol == unit).all()
    ol, st = gpu_ctx.mtf_decode_streams(d_rk, off, n, d_back, off, n)
    assert (st == 0).all() and torch.equal(d_back, d_raw)
    assert bytes(d_rk[:unit].cpu().numpy()) == bytes(oracle.mtf_encode(raw[:unit].tobytes()))
