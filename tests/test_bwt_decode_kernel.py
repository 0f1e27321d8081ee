"""Inverse-BWT kernels (csrc/bwt_decode.cu) against the oracle (bwt/mod.rs:223-294): bit-exact output and length,
including blocks that are not valid BWTs (the reference's iterator then stops early) and malformed origins."""
import numpy as np
import pytest

from conftest import golden
from util import pack

TXT = golden("ref_test.txt")


def _run(ctx, blocks, device=False):
    inb, in_off, n = pack([l for l, _ in blocks], pad_front=3, gap=5)
    origin = np.array([o for _, o in blocks], dtype=np.uint32)
    if device:
        import torch
        d_in = torch.from_numpy(inb).cuda()
        d_out = torch.zeros(len(inb), dtype=torch.uint8, device="cuda")
        out_len, status = ctx.bwt_decode_blocks(d_in, in_off, n, origin, d_out, in_off)
        outb = d_out.cpu().numpy()
    else:
        outb = np.zeros(len(inb), dtype=np.uint8)
        out_len, status = ctx.bwt_decode_blocks(inb, in_off, n, origin, outb, in_off)
    return [(int(s), outb[int(o): int(o) + int(l)].tobytes()) for s, o, l in zip(status, in_off, out_len)]


def _check(ctx, oracle, blocks, **kw):
    got = _run(ctx, blocks, **kw)
    for i, (l, og) in enumerate(blocks):
        st, ref = oracle.bwt_decode(l, og)
        assert got[i][0] == st, (i, got[i][0], st)
        if st == 0:
            assert got[i][1] == ref, "block %d differs" % i


def _enc(oracle, d):
    st, l, og = oracle.bwt_encode(d)
    assert st == 0
    return (l, og)


def _cases(oracle, gen, big):
    rs = np.random.RandomState(5)
    invalid = []
    for _ in range(8):
        n = int(rs.randint(1, 5000))
        invalid.append((bytes(rs.randint(0, 4, size=n, dtype=np.uint8)), int(rs.randint(0, n))))
    cases = {
        "abracadabra": [_enc(oracle, b"abracadabra")],
        "single_byte": [_enc(oracle, b"a")],
        "reference_roundtrips": [_enc(oracle, b"test"), _enc(oracle, TXT)],       # bwt/mod.rs:541-547
        "ragged_batch": [_enc(oracle, b"banana"), _enc(oracle, b"test"), _enc(oracle, TXT[:1000]), _enc(oracle, gen.one("hextext", 1, 40000))],
        "random": [_enc(oracle, gen.one("random", 1, big))],
        "hextext": [_enc(oracle, gen.one("hextext", 2, big))],
        "zeros": [_enc(oracle, bytes(70000))],
        "period2": [_enc(oracle, b"ab" * 20000)],
        "malformed_origin_and_empty": [(b"abc", 3), (b"", 0), _enc(oracle, b"hello world")],
        "not_a_bwt": invalid,
        "origin_on_sampled_row": [(l, og) for l, og in [_enc(oracle, gen.one("random", 40 + k, 3000)) for k in range(12)]],
    }
    return cases


NAMES = ["abracadabra", "single_byte", "reference_roundtrips", "ragged_batch", "random", "hextext", "zeros", "period2",
         "malformed_origin_and_empty", "not_a_bwt", "origin_on_sampled_row"]


@pytest.mark.parametrize("name", NAMES)
def test_ibwt_emu(emu_ctx, oracle, gen, name):
    _check(emu_ctx, oracle, _cases(oracle, gen, 300000)[name])


def test_ibwt_emu_origin_exactly_on_sample(emu_ctx, oracle):
    """origin a multiple of the sampling stride (256): the sampled chain at that row is unreachable by construction."""
    rs = np.random.RandomState(11)
    blocks = []
    for n in (257, 1024, 5000):
        l = bytes(rs.randint(0, 3, size=n, dtype=np.uint8))
        for og in (0, 256):
            if og < n:
                blocks.append((l, og))
    _check(emu_ctx, oracle, blocks)


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_ibwt_gpu_cases(gpu_ctx, oracle, gen, device):
    cases = _cases(oracle, gen, 1 << 20)
    for name in NAMES:
        _check(gpu_ctx, oracle, cases[name], device=device)


@pytest.mark.gpu
def test_ibwt_gpu_4mib_blocks(gpu_ctx, oracle, gen):
    """BASELINE config 3 decode leg at reduced count: 4 x 4 MiB random blocks; L columns from the oracle."""
    import torch
    unit, count = 4 << 20, 4
    raw = gen.units("random", gen.unit_seed(3, 0), unit, count)
    off = np.arange(count, dtype=np.uint64) * unit
    n = np.full(count, unit, dtype=np.uint64)
    l_buf = np.zeros(unit * count + 64, dtype=np.uint8)
    origin, st = oracle.bwt_encode_blocks_mt(raw, off, n, l_buf, 4)
    assert (st == 0).all()
    d_out = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    out_len, status = gpu_ctx.bwt_decode_blocks(torch.from_numpy(l_buf).cuda(), off, n, origin, d_out, off)
    assert (status == 0).all() and (out_len == unit).all()
    assert bytes(d_out.cpu().numpy()) == raw.tobytes()
