"""Suffix-sort + L-column kernels (csrc/bwt_encode.cu) against the oracle (bwt/mod.rs:136-219): L bytes and origin bit-exact.
The suffix array is unique, so this pins the GPU sort against the reference's comparison sort; adversarial inputs
(long repeats, where the reference itself is quadratic) are kept small for the CPU oracle."""
import numpy as np
import pytest

from conftest import golden
from util import pack

TXT = golden("ref_test.txt")


def _run(ctx, blocks, device=False):
    inb, in_off, n = pack(blocks, pad_front=3, gap=5)
    if device:
        import torch
        d_out = torch.zeros(len(inb), dtype=torch.uint8, device="cuda")
        origin, status = ctx.bwt_encode_blocks(torch.from_numpy(inb).cuda(), in_off, n, d_out, in_off)
        outb = d_out.cpu().numpy()
    else:
        outb = np.zeros(len(inb), dtype=np.uint8)
        origin, status = ctx.bwt_encode_blocks(inb, in_off, n, outb, in_off)
    return [(int(s), outb[int(o): int(o) + len(b)].tobytes(), int(og)) for s, o, og, b in zip(status, in_off, origin, blocks)]


def _check(ctx, oracle, blocks, **kw):
    got = _run(ctx, blocks, **kw)
    for i, b in enumerate(blocks):
        st, l, og = oracle.bwt_encode(b)
        assert got[i][0] == st, (i, got[i][0], st)
        if st == 0:
            assert got[i][2] == og, (i, got[i][2], og)
            assert got[i][1] == l, "block %d: L column differs" % i


def _cases(gen, big):
    rs = np.random.RandomState(2)
    return {
        "appendix_c": [b"abracadabra", b"banana", b"test"],                      # rdarcaaaabb/2, nnbaaa/3, test/3
        "reference_roundtrips": [b"test", TXT],                                    # bwt/mod.rs:541-547
        "tiny_and_empty": [b"a", b"", b"ab", b"ba", b"aa", b"aaa"],
        "random": [gen.one("random", 1, big)],
        "hextext": [gen.one("hextext", 2, big)],
        "lzsyn": [gen.one("lzsyn", 3, big)],
        "runs": [gen.one("runs", 4, big // 4)],
        "zeros": [bytes(6000)],
        "periodic": [b"ab" * 3000, b"abc" * 2000 + b"abd", bytes(range(256)) * 20],
        "binary_small_alphabet": [bytes(rs.randint(0, 2, size=n, dtype=np.uint8)) for n in (1, 2, 3, 7, 8, 9, 63, 64, 65, 1000, 20000)],
        "ragged_batch": [gen.one("random", 10 + i, 1000 + 977 * i) for i in range(9)] + [b"", TXT[:100]],
        "short_suffix_vs_zero_bytes": [bytes([1, 0, 0, 0, 0, 0, 0, 0, 0, 0]), bytes(9) + bytes([1]) + bytes(9), bytes([0, 1] * 9 + [0])],
    }


NAMES = ["appendix_c", "reference_roundtrips", "tiny_and_empty", "random", "hextext", "lzsyn", "runs", "zeros", "periodic",
         "binary_small_alphabet", "ragged_batch", "short_suffix_vs_zero_bytes"]


@pytest.mark.parametrize("name", NAMES)
def test_bwt_encode_emu(emu_ctx, oracle, gen, name):
    _check(emu_ctx, oracle, _cases(gen, 40000)[name])


def test_bwt_encode_emu_appendix_c(emu_ctx):
    got = _run(emu_ctx, [b"abracadabra", b"banana", b"test"])
    assert got == [(0, b"rdarcaaaabb", 2), (0, b"nnbaaa", 3), (0, b"test", 3)]


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_bwt_encode_gpu_cases(gpu_ctx, oracle, gen, device):
    cases = _cases(gen, 300000)
    for name in NAMES:
        _check(gpu_ctx, oracle, cases[name], device=device)


@pytest.mark.gpu
def test_bwt_gpu_roundtrip_4mib_blocks(gpu_ctx, oracle, gen):
    """BASELINE config 3 shape at reduced count: 8 x 4 MiB random blocks, encode then decode on the device; the first
    block's L column and origin are also checked against the oracle."""
    import torch
    unit, count = 4 << 20, 8
    raw = gen.units("random", gen.unit_seed(3, 0), unit, count)
    d_raw = torch.from_numpy(raw).cuda()
    off = np.arange(count, dtype=np.uint64) * unit
    n = np.full(count, unit, dtype=np.uint64)
    d_l = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    origin, st = gpu_ctx.bwt_encode_blocks(d_raw, off, n, d_l, off)
    assert (st == 0).all()
    d_back = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
    out_len, st = gpu_ctx.bwt_decode_blocks(d_l, off, n, origin, d_back, off)
    assert (st == 0).all() and (out_len == unit).all()
    assert torch.equal(d_back, d_raw)
    ost, ol, oog = oracle.bwt_encode(raw[:unit].tobytes())
    assert ost == 0 and int(origin[0]) == oog and bytes(d_l[:unit].cpu().numpy()) == ol
