"""Distance-coding kernels (csrc/dc.cu) against the oracle (bwt/dc.rs:110-233): init[256] + distances bit-exact on
encode, bytes and status on decode.  SURVEY Appendix C vectors included (the reference pins DC by roundtrip only)."""
import numpy as np
import pytest

from conftest import golden

TXT = golden("ref_test.txt")


def _layout(sizes, gap=3):
    off, cur = [], 5
    for s in sizes:
        off.append(cur)
        cur += s + gap
    return np.array(off, dtype=np.uint64), cur + 64


def _encode(ctx, blocks, device=False, cap_slack=0):
    in_off, total = _layout([len(b) for b in blocks])
    inb = np.full(total, 0xAA, dtype=np.uint8)
    for o, b in zip(in_off, blocks):
        inb[int(o): int(o) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    caps = [256 + len(b) + cap_slack for b in blocks]
    out_off, ototal = _layout(caps)
    n = np.array([len(b) for b in blocks], dtype=np.uint64)
    if device:
        import torch
        d_out = torch.zeros(ototal, dtype=torch.int32, device="cuda")
        out_len, status = ctx.dc_encode_blocks(torch.from_numpy(inb).cuda(), in_off, n, d_out, out_off, np.array(caps, dtype=np.uint64))
        outb = d_out.cpu().numpy().view(np.uint32)
    else:
        outb = np.zeros(ototal, dtype=np.uint32)
        out_len, status = ctx.dc_encode_blocks(inb, in_off, n, outb, out_off, np.array(caps, dtype=np.uint64))
    return [(int(s), outb[int(o): int(o) + int(l)].copy()) for s, o, l in zip(status, out_off, out_len)]


def _decode(ctx, streams, ns, device=False):
    in_off, total = _layout([len(s) for s in streams])
    inb = np.zeros(total, dtype=np.uint32)
    for o, s in zip(in_off, streams):
        inb[int(o): int(o) + len(s)] = s
    out_off, ototal = _layout(ns)
    in_len = np.array([len(s) for s in streams], dtype=np.uint64)
    n = np.array(ns, dtype=np.uint64)
    if device:
        import torch
        d_out = torch.zeros(ototal, dtype=torch.uint8, device="cuda")
        status = ctx.dc_decode_blocks(torch.from_numpy(inb.view(np.int32)).cuda(), in_off, in_len, d_out, out_off, n)
        outb = d_out.cpu().numpy()
    else:
        outb = np.zeros(ototal, dtype=np.uint8)
        status = ctx.dc_decode_blocks(inb, in_off, in_len, outb, out_off, n)
    return [(int(s), outb[int(o): int(o) + int(k)].tobytes()) for s, o, k in zip(status, out_off, ns)]


def _blocks(oracle, gen, big):
    def bw(d):
        return oracle.bwt_encode(d)[1]
    return [b"teeesst_dc", b"abracadabra", b"", b"a", b"aaaaaaa", b"ab", TXT, b"../data/test.txt", bw(TXT), bytes(range(256)) * 3,
            bytes(range(255, -1, -1)) + bytes(range(256)), bw(gen.one("hextext", 4, big)), gen.one("random", 5, big // 2),
            gen.one("runs", 6, big), bytes(40000), bw(gen.one("lzsyn", 7, big // 2))]


def _check(ctx, oracle, gen, big, device=False):
    blocks = _blocks(oracle, gen, big)
    got = _encode(ctx, blocks, device=device)
    streams = []
    for i, b in enumerate(blocks):
        st, init, dist = oracle.dc_encode(b)
        assert st == 0 and got[i][0] == 0, i
        ref = np.concatenate([init, dist]).astype(np.uint32)
        assert np.array_equal(got[i][1], ref), "dc encode block %d differs" % i
        streams.append(ref)
    dec = _decode(ctx, streams, [len(b) for b in blocks], device=device)
    for i, b in enumerate(blocks):
        assert dec[i] == (0, b), "dc decode block %d differs" % i
    # error parity: truncated distance list (dc.rs:245-246), corrupted distances / init (asserts at dc.rs:213, :230)
    rs = np.random.RandomState(3)
    bad, ns = [], []
    for src in (6, 8, 11):
        s = streams[src]
        bad.append(s[:-1].copy()); ns.append(len(blocks[src]))
        bad.append(s[: 256 + (len(s) - 256) // 2].copy()); ns.append(len(blocks[src]))
        for _ in range(4):
            t = s.copy()
            t[256 + rs.randint(0, len(t) - 256)] += np.uint32(rs.randint(1, 50))
            bad.append(t); ns.append(len(blocks[src]))
        t = s.copy()
        t[rs.randint(0, 256)] = 3
        bad.append(t); ns.append(len(blocks[src]))
    bad.append(streams[0][:100].copy()); ns.append(10)
    dec = _decode(ctx, bad, ns, device=device)
    for i, (s, k) in enumerate(zip(bad, ns)):
        if len(s) < 256:
            assert dec[i][0] == oracle.E_UNEXPECTED_EOF
            continue
        ost, oout, used = oracle.dc_decode(k, s[:256], s[256:])
        assert dec[i][0] == ost, (i, dec[i][0], ost)
        if ost == 0:
            assert dec[i][1] == oout
    # output too small on encode
    small = _encode(ctx, [TXT], device=device, cap_slack=-3000)
    assert small[0][0] == oracle.E_OUTPUT_FULL


def test_dc_emu(emu_ctx, oracle, gen):
    _check(emu_ctx, oracle, gen, 70000)


def _alphabet_case(ctx, oracle, device=False, n=30000):
    """decode keeps the whole symbol list in registers, 1 / 2 / 4 / 8 ranks per lane by alphabet size: sizes at and around every
    class boundary, skewed symbol frequencies so that re-entries reach every depth"""
    rs = np.random.RandomState(17)
    blocks, streams = [], []
    for a in (2, 3, 31, 32, 33, 63, 64, 65, 96, 97, 127, 128, 129, 160, 161, 192, 193, 224, 225, 255, 256):
        syms = rs.permutation(256)[:a].astype(np.uint8)
        w = 1.0 / (1.0 + np.arange(a)) ** 0.7
        b = syms[rs.choice(a, size=n, p=w / w.sum())]
        b[:a] = syms                                                # every symbol present
        blocks.append(b.tobytes())
        st, init, dist = oracle.dc_encode(blocks[-1])
        assert st == 0
        streams.append(np.concatenate([init, dist]).astype(np.uint32))
    dec = _decode(ctx, streams, [len(b) for b in blocks], device=device)
    for i, b in enumerate(blocks):
        assert dec[i] == (0, b), "alphabet case %d differs" % i
    # one corrupted distance each: status parity, bytes when it still decodes
    bad = []
    for s in streams:
        t = s.copy()
        t[256 + rs.randint(0, len(t) - 256)] += np.uint32(rs.randint(1, 9))
        bad.append(t)
    # two symbols with one first position (a damaged init[] table): the list has a tie, the decoder must take its exact path
    # (leading-hits rank, dc.rs:215-218) and agree with the oracle in status and, where it still decodes, in bytes
    for s_ in streams:
        present = np.nonzero(s_[:256] < n)[0]
        for k in (1, len(present) // 2, len(present) - 1):
            t = s_.copy()
            t[present[k]] = t[present[k - 1]]
            bad.append(t)
    # a symbol that never occurs whose init[] entry is not in [n, n + alphabet): the reference's closing assert (dc.rs:230) looks at
    # all 256 entries (found by tools/stress_emu.py)
    for s_ in streams[:6]:
        absent = np.nonzero(s_[:256] >= n)[0]
        for v in (0xFFFFFFFF, n + 256, n + 1):
            if len(absent):
                t = s_.copy()
                t[absent[len(absent) // 2]] = v
                bad.append(t)
    ns = [len(blocks[0])] * len(bad)
    dec = _decode(ctx, bad, ns, device=device)
    for i, t in enumerate(bad):
        ost, oout, used = oracle.dc_decode(ns[i], t[:256], t[256:])
        assert dec[i][0] == ost, (i, dec[i][0], ost)
        if ost == 0:
            assert dec[i][1] == oout


def test_dc_emu_alphabet_classes(emu_ctx, oracle):
    _alphabet_case(emu_ctx, oracle)


@pytest.mark.gpu
def test_dc_gpu_alphabet_classes(gpu_ctx, oracle):
    _alphabet_case(gpu_ctx, oracle, device=True, n=200000)


def test_dc_emu_appendix_c(emu_ctx):
    got = _encode(emu_ctx, [b"teeesst_dc", b"abracadabra"])
    init = got[0][1][:256]
    assert [int(init[ord(c)]) for c in "tes_dc"] == [0, 1, 4, 7, 8, 9] and int(init[0]) == 10
    assert got[0][1][256:].tolist() == [3, 1, 0, 0, 0, 0, 0]
    init = got[1][1][:256]
    assert [int(init[ord(c)]) for c in "abrcd"] == [0, 1, 2, 4, 6]
    assert got[1][1][256:].tolist() == [0, 2, 2, 0, 2, 0, 1, 0, 0, 0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
def test_dc_gpu(gpu_ctx, oracle, gen, device):
    _check(gpu_ctx, oracle, gen, 1 << 20, device=device)
