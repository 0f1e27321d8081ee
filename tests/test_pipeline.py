"""BASELINE configs[4]: bwt -> dc -> entropy::ari chained on the device (csrc/pipeline.cu) against the oracle's composition of
the restated stages (oracle/pipeline.cpp: bwt/mod.rs:136-204, dc.rs:62-159, table.rs:203-219 and their inverses).

The composition / container is this project's (the reference never composes dc with ari); parity is per stage:
  - every stage kernel against its oracle stage on the same blocks (bwt L + origin, dc init + distances, ari code bytes)
  - the container the chained call writes == the container the oracle composition writes, byte for byte
  - decode(encode(x)) == x, and the chained decoder reads the oracle's containers
ari / dc encode bytes are "parity unpinned" (the reference holds round-trip tests only for them).
"""
import os

import numpy as np
import pytest

from conftest import golden

TXT = golden("ref_test.txt")


def _layout(sizes, gap=5, start=7):
    off, cur = [], start
    for s in sizes:
        off.append(cur)
        cur += s + gap
    return np.array(off, dtype=np.uint64), cur + 64


def _cap(n, chunk):
    """upper bound of a container: header + per-stream 2x + 64 (the ByteEncoder bound the stage tests use)"""
    ser = 4 * (256 + n)
    ns = (ser + chunk - 1) // chunk if chunk else 1
    return 24 + 4 * ns + 2 * ser + 64 * ns


def _dev(x, device):
    if not device:
        return x
    import torch
    return torch.from_numpy(x).cuda()


def _host(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else x


def _encode(ctx, blocks, chunk, device, async_=False):
    in_off, total = _layout([len(b) for b in blocks])
    inb = np.full(total, 0xAA, dtype=np.uint8)
    for o, b in zip(in_off, blocks):
        inb[int(o): int(o) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    caps = [_cap(len(b), chunk) for b in blocks]
    out_off, ototal = _layout(caps)
    n = np.array([len(b) for b in blocks], dtype=np.uint64)
    outb = _dev(np.zeros(ototal, dtype=np.uint8), device)
    out_len, origin, status = ctx.bwt_dc_ari_encode_blocks(_dev(inb, device), in_off, n, outb, out_off, np.array(caps, dtype=np.uint64),
                                                           ari_chunk=chunk, async_=async_)
    outb, out_len, origin, status = map(_host, (outb, out_len, origin, status))
    return [(int(s), outb[int(o): int(o) + int(l)].tobytes(), int(org)) for s, o, l, org in zip(status, out_off, out_len, origin)]


def _decode(ctx, containers, ns, chunk, device, async_=False):
    in_off, total = _layout([len(c) for c in containers])
    inb = np.full(total, 0x55, dtype=np.uint8)
    for o, c in zip(in_off, containers):
        inb[int(o): int(o) + len(c)] = np.frombuffer(c, dtype=np.uint8)
    out_off, ototal = _layout(ns)
    outb = _dev(np.zeros(ototal, dtype=np.uint8), device)
    out_len, status = ctx.bwt_dc_ari_decode_blocks(_dev(inb, device), in_off, np.array([len(c) for c in containers], dtype=np.uint64), outb, out_off,
                                                   np.array(ns, dtype=np.uint64), ari_chunk=chunk, async_=async_)
    outb, out_len, status = map(_host, (outb, out_len, status))
    return [(int(s), outb[int(o): int(o) + int(l)].tobytes()) for s, o, l in zip(status, out_off, out_len)]


def _oracle_encode(oracle, blocks, chunk, nthreads=8):
    in_off, total = _layout([len(b) for b in blocks])
    inb = np.zeros(total, dtype=np.uint8)
    for o, b in zip(in_off, blocks):
        inb[int(o): int(o) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    caps = [_cap(len(b), chunk) for b in blocks]
    out_off, ototal = _layout(caps)
    outb = np.zeros(ototal, dtype=np.uint8)
    out_len, origin, status = oracle.bda_encode_blocks_mt(inb, in_off, [len(b) for b in blocks], chunk, outb, out_off, caps, nthreads)
    return [(int(s), outb[int(o): int(o) + int(l)].tobytes(), int(org)) for s, o, l, org in zip(status, out_off, out_len, origin)]


def _blocks(gen, big):
    return [b"abracadabra", TXT, b"a", b"aaaaaaaaaaaaaaaa", gen.one("hextext", 11, big), gen.one("random", 12, big // 3), gen.one("runs", 13, big),
            bytes(range(256)) * 5, gen.one("lzsyn", 14, big // 2), gen.one("hextext", 15, 70001)]


def _check(ctx, oracle, gen, big, chunk, device):
    blocks = _blocks(gen, big)
    ref = _oracle_encode(oracle, blocks, chunk)
    got = _encode(ctx, blocks, chunk, device)
    for i, b in enumerate(blocks):
        assert ref[i][0] == 0 and got[i][0] == 0, (i, ref[i][0], got[i][0])
        assert got[i][2] == ref[i][2], "origin of block %d" % i
        assert got[i][1] == ref[i][1], "container of block %d differs from the oracle composition" % i
    dec = _decode(ctx, [r[1] for r in ref], [len(b) for b in blocks], chunk, device)
    for i, b in enumerate(blocks):
        assert dec[i] == (0, b), "decode of block %d" % i
    # oracle decodes what the device wrote (same bytes, so this pins the oracle composition's own inverse)
    for i in (0, 1, 4):
        c = np.frombuffer(got[i][1], dtype=np.uint8)
        out = np.zeros(len(blocks[i]) + 64, dtype=np.uint8)
        ol, st = oracle.bda_decode_blocks_mt(c, [0], [len(c)], chunk, out, [0], [len(blocks[i])], 1)
        assert int(st[0]) == 0 and out[: int(ol[0])].tobytes() == blocks[i]


def _check_errors(ctx, oracle, gen, chunk, device):
    """header / payload damage: same status from the chained decoder and the oracle composition"""
    b = gen.one("hextext", 21, 9000)
    good = _oracle_encode(oracle, [b], chunk)[0][1]
    n = len(b)

    def mut(f):
        c = bytearray(good)
        f(c)
        return bytes(c)
    cases = [
        ("empty block", b"", 0),
        ("short header", good[:20], n),
        ("bad magic", mut(lambda c: c.__setitem__(0, c[0] ^ 1)), n),
        ("n mismatch", good, n + 1),
        ("nsym too small", mut(lambda c: c.__setitem__(slice(12, 16), (100).to_bytes(4, "little"))), n),
        ("nstreams wrong", mut(lambda c: c.__setitem__(slice(20, 24), (77).to_bytes(4, "little"))), n),
        ("truncated code", good[: len(good) - 9], n),
        ("origin out of range", mut(lambda c: c.__setitem__(slice(8, 12), (n + 5).to_bytes(4, "little"))), n),
        ("corrupt code byte", mut(lambda c: c.__setitem__(len(c) // 2, c[len(c) // 2] ^ 0x5A)), n),
        ("first length too short", mut(lambda c: c.__setitem__(slice(24, 28), (3).to_bytes(4, "little"))), n),
    ]
    conts = [c for _, c, _ in cases]
    ns = [k for _, _, k in cases]
    got = _decode(ctx, conts, ns, chunk, device)
    for (name, c, k), (st, data) in zip(cases, got):
        cb = np.frombuffer(c + b"\0" * 8, dtype=np.uint8)
        out = np.zeros(k + 64, dtype=np.uint8)
        ol, ost = oracle.bda_decode_blocks_mt(cb, [0], [len(c)], chunk, out, [0], [k], 1)
        if name == "corrupt code byte":          # which assert a damaged range-coder stream trips first is not pinned; both must fail or both give bytes
            assert (st == 0) == (int(ost[0]) == 0), name
            if st == 0:
                assert data == out[: int(ol[0])].tobytes(), name
            continue
        assert st == int(ost[0]), (name, st, int(ost[0]))
        assert st != 0 and data == b"", name


@pytest.mark.parametrize("chunk", [0, 4096])
def test_pipeline_emu(emu_ctx, oracle, gen, chunk):
    _check(emu_ctx, oracle, gen, 12000, chunk, device=False)


def test_pipeline_emu_device_kinds(emu_ctx, oracle, gen):
    """the DEVICE flavour of the call on the emulator (host arrays for the results, 'device' data pointers)"""
    blocks = [TXT, gen.one("hextext", 3, 20000)]
    ref = _oracle_encode(oracle, blocks, 4096)
    in_off, total = _layout([len(b) for b in blocks])
    inb = np.zeros(total, dtype=np.uint8)
    for o, b in zip(in_off, blocks):
        inb[int(o): int(o) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    caps = np.array([_cap(len(b), 4096) for b in blocks], dtype=np.uint64)
    out_off, ototal = _layout(caps.tolist())
    outb = np.zeros(ototal, dtype=np.uint8)
    ol, org, st = emu_ctx.bwt_dc_ari_encode_blocks(inb, in_off, [len(b) for b in blocks], outb, out_off, caps, ari_chunk=4096, async_="emu-device")
    for i in range(2):
        assert int(st[i]) == 0 and outb[int(out_off[i]): int(out_off[i]) + int(ol[i])].tobytes() == ref[i][1]


def test_pipeline_emu_errors(emu_ctx, oracle, gen):
    _check_errors(emu_ctx, oracle, gen, 4096, device=False)
    _check_errors(emu_ctx, oracle, gen, 0, device=False)


def test_pipeline_emu_output_full(emu_ctx, oracle, gen):
    b = gen.one("hextext", 5, 20000)
    ref = _oracle_encode(oracle, [b], 4096)[0][1]
    inb = np.frombuffer(b + bytes(64), dtype=np.uint8).copy()
    for cap, want in ((len(ref), 0), (len(ref) - 1, -5), (10, -5)):
        outb = np.zeros(len(ref) + 64, dtype=np.uint8)
        ol, org, st = emu_ctx.bwt_dc_ari_encode_blocks(inb, [0], [len(b)], outb, [0], [cap], ari_chunk=4096)
        assert int(st[0]) == want
        assert int(ol[0]) == (len(ref) if want == 0 else 0)
        if want == 0:
            assert outb[: len(ref)].tobytes() == ref


def test_pipeline_bad_chunk(emu_ctx, rcz):
    z = np.zeros(64, dtype=np.uint8)
    for chunk in (3, 512, 4098):
        with pytest.raises(rcz.RczError):
            emu_ctx.bwt_dc_ari_encode_blocks(z, [0], [8], z, [0], [64], ari_chunk=chunk)


@pytest.mark.gpu
@pytest.mark.parametrize("device", [True, False])
@pytest.mark.parametrize("chunk", [0, 65536])
def test_pipeline_gpu(gpu_ctx, oracle, gen, device, chunk):
    _check(gpu_ctx, oracle, gen, 300000, chunk, device)
    _check_errors(gpu_ctx, oracle, gen, chunk, device)


@pytest.mark.gpu
def test_pipeline_gpu_stage_parity_4mib_text_blocks(gpu_ctx, oracle, gen):
    """BASELINE configs[4] at size: 256 text blocks of 4 MiB.  Every stage kernel's output equals the oracle's stage output for
    every block, the chained call's containers equal the oracle composition's, and the chained decoder returns the input."""
    import torch
    nb, unit, chunk = int(os.environ.get("RCZ_C5_BLOCKS", "256")), 4 << 20, 65536
    threads = os.cpu_count() or 8
    raw = gen.units("hextext", gen.unit_seed(5, 0), unit, nb, nthreads=threads)
    off = np.arange(nb, dtype=np.uint64) * unit
    n = np.full(nb, unit, dtype=np.uint64)
    d_raw = torch.from_numpy(raw).cuda()
    # ---- oracle, stage by stage
    l_ref = np.zeros(unit * nb + 64, dtype=np.uint8)
    org_ref, st = oracle.bwt_encode_blocks_mt(raw, off, n, l_ref, threads)
    assert (st == 0).all()
    caps = np.full(nb, _cap(unit, chunk), dtype=np.uint64)
    c_off = np.arange(nb, dtype=np.uint64) * int(caps[0])
    cont_ref = np.zeros(int(caps[0]) * nb + 64, dtype=np.uint8)
    len_ref, org2, st = oracle.bda_encode_blocks_mt(raw, off, n, chunk, cont_ref, c_off, caps, threads)
    assert (st == 0).all() and (org2 == org_ref).all()
    # ---- stage 1: forward BWT
    d_l = torch.zeros(unit * nb + 64, dtype=torch.uint8, device="cuda")
    org, st = gpu_ctx.bwt_encode_blocks(d_raw, off, n, d_l, off)
    assert (st == 0).all() and (org == org_ref).all()
    assert torch.equal(d_l.cpu(), torch.from_numpy(l_ref)), "BWT L columns differ from the oracle"
    # ---- stage 2: distance coding (checked through the nsym field and, for 8 blocks, word by word)
    dcap = 256 + unit
    d_dc = torch.zeros(dcap * nb, dtype=torch.int32, device="cuda")
    dc_off = np.arange(nb, dtype=np.uint64) * dcap
    dlen, st = gpu_ctx.dc_encode_blocks(d_l, off, n, d_dc, dc_off, np.full(nb, dcap, np.uint64))
    assert (st == 0).all()
    nsym_ref = np.array([int.from_bytes(cont_ref[int(o) + 12: int(o) + 16].tobytes(), "little") for o in c_off], dtype=np.uint64)
    assert (dlen == nsym_ref).all()
    dc_host = d_dc.cpu().numpy().view(np.uint32)
    for i in range(0, nb, max(1, nb // 8)):
        s, init, dist = oracle.dc_encode(l_ref[int(off[i]): int(off[i]) + unit])
        assert np.array_equal(dc_host[int(dc_off[i]): int(dc_off[i]) + int(dlen[i])], np.concatenate([init, dist]))
    del d_dc
    # ---- stages 1-3 chained: containers byte for byte (this covers every block's dc words and ari code bytes)
    d_cont = torch.zeros(int(caps[0]) * nb + 64, dtype=torch.uint8, device="cuda")
    clen, corg, st = gpu_ctx.bwt_dc_ari_encode_blocks(d_raw, off, n, d_cont, c_off, caps, ari_chunk=chunk)
    assert (st == 0).all() and (clen == len_ref).all() and (corg == org_ref).all()
    enc_ms = gpu_ctx.last_stage_ms()
    got = d_cont.cpu().numpy()
    for i in range(nb):
        a, b = int(c_off[i]), int(c_off[i]) + int(clen[i])
        assert np.array_equal(got[a:b], cont_ref[a:b]), "container %d differs from the oracle composition" % i
    # ---- chained decode
    d_back = torch.zeros(unit * nb, dtype=torch.uint8, device="cuda")
    olen, st = gpu_ctx.bwt_dc_ari_decode_blocks(d_cont, c_off, clen, d_back, off, n, ari_chunk=chunk)
    assert (st == 0).all() and (olen == n).all()
    assert torch.equal(d_back, d_raw)
    print("C5 stage ms: encode", enc_ms, "decode", gpu_ctx.last_stage_ms())
