"""Randomised stress of the kernels changed late in round 2, on the CPU SIMT emulator against the oracle (no GPU needed):
   python tools/stress_emu.py [seconds per op] [op ...]
dc decode (rank-free list update: alphabets, run structure, long runs), inflate (deferred copy stores: every zlib level / strategy,
stored and fixed blocks, overlapping matches), ari (prefetching byte reader: stream lengths around the 16-byte refills)."""
import importlib
import os
import random
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import oracle          # noqa: E402
from util import pack, out_layout  # noqa: E402

rcz = importlib.import_module("rust-compress_b200")


def rand_block(rs, n):
    kind = rs.randrange(5)
    a = rs.choice([1, 2, 3, 5, 17, 31, 32, 33, 64, 65, 96, 128, 129, 200, 256])
    syms = rs.sample(range(256), a)
    if kind == 0:      # iid skewed
        w = [1.0 / (1 + i) ** rs.uniform(0, 2) for i in range(a)]
        return bytes(rs.choices(syms, weights=w, k=n))
    if kind == 1:      # runs, some longer than a warp store
        out = bytearray()
        while len(out) < n:
            out += bytes([rs.choice(syms)]) * rs.choice([1, 1, 2, 3, 7, 31, 32, 33, 64, 100, 1000])
        return bytes(out[:n])
    if kind == 2:      # bwt of text-like data
        words = [bytes(rs.choices(syms, k=rs.randrange(1, 9))) for _ in range(40)]
        t = b" ".join(rs.choice(words) for _ in range(n // 4 + 1))[:n]
        return oracle.bwt_encode(t)[1] if t else t
    if kind == 3:      # periodic
        p = bytes(rs.choices(syms, k=rs.randrange(1, 40)))
        return (p * (n // len(p) + 1))[:n]
    return bytes(rs.choices(syms, k=n))


def stress_dc(ctx, rs, secs):
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        blocks = [rand_block(rs, rs.choice([1, 2, 31, 33, 100, 1000, 5000, 20000])) for _ in range(6)]
        streams = []
        for b in blocks:
            st, init, dist = oracle.dc_encode(b)
            assert st == 0
            s = np.concatenate([init, dist]).astype(np.uint32)
            if rs.random() < 0.3 and len(s) > 257:      # damage: status (and bytes when it still decodes) must follow the oracle
                s = s.copy()
                s[rs.randrange(len(s))] = rs.choice([0, 1, 3, len(b), len(b) + 1, 0xFFFFFFFF, int(s[rs.randrange(len(s))])])
            streams.append(s)
        off, cur = [], 3
        for s in streams:
            off.append(cur); cur += len(s) + 2
        inb = np.zeros(cur + 64, dtype=np.uint32)
        for o, s in zip(off, streams):
            inb[o: o + len(s)] = s
        ns = [len(b) for b in blocks]
        o_off, o_cap, tot = out_layout(ns, gap=3)
        outb = np.zeros(tot, dtype=np.uint8)
        status = ctx.dc_decode_blocks(inb, np.array(off, dtype=np.uint64), np.array([len(s) for s in streams], dtype=np.uint64), outb, o_off, np.array(ns, dtype=np.uint64))
        for i, s in enumerate(streams):
            ost, oout, used = oracle.dc_decode(ns[i], s[:256], s[256:])
            assert int(status[i]) == ost, ("dc status", i, int(status[i]), ost)
            if ost == 0:
                assert outb[int(o_off[i]): int(o_off[i]) + ns[i]].tobytes() == oout, ("dc bytes", i)
        cases += len(blocks)
    return cases


def stress_flate(ctx, rs, secs):
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        raws, units = [], []
        for _ in range(8):
            r = rand_block(rs, rs.choice([0, 1, 10, 300, 5000, 40000]))
            co = zlib.compressobj(rs.randrange(0, 10), zlib.DEFLATED, -15, rs.randrange(1, 10), rs.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
            z = co.compress(r[: len(r) // 2]) + (co.flush(zlib.Z_FULL_FLUSH) if rs.random() < 0.5 else b"") + co.compress(r[len(r) // 2:]) + co.flush()
            if rs.random() < 0.2 and len(z) > 4:
                z = bytearray(z); z[rs.randrange(len(z))] ^= 1 << rs.randrange(8); z = bytes(z)
            raws.append(r); units.append(z)
        caps = [max(len(r) + rs.choice([0, 0, 0, -1, 50]), 0) for r in raws]
        zb, z_off, z_len = pack(units, pad_front=rs.randrange(4), gap=rs.randrange(4), align=1)
        o_off, o_cap, tot = out_layout(caps, gap=rs.randrange(5))
        out = np.zeros(tot, dtype=np.uint8)
        out_len, status, used, detail = ctx.flate_decode_streams(zb, z_off, z_len, out, o_off, o_cap)
        for i in range(len(units)):
            ref = oracle.flate_decode(units[i], caps[i])
            assert int(status[i]) == ref[0], ("flate status", i, int(status[i]), ref[0])
            k = min(int(out_len[i]), caps[i])
            assert out[int(o_off[i]): int(o_off[i]) + k].tobytes() == bytes(ref[1][:k]), ("flate bytes", i)
        cases += len(units)
    return cases


def stress_ari(ctx, rs, secs):
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        raws = [rand_block(rs, rs.choice([0, 1, 15, 16, 17, 31, 32, 33, 47, 48, 49, 500, 4000])) for _ in range(8)]
        inb, in_off, in_len = pack(raws, pad_front=rs.randrange(17), gap=rs.randrange(3), align=1)
        caps = [len(r) + 64 + len(r) // 2 for r in raws]
        o_off, o_cap, tot = out_layout(caps, gap=1)
        enc = np.zeros(tot, dtype=np.uint8)
        res = ctx.ari_encode_streams(inb, in_off, in_len, enc, o_off, o_cap)
        out_len, status = res[0], res[1]
        codes = []
        for i, r in enumerate(raws):
            ref = oracle.ari_encode(r)
            ref = ref[1] if isinstance(ref, tuple) else ref
            got = enc[int(o_off[i]): int(o_off[i]) + int(out_len[i])].tobytes()
            assert status[i] == 0 and got == bytes(ref), ("ari encode", i)
            codes.append(got)
        cb, c_off, c_len = pack(codes, pad_front=rs.randrange(17), gap=rs.randrange(3), align=1)
        d_off, d_cap, dtot = out_layout([len(r) for r in raws], gap=2)
        dec = np.zeros(dtot, dtype=np.uint8)
        res = ctx.ari_decode_streams(cb, c_off, c_len, dec, d_off, d_cap)
        out_len, status = res[0], res[1]
        for i, r in enumerate(raws):
            assert status[i] == 0 and dec[int(d_off[i]): int(d_off[i]) + len(r)].tobytes() == r, ("ari decode", i)
        # damaged code streams: what decodes in the oracle must decode to the same bytes; what fails there must fail here
        # (which of the reference's asserts a damaged stream trips first is not pinned)
        bad = []
        for c in codes:
            b = bytearray(c)
            if b and rs.random() < 0.7:
                b[rs.randrange(len(b))] ^= 1 << rs.randrange(8)
            bad.append(bytes(b[: rs.randrange(len(b) + 1)]) if rs.random() < 0.3 else bytes(b))
        cb, c_off, c_len = pack(bad, pad_front=rs.randrange(17), gap=rs.randrange(3), align=1)
        bcaps = [len(r) + 8 for r in raws]
        d_off, d_cap, dtot = out_layout(bcaps, gap=2)
        dec = np.zeros(dtot, dtype=np.uint8)
        res = ctx.ari_decode_streams(cb, c_off, c_len, dec, d_off, d_cap)
        out_len, status = res[0], res[1]
        for i in range(len(bad)):
            ref = oracle.ari_decode(bad[i], bcaps[i])
            if ref[0] == 0:
                assert status[i] == 0 and dec[int(d_off[i]): int(d_off[i]) + int(out_len[i])].tobytes() == bytes(ref[1]), ("ari damaged decode", i)
            else:
                assert status[i] != 0, ("ari damaged status", i, ref[0])
        cases += 2 * len(raws)
    return cases


def stress_lz4(ctx, rs, secs):
    from tools import gen
    os.environ["RCZ_LZ4_CHUNK_BYTES"] = "30000"                  # >= 8 blocks go through the chunked host pipeline
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        nb = rs.choice([1, 3, 9, 14])
        raws = [rand_block(rs, rs.choice([0, 1, 12, 13, 70, 3000, 20000, 70000])) for _ in range(nb)]
        units = []
        for r in raws:
            c = gen.lz4_compress(r, hc=rs.random() < 0.3) if rs.random() < 0.8 else bytes(oracle.lz4_encode_block(r))
            if rs.random() < 0.25 and len(c) > 1:
                c = bytearray(c)
                for _ in range(rs.randrange(1, 3)):
                    c[rs.randrange(len(c))] = rs.randrange(256)
                c = bytes(c[: rs.randrange(1, len(c) + 1)]) if rs.random() < 0.3 else bytes(c)
            units.append(c)
        roomy = rs.random() < 0.15                                # capacities far beyond the data: the host pipeline's length-first copy path
        caps = [max(len(r) + rs.choice([0, 0, 0, 0, -1, -17, 100]), 0) + (rs.choice([0, 300000, 2000000]) if roomy else 0) for r in raws]
        inb, in_off, in_len = pack(units, pad_front=rs.randrange(20), gap=rs.randrange(4), align=1)
        o_off, o_cap, tot = out_layout(caps, gap=rs.randrange(9))
        out = np.zeros(tot, dtype=np.uint8)
        ref = np.zeros(tot, dtype=np.uint8)
        out_len, status = ctx.lz4_decode_blocks(inb, in_off, in_len, out, o_off, o_cap)
        ref_len, ref_st = oracle.lz4_decode_blocks_mt(inb, in_off, in_len, ref, o_off, o_cap, 1)
        for i in range(nb):
            assert int(status[i]) == int(ref_st[i]), ("lz4 status", i, int(status[i]), int(ref_st[i]))
            if ref_st[i] == 0:
                assert int(out_len[i]) == int(ref_len[i]), ("lz4 len", i)
                k = int(ref_len[i])
                assert out[int(o_off[i]): int(o_off[i]) + k].tobytes() == ref[int(o_off[i]): int(o_off[i]) + k].tobytes(), ("lz4 bytes", i)
            else:
                assert int(out_len[i]) == 0, ("lz4 failed block length", i)
        cases += nb
    return cases


def stress_bwt(ctx, rs, secs):
    os.environ["RCZ_HOST_CHUNK_BYTES"] = "20000"
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        nb = rs.choice([1, 2, 7])
        raws = [rand_block(rs, rs.choice([1, 2, 3, 9, 100, 1000, 6000, 30000])) for _ in range(nb)]
        inb, in_off, in_len = pack(raws, pad_front=rs.randrange(5), gap=rs.randrange(4), align=1)
        o_off, o_cap, tot = out_layout([len(r) for r in raws], gap=rs.randrange(5))
        out = np.zeros(tot, dtype=np.uint8)
        origin, status = ctx.bwt_encode_blocks(inb, in_off, in_len, out, o_off)
        ls = []
        for i, r in enumerate(raws):
            st, l, org = oracle.bwt_encode(r)
            assert status[i] == 0 and int(origin[i]) == org and out[int(o_off[i]): int(o_off[i]) + len(r)].tobytes() == l, ("bwt encode", i, len(r))
            ls.append(l)
        lb, l_off, l_len = pack(ls, pad_front=rs.randrange(5), gap=rs.randrange(4), align=1)
        back = np.zeros(tot, dtype=np.uint8)
        org2 = np.array([int(o) for o in origin], dtype=np.uint32)
        bad = rs.randrange(nb) if rs.random() < 0.3 else -1
        if bad >= 0:
            org2[bad] = rs.choice([len(raws[bad]), len(raws[bad]) + 5, 0xFFFFFF])
        out_len, status = ctx.bwt_decode_blocks(lb, l_off, l_len, org2, back, o_off)
        for i, r in enumerate(raws):
            ref = oracle.bwt_decode(ls[i], int(org2[i]))
            assert int(status[i]) == ref[0], ("bwt decode status", i, int(status[i]), ref[0])
            if ref[0] == 0:
                assert back[int(o_off[i]): int(o_off[i]) + len(r)].tobytes() == bytes(ref[1]), ("bwt decode", i)
        cases += nb
    return cases


def stress_mtf_rle(ctx, rs, secs):
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        raws = [rand_block(rs, rs.choice([0, 1, 31, 32, 33, 500, 5000])) for _ in range(8)]
        inb, in_off, in_len = pack(raws, pad_front=rs.randrange(5), gap=rs.randrange(3), align=1)
        o_off, o_cap, tot = out_layout([2 * len(r) + 16 for r in raws], gap=1)
        for enc, dec, oenc in ((ctx.mtf_encode_streams, ctx.mtf_decode_streams, oracle.mtf_encode), (ctx.rle_encode_streams, ctx.rle_decode_streams, oracle.rle_encode)):
            e = np.zeros(tot, dtype=np.uint8)
            res = enc(inb, in_off, in_len, e, o_off, o_cap)
            out_len, status = res[0], res[1]
            codes = []
            for i, r in enumerate(raws):
                got = e[int(o_off[i]): int(o_off[i]) + int(out_len[i])].tobytes()
                assert status[i] == 0 and got == bytes(oenc(r)), ("encode", enc.__name__, i)
                codes.append(got)
            cb, c_off, c_len = pack(codes, pad_front=rs.randrange(5), gap=rs.randrange(3), align=1)
            d_off, d_cap, dtot = out_layout([len(r) for r in raws], gap=2)
            d = np.zeros(dtot, dtype=np.uint8)
            res = dec(cb, c_off, c_len, d, d_off, d_cap)
            out_len, status = res[0], res[1]
            for i, r in enumerate(raws):
                assert status[i] == 0 and d[int(d_off[i]): int(d_off[i]) + len(r)].tobytes() == r, ("decode", dec.__name__, i)
        # damaged / truncated rle streams: status and the bytes delivered before the error follow the oracle (rle.rs:151-154, lib.rs:53-62)
        rcodes = [bytes(oracle.rle_encode(r)) for r in raws]
        bad = []
        for c in rcodes:
            b = bytearray(c)
            if b and rs.random() < 0.7:
                b[rs.randrange(len(b))] = rs.randrange(256)
            bad.append(bytes(b[: rs.randrange(len(b) + 1)]) if rs.random() < 0.4 else bytes(b))
        cb, c_off, c_len = pack(bad, pad_front=rs.randrange(5), gap=rs.randrange(3), align=1)
        bcaps = [len(r) + rs.choice([0, 40, 400]) for r in raws]
        d_off, d_cap, dtot = out_layout(bcaps, gap=2)
        d = np.zeros(dtot, dtype=np.uint8)
        res = ctx.rle_decode_streams(cb, c_off, c_len, d, d_off, d_cap)
        out_len, status = res[0], res[1]
        for i in range(len(bad)):
            ost, oout = oracle.rle_decode(bad[i], bcaps[i])
            assert int(status[i]) == ost, ("rle damaged status", i, int(status[i]), ost)
            k = min(int(out_len[i]), bcaps[i])
            assert d[int(d_off[i]): int(d_off[i]) + k].tobytes() == bytes(oout[:k]), ("rle damaged bytes", i)
        cases += 3 * len(raws)
    return cases


def stress_pipeline(ctx, rs, secs):
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        nb = rs.choice([1, 3, 5])
        chunk = rs.choice([0, 1024, 4096, 65536])
        raws = [rand_block(rs, rs.choice([1, 2, 50, 3000, 20000])) for _ in range(nb)]
        inb, in_off, in_len = pack(raws, pad_front=rs.randrange(5), gap=rs.randrange(4), align=1)
        caps = [24 + 4 * 300 + 3 * 4 * (256 + len(r)) + 64 * (2 + (4 * (256 + len(r))) // max(chunk, 1024)) for r in raws]
        c_off, c_cap, ctot = out_layout(caps, gap=3)
        cont, ref = np.zeros(ctot, dtype=np.uint8), np.zeros(ctot, dtype=np.uint8)
        clen, org, st = ctx.bwt_dc_ari_encode_blocks(inb, in_off, in_len, cont, c_off, c_cap, ari_chunk=chunk)
        rlen, rorg, rst = oracle.bda_encode_blocks_mt(inb, in_off, in_len, chunk, ref, c_off, c_cap, 1)
        for i in range(nb):
            assert int(st[i]) == int(rst[i]) and int(clen[i]) == int(rlen[i]), ("pipeline encode", i, int(st[i]), int(rst[i]))
            assert cont[int(c_off[i]): int(c_off[i]) + int(clen[i])].tobytes() == ref[int(c_off[i]): int(c_off[i]) + int(clen[i])].tobytes(), ("container", i)
        if rs.random() < 0.4:                                    # damage one container byte: statuses must follow the oracle
            i = rs.randrange(nb)
            if clen[i] > 0:
                cont[int(c_off[i]) + rs.randrange(int(clen[i]))] ^= 1 << rs.randrange(8)
        o_off, o_cap, tot = out_layout([len(r) for r in raws], gap=2)
        back, rback = np.zeros(tot, dtype=np.uint8), np.zeros(tot, dtype=np.uint8)
        out_len, status = ctx.bwt_dc_ari_decode_blocks(cont, c_off, clen, back, o_off, in_len, ari_chunk=chunk)
        r_len, r_st = oracle.bda_decode_blocks_mt(cont, c_off, clen, chunk, rback, o_off, in_len, 1)
        for i in range(nb):
            same = int(status[i]) == int(r_st[i]) or (int(status[i]) != 0 and int(r_st[i]) != 0)    # which assert a damaged range-coder stream trips first is not pinned
            assert same, ("pipeline decode status", i, int(status[i]), int(r_st[i]))
            if r_st[i] == 0:
                assert back[int(o_off[i]): int(o_off[i]) + len(raws[i])].tobytes() == rback[int(o_off[i]): int(o_off[i]) + len(raws[i])].tobytes(), ("pipeline decode", i)
        cases += nb
    return cases


def stress_zlib_adler(ctx, rs, secs):
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        raws = [rand_block(rs, rs.choice([0, 1, 10, 300, 5000, 40000])) for _ in range(8)]
        units = []
        for r in raws:
            z = zlib.compress(r, rs.randrange(0, 10))
            if rs.random() < 0.15:
                z += zlib.compressobj(6, zlib.DEFLATED, -15).flush()[:0]
            k = rs.random()
            if k < 0.1:
                z = z[:-rs.randrange(1, 5)]                      # trailer cut
            elif k < 0.25 and len(z) > 6:
                z = bytearray(z); z[rs.choice([0, 1, len(z) - 1, len(z) - 4, rs.randrange(len(z))])] ^= 1 << rs.randrange(8); z = bytes(z)
            units.append(z)
        caps = [len(r) + rs.choice([0, 0, 7]) for r in raws]
        zb, z_off, z_len = pack(units, pad_front=rs.randrange(4), gap=rs.randrange(4), align=1)
        o_off, o_cap, tot = out_layout(caps, gap=rs.randrange(5))
        out = np.zeros(tot, dtype=np.uint8)
        out_len, status, used, detail, adler = ctx.zlib_decode_streams(zb, z_off, z_len, out, o_off, o_cap)
        for i in range(len(units)):
            ref = oracle.zlib_decode(units[i], caps[i])
            assert int(status[i]) == ref[0], ("zlib status", i, int(status[i]), ref[0])
            k = min(int(out_len[i]), caps[i])
            assert out[int(o_off[i]): int(o_off[i]) + k].tobytes() == bytes(ref[1][:k]), ("zlib bytes", i)
            if ref[0] == 0:
                assert int(adler[i]) == ref[4] and int(used[i]) == ref[2], ("zlib adler / used", i)
        inb, in_off, in_len = pack(raws, pad_front=rs.randrange(17), gap=rs.randrange(3), align=1)
        ad = ctx.adler32_streams(inb, in_off, in_len)
        for i, r in enumerate(raws):
            assert int(ad[i]) == (zlib.adler32(r) & 0xFFFFFFFF) == oracle.adler32(r), ("adler32", i)
        cases += 2 * len(units)
    return cases


def stress_encoders(ctx, rs, secs):
    t0, cases = time.time(), 0
    while time.time() - t0 < secs:
        raws = [rand_block(rs, rs.choice([0, 1, 11, 12, 13, 100, 3000, 30000, 70000])) for _ in range(5)]
        inb, in_off, in_len = pack(raws, pad_front=rs.randrange(9), gap=rs.randrange(4), align=1)
        caps = [oracle.lz4_compression_bound(len(r)) + 16 for r in raws]
        o_off, o_cap, tot = out_layout(caps, gap=2)
        out = np.zeros(tot, dtype=np.uint8)
        out_len, status = ctx.lz4_encode_blocks(inb, in_off, in_len, out, o_off, o_cap)
        for i, r in enumerate(raws):
            assert status[i] == 0 and out[int(o_off[i]): int(o_off[i]) + int(out_len[i])].tobytes() == bytes(oracle.lz4_encode_block(r)), ("lz4 encode", i, len(r))
        dcaps = [256 + len(r) for r in raws]
        d_off, d_cap, dtot = out_layout(dcaps, gap=1)
        dout = np.zeros(dtot, dtype=np.uint32)
        out_len, status = ctx.dc_encode_blocks(inb, in_off, in_len, dout, d_off, d_cap)
        for i, r in enumerate(raws):
            st, init, dist = oracle.dc_encode(r)
            ref = np.concatenate([init, dist]).astype(np.uint32)
            assert status[i] == 0 and np.array_equal(dout[int(d_off[i]): int(d_off[i]) + int(out_len[i])], ref), ("dc encode", i, len(r))
        cases += 2 * len(raws)
    return cases


def main():
    import faulthandler
    faulthandler.enable()                                         # a crash inside the emulated kernels still says where Python was
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    seed = int(os.environ.get("STRESS_SEED", "1"))
    ctx = rcz.Context(emu=True)
    only = [a for a in sys.argv[2:]]
    for name, fn in (("dc", stress_dc), ("flate", stress_flate), ("ari", stress_ari), ("lz4", stress_lz4), ("bwt", stress_bwt), ("mtf_rle", stress_mtf_rle), ("pipeline", stress_pipeline),
                     ("zlib_adler", stress_zlib_adler), ("encoders", stress_encoders)):
        if only and name not in only:
            continue
        rs = random.Random("%d/%s" % (seed, name))               # every op has its own sequence: a failure does not depend on what ran before
        n = fn(ctx, rs, secs)
        print("%s: %d random cases equal the oracle (seed %d)" % (name, n, seed), flush=True)


if __name__ == "__main__":
    main()
