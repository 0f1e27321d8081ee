#!/usr/bin/env python
"""Per-op device timings at (scaled) BASELINE config sizes — a tuning aid, not the judged bench (that is bench.py).

    python tools/opbench.py [lz4] [ibwt] [bwt] [flate] [zlib] [mtf] [dc] [ari] [rle] [--blocks N] [--reps R]

Prints one JSON line per op: uncompressed GB/s from CUDA events around the C-ABI call (device-resident buffers)."""
import importlib
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = 4 << 20


def timed(fn, reps):
    import torch
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


def main():
    import torch
    from oracle import oracle
    from tools import gen
    rcz = importlib.import_module("rust-compress_b200")
    opt = {sys.argv[i][2:]: int(sys.argv[i + 1]) for i in range(1, len(sys.argv) - 1) if sys.argv[i].startswith("--")}
    args = [a for a in sys.argv[1:] if not a.startswith("--") and not a.isdigit()]
    ops = args or ["lz4", "lz4enc", "ibwt", "bwt", "flate", "zlib", "mtf", "dc", "ari", "rle"]
    nblk, reps = opt.get("blocks", 32), opt.get("reps", 5)
    ctx = rcz.Context(device=0)
    ctx.set_stream(torch.cuda.current_stream())
    off = np.arange(nblk, dtype=np.uint64) * UNIT
    n = np.full(nblk, UNIT, dtype=np.uint64)

    def report(op, U, ms, extra=None):
        line = {"op": op, "units": nblk, "U_bytes": U, "ms_median": ms[0], "ms_best": ms[1], "GBps": U / ms[0] / 1e6}
        line.update(extra or {})
        print(json.dumps(line), flush=True)

    if "lz4" in ops:
        raw = gen.units("lzsyn", gen.unit_seed(2, 0), UNIT, nblk)
        packed, ioff, ilen = gen.lz4_compress_units(raw, UNIT, nblk)
        d_in, d_out = torch.from_numpy(packed).cuda(), torch.zeros(UNIT * nblk, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: ctx.lz4_decode_blocks(d_in, ioff, ilen, d_out, off, n, async_=True), reps)
        assert torch.equal(d_out, torch.from_numpy(raw).cuda())
        report("lz4_decode", UNIT * nblk, ms, {"C_bytes": int(ilen.sum()), "stage_ms": ctx.last_stage_ms()})
        for kind in ("hextext", "runs", "random"):          # other shapes of input: 9-byte sequences with in-tile chains, long matches, one literal run
            raw = gen.units(kind, gen.unit_seed(2, 0), UNIT, nblk)
            packed, ioff, ilen = gen.lz4_compress_units(raw, UNIT, nblk)
            d_in = torch.from_numpy(packed).cuda()
            ms = timed(lambda: ctx.lz4_decode_blocks(d_in, ioff, ilen, d_out, off, n, async_=True), reps)
            assert torch.equal(d_out, torch.from_numpy(raw).cuda())
            report("lz4_decode_" + kind, UNIT * nblk, ms, {"C_bytes": int(ilen.sum()), "stage_ms": ctx.last_stage_ms()})
    if "lz4enc" in ops:
        for kind in ("lzsyn", "hextext", "random"):
            raw = gen.units(kind, gen.unit_seed(2, 0), UNIT, nblk)
            bound = oracle.lz4_compression_bound(UNIT)
            stride = (bound + 15) // 16 * 16
            coff = np.arange(nblk, dtype=np.uint64) * stride
            d_raw = torch.from_numpy(raw).cuda()
            d_enc = torch.zeros(stride * nblk + 64, dtype=torch.uint8, device="cuda")
            r5 = {}

            def le():
                r5["o"] = ctx.lz4_encode_blocks(d_raw, off, n, d_enc, coff, np.full(nblk, bound, np.uint64))
            ms = timed(le, max(2, reps // 2))
            clen, st = r5["o"]
            assert (st == 0).all()
            d_back = torch.zeros(UNIT * nblk, dtype=torch.uint8, device="cuda")
            ctx.lz4_decode_blocks(d_enc, coff, clen, d_back, off, n)
            assert torch.equal(d_back, d_raw)
            report("lz4_encode_" + kind, UNIT * nblk, ms, {"C_bytes": int(clen.sum())})
    if "ibwt" in ops or "bwt" in ops:
        for kind in ("random", "hextext"):
            raw = gen.units(kind, gen.unit_seed(3, 0), UNIT, nblk)
            d_raw = torch.from_numpy(raw).cuda()
            d_l = torch.zeros(UNIT * nblk, dtype=torch.uint8, device="cuda")
            res = {}

            def enc():
                res["o"] = ctx.bwt_encode_blocks(d_raw, off, n, d_l, off)
            ms = timed(enc, max(2, reps // 2))
            origin, st = res["o"]
            assert (st == 0).all()
            report("bwt_encode_" + kind, UNIT * nblk, ms)
            d_back = torch.zeros(UNIT * nblk, dtype=torch.uint8, device="cuda")
            ms = timed(lambda: ctx.bwt_decode_blocks(d_l, off, n, origin, d_back, off, async_=True), reps)
            assert torch.equal(d_back, d_raw)
            report("bwt_decode_" + kind, UNIT * nblk, ms)
            if kind == "hextext" and "dc" in ops:
                cap = 256 + UNIT
                d_dc = torch.zeros(cap * nblk, dtype=torch.int32, device="cuda")
                coff = np.arange(nblk, dtype=np.uint64) * cap
                r2 = {}

                def dce():
                    r2["o"] = ctx.dc_encode_blocks(d_l, off, n, d_dc, coff, np.full(nblk, cap, np.uint64))
                ms = timed(dce, reps)
                dlen, st = r2["o"]
                assert (st == 0).all()
                report("dc_encode_bwt_hextext", UNIT * nblk, ms, {"distances": int(dlen.sum()) - 256 * nblk})
                d_l2 = torch.zeros(UNIT * nblk, dtype=torch.uint8, device="cuda")
                ms = timed(lambda: ctx.dc_decode_blocks(d_dc, coff, dlen, d_l2, off, n), reps)
                assert torch.equal(d_l2, d_l)
                report("dc_decode_bwt_hextext", UNIT * nblk, ms)
    if "flate" in ops:
        unit, count = 65536, nblk * 64
        raw = gen.units("hextext", gen.unit_seed(4, 0), unit, count)
        from concurrent.futures import ThreadPoolExecutor

        def comp(i):
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            return c.compress(raw[i * unit: (i + 1) * unit].tobytes()) + c.flush()
        with ThreadPoolExecutor(os.cpu_count()) as ex:
            cs = list(ex.map(comp, range(count)))
        lens = np.array([len(c) for c in cs], dtype=np.uint64)
        foff = np.zeros(count, dtype=np.uint64); foff[1:] = np.cumsum(lens)[:-1]
        d_in = torch.from_numpy(np.frombuffer(b"".join(cs) + bytes(64), dtype=np.uint8).copy()).cuda()
        d_out = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
        ooff = np.arange(count, dtype=np.uint64) * unit
        caps = np.full(count, unit, dtype=np.uint64)
        ms = timed(lambda: ctx.flate_decode_streams(d_in, foff, lens, d_out, ooff, caps, async_=True), reps)
        assert torch.equal(d_out, torch.from_numpy(raw).cuda())
        report("inflate_64k_streams", unit * count, ms, {"streams": count, "C_bytes": int(lens.sum())})
    if "zlib" in ops:
        unit, count = 65536, nblk * 64
        raw = gen.units("hextext", gen.unit_seed(4, 0), unit, count)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(os.cpu_count()) as ex:
            cs = list(ex.map(lambda i: zlib.compress(raw[i * unit: (i + 1) * unit].tobytes(), 6), range(count)))
        lens = np.array([len(c) for c in cs], dtype=np.uint64)
        foff = np.zeros(count, dtype=np.uint64); foff[1:] = np.cumsum(lens)[:-1]
        d_in = torch.from_numpy(np.frombuffer(b"".join(cs) + bytes(64), dtype=np.uint8).copy()).cuda()
        d_out = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
        ooff = np.arange(count, dtype=np.uint64) * unit
        caps = np.full(count, unit, dtype=np.uint64)
        ms = timed(lambda: ctx.zlib_decode_streams(d_in, foff, lens, d_out, ooff, caps, async_=True), reps)
        res = ctx.zlib_decode_streams(d_in, foff, lens, d_out, ooff, caps)
        assert (res[1] == 0).all() and torch.equal(d_out, torch.from_numpy(raw).cuda())
        report("zlib_64k_streams (inflate + adler32 + trailer)", unit * count, ms, {"streams": count, "C_bytes": int(lens.sum())})
        ms = timed(lambda: ctx.adler32_streams(d_out, ooff, caps, async_=True), reps)
        report("adler32_64k_streams", unit * count, ms, {"streams": count})
    if "mtf" in ops:
        # (i) C5 shape: one stream per 4 MiB block (the L column of a BWT would come here; hexdump text has the same alphabet);
        # (ii) many short streams: the coder is one dependency chain per stream, throughput comes from the number of streams
        for unit, count, tag in ((UNIT, nblk, "4mib_blocks"), (65536, nblk * 64, "64k_streams")):
            raw = gen.units("hextext", gen.unit_seed(5, 0), unit, count)
            d_raw = torch.from_numpy(raw).cuda()
            d_rk = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
            d_back = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
            so = np.arange(count, dtype=np.uint64) * unit
            sn = np.full(count, unit, dtype=np.uint64)
            ms = timed(lambda: ctx.mtf_encode_streams(d_raw, so, sn, d_rk, so, sn, async_=True), reps)
            report("mtf_encode_" + tag, unit * count, ms, {"streams": count})
            ms = timed(lambda: ctx.mtf_decode_streams(d_rk, so, sn, d_back, so, sn, async_=True), reps)
            assert torch.equal(d_back, d_raw)
            report("mtf_decode_" + tag, unit * count, ms, {"streams": count})
    if "ari" in ops:
        unit, count = 65536, nblk * 64
        raw = gen.units("hextext", gen.unit_seed(5, 0), unit, count)
        d_raw = torch.from_numpy(raw).cuda()
        cap = 2 * unit + 64
        d_enc = torch.zeros(cap * count, dtype=torch.uint8, device="cuda")
        aoff = np.arange(count, dtype=np.uint64) * unit
        coff = np.arange(count, dtype=np.uint64) * cap
        r3 = {}

        def ae():
            r3["o"] = ctx.ari_encode_streams(d_raw, aoff, np.full(count, unit, np.uint64), d_enc, coff, np.full(count, cap, np.uint64))
        ms = timed(ae, reps)
        elen, st = r3["o"]
        assert (st == 0).all()
        report("ari_encode_64k_streams", unit * count, ms, {"C_bytes": int(elen.sum())})
        d_dec = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: ctx.ari_decode_streams(d_enc, coff, elen, d_dec, aoff, np.full(count, unit, np.uint64), async_=True), reps)
        assert torch.equal(d_dec, d_raw)
        report("ari_decode_64k_streams", unit * count, ms)
    if "rle" in ops:
        unit, count = 1 << 20, nblk * 4
        raw = gen.units("runs", gen.unit_seed(1, 0), unit, count)
        d_raw = torch.from_numpy(raw).cuda()
        cap = 2 * unit + 16
        d_enc = torch.zeros(cap * count, dtype=torch.uint8, device="cuda")
        roff = np.arange(count, dtype=np.uint64) * unit
        coff = np.arange(count, dtype=np.uint64) * cap
        r4 = {}

        def re_():
            r4["o"] = ctx.rle_encode_streams(d_raw, roff, np.full(count, unit, np.uint64), d_enc, coff, np.full(count, cap, np.uint64))
        ms = timed(re_, reps)
        elen, st = r4["o"]
        assert (st == 0).all()
        report("rle_encode_1m_streams", unit * count, ms, {"C_bytes": int(elen.sum())})
        d_dec = torch.zeros(unit * count, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: ctx.rle_decode_streams(d_enc, coff, elen, d_dec, roff, np.full(count, unit, np.uint64), async_=True), reps)
        assert torch.equal(d_dec, d_raw)
        report("rle_decode_1m_streams", unit * count, ms)


if __name__ == "__main__":
    main()
