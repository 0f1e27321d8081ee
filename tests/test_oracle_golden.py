"""The oracle against the reference's own fixtures and inline known-answer vectors.
CPU only.  These tests pin the oracle; the GPU parity tests then compare librcz with the oracle."""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, REFDATA, golden

MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
TXT = golden("ref_test.txt")


def test_fixture_integrity():
    for name, sha in MAN["files"].items():
        assert hashlib.sha256(golden(name)).hexdigest() == sha, name
    assert len(TXT) == 3050


# ---- rle.rs:320-352 inline KATs -------------------------------------------------------------------------
RLE_KATS = [
    (b"", b""), (b"a", b"a"), (b"abca123", b"abca123"),
    (bytes([20] * 5 + [15]), bytes([20, 20, 5 - 2 + 128, 15])),
    (bytes([0, 0]), bytes([0, 0, 2 - 2 + 128])),
    (bytes([5] * 129), bytes([5, 5, 255])),
    (bytes([1, 3, 4, 4] + [100] * (2 + 52 + 128)), bytes([1, 3, 4, 4, 0 + 128, 100, 100, 52, 1 + 128])),
]


@pytest.mark.parametrize("raw,enc", RLE_KATS)
def test_rle_kats(oracle, raw, enc):
    assert oracle.rle_encode(raw) == enc
    assert oracle.rle_decode(enc) == (0, raw)


def test_rle_roundtrips_and_errors(oracle, gen):
    for seed in range(20):
        d = gen.one("random", seed, 13579)
        assert oracle.rle_decode(oracle.rle_encode(d)) == (0, d)
    d = gen.one("runs", 3, 200000)
    e = oracle.rle_encode(d)
    assert len(e) < len(d) and oracle.rle_decode(e) == (0, d)
    assert len(oracle.rle_encode(TXT)) == MAN["vectors"]["rle_txt_len"] == 3084
    # rle.rs:151-154: a 10th length byte is "Overly long run"
    st, _ = oracle.rle_decode(bytes([7, 7] + [1] * 10))
    assert st == oracle.E_OVERLONG_RUN
    # rle.rs:247-256: input exhausted mid-run flushes the partial run
    assert oracle.rle_decode(bytes([7, 7, 3])) == (0, bytes([7] * 5))
    assert oracle.rle_decode(bytes([7, 7])) == (0, bytes([7] * 2))


# ---- lz4.rs:647-726 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(1, 10))
def test_lz4_frame_fixtures(oracle, i):
    frame = golden("ref_test.lz4.%d" % i)
    st, out, used = oracle.lz4_frame_decode(frame, 1 << 16)
    assert st == 0 and out == TXT
    assert used == len(frame) - 4          # the content checksum is never read (lz4.rs:384)


def test_lz4_block_roundtrip_and_liblz4(oracle, gen):
    enc = oracle.lz4_encode_block(TXT)
    v = MAN["vectors"]["lz4_encode_block_txt"]
    assert len(enc) == v["len"] and hashlib.sha256(enc).hexdigest() == v["sha256"]
    assert oracle.lz4_decode_block(enc, 4096) == (0, TXT)
    # independent cross-check: liblz4 can decode the restated encoder's output, and the oracle decodes liblz4's
    n, dec = gen.lz4_decompress(enc, 4096)
    assert dec == TXT
    for kind, seed, size in [("lzsyn", 1, 300000), ("hextext", 2, 100000), ("random", 3, 70000), ("runs", 4, 90000)]:
        d = gen.one(kind, seed, size)
        assert oracle.lz4_decode_block(gen.lz4_compress(d), size) == (0, d)
        e = oracle.lz4_encode_block(d)
        assert gen.lz4_decompress(e, size)[1] == d
    assert oracle.lz4_compression_bound(100) == 100 + 0 + 20
    assert oracle.lz4_compression_bound(0x7e000001) is None


def test_lz4_malformed(oracle):
    assert oracle.lz4_decode_block(bytes([0x10, 0x41, 0, 0, 0x00]), 64)[0] == oracle.E_MALFORMED     # offset 0
    assert oracle.lz4_decode_block(bytes([0x10, 0x41, 5, 0, 0x00]), 64)[0] == oracle.E_MALFORMED     # offset before start
    assert oracle.lz4_decode_block(bytes([0xf0, 0xff, 0xff]), 4096)[0] == oracle.E_MALFORMED          # truncated length
    assert oracle.lz4_decode_block(bytes([0x14, 0x41, 1, 0]), 64) == (0, b"A" * 9)                    # ends after a match
    assert oracle.lz4_frame_decode(b"\x00\x01\x02\x03\x04", 16)[0] == oracle.E_INVALID_INPUT           # bad magic


# ---- flate.rs:528-548 -------------------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(10))
def test_flate_fixtures(oracle, i):
    z = golden("ref_test.z.%d" % i)
    st, out, used, detail = oracle.flate_decode(z[2:-4], 1 << 16)       # fixup(): strip zlib header/trailer
    assert st == 0 and out == TXT and used == len(z) - 6
    assert zlib.decompress(z) == TXT
    assert int.from_bytes(z[-4:], "big") == oracle.adler32(TXT) == 0xFB4FCFA6


def test_flate_go_fixture_blocks(oracle):
    st, out, used, detail, blocks = oracle.flate_decode(golden("ref_test.z.go"), 1 << 16, blocks=True)
    assert st == 0 and out == TXT
    assert list(blocks) == [3050, 0, 0]        # two empty stored blocks: where the reference's read() returns Ok(0)


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDATA, "test.large.z.5")), reason="large fixture lives only in /root/reference")
def test_flate_large_fixture(oracle):
    z = open(os.path.join(REFDATA, "test.large.z.5"), "rb").read()
    big = open(os.path.join(REFDATA, "test.large"), "rb").read()
    assert hashlib.sha256(z).hexdigest() == MAN["large"]["test.large.z.5"]["sha256"]
    st, out, used, detail, blocks = oracle.flate_decode(z[2:-4], len(big) + 16, blocks=True)
    assert st == 0 and out == big and len(blocks) == 145


def test_flate_vs_zlib_all_block_types(oracle, gen):
    for kind, seed, size in [("hextext", 1, 65536), ("lzsyn", 2, 200000), ("random", 3, 30000), ("runs", 4, 100000)]:
        d = gen.one(kind, seed, size)
        for level, strategy in [(0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY)]:
            co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
            z = co.compress(d) + co.flush()
            st, out, used, detail = oracle.flate_decode(z, size + 16)
            assert (st, out, used) == (0, d, len(z)), (kind, level, strategy)


def test_flate_errors(oracle):
    E = oracle
    assert E.flate_decode(bytes([0x07]), 16)[::3] == (E.E_INVALID_INPUT, 2)                      # BTYPE 3
    assert E.flate_decode(bytes([0x01, 0x05, 0x00, 0x00, 0x00]), 16)[::3] == (E.E_INVALID_INPUT, 7)   # LEN/NLEN mismatch
    assert E.flate_decode(bytes([0x01, 0x05, 0x00, 0xfa, 0xff, 1, 2]), 16)[0] == E.E_UNEXPECTED_EOF
    assert E.flate_decode(b"", 16)[0] == E.E_UNEXPECTED_EOF
    # distance beyond the history (flate.rs:314): fixed block, length code 257 (len 3), distance code 0 (dist 1), nothing output yet
    bits = "1" + "01" + "0000001" + "00000"       # BFINAL=1, BTYPE=01 (LSB first: '1','0'), then codes MSB-first
    bits = "1" + "10" + "0000001" + "00000"
    by = int(bits[::-1].zfill(16), 2).to_bytes(2, "little")
    assert E.flate_decode(by, 16)[::3] == (E.E_INVALID_INPUT, 6)


# ---- bwt/mod.rs:541-551 + Appendix C ------------------------------------------------------------------------
def test_bwt_vectors(oracle):
    v = MAN["vectors"]["bwt_abracadabra"]
    assert (v["L"], v["origin"]) == ("rdarcaaaabb", 2)
    st, l, org = oracle.bwt_encode(b"abracadabra")
    assert (st, l, org) == (0, b"rdarcaaaabb", 2)
    assert list(oracle.bwt_suffixes(b"abracadabra")) == [10, 7, 0, 3, 5, 8, 1, 4, 6, 9, 2] == v["sa"]
    assert list(oracle.bwt_inversion_table(l, org)[1]) == [0, 6, 7, 8, 9, 10, 11, 5, 2, 1, 4] == v["table"]
    assert oracle.bwt_encode(b"banana")[1:] == (b"nnbaaa", 3)
    assert oracle.bwt_encode(b"test")[1:] == (b"test", 3)
    assert oracle.bwt_stream_encode(b"abracadabra", 1024)[1].hex() == "000400000b000000726461726361616161626202000000"
    st, s = oracle.bwt_stream_encode(TXT, 1024)
    assert len(s) == 3078 and hashlib.sha256(s).hexdigest() == MAN["vectors"]["bwt_stream_txt"]["sha256"]
    assert oracle.bwt_stream_decode(s) == (0, TXT)


def test_bwt_definition_and_roundtrips(oracle, gen):
    # the suffix array is unique: check the oracle against an independent O(n^2 log n) Python sort
    for d in [b"test", b"mississippi", gen.one("hextext", 5, 700), bytes(50), b"ab" * 40, gen.one("random", 6, 500)]:
        sa = sorted(range(len(d)), key=lambda i: d[i:])
        assert list(oracle.bwt_suffixes(d)) == sa
        st, l, org = oracle.bwt_encode(d)
        assert l == bytes(d[i - 1] for i in sa) and sa[org] == 0
        assert oracle.bwt_decode(l, org) == (0, d)
    # some_roundtrips (bwt/mod.rs:541-547): b"test", b"", test.txt with block 1<<10
    for d in [b"test", b"", TXT]:
        st, s = oracle.bwt_stream_encode(d, 1 << 10)
        assert oracle.bwt_stream_decode(s) == (0, d)
    # truncated n field is a clean EOF (bwt/mod.rs:374-378); truncated payload is an error
    st, s = oracle.bwt_stream_encode(b"hello world", 1024)
    assert oracle.bwt_stream_decode(s + b"\x01\x02") == (0, b"hello world")
    assert oracle.bwt_stream_decode(s[:-6])[0] == oracle.E_UNEXPECTED_EOF
    assert oracle.bwt_decode(b"abc", 3)[0] == oracle.E_MALFORMED


def test_mtf(oracle, gen):
    d = gen.one("hextext", 1, 5000)
    r = oracle.mtf_encode(d)
    assert oracle.mtf_decode(r) == d
    assert oracle.mtf_encode(b"aaa") == bytes([97, 0, 0])


# ---- dc.rs:291-302 + Appendix C -------------------------------------------------------------------------------
def test_dc_vectors_and_context_equality(oracle, gen):
    st, init, dist, ctx = oracle.dc_encode(b"teeesst_dc", True)
    assert list(dist) == [3, 1, 0, 0, 0, 0, 0]
    assert {chr(i): int(v) for i, v in enumerate(init) if v < 10} == {"t": 0, "e": 1, "s": 4, "_": 7, "d": 8, "c": 9}
    assert [(chr(a), int(b), int(c)) for a, b, c in zip(*ctx)] == [("t", 0, 10), ("e", 0, 7), ("s", 0, 5), ("t", 2, 4), ("_", 0, 3), ("d", 0, 2), ("c", 0, 1)]
    assert list(oracle.dc_encode(b"abracadabra")[2]) == [0, 2, 2, 0, 2, 0, 1, 0, 0, 0, 0]
    for d in [b"teeesst_dc", b"", TXT, b"../data/test.txt", b"aaaa", gen.one("hextext", 3, 20000), oracle.bwt_encode(gen.one("hextext", 4, 30000))[1]]:
        st, init, dist, ctx = oracle.dc_encode(d, True)
        assert st == 0
        st2, out, used, ctx2 = oracle.dc_decode(len(d), init, dist, True)
        assert st2 == 0 and out == d
        if len(set(d)) > 1:               # a one-symbol block returns before reading any distance (dc.rs:180-187)
            assert used == len(dist)
            for a, b in zip(ctx, ctx2):   # roundtrips_context (dc.rs:268-289): encoder Context == decoder Context
                assert (a == b).all()


# ---- ari/test.rs:185-212 + Appendix C -------------------------------------------------------------------------
def test_ari_vectors_and_roundtrips(oracle, gen):
    assert oracle.ari_encode(b"abracadabra").hex() == "6101aba17aa9d5cc68d39733f600"
    assert oracle.ari_encode(b"").hex() == "ff00ff0000"
    e = oracle.ari_encode(TXT)
    assert len(e) == 1861 and hashlib.sha256(e).hexdigest() == MAN["vectors"]["ari_txt"]["sha256"]
    for d in [b"abracadabra", b"", TXT, gen.one("random", 1, 20000), bytes(10000), gen.one("runs", 2, 30000)]:
        e = oracle.ari_encode(d)
        st, out, c_read, c_fin = oracle.ari_decode(e, len(d) + 8)
        assert (st, out) == (0, d) and c_fin == len(e)
    # roundtrips_term (ari/test.rs:52-89): two terminated streams back to back; finish() re-syncs the reader
    e1, e2 = oracle.ari_encode(b"abra"), oracle.ari_encode(b"cadabra")
    st, out, c_read, c_fin = oracle.ari_decode(e1 + e2, 64)
    assert (st, out, c_fin) == (0, b"abra", len(e1))
    assert oracle.ari_decode((e1 + e2)[c_fin:], 64)[:2] == (0, b"cadabra")
    assert oracle.ari_decode(e1[:3], 64)[0] == oracle.E_MALFORMED       # truncated stream: feed().unwrap() panics
