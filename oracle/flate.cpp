// oracle/flate.cpp — CPU restatement of /root/reference/src/flate.rs (TEST INFRASTRUCTURE ONLY, see oracle.h)
#include "oracle.h"
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {
const int MAXBITS = 15;           // flate.rs:36
const int MAXLCODES = 286, MAXDCODES = 30, MAXCODES = 316;   // flate.rs:37-39
const size_t HISTORY = 32 * 1024; // flate.rs:40

struct Err { int status; int detail; };
#define FAIL(st, dt) do { e.status = (st); e.detail = (dt); return false; } while (0)

// flate.rs:69-147 HuffmanTree
struct Tree {
    uint16_t count[MAXBITS + 1];
    uint16_t symbol[MAXCODES];
    // flate.rs:83-120 construct
    bool construct(const uint16_t* lens, size_t n, Err& e) {
        memset(count, 0, sizeof count); memset(symbol, 0, sizeof symbol);
        for (size_t i = 0; i < n; ++i) count[lens[i]]++;
        if (count[0] == n) return true;                                  // flate.rs:93
        long left = 1;
        for (int i = 1; i <= MAXBITS; ++i) {                             // flate.rs:98-103: only over-subscription errors
            left *= 2; left -= count[i];
            if (left < 0) FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_HUFFMAN_TREE);
        }
        uint16_t offs[MAXBITS + 1]; memset(offs, 0, sizeof offs);
        for (int i = 1; i < MAXBITS; ++i) offs[i + 1] = offs[i] + count[i];
        for (size_t s = 0; s < n; ++s) if (lens[s] != 0) symbol[offs[lens[s]]++] = (uint16_t)s;
        return true;
    }
};

struct Dec {
    const uint8_t* in; size_t n; size_t p = 0;
    size_t bitbuf = 0, bitcnt = 0;
    uint8_t* out; size_t cap; size_t o = 0;        // whole-stream output (== concatenation of `block`s)
    bool eof = false;
    Err e{ORC_OK, 0};

    // flate.rs:250-260 bits(): LSB-first, byte-at-a-time refill
    bool bits(size_t cnt, uint16_t& v) {
        while (bitcnt < cnt) {
            if (p >= n) FAIL(ORC_E_UNEXPECTED_EOF, 0);                   // raw UnexpectedEof from read_u8
            bitbuf |= (size_t)in[p++] << bitcnt;
            bitcnt += 8;
        }
        v = (uint16_t)(bitbuf & (((size_t)1 << cnt) - 1));
        bitbuf >>= cnt; bitcnt -= cnt;
        return true;
    }
    // flate.rs:129-146 HuffmanTree::decode — one bit per iteration
    bool decode(const Tree& t, uint16_t& sym) {
        uint16_t code = 0, first = 0, index = 0;
        for (int len = 1; len <= MAXBITS; ++len) {
            uint16_t b; if (!bits(1, b)) return false;
            code |= b;
            uint16_t count = t.count[len];
            if (code < first + count) { sym = t.symbol[index + (code - first)]; return true; }
            index += count; first += count; first <<= 1; code <<= 1;
        }
        FAIL(ORC_E_INVALID_INPUT, ORC_FL_NOT_ENOUGH_BITS);
    }
    bool push(uint8_t b) { if (o >= cap) FAIL(ORC_E_OUTPUT_FULL, 0); out[o++] = b; return true; }

    // flate.rs:237-246 statik: LEN/NLEN straight from the byte stream (bit buffer never holds a whole byte)
    bool statik() {
        if (p + 2 > n) { p = n; FAIL(ORC_E_UNEXPECTED_EOF, 0); }
        uint16_t len = (uint16_t)(in[p] | (in[p + 1] << 8)); p += 2;
        if (p + 2 > n) { p = n; FAIL(ORC_E_UNEXPECTED_EOF, 0); }
        uint16_t nlen = (uint16_t)(in[p] | (in[p + 1] << 8)); p += 2;
        if ((uint16_t)~nlen != len) FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_STATIC_SIZE);
        if (p + len > n) { p = n; FAIL(ORC_E_UNEXPECTED_EOF, 0); }       // push_exactly, lib.rs:111-118
        if (len > cap - o) FAIL(ORC_E_OUTPUT_FULL, 0);
        memcpy(out + o, in + p, len); o += len; p += len;
        bitcnt = 0; bitbuf = 0;
        return true;
    }
    // flate.rs:262-341 codes
    bool codes(const Tree& lens, const Tree& dist) {
        static const uint16_t EXTRALENS[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51,
                                               59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint16_t EXTRABITS[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                               4, 5, 5, 5, 5, 0};
        static const uint16_t EXTRADIST[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385,
                                               513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint16_t EXTRADBITS[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9,
                                                10, 10, 11, 11, 12, 12, 13, 13};
        for (;;) {
            uint16_t sym; if (!decode(lens, sym)) return false;
            if (sym < 256) { if (!push((uint8_t)sym)) return false; }
            else if (sym == 256) break;
            else if (sym < 290) {
                uint16_t k = sym - 257;
                if (k > 29) FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_HUFFMAN_CODE);   // flate.rs:294 (off-by-one kept)
                if (k == 29) FAIL(ORC_E_MALFORMED, 0);                   // EXTRALENS[29] index panic (App. B #5)
                uint16_t x; if (!bits(EXTRABITS[k], x)) return false;
                size_t len = (size_t)EXTRALENS[k] + x;
                uint16_t ds; if (!decode(dist, ds)) return false;
                if (ds >= 30) FAIL(ORC_E_MALFORMED, 0);                  // EXTRADIST index panic
                if (!bits(EXTRADBITS[ds], x)) return false;
                size_t d = (size_t)EXTRADIST[ds] + x;
                // flate.rs:314: dist > output.len() where output is the <=32 KiB history of everything so far
                size_t hist = o < HISTORY ? o : HISTORY;
                if (d > hist) FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_HUFFMAN_CODE);
                if (len > cap - o) FAIL(ORC_E_OUTPUT_FULL, 0);
                for (size_t i = 0; i < len; ++i) { out[o] = out[o - d]; ++o; }   // flate.rs:325-334 (ring + overlap == byte-forward copy)
            } else FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_HUFFMAN_CODE);
        }
        return true;
    }
    // flate.rs:343-395 fixed (static trees; symbols 286/287 present, distance codes 30/31 absent)
    bool fixed() {
        uint16_t l[288];
        for (int i = 0; i < 144; ++i) l[i] = 8;
        for (int i = 144; i < 256; ++i) l[i] = 9;
        for (int i = 256; i < 280; ++i) l[i] = 7;
        for (int i = 280; i < 288; ++i) l[i] = 8;
        uint16_t d[30]; for (auto& x : d) x = 5;
        Tree lt, dt;
        if (!lt.construct(l, 288, e) || !dt.construct(d, 30, e)) return false;
        return codes(lt, dt);
    }
    // flate.rs:397-450 dynamic
    bool dynamic() {
        uint16_t v;
        if (!bits(5, v)) return false;
        uint16_t hlit = v + 257;
        if (!bits(5, v)) return false;
        uint16_t hdist = v + 1;
        if (!bits(4, v)) return false;
        uint16_t hclen = v + 4;
        if (hlit > MAXLCODES || hdist > MAXDCODES) FAIL(ORC_E_INVALID_INPUT, ORC_FL_HUFFMAN_TREE_TOO_LARGE);
        static const int ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint16_t cl[19]; memset(cl, 0, sizeof cl);
        for (int i = 0; i < hclen; ++i) { if (!bits(3, v)) return false; cl[ORDER[i]] = v; }
        Tree tree; if (!tree.construct(cl, 19, e)) return false;
        uint16_t lengths[MAXCODES]; memset(lengths, 0, sizeof lengths);
        unsigned i = 0;
        while (i < (unsigned)hlit + hdist) {
            uint16_t s; if (!decode(tree, s)) return false;
            if (s < 16) { lengths[i++] = s; }
            else if (s == 16) {
                if (i == 0) FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_HUFFMAN_HEADER_SYMBOL);
                uint16_t prev = lengths[i - 1];
                if (!bits(2, v)) return false;
                for (unsigned k = 0; k < (unsigned)v + 3; ++k) {
                    if (i >= (unsigned)MAXCODES) FAIL(ORC_E_MALFORMED, 0);   // lengths[i] index panic
                    lengths[i++] = prev;
                }
            }
            else if (s == 17) { if (!bits(3, v)) return false; i += v + 3; }
            else if (s == 18) { if (!bits(7, v)) return false; i += v + 11; }
            else FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_HUFFMAN_HEADER_SYMBOL);
        }
        if (i > (unsigned)hlit + hdist) FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_HUFFMAN_TREE_HEADER);
        Tree lt, dt;
        if (!lt.construct(lengths, hlit, e)) return false;
        if (!dt.construct(lengths + hlit, hdist, e)) return false;
        return codes(lt, dt);
    }
    // flate.rs:195-206 block
    bool block() {
        uint16_t v;
        if (!bits(1, v)) return false;
        if (v == 1) eof = true;
        if (!bits(2, v)) return false;
        switch (v) {
        case 0: return statik();
        case 1: return fixed();
        case 2: return dynamic();
        default: FAIL(ORC_E_INVALID_INPUT, ORC_FL_INVALID_BLOCK_CODE);
        }
    }
};
}  // namespace

// Decode DEFLATE blocks until BFINAL (what repeated Read::read calls deliver, flate.rs:468-488),
// recording every block's output size (an empty block is where the reference's read() returns Ok(0)).
extern "C" int orc_flate_decode_blocks(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                                       size_t* consumed, int* detail, uint32_t* blk_sizes, size_t blk_cap,
                                       size_t* nblk) {
    Dec d; d.in = in; d.n = n; d.out = out; d.cap = cap;
    size_t nb = 0;
    bool ok = true;
    while (!d.eof) {
        size_t before = d.o;
        if (!d.block()) { ok = false; break; }
        if (blk_sizes && nb < blk_cap) blk_sizes[nb] = (uint32_t)(d.o - before);
        ++nb;
    }
    *out_len = d.o;
    if (consumed) *consumed = d.p;
    if (detail) *detail = d.e.detail;
    if (nblk) *nblk = nb;
    return ok ? ORC_OK : d.e.status;
}

extern "C" int orc_flate_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                                size_t* consumed, int* detail) {
    return orc_flate_decode_blocks(in, n, out, cap, out_len, consumed, detail, nullptr, 0, nullptr);
}

extern "C" int orc_flate_decode_streams_mt(const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                           uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                                           uint64_t* out_len, int32_t* status, size_t nstreams, int nthreads) {
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < nstreams;) {
            size_t got = 0;
            status[i] = orc_flate_decode(in_base + in_off[i], in_len[i], out_base + out_off[i], out_cap[i], &got,
                                         nullptr, nullptr);
            out_len[i] = got;
        }
    };
    if (nthreads <= 1) { work(); return ORC_OK; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
    return ORC_OK;
}

// zlib.rs:55-126  zlib::Decoder as driven by read_to_end: two header bytes (validate_header, zlib.rs:55-84), the DEFLATE
// stream block by block, and the big-endian Adler-32 trailer only where the reference reads it (zlib.rs:104-117; checksum/adler.rs:34-44).
// detail: ORC_ZL_* for the wrapper's own errors, ORC_FL_* when the inner flate::Decoder failed.
extern "C" uint32_t orc_adler32(const uint8_t* in, size_t n);
extern "C" int orc_zlib_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len, size_t* consumed,
                               int* detail, uint32_t* adler) {
    *out_len = 0;
    if (consumed) *consumed = 0;
    if (detail) *detail = ORC_FL_NONE;
    if (adler) *adler = 1;
    if (n < 2) return ORC_E_UNEXPECTED_EOF;                                   // read_u8: raw UnexpectedEof
    const uint8_t cmf = in[0], flg = in[1];
    int zd = 0;
    if ((cmf & 0xf) != 0x8) zd = ORC_ZL_UNSUPPORTED_FORMAT;                   // zlib.rs:58-63
    else if ((cmf & 0xf0) != 0x70) zd = ORC_ZL_UNSUPPORTED_WINDOW;            // zlib.rs:65-70
    else if (flg & 0x20) zd = ORC_ZL_PRESET_DICTIONARY;                       // zlib.rs:72-77
    else if ((((unsigned)cmf << 8) + flg) % 31 != 0) zd = ORC_ZL_BAD_HEADER_CHECKSUM;   // zlib.rs:79-84
    if (zd) { if (detail) *detail = zd; if (consumed) *consumed = 2; return ORC_E_INVALID_INPUT; }
    // zlib.rs:99-124 under read_to_end.  Every read() hands out (part of) one DEFLATE block; once the inner decoder has drained its
    // final block `self.inner.eof()` answers Ok(0) at zlib.rs:104-105 and the trailer is NEVER read.  The trailer is only read
    // (zlib.rs:109) when `inner.read` itself returns Ok(0), i.e. when a block decodes to zero bytes (flate.rs:474-476: an empty
    // final block, or an empty stored block in mid-stream) — from wherever the byte reader then stands.
    Dec d; d.in = in + 2; d.n = n - 2; d.out = out; d.cap = cap;
    bool empty_block = false;
    for (;;) {
        const size_t before = d.o;
        if (!d.block()) {
            *out_len = d.o;
            if (consumed) *consumed = 2 + d.p;
            if (detail) *detail = d.e.detail;
            return d.e.status;
        }
        if (d.o == before) { empty_block = true; break; }
        if (d.eof) break;
    }
    const size_t got = d.o, used = d.p;
    *out_len = got;
    if (consumed) *consumed = 2 + used;
    const uint32_t a = orc_adler32(out, got);
    if (adler) *adler = a;
    if (!empty_block) return ORC_OK;                                          // zlib.rs:104-105: eof() -> Ok(0), no trailer check
    if (n - 2 - used < 4) return ORC_E_UNEXPECTED_EOF;                        // read_u32::<BigEndian>
    const uint8_t* t = in + 2 + used;
    const uint32_t ck = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
    if (consumed) *consumed = 2 + used + 4;
    if (ck != a) { if (detail) *detail = ORC_ZL_BAD_CHECKSUM; return ORC_E_INVALID_INPUT; }   // zlib.rs:108-113
    return ORC_OK;
}

