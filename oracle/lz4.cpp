// oracle/lz4.cpp — CPU restatement of /root/reference/src/lz4.rs (TEST INFRASTRUCTURE ONLY, see oracle.h)
#include "oracle.h"
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

// lz4.rs:64-162  BlockDecoder::decode.  Same control flow: token, literal run, (break if input
// consumed), u16 LE offset, match with the DECR dance for offsets < 4, forward byte copy.
// The reference has no bounds checks (it panics on OOB / reads uninitialised bytes for offset 0,
// SURVEY App. B #8/#9): every such case is ORC_E_MALFORMED here.
extern "C" int orc_lz4_decode_block(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    size_t cur = 0, end = 0;
    *out_len = 0;
    while (cur < n) {                                              // lz4.rs:68
        uint8_t code = in[cur++];                                  // lz4.rs:69 bump
        // lz4.rs:112-122 length()
        size_t len = code >> 4;
        if (len == 0xf) {
            for (;;) {
                if (cur >= n) return ORC_E_MALFORMED;              // bump() index panic
                uint8_t t = in[cur++]; len += t;
                if (t != 0xff) break;
            }
        }
        if (len > 0) {                                             // lz4.rs:75-85
            if (cur >= n || len > n - cur) return ORC_E_MALFORMED; // &input[cur] panic / OOB read
            if (len > cap - end) return ORC_E_OUTPUT_FULL;
            memcpy(out + end, in + cur, len);
            end += len; cur += len;
        }
        if (cur == n) break;                                       // lz4.rs:87
        if (cur + 2 > n) return ORC_E_MALFORMED;                   // bump() panic
        size_t back = (size_t)in[cur] | ((size_t)in[cur + 1] << 8);// lz4.rs:91
        cur += 2;
        if (back > end) return ORC_E_MALFORMED;                    // lz4.rs:93 usize underflow
        if (back == 0) return ORC_E_MALFORMED;                     // self-copy of uninitialised bytes (App. B #9)
        size_t start = end - back;
        size_t mlen = code & 0xf;                                  // lz4.rs:98
        if (mlen == 0xf) {
            for (;;) {
                if (cur >= n) return ORC_E_MALFORMED;
                uint8_t t = in[cur++]; mlen += t;
                if (t != 0xff) break;
            }
        }
        // lz4.rs:99-106: literal<4 -> cp(4, DECR[literal]) then cp(len,0); else cp(len+4, 0)
        static const size_t DECR[4] = {0, 3, 2, 3};
        auto cp = [&](size_t l, size_t decr) -> bool {             // lz4.rs:131-140
            if (l > cap - end) return false;
            for (size_t i = 0; i < l; ++i) out[end + i] = out[start + i];
            end += l; start += l - decr;
            return true;
        };
        if (back < 4) { if (!cp(4, DECR[back])) return ORC_E_OUTPUT_FULL; }
        else mlen += 4;
        if (!cp(mlen, 0)) return ORC_E_OUTPUT_FULL;
        *out_len = end;
    }
    *out_len = end;
    return ORC_OK;
}

// lz4.rs:175-181
extern "C" int64_t orc_lz4_compression_bound(uint32_t size) {
    if (size > 0x7e000000u) return -1;
    return (int64_t)size + (size / 255) + 16 + 4;
}

// lz4.rs:183-311  BlockEncoder::encode (greedy single-probe hash compressor).
extern "C" int orc_lz4_encode_block(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    *out_len = 0;
    int64_t bound = orc_lz4_compression_bound((uint32_t)n);
    if (n > 0x7e000000u || bound < 0) return ORC_OK;               // lz4.rs:229-230 -> returns 0
    std::vector<uint8_t> buf((size_t)bound);
    std::vector<uint32_t> table(1u << 17, 0);                      // lz4.rs:620
    const uint32_t UNINIT = 0x88888888u;
    uint32_t input_len = (uint32_t)n, pos = 0, anchor = 0, dest = 0;
    auto seq_at = [&](uint32_t p) -> uint32_t {                    // lz4.rs:185-190
        return (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24);
    };
    auto write_literals = [&](uint32_t len, uint32_t ml_len, uint32_t p) {   // lz4.rs:192-224
        uint32_t ln = len;
        uint8_t code = ln > 14 ? 15 : (uint8_t)ln;
        buf[dest++] = ml_len > 14 ? (uint8_t)((code << 4) + 15) : (uint8_t)((code << 4) + ml_len);
        if (code == 15) {
            ln -= 15;
            while (ln > 254) { buf[dest++] = 255; ln -= 255; }
            buf[dest++] = (uint8_t)ln;
        }
        for (uint32_t i = 0; i < len; ++i) buf[dest + i] = in[p + i];
        dest += len;
    };
    uint32_t step = 1, limit = 128;                                // lz4.rs:239-240
    for (;;) {
        if (pos + 12 > input_len) {                                // lz4.rs:243-248
            write_literals(input_len - anchor, 0, anchor);
            break;
        }
        uint32_t seq = seq_at(pos);
        uint32_t hash = (uint32_t)(seq * 2654435761u) >> 15;       // lz4.rs:251 (HASH_SHIFT = 32-17)
        uint32_t r = table[hash] + UNINIT;                         // lz4.rs:252 wrapping
        table[hash] = pos - UNINIT;                                // lz4.rs:253 wrapping
        if (((uint32_t)(pos - r) >> 16) != 0 || seq != seq_at(r)) {// lz4.rs:255
            if (pos - anchor > limit) { limit <<= 1; step += 1 + (step >> 2); }
            pos += step;
            continue;
        }
        if (step > 1) {                                            // lz4.rs:264-269
            table[hash] = r - UNINIT;
            pos -= step - 1;
            step = 1;
            continue;
        }
        limit = 128;                                               // lz4.rs:271
        uint32_t ln = pos - anchor, back = pos - r, anc = anchor;
        pos += 4; r += 4; anchor = pos;                            // lz4.rs:277-279
        while (pos < input_len - 5 && in[pos] == in[r]) { ++pos; ++r; }   // lz4.rs:281-284
        uint32_t ml_len = pos - anchor;
        write_literals(ln, ml_len, anc);                           // lz4.rs:288
        buf[dest] = (uint8_t)back; buf[dest + 1] = (uint8_t)(back >> 8); dest += 2;
        if (ml_len > 14) {                                         // lz4.rs:293-304
            ml_len -= 15;
            while (ml_len > 254) { ml_len -= 255; buf[dest++] = 255; }
            buf[dest++] = (uint8_t)ml_len;
        }
        anchor = pos;                                              // lz4.rs:306
    }
    *out_len = dest;
    if (dest > cap) return ORC_E_OUTPUT_FULL;
    memcpy(out, buf.data(), dest);
    return ORC_OK;
}

// lz4.rs:363-500  frame Decoder: read_header + decode_block loop, as driven by read_to_end.
extern "C" int orc_lz4_frame_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                                    size_t* consumed) {
    size_t p = 0, o = 0;
    *out_len = 0; if (consumed) *consumed = 0;
    auto rd32 = [&](uint32_t& v) -> bool {
        if (p + 4 > n) { p = n; return false; }
        v = (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24);
        p += 4; return true;
    };
    uint32_t magic;
    if (!rd32(magic)) return ORC_E_UNEXPECTED_EOF;                 // lz4.rs:365 (raw UnexpectedEof)
    if (magic != 0x184d2204u) return ORC_E_INVALID_INPUT;
    uint8_t flg = 0, bd = 0;                                       // lz4.rs:369-372: one read() of 2 bytes, result ignored
    if (p < n) flg = in[p++];
    if (p < n) bd = in[p++];
    (void)bd;
    if ((flg >> 6) != 1) return ORC_E_INVALID_INPUT;               // lz4.rs:375
    bool blk_checksum = (flg & 0x10) != 0;
    bool stream_size = (flg & 0x08) != 0;
    bool preset = (flg & 0x01) != 0;
    if (stream_size) { if (p + 8 > n) return ORC_E_UNEXPECTED_EOF; p += 8; }   // lz4.rs:402-406
    if (preset) return ORC_E_MALFORMED;                            // lz4.rs:407 assert!
    if (p + 1 > n) return ORC_E_UNEXPECTED_EOF;                    // lz4.rs:417 header checksum, ignored
    p += 1;
    for (;;) {                                                     // lz4.rs:422-464
        uint32_t sz;
        if (!rd32(sz)) { *out_len = o; if (consumed) *consumed = p; return ORC_E_UNEXPECTED_EOF; }
        if (sz == 0) break;                                        // end mark
        if (sz & 0x80000000u) {                                    // raw block
            size_t amt = sz & 0x7fffffffu;
            if (p + amt > n) { *out_len = o; return ORC_E_UNEXPECTED_EOF; }   // push_exactly, lib.rs:111-118
            if (amt > cap - o) return ORC_E_OUTPUT_FULL;
            memcpy(out + o, in + p, amt); o += amt; p += amt;
        } else {
            if (p + sz > n) { *out_len = o; return ORC_E_UNEXPECTED_EOF; }
            size_t got = 0;
            int st = orc_lz4_decode_block(in + p, sz, out + o, cap - o, &got);
            if (st != ORC_OK) { *out_len = o; return st; }
            o += got; p += sz;
        }
        if (blk_checksum) { uint32_t ck; if (!rd32(ck)) { *out_len = o; return ORC_E_UNEXPECTED_EOF; } }  // lz4.rs:459-462
    }
    *out_len = o;
    if (consumed) *consumed = p;                                   // content checksum is never read (lz4.rs:384)
    return ORC_OK;
}

template <class F> static void parallel_for(size_t n, int nthreads, F f) {
    if (nthreads <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto& t : th) t.join();
}

extern "C" int orc_lz4_decode_blocks_mt(const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                        uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                                        uint64_t* out_len, int32_t* status, size_t nblocks, int nthreads) {
    parallel_for(nblocks, nthreads, [&](size_t i) {
        size_t got = 0;
        status[i] = orc_lz4_decode_block(in_base + in_off[i], in_len[i], out_base + out_off[i], out_cap[i], &got);
        out_len[i] = got;
    });
    return ORC_OK;
}
