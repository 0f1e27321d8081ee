// oracle/ari.cpp — CPU restatement of /root/reference/src/entropy/ari/mod.rs and table.rs
// (TEST INFRASTRUCTURE ONLY, see oracle.h)
#include "oracle.h"
#include <cstring>

namespace {
typedef uint32_t Border;                                 // ari/mod.rs:55
const Border SYMBOL_MASK = 0xFF000000u;                  // ari/mod.rs:59
const Border THRESHOLD = 1u << 14;                       // ari/mod.rs:61

// ari/mod.rs:67-169 RangeEncoder
struct Range {
    Border low = 0, hai = 0xFFFFFFFFu;
    // ari/mod.rs:117-150 process: returns bytes produced (0..4)
    int process(Border total, Border from, Border to, uint8_t* output) {
        Border range = (hai - low) / total;
        Border lo = low + range * from;
        Border hi = low + range * to;
        int num = 0;
        for (;;) {
            if ((lo ^ hi) & SYMBOL_MASK) {
                if (hi - lo > THRESHOLD) break;
                Border lim = hi & SYMBOL_MASK;
                if (hi - lim >= lim - lo) lo = lim; else hi = lim - 1;
            }
            if (num >= 4) return -1;                     // output[num_shift] index panic (cannot happen for sane input)
            output[num++] = (uint8_t)(lo >> 24);
            lo <<= 8; hi <<= 8;
        }
        low = lo; hai = hi;
        return num;
    }
    // ari/mod.rs:153-159 query
    bool query(Border total, Border code, Border& offset) const {
        Border range = (hai - low) / total;
        if (range == 0) return false;                    // division by zero panic
        offset = (Border)(code - low) / range;
        return true;
    }
};

// ari/table.rs:20-122 Model (257 u16 counters), as configured by ByteEncoder/ByteDecoder (table.rs:194-199)
struct Model {
    Border total; uint16_t table[257]; Border cut_threshold = THRESHOLD >> 2;
    Model() { for (auto& f : table) f = 1; total = 257; }
    void downscale() {                                   // table.rs:82-91
        total = 0;
        for (auto& f : table) { f = (uint16_t)((f + 1) >> 1); total += f; }
    }
    bool update(size_t value) {                          // table.rs:69-79 with add_log=10, add_const=1
        Border add = (total >> 10) + 1;
        if (!(add < 2 * cut_threshold)) return false;
        table[value] = (uint16_t)(table[value] + add);
        total += add;
        if (total >= cut_threshold) { downscale(); if (!(total < cut_threshold)) return false; }
        return true;
    }
    void get_range(size_t value, Border& lo, Border& hi) const {   // table.rs:100-103
        lo = 0; for (size_t i = 0; i < value; ++i) lo += table[i];
        hi = lo + table[value];
    }
    bool find_value(Border offset, size_t& value, Border& lo, Border& hi) const {   // table.rs:105-117
        if (!(offset < total)) return false;             // assert!
        value = 0; lo = 0;
        while ((hi = lo + table[value]) <= offset) { lo = hi; ++value; }
        return true;
    }
};
}  // namespace

// table.rs:210-219 ByteEncoder::write over the whole input + table.rs:203-207 finish
// (terminator 256 coded with the live model, then ari/mod.rs:230-237: `low` as u32 big-endian).
extern "C" int orc_ari_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    Range re; Model m;
    size_t o = 0; bool full = false;
    auto put = [&](uint8_t b) { if (o < cap) out[o] = b; else full = true; ++o; };
    uint8_t buf[4];
    for (size_t i = 0; i <= n; ++i) {
        size_t value = i < n ? in[i] : 256;
        Border lo, hi; m.get_range(value, lo, hi);
        int k = re.process(m.total, lo, hi, buf);        // ari/mod.rs:184-189
        if (k < 0) return ORC_E_MALFORMED;
        for (int j = 0; j < k; ++j) put(buf[j]);
        if (i < n && !m.update(value)) return ORC_E_MALFORMED;
    }
    Border tail = re.low;                                // ari/mod.rs:163-168
    put((uint8_t)(tail >> 24)); put((uint8_t)(tail >> 16)); put((uint8_t)(tail >> 8)); put((uint8_t)tail);
    *out_len = o;
    return full ? ORC_E_OUTPUT_FULL : ORC_OK;
}

// table.rs:255-272 ByteDecoder::read driven until the terminator + ari/mod.rs:271-292 Decoder.
// consumed_read   = stream bytes pulled by the time the terminator is seen (what read_to_end leaves behind it)
// consumed_finish = consumed_read + the pending shift bytes that only finish() consumes (ari/mod.rs:289-292)
extern "C" int orc_ari_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                              size_t* consumed_read, size_t* consumed_finish) {
    Range re; Model m;
    Border code = 0; size_t pending = 4, p = 0, o = 0;
    *out_len = 0;
    for (;;) {
        while (pending) {                                // feed(), ari/mod.rs:271-278 (.unwrap() => panic on EOF)
            if (p >= n) { *out_len = o; return ORC_E_MALFORMED; }
            code = (code << 8) + in[p++]; --pending;
        }
        Border offset;
        if (!re.query(m.total, code, offset)) { *out_len = o; return ORC_E_MALFORMED; }
        size_t value; Border lo, hi;
        if (!m.find_value(offset, value, lo, hi)) { *out_len = o; return ORC_E_MALFORMED; }
        uint8_t tmp[4];
        int shift = re.process(m.total, lo, hi, tmp);    // ari/mod.rs:199
        if (shift < 0) { *out_len = o; return ORC_E_MALFORMED; }
        pending = (size_t)shift;
        if (value == 256) break;                         // table.rs:263-266
        if (!m.update(value)) { *out_len = o; return ORC_E_MALFORMED; }
        if (o >= cap) { *out_len = o; return ORC_E_OUTPUT_FULL; }
        out[o++] = (uint8_t)value;
    }
    *out_len = o;
    if (consumed_read) *consumed_read = p;
    if (consumed_finish) *consumed_finish = p + pending;   // may exceed n: finish() would then report EOF
    return ORC_OK;
}

// checksum/adler.rs:34-44
extern "C" uint32_t orc_adler32(const uint8_t* in, size_t n) {
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; ++i) { a = (a + in[i]) % 65521; b = (b + a) % 65521; }
    return (b << 16) | a;
}
