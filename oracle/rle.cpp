// oracle/rle.cpp — CPU restatement of /root/reference/src/rle.rs (TEST INFRASTRUCTURE ONLY, see oracle.h)
#include "oracle.h"
#include <cstring>
#include <vector>

namespace {
struct Sink {
    uint8_t* out; size_t cap; size_t len = 0; bool full = false;
    void put(uint8_t b) { if (len < cap) out[len] = b; else full = true; ++len; }
};

// rle.rs:96-122  Encoder::flush — emit the pending (byte, reps) pair.
void flush_run(Sink& s, uint8_t byte, uint64_t reps) {
    if (reps == 1) {
        s.put(byte);
    } else if (reps > 1) {
        uint64_t v = reps - 2;            // rle.rs:101
        s.put(byte); s.put(byte);         // rle.rs:103-104
        for (;;) {                        // rle.rs:106-116: 7-bit LE groups, MSB set on the LAST group
            uint8_t g = (uint8_t)(v & 0x7f);
            v >>= 7;
            if (v == 0) { s.put(g | 0x80); break; }
            s.put(g);
        }
    }
}
}  // namespace

// rle.rs:40-123: Encoder driven by ONE write() call over the whole buffer followed by finish()
// (the reference drops the first byte of later write() calls — SURVEY App. B #11 — so the
// well-defined behaviour is the single-call one, which is what its own tests exercise, rle.rs:292-298).
extern "C" int orc_rle_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    Sink s{out, cap};
    uint64_t reps = 0; uint8_t byte = 0;
    if (n > 0) { byte = in[0]; reps = 1; }              // rle.rs:83-87
    for (size_t i = 1; i < n; ++i) {                    // rle.rs:89-91 -> process_byte rle.rs:68-78
        if (in[i] == byte) { ++reps; }
        else { flush_run(s, byte, reps); reps = 1; byte = in[i]; }
    }
    flush_run(s, byte, reps);                           // finish -> flush, rle.rs:62-66
    *out_len = s.len;
    return s.full ? ORC_E_OUTPUT_FULL : ORC_OK;
}

// rle.rs:176-281: Decoder::read_run state machine restated over a whole input slice.
extern "C" int orc_rle_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    enum { CLEAN, SINGLE, RUN } state = CLEAN;
    uint8_t cur = 0;
    uint8_t slice[9]; int count = 0;
    size_t len = 0; bool full = false;
    auto emit = [&](uint8_t b, uint64_t reps) {
        uint64_t room = len < cap ? cap - len : 0;
        uint64_t w = reps < room ? reps : room;
        if (w) memset(out + len, b, (size_t)w);
        if (w < reps) full = true;
        len = (len + reps < len) ? SIZE_MAX : len + (size_t)reps;   // saturating virtual length
    };
    auto to_run = [&]() -> uint64_t {                   // rle.rs:140-149
        uint64_t v = 0;
        for (int i = 0; i < 9; ++i) v |= (uint64_t)(slice[i] & 0x7f) << (i * 7);
        return 2 + v;
    };
    for (size_t i = 0; i < n; ++i) {
        uint8_t b = in[i];
        switch (state) {
        case CLEAN: state = SINGLE; cur = b; break;                          // rle.rs:219-221
        case SINGLE:
            if (b == cur) { state = RUN; for (auto& x : slice) x = 0; count = 0; }   // rle.rs:223-224
            else { emit(cur, 1); cur = b; }                                  // rle.rs:226-228
            break;
        case RUN:
            if (count >= 9) { *out_len = len; return ORC_E_OVERLONG_RUN; }   // rle.rs:151-154
            slice[count++] = b;                                              // rle.rs:155-156
            if (b & 0x80) { emit(cur, to_run()); state = CLEAN; }            // rle.rs:234-238,243-245
            break;
        }
    }
    // rle.rs:247-256: input exhausted -> flush partial state
    if (state == SINGLE) emit(cur, 1);
    else if (state == RUN) emit(cur, to_run());
    *out_len = len;
    return full ? ORC_E_OUTPUT_FULL : ORC_OK;
}
