// oracle/bwt.cpp — CPU restatement of /root/reference/src/bwt/mod.rs and bwt/mtf.rs::MTF
// (TEST INFRASTRUCTURE ONLY, see oracle.h)
#include "oracle.h"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {
// bwt/mod.rs:73-130  Radix: 257 counters, gather / accumulate / place / shift
struct Radix {
    size_t freq[257];
    Radix() { memset(freq, 0, sizeof freq); }
    void gather(const uint8_t* in, size_t n) { for (size_t i = 0; i < n; ++i) freq[in[i]]++; }
    void accumulate() { size_t n = 0; for (auto& f : freq) { size_t t = f; f = n; n += t; } }
    bool place(uint8_t b, size_t& pos) {           // bwt/mod.rs:112-119 (assert! -> false)
        pos = freq[b];
        if (!(freq[b] < freq[(size_t)b + 1])) return false;
        freq[b]++;
        return true;
    }
    void shift() { for (int i = 255; i >= 0; --i) freq[i + 1] = freq[i]; freq[0] = 0; }
};

// Rust slice Ord (`input[a..].cmp(&input[b..])`, bwt/mod.rs:160-162): lexicographic, a strict prefix sorts first.
inline bool suffix_less(const uint8_t* in, size_t n, size_t a, size_t b) {
    size_t la = n - a, lb = n - b, m = la < lb ? la : lb;
    int c = memcmp(in + a, in + b, m);
    if (c != 0) return c < 0;
    return la < lb;
}
}  // namespace

// bwt/mod.rs:136-166  compute_suffixes: bucket by first byte, then comparison-sort every bucket.
extern "C" int orc_bwt_suffixes(const uint8_t* in, size_t n, uint32_t* sa) {
    Radix radix;
    radix.gather(in, n);
    radix.accumulate();
    for (size_t i = 0; i < n; ++i) { size_t p; if (!radix.place(in[i], p)) return ORC_E_MALFORMED; sa[p] = (uint32_t)i; }
    radix.shift();
    for (int c = 0; c < 256; ++c) {
        size_t lo = radix.freq[c], hi = radix.freq[c + 1];
        if (lo == hi) continue;
        std::sort(sa + lo, sa + hi, [&](uint32_t a, uint32_t b) { return suffix_less(in, n, a, b); });
    }
    return ORC_OK;
}

// bwt/mod.rs:193-219  TransformIterator / encode: L[i] = in[SA[i]-1], or in[n-1] at SA[i]==0 (that i is `origin`).
extern "C" int orc_bwt_encode(const uint8_t* in, size_t n, uint8_t* out_l, uint32_t* origin) {
    if (n == 0) return ORC_E_MALFORMED;            // get_origin().unwrap() on None (bwt/mod.rs:186-188)
    std::vector<uint32_t> sa(n);
    int st = orc_bwt_suffixes(in, n, sa.data());
    if (st) return st;
    for (size_t i = 0; i < n; ++i) {
        if (sa[i] == 0) { *origin = (uint32_t)i; out_l[i] = in[n - 1]; }
        else out_l[i] = in[sa[i] - 1];
    }
    return ORC_OK;
}

// bwt/mod.rs:223-239  compute_inversion_table: stable counting sort with `origin` placed first;
// entries are index+1, the origin's entry is 0.
extern "C" int orc_bwt_inversion_table(const uint8_t* l, size_t n, size_t origin, uint32_t* table) {
    if (origin >= n) return ORC_E_MALFORMED;       // input[origin] index panic (also n == 0)
    Radix radix;
    radix.gather(l, n);
    radix.accumulate();
    size_t p;
    if (!radix.place(l[origin], p)) return ORC_E_MALFORMED;
    table[p] = 0;
    for (size_t i = 0; i < origin; ++i) { if (!radix.place(l[i], p)) return ORC_E_MALFORMED; table[p] = (uint32_t)(i + 1); }
    for (size_t i = 0; i + origin + 1 < n; ++i) { if (!radix.place(l[origin + 1 + i], p)) return ORC_E_MALFORMED; table[p] = (uint32_t)(origin + 2 + i); }
    return ORC_OK;
}

// bwt/mod.rs:266-294  InverseIterator driven to exhaustion (the stream decoder's `for ch in decode(..)`,
// bwt/mod.rs:391-393): follow cur = table[cur]-1 from origin; the hop that lands on the 0 entry emits
// input[origin] and ends the iteration.
extern "C" int orc_bwt_decode(const uint8_t* l, size_t n, size_t origin, uint8_t* out, size_t* out_len) {
    *out_len = 0;
    std::vector<uint32_t> table(n);
    int st = orc_bwt_inversion_table(l, n, origin, table.data());
    if (st) return st;
    size_t cur = origin, k = 0;
    while (cur != SIZE_MAX) {
        cur = (size_t)table[cur] - 1;              // wrapping_sub(1)
        size_t p = cur != SIZE_MAX ? cur : origin;
        if (k >= n) return ORC_E_MALFORMED;        // cannot happen: table is injective
        out[k++] = l[p];
    }
    *out_len = k;
    return ORC_OK;
}

static void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }

// bwt/mod.rs:461-518  Encoder: u32 LE block_size header on the first write, then per block
// [u32 n][n bytes L][u32 origin].  One write() over the whole input followed by finish().
extern "C" int orc_bwt_stream_encode(const uint8_t* in, size_t n, uint32_t block_size, uint8_t* out, size_t cap,
                                     size_t* out_len) {
    std::vector<uint8_t> v;
    put32(v, block_size);                                          // bwt/mod.rs:493-496
    if (block_size == 0 && n > 0) return ORC_E_ARG;                // reference would loop forever
    for (size_t off = 0; off < n; off += block_size) {
        size_t m = std::min<size_t>(block_size, n - off);
        put32(v, (uint32_t)m);                                     // bwt/mod.rs:463
        size_t base = v.size();
        v.resize(base + m);
        uint32_t origin = 0;
        int st = orc_bwt_encode(in + off, m, v.data() + base, &origin);
        if (st) return st;
        put32(v, origin);                                          // bwt/mod.rs:475
    }
    *out_len = v.size();
    if (v.size() > cap) return ORC_E_OUTPUT_FULL;
    memcpy(out, v.data(), v.size());
    return ORC_OK;
}

// bwt/mod.rs:362-432  Decoder as driven by read_to_end.
extern "C" int orc_bwt_stream_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
    size_t p = 0, o = 0;
    *out_len = 0;
    auto rd32 = [&](uint32_t& v) -> bool {
        if (p + 4 > n) { p = n; return false; }
        v = (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24);
        p += 4; return true;
    };
    uint32_t max_block;
    if (!rd32(max_block)) return ORC_E_UNEXPECTED_EOF;             // bwt/mod.rs:362-371 (mapped to Other "unexpected end of file")
    for (;;) {
        uint32_t m;
        if (!rd32(m)) break;                                       // bwt/mod.rs:374-378: ANY short read here is a clean EOF
        if (p + m > n) { *out_len = o; return ORC_E_UNEXPECTED_EOF; }          // push_exactly
        const uint8_t* l = in + p; p += m;
        uint32_t origin;
        if (!rd32(origin)) { *out_len = o; return ORC_E_UNEXPECTED_EOF; }      // bwt/mod.rs:384 raw UnexpectedEof
        if (m > cap - o) return ORC_E_OUTPUT_FULL;
        size_t got = 0;
        int st = orc_bwt_decode(l, m, origin, out + o, &got);
        if (st) { *out_len = o; return st; }
        o += got;
    }
    *out_len = o;
    return ORC_OK;
}

template <class F> static void parallel_for(size_t n, int nthreads, F f) {
    if (nthreads <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto& t : th) t.join();
}

extern "C" int orc_bwt_decode_blocks_mt(const uint8_t* l_base, const uint64_t* off, const uint64_t* n,
                                        const uint32_t* origin, uint8_t* out_base, uint64_t* out_len,
                                        int32_t* status, size_t nblocks, int nthreads) {
    parallel_for(nblocks, nthreads, [&](size_t i) {
        size_t got = 0;
        status[i] = orc_bwt_decode(l_base + off[i], n[i], origin[i], out_base + off[i], &got);
        out_len[i] = got;
    });
    return ORC_OK;
}

extern "C" int orc_bwt_encode_blocks_mt(const uint8_t* in_base, const uint64_t* off, const uint64_t* n,
                                        uint8_t* out_base, uint32_t* origin, int32_t* status, size_t nblocks,
                                        int nthreads) {
    parallel_for(nblocks, nthreads, [&](size_t i) {
        status[i] = orc_bwt_encode(in_base + off[i], n[i], out_base + off[i], &origin[i]);
    });
    return ORC_OK;
}

// bwt/mtf.rs:44-91  MTF::{encode,decode} with the stream coders' alphabetical start (mtf.rs:100-109,140-147)
extern "C" void orc_mtf_encode(const uint8_t* in, size_t n, uint8_t* ranks) {
    uint8_t sym[256];
    for (int i = 0; i < 256; ++i) sym[i] = (uint8_t)i;
    for (size_t k = 0; k < n; ++k) {
        uint8_t s = in[k], next = sym[0];
        if (next == s) { ranks[k] = 0; continue; }
        unsigned rank = 1;
        for (;;) { std::swap(sym[rank], next); if (next == s) break; ++rank; }
        sym[0] = s;
        ranks[k] = (uint8_t)rank;
    }
}
extern "C" void orc_mtf_decode(const uint8_t* ranks, size_t n, uint8_t* out) {
    uint8_t sym[256];
    for (int i = 0; i < 256; ++i) sym[i] = (uint8_t)i;
    for (size_t k = 0; k < n; ++k) {
        unsigned r = ranks[k]; uint8_t s = sym[r];
        for (unsigned i = r; i > 0; --i) sym[i] = sym[i - 1];
        sym[0] = s; out[k] = s;
    }
}
