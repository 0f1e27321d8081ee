// flate_decode.cu — K6: raw-DEFLATE decode, one warp per independent stream.
//
// Replaces /root/reference/src/flate.rs: `Decoder::block` (:195-206), `statik` (:237-246), `fixed` (:343-395),
// `dynamic` (:397-450), `codes` (:262-341) and the bit-serial `HuffmanTree::decode` (:129-146), driven to BFINAL
// for every stream (what repeated `Read::read` calls deliver, :468-488).
//
// Shape on the GPU: DEFLATE blocks inside a stream are bit-aligned and share a 32 KiB history, so the parallel
// unit is the stream (BASELINE config 4: 131,072 of them).  All 32 lanes of a warp track the same bit-reader
// state; canonical codes are decoded through per-warp look-up tables in shared memory (10 bits literal/length,
// 9 bits distance) built cooperatively from the code lengths — results identical to the reference's one-bit-
// per-iteration canonical decode, including its acceptance of incomplete codes (:92-103), rejection of
// over-subscribed ones, and `NotEnoughBits` after 15 unmatched bits (:145); longer or unmatched codes take the
// bit-serial path over the same count[]/symbol[] arrays.  Literals are queued 32 at a time and leave as one
// warp-wide store; LZ77 copies are warp-wide, with the overlap rule of :325-334 folded as
// byte j <- byte (j mod dist) of the `dist` bytes before the match.
#include "rcz_internal.h"
#include <algorithm>

namespace flk {

constexpr int NT = 128, WPB = NT / 32;
constexpr int LB = 9, DB = 8;                        // primary LUT bits (32-bit entries: same shared memory as 10 / 9 bits of 16-bit ones)
constexpr int MAXBITS = 15, MAXLCODES = 286, MAXDCODES = 30, MAXCODES = 316;   // flate.rs:36-39
constexpr unsigned HISTORY = 32 * 1024;              // flate.rs:40

struct Tree { unsigned count[16]; unsigned offs[16]; unsigned first[16]; unsigned run[16]; };
struct WarpSmem {
    __align__(16) uint32_t llut[1 << LB];             // see ll_entry / dd_entry
    __align__(16) uint32_t dlut[1 << DB];
    uint16_t lsym[288], dsym[32], csym[20];
    Tree lt, dt, ct;
    uint8_t lens[MAXCODES + 4];
    uint8_t clen[20];
};

__constant__ uint16_t EXTRALENS[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t EXTRABITS[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t EXTRADIST[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
                                       6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t EXTRADBITS[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// LSB-first bit reader (flate.rs:250-260).  The reference refills one byte at a time, so its byte position is always
// ceil(bits consumed / 8), and reading past the input is its UnexpectedEof.  Here the buffer is refilled 32 bits at a time from
// aligned words (the next one already in flight); words past the end read as zero and are counted in `pad`, so "consumed more than
// the input holds" is simply bc < pad — no 64-bit position is carried through the symbol loop.
struct Bits {
    const uint8_t* in; const uint8_t* end;
    const uint32_t* wp;                                  // next word to load
    unsigned long long bb; unsigned bc;                  // bit buffer, valid + padding bits in it
    unsigned pad;                                        // padding bits that entered the buffer
    unsigned nxt;                                        // the word at wp[-1], loaded ahead of its use
    unsigned long long base_bits;                        // bit position (in the stream) of the first bit of the word at wp0
    const uint32_t* wp0;
    __device__ __forceinline__ unsigned fetch() {        // returns the word at wp, masked beyond the end; advances
        unsigned w = 0;
        const uint8_t* p = (const uint8_t*)wp;
        if (p + 4 <= end) w = __ldg(wp);
        else if (p < end) { w = __ldg(wp) & (0xffffffffu >> (8u * (unsigned)(p + 4 - end))); }
        ++wp;
        return w;
    }
    __device__ __forceinline__ unsigned padbits_of(const uint8_t* p) const {     // padding bits inside the word that starts at p
        return p + 4 <= end ? 0u : p >= end ? 32u : 8u * (unsigned)(p + 4 - end);
    }
    __device__ __forceinline__ void seek(unsigned long long bytepos) {
        const uint8_t* p = in + bytepos;
        const unsigned mis = (unsigned)((uintptr_t)p & 3);
        wp = reinterpret_cast<const uint32_t*>(p - mis);
        wp0 = wp;
        base_bits = (bytepos - mis) * 8;                 // may "underflow" by up to 24 bits below zero at bytepos 0: only differences are used
        pad = padbits_of((const uint8_t*)wp);             // padding bits (beyond the input's end) that entered the buffer, cumulative
        bb = (unsigned long long)fetch() >> (8 * mis);
        bc = 32 - 8 * mis;
        nxt = fetch();                                    // not in the buffer yet: its padding is counted when it enters (refill)
    }
    __device__ __forceinline__ void init(const uint8_t* p, unsigned long long n) { in = p; end = p + n; seek(0); }
    __device__ __forceinline__ void refill() {           // bc >= 33 afterwards
        if (bc <= 32) {
            bb |= (unsigned long long)nxt << bc; bc += 32;
            pad += padbits_of((const uint8_t*)(wp - 1));
            nxt = fetch();
        }
    }
    __device__ __forceinline__ unsigned peek(unsigned k) const { return (unsigned)bb & ((1u << k) - 1u); }
    __device__ __forceinline__ void consume(unsigned k) { bb >>= k; bc -= k; }
    __device__ __forceinline__ bool eof() const { return bc < pad; }
    __device__ __forceinline__ unsigned take(unsigned k) { const unsigned v = peek(k); consume(k); return v; }
    // bits consumed so far = bits that entered the buffer - bits still in it
    __device__ __forceinline__ unsigned long long bitpos() const {
        return base_bits + 32ull * (unsigned long long)((wp - 1) - wp0) - bc;    // (wp - 1): `nxt` has not entered the buffer
    }
    __device__ __forceinline__ unsigned long long endbits() const { return 8ull * (unsigned long long)(end - in); }
    __device__ __forceinline__ unsigned long long bytepos() const { return (bitpos() + 7) >> 3; }
    __device__ __forceinline__ void set_bytepos(unsigned long long p) { seek(p); }
};

enum { F_OK = 0, F_EOF = 1, F_INVALID = 2, F_MALFORMED = 3, F_FULL = 4 };
constexpr unsigned long long ZL_EMPTY_BLOCK = 1ull << 63;   // zlib mode: in_used flag "the stream ended on a block of zero bytes"

// Look-up entries (32 bit).  bits 0-3: code length (0 = not in the table); bits 4-5: kind; bits 8-11: extra bits; bits 16-31: value.
//   literal/length table: kind 0 = literal (value = byte), 1 = length (value = base, flate.rs:296), 2 = end of block,
//                         a literal entry with bit 6 set holds TWO literals whose codes fit the table index together: bits 0-3 = both
//                         lengths, bits 8-11 = length of the first, bits 16-23 / 24-31 = first / second byte
//                         3 = bad: value 0 -> InvalidHuffmanCode (flate.rs:294, symbols 287.. and the kept off-by-one), 1 -> EXTRALENS[29] index panic
//   distance table:       kind 0 = distance (value = base, flate.rs:307), 3 = bad (EXTRADIST index panic)
enum { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_BAD = 3 };
constexpr unsigned E_PAIR = 1u << 6;
__device__ __forceinline__ unsigned ll_entry(unsigned sym, unsigned len) {
    if (sym < 256u) return len | (K_LIT << 4) | (sym << 16);
    if (sym == 256u) return len | (K_EOB << 4);
    const unsigned k = sym - 257u;
    if (k > 29u) return len | (K_BAD << 4);                                              // flate.rs:294 (off-by-one kept)
    if (k == 29u) return len | (K_BAD << 4) | (1u << 16);                                // EXTRALENS[29] index panic
    return len | (K_LEN << 4) | ((unsigned)EXTRABITS[k] << 8) | ((unsigned)EXTRALENS[k] << 16);
}
__device__ __forceinline__ unsigned dd_entry(unsigned sym, unsigned len) {
    if (sym >= 30u) return len | (K_BAD << 4);                                           // EXTRADIST index panic
    return len | ((unsigned)EXTRADBITS[sym] << 8) | ((unsigned)EXTRADIST[sym] << 16);
}

// flate.rs:83-120 HuffmanTree::construct, warp-cooperative.  lens[0..n) in shared memory.  Returns false when the
// set is over-subscribed (:98-103).  KIND 0 builds only count[]/symbol[] (code-length tree), 1 the literal/length table, 2 the distance table.
template <int KIND>
__device__ bool build_tree(Tree& t, uint16_t* symtab, uint32_t* lut, const uint8_t* lens, unsigned n) {
    constexpr int lutbits = KIND == 1 ? LB : KIND == 2 ? DB : 0;
    const unsigned lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < 16) t.count[lane] = 0;
    if (lutbits) {
        uint4* l4 = reinterpret_cast<uint4*>(lut);
        for (unsigned i = lane; i < (4u << lutbits) / 16; i += 32) l4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    for (unsigned s = lane; s < n; s += 32) atomicAdd(&t.count[lens[s]], 1u);
    __syncwarp();
    if (t.count[0] == n) { __syncwarp(); if (lane < 16) t.count[lane] = 0; __syncwarp(); return true; }   // :93 no codes: every decode fails
    int left = 1; unsigned off = 0, first = 0;
    bool over = false;
    for (int len = 1; len <= MAXBITS; ++len) {
        const unsigned cnt = t.count[len];
        left = left * 2 - (int)cnt;
        if (left < 0) over = true;
        if (lane == 0) { t.offs[len] = off; t.run[len] = off; t.first[len] = first; }
        off += cnt;
        first = (first + cnt) << 1;
    }
    if (over) return false;
    __syncwarp();
    for (unsigned base = 0; base < n; base += 32) {
        const unsigned s = base + lane;
        const unsigned l = s < n ? (unsigned)lens[s] : 0u;
        const bool act = l != 0;
        const unsigned m = __match_any_sync(RCZ_FULL, act ? l : 100u + lane);
        const unsigned r = __popc(m & ((1u << lane) - 1u));
        const unsigned idx = act ? t.run[l] : 0u;
        __syncwarp();
        if (act && r == 0) t.run[l] = idx + __popc(m);
        __syncwarp();
        if (act) {
            const unsigned pos = idx + r;
            symtab[pos] = (uint16_t)s;
            if ((int)l <= lutbits) {
                const unsigned code = t.first[l] + (pos - t.offs[l]);
                const unsigned rev = __brev(code) >> (32 - l);
                const unsigned e = KIND == 1 ? ll_entry(s, l) : dd_entry(s, l);
                for (unsigned j = rev; j < (1u << lutbits); j += 1u << l) lut[j] = e;
            }
        }
    }
    __syncwarp();
    if (KIND == 1) {
        // second pass: an index whose low bits are a literal code and whose remaining bits decide another literal code gets both.
        // In-place is safe: a converted entry still tells its first literal (length in bits 8-11, byte in bits 16-23).
        for (unsigned j = lane; j < (1u << lutbits); j += 32) {                        // 32 entries per step: all reads, then all writes
            const unsigned e1 = lut[j];
            unsigned ne = e1;
            if ((e1 & 0x70u) == 0 && (e1 & 15u) != 0 && (e1 & 15u) < (unsigned)lutbits) {  // a single literal with index bits to spare
                const unsigned l1 = e1 & 15u;
                const unsigned e2 = lut[j >> l1];
                const unsigned l2 = (e2 & E_PAIR) ? (e2 >> 8) & 15u : e2 & 15u;
                if ((e2 & 0x30u) == 0 && (e2 & 15u) != 0 && l1 + l2 <= (unsigned)lutbits)  // the next code is a literal decided by the bits that are left
                    ne = (l1 + l2) | E_PAIR | (l1 << 8) | (e1 & 0x00ff0000u) | (((e2 >> 16) & 255u) << 24);
            }
            __syncwarp();
            if (ne != e1) lut[j] = ne;
            __syncwarp();
        }
        __syncwarp();
    }
    return true;
}

// flate.rs:129-146 on the next (up to) 15 bits; -1: no code within 15 bits
__device__ __forceinline__ int slow_decode(const Tree& t, const uint16_t* symtab, unsigned bits15, unsigned& len_out) {
    unsigned code = 0, first = 0, index = 0;
    for (unsigned len = 1; len <= (unsigned)MAXBITS; ++len) {
        code |= (bits15 >> (len - 1)) & 1u;
        const unsigned cnt = t.count[len];
        if (code - first < cnt) { len_out = len; return (int)symtab[index + (code - first)]; }
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    return -1;
}

struct St {
    Bits br;
    uint8_t* out; unsigned cap, o;      // output cursor: streams and their outputs are < 2 GiB (rcz.h)
    int detail;
};

// a code that is not in the table: the bit-serial path over count[]/symbol[] (out of line; takes nothing by reference, so the
// decoder state stays in registers).  Returns the table entry of the decoded symbol, 0 when no code matches within 15 bits.
template <int KIND>
__device__ __noinline__ unsigned huff_slow(const Tree* t, const uint16_t* symtab, unsigned bits15) {
    unsigned l = 0;
    const int v = slow_decode(*t, symtab, bits15, l);
    if (v < 0) return 0u;
    return KIND == 1 ? ll_entry((unsigned)v, l) : dd_entry((unsigned)v, l);
}

// flate.rs:262-341 codes.  Every lane of the warp tracks the same bit-reader state (the symbol chain is serial); what the lanes share out
// is the LZ77 copy.  One table load gives the symbol, its extra-bit count and its base value; literals leave through lane 0.
// A match copy is a load from the stream's own recent output (L2) followed by a store of what was loaded; issued back to back the
// store waits out the load's round trip with the whole warp behind it (15 % of the kernel's stall samples).  So a short,
// non-overlapping copy only LOADS, and its store is issued when the next match (whose source may lie in the copied bytes) or
// the end of the block is reached — one or more symbol decodes later, when the data has arrived.
struct PendingCopy { unsigned o, n, v; };                                                // n bytes at out[o + lane], this lane's byte in v
__device__ __forceinline__ void commit(const St& s, PendingCopy& pc, unsigned lane) {
    if (lane < pc.n) s.out[pc.o + lane] = (uint8_t)pc.v;
    pc.n = 0;
}
__device__ __forceinline__ int codes_body(St& s, WarpSmem& w, unsigned lane, PendingCopy& pc) {
    for (;;) {
        s.br.refill();                                                                  // >= 33 bits: code (<= 15) + length extra (<= 5) ...
        unsigned e = w.llut[s.br.peek(LB)];
        // ---- literal batch: while no padding bit has entered the buffer (no EOF possible) and the output has room, up to three table
        // look-ups (<= 27 of the >= 33 buffered bits), each one or two literals, go out with ONE store (lane k writes byte k)
        if (s.br.pad == 0 && s.cap - s.o >= 6u) {
            unsigned long long acc = 0; unsigned cnt = 0, k = 0;
#pragma unroll
            for (; k < 3; ++k) {
                if ((e & 0x30u) != 0 || (e & 15u) == 0) break;                          // not a literal entry of the table
                s.br.consume(e & 15u);
                acc |= (unsigned long long)(e >> 16) << (8u * cnt);
                cnt += 1u + ((e >> 6) & 1u);
                if (k < 2) e = w.llut[s.br.peek(LB)];                                   // (>= 15 bits are left after two look-ups)
            }
            if (cnt) { if (lane < cnt) s.out[s.o + lane] = (uint8_t)(acc >> (8u * lane)); s.o += cnt; }
            if (k == 3) continue;
            s.br.refill();                                                              // the symbol that ended the batch takes the careful path
        }
        if ((e & 15u) == 0) {
            e = huff_slow<1>(&w.lt, w.lsym, s.br.peek(15));
            if (!e) { if (s.br.bitpos() + 15 > s.br.endbits()) return F_EOF; s.detail = RCZ_FL_NOT_ENOUGH_BITS; return F_INVALID; }
        }
        const unsigned kind = (e >> 4) & 3u;
        if (kind == K_LIT) {                                                            // one literal at a time (a pair entry gives its first)
            s.br.consume((e & E_PAIR) ? (e >> 8) & 15u : e & 15u);
            if (s.br.eof()) return F_EOF;
            if (s.o >= s.cap) return F_FULL;
            if (lane == 0) s.out[s.o] = (uint8_t)(e >> 16);
            ++s.o;
            continue;
        }
        s.br.consume(e & 15u);
        if (s.br.eof()) return F_EOF;
        if (kind == K_EOB) return F_OK;
        if (kind == K_BAD) { if (e >> 16) return F_MALFORMED; s.detail = RCZ_FL_INVALID_HUFFMAN_CODE; return F_INVALID; }
        const unsigned len = (e >> 16) + s.br.take((e >> 8) & 15u);                     // flate.rs:296-297
        if (s.br.eof()) return F_EOF;
        s.br.refill();                                                                  // distance code (<= 15) + extra (<= 13)
        unsigned de = w.dlut[s.br.peek(DB)];
        if ((de & 15u) == 0) {
            de = huff_slow<2>(&w.dt, w.dsym, s.br.peek(15));
            if (!de) { if (s.br.bitpos() + 15 > s.br.endbits()) return F_EOF; s.detail = RCZ_FL_NOT_ENOUGH_BITS; return F_INVALID; }
        }
        s.br.consume(de & 15u);
        if (s.br.eof()) return F_EOF;
        if (((de >> 4) & 3u) == K_BAD) return F_MALFORMED;                              // EXTRADIST index panic
        const unsigned d = (de >> 16) + s.br.take((de >> 8) & 15u);                     // flate.rs:307-308
        if (s.br.eof()) return F_EOF;
        const unsigned hist = s.o < HISTORY ? s.o : HISTORY;                            // flate.rs:314
        if (d > hist) { s.detail = RCZ_FL_INVALID_HUFFMAN_CODE; return F_INVALID; }
        if (len > s.cap - s.o) return F_FULL;
        commit(s, pc, lane);                                                            // the previous copy's bytes may be this one's source
        __syncwarp();                                                                   // the literal stores and that copy are ordered before the loads
        const uint8_t* src = s.out + s.o - d;
        uint8_t* dst = s.out + s.o;
        if (d >= len && len <= 32u) {                                                   // the common case: load now, store later
            pc.o = s.o; pc.n = len; pc.v = lane < len ? (unsigned)src[lane] : 0u;
        } else if (d >= len) {                                                          // source and destination do not overlap
            for (unsigned j = lane; j < len; j += 32) dst[j] = src[j];
        } else {
            for (unsigned j = lane; j < len; j += 32) dst[j] = src[j % d];              // flate.rs:325-334: the d bytes before the match, repeated
        }
        __syncwarp();
        s.o += len;
    }
}

__device__ int codes(St& s, WarpSmem& w, unsigned lane) {
    PendingCopy pc{0, 0, 0};
    const int r = codes_body(s, w, lane, pc);
    commit(s, pc, lane);
    __syncwarp();
    return r;
}

// flate.rs:237-246 statik
__device__ int stored(St& s, unsigned lane) {
    const unsigned long long n = (unsigned long long)(s.br.end - s.br.in);
    unsigned long long p = s.br.bytepos();
    if (p + 2 > n) return F_EOF;
    const unsigned len = (unsigned)s.br.in[p] | ((unsigned)s.br.in[p + 1] << 8);
    p += 2;
    if (p + 2 > n) return F_EOF;
    const unsigned nlen = (unsigned)s.br.in[p] | ((unsigned)s.br.in[p + 1] << 8);
    p += 2;
    if (((~nlen) & 0xffffu) != len) { s.br.set_bytepos(p); s.detail = RCZ_FL_INVALID_STATIC_SIZE; return F_INVALID; }
    if (p + len > n) return F_EOF;
    if (len > s.cap - s.o) { s.br.set_bytepos(p); return F_FULL; }
    __syncwarp();
    for (unsigned j = lane; j < len; j += 32) s.out[s.o + j] = s.br.in[p + j];
    __syncwarp();
    s.o += len;
    s.br.seek(p + len);
    return F_OK;
}

__device__ int fixed_block(St& s, WarpSmem& w, unsigned lane) {                          // flate.rs:343-395
    __syncwarp();
    for (unsigned i = lane; i < 288; i += 32) w.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
    __syncwarp();
    build_tree<1>(w.lt, w.lsym, w.llut, w.lens, 288);
    __syncwarp();
    if (lane < 30) w.lens[lane] = 5;
    __syncwarp();
    build_tree<2>(w.dt, w.dsym, w.dlut, w.lens, 30);
    return codes(s, w, lane);
}

__device__ int dynamic_block(St& s, WarpSmem& w, unsigned lane) {                        // flate.rs:397-450
    s.br.refill();
    const unsigned hlit = s.br.take(5) + 257; if (s.br.eof()) return F_EOF;
    const unsigned hdist = s.br.take(5) + 1; if (s.br.eof()) return F_EOF;
    const unsigned hclen = s.br.take(4) + 4; if (s.br.eof()) return F_EOF;
    if (hlit > (unsigned)MAXLCODES || hdist > (unsigned)MAXDCODES) { s.detail = RCZ_FL_HUFFMAN_TREE_TOO_LARGE; return F_INVALID; }
    __syncwarp();
    if (lane < 19) w.clen[lane] = 0;
    for (unsigned i = lane; i < (unsigned)MAXCODES; i += 32) w.lens[i] = 0;
    __syncwarp();
    for (unsigned i = 0; i < hclen; ++i) {
        s.br.refill();
        const unsigned v = s.br.take(3);
        if (s.br.eof()) return F_EOF;
        if (lane == 0) w.clen[CL_ORDER[i]] = (uint8_t)v;
    }
    __syncwarp();
    if (!build_tree<0>(w.ct, w.csym, nullptr, w.clen, 19)) { s.detail = RCZ_FL_INVALID_HUFFMAN_TREE; return F_INVALID; }
    const unsigned total = hlit + hdist;
    unsigned i = 0, prev = 0;
    while (i < total) {
        s.br.refill();
        unsigned l;
        const int v = slow_decode(w.ct, w.csym, s.br.peek(15), l);
        if (v < 0) {
            if (s.br.bitpos() + 15 > s.br.endbits()) return F_EOF;
            s.detail = RCZ_FL_NOT_ENOUGH_BITS; return F_INVALID;
        }
        s.br.consume(l);
        if (s.br.eof()) return F_EOF;
        if (v < 16) { if (lane == 0) w.lens[i] = (uint8_t)v; prev = (unsigned)v; ++i; }
        else if (v == 16) {
            if (i == 0) { s.detail = RCZ_FL_INVALID_HUFFMAN_HEADER_SYMBOL; return F_INVALID; }
            const unsigned rep = s.br.take(2) + 3;
            if (s.br.eof()) return F_EOF;
            for (unsigned k = 0; k < rep; ++k) {
                if (i >= (unsigned)MAXCODES) return F_MALFORMED;                         // lengths[i] index panic
                if (lane == 0) w.lens[i] = (uint8_t)prev;
                ++i;
            }
        } else if (v == 17) { const unsigned z = s.br.take(3); if (s.br.eof()) return F_EOF; i += z + 3; prev = 0; }
        else if (v == 18) { const unsigned z = s.br.take(7); if (s.br.eof()) return F_EOF; i += z + 11; prev = 0; }
        else { s.detail = RCZ_FL_INVALID_HUFFMAN_HEADER_SYMBOL; return F_INVALID; }
    }
    if (i > total) { s.detail = RCZ_FL_INVALID_HUFFMAN_TREE_HEADER; return F_INVALID; }
    __syncwarp();
    if (!build_tree<1>(w.lt, w.lsym, w.llut, w.lens, hlit)) { s.detail = RCZ_FL_INVALID_HUFFMAN_TREE; return F_INVALID; }
    if (!build_tree<2>(w.dt, w.dsym, w.dlut, w.lens + hlit, hdist)) { s.detail = RCZ_FL_INVALID_HUFFMAN_TREE; return F_INVALID; }
    return codes(s, w, lane);
}

__global__ void __launch_bounds__(NT)
inflate_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
               uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint64_t* __restrict__ out_cap,
               uint64_t* __restrict__ out_len, uint64_t* __restrict__ in_used, int32_t* __restrict__ status, int32_t* __restrict__ detail,
               unsigned nstreams, unsigned zlib_mode) {
    RCZ_DYN_SMEM(raw);
    const unsigned lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    WarpSmem& w = reinterpret_cast<WarpSmem*>(raw)[wi];
    for (unsigned sidx = blockIdx.x * WPB + wi; sidx < nstreams; sidx += gridDim.x * WPB) {
        St s;
        const unsigned long long n = in_len[sidx];
        s.br.init(in_base + in_off[sidx], n);
        s.out = out_base + out_off[sidx];
        { const unsigned long long c64 = out_cap[sidx]; s.cap = c64 > 0x7fffffffull ? 0x7fffffffu : (unsigned)c64; }
        s.o = 0; s.detail = 0;
        int r = F_OK;
        bool empty_block = false;
        for (;;) {                                                                   // flate.rs:195-206 block
            const unsigned before = s.o;
            s.br.refill();
            const unsigned bfinal = s.br.take(1);
            if (s.br.eof()) { r = F_EOF; break; }
            const unsigned type = s.br.take(2);
            if (s.br.eof()) { r = F_EOF; break; }
            if (type == 0) r = stored(s, lane);
            else if (type == 1) r = fixed_block(s, w, lane);
            else if (type == 2) r = dynamic_block(s, w, lane);
            else { s.detail = RCZ_FL_INVALID_BLOCK_CODE; r = F_INVALID; }
            if (r != F_OK || bfinal) { empty_block = r == F_OK && s.o == before; break; }
            // zlib.rs:106-109: under zlib::Decoder a block of zero bytes makes flate's read() return Ok(0) and ends the stream there
            if (zlib_mode && s.o == before) { empty_block = true; break; }
        }
        __syncwarp();
        if (lane == 0) {
            out_len[sidx] = s.o;
            status[sidx] = r == F_OK ? RCZ_OK : r == F_EOF ? RCZ_E_UNEXPECTED_EOF : r == F_INVALID ? RCZ_E_INVALID_INPUT
                           : r == F_MALFORMED ? RCZ_E_MALFORMED : RCZ_E_OUTPUT_FULL;
            if (detail) detail[sidx] = r == F_INVALID ? s.detail : 0;
            if (in_used) {
                const unsigned long long bp = s.br.bytepos();
                in_used[sidx] = ((r == F_EOF || bp > n) ? n : bp) | ((zlib_mode && empty_block) ? ZL_EMPTY_BLOCK : 0ull);
            }
        }
    }
}

}  // namespace flk

extern "C" int rcz_flate_decode_streams(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                        const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, uint64_t* in_used,
                                        int32_t* status, int32_t* detail, size_t n, int mem_kind) {
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (n == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !status || n > 0x7fffffffu) return RCZ_E_ARG;
    rt_set_device(c->device);
    if (!rcz_spans_ok(in_off, in_len, n) || !rcz_spans_ok(out_off, out_cap, n)) return RCZ_E_ARG;
    if (mem_kind == RCZ_MEM_HOST) {                                            // big host batches: chunks of ~1 GiB of output (>= 16 K streams of 64 KiB fill the GPU)
        bool handled = false;
        const int st = host_chunked(c, n, 1ull << 30, in_base, in_off, in_len, 1, out_base, out_off, out_cap, 1,
            [&](size_t b0, size_t nb, const uint8_t* din, uint8_t* dout) {
                return rcz_flate_decode_streams(c, din, in_off + b0, in_len + b0, dout, out_off + b0, out_cap + b0, out_len + b0, in_used ? in_used + b0 : nullptr,
                                                status + b0, detail ? detail + b0 : nullptr, nb, RCZ_MEM_DEVICE);
            },
            [&](size_t i) { return out_len[i] < out_cap[i] ? out_len[i] : out_cap[i]; }, &handled);
        if (st || handled) return st;
    }
    DescStager ds(c, mem_kind, n);
    ds.add_in(in_off, n * 8); ds.add_in(in_len, n * 8); ds.add_in(out_off, n * 8); ds.add_in(out_cap, n * 8);
    ds.add_out(out_len, n * 8); ds.add_out(status, n * 4);
    const size_t o_used = ds.add_out(in_used, in_used ? n * 8 : 0);
    const size_t o_det = ds.add_out(detail, detail ? n * 4 : 0);
    int st = ds.upload(); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, n, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, n, 1, &dout); if (st) return st;
    }
    const size_t smem = sizeof(flk::WarpSmem) * flk::WPB;
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(flk::inflate_kernel, smem));
    const unsigned grid = (unsigned)std::min<size_t>((n + flk::WPB - 1) / flk::WPB, (size_t)c->sm_count * 12);
    st = ctx_timer_begin(c); if (st) return st;
    RCZ_KLAUNCH(c, flk::inflate_kernel, grid, flk::NT, smem, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2),
                ds.in_ptr<uint64_t>(3), ds.out_ptr<uint64_t>(0), in_used ? ds.out_ptr<uint64_t>(o_used) : (uint64_t*)nullptr,
                ds.out_ptr<int32_t>(1), detail ? ds.out_ptr<int32_t>(o_det) : (int32_t*)nullptr, (unsigned)n, 0u);
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> clipped(n);
        for (size_t i = 0; i < n; ++i) clipped[i] = out_len[i] < out_cap[i] ? out_len[i] : out_cap[i];
        st = unstage_span_out(c, out_base, dout, out_off, clipped.data(), n, 1); if (st) return st;
    }
    return RCZ_OK;
}

// ======================================================================================================
// zlib wrapper: zlib.rs:55-117 header checks + the inflate kernel above + Adler-32
// (checksum/adler.rs:34-44) of the output; the big-endian trailer is compared exactly where `zlib::Decoder::read` compares it.
// ======================================================================================================
namespace zlk {

constexpr unsigned MOD_ADLER = 65521u;   // adler.rs:18
constexpr int NT = 256, WPB = NT / 32;

// Adler-32 of p[0..n) by one warp.  a = 1 + sum(b_i), b = n + sum((n - i) * b_i) (mod 65521): the serial recurrence of
// adler.rs:34-39 in closed form, so pieces of <= 4096 bytes reduce independently (warp-shuffle sums) and fold with their
// offset.  The body runs on aligned 16-byte loads and DP4A (byte sums and index-weighted byte sums, four bytes per instruction).
__device__ __forceinline__ void adler_fold(unsigned long long& S1, unsigned long long& S2, unsigned long long g, unsigned s1, unsigned s2) {
    s1 = warp_reduce_add(s1); s2 = warp_reduce_add(s2);               // <= 4096 * 255 and < 2^31.1: no overflow
    S2 = (S2 + (g % MOD_ADLER) * (unsigned long long)(s1 % MOD_ADLER) + s2) % MOD_ADLER;
    S1 = (S1 + s1) % MOD_ADLER;
}
__device__ __forceinline__ unsigned warp_adler32(const uint8_t* __restrict__ p, unsigned long long n) {
    const unsigned lane = threadIdx.x & 31;
    unsigned long long S1 = 0, S2 = 0, g = 0;                         // sum b_i, sum i * b_i (mod 65521); bytes done
    {   // head: up to the first 16-byte boundary
        const unsigned long long h64 = (16u - (unsigned)((uintptr_t)p & 15u)) & 15u;
        const unsigned h = (unsigned)(h64 < n ? h64 : n);
        unsigned s1 = 0, s2 = 0;
        if (lane < h) { s1 = p[lane]; s2 = lane * s1; }
        adler_fold(S1, S2, 0, s1, s2);
        g = h;
    }
    while (n - g >= 16) {                                             // body: pieces of <= 4096 bytes, 16 per lane and round
        const unsigned long long rest = n - g;
        const unsigned nv = rest >= 4096 ? 256u : (unsigned)(rest >> 4);
        const uint4* v = reinterpret_cast<const uint4*>(p + g);
        unsigned s1 = 0, s2 = 0;
        for (unsigned k = lane; k < nv; k += 32) {
            const uint4 q = __ldg(v + k);
            const unsigned a0 = __dp4a(q.x, 0x01010101u, 0u), a1 = __dp4a(q.y, 0x01010101u, 0u), a2 = __dp4a(q.z, 0x01010101u, 0u), a3 = __dp4a(q.w, 0x01010101u, 0u);
            unsigned w = __dp4a(q.x, 0x03020100u, 0u);
            w = __dp4a(q.y, 0x07060504u, w); w = __dp4a(q.z, 0x0b0a0908u, w); w = __dp4a(q.w, 0x0f0e0d0cu, w);
            const unsigned sv = a0 + a1 + a2 + a3;
            s1 += sv; s2 += 16u * k * sv + w;
        }
        adler_fold(S1, S2, g, s1, s2);
        g += 16ull * nv;
    }
    if (g < n) {                                                      // tail: < 16 bytes
        const unsigned t = (unsigned)(n - g);
        unsigned s1 = 0, s2 = 0;
        if (lane < t) { s1 = p[g + lane]; s2 = lane * s1; }
        adler_fold(S1, S2, g, s1, s2);
    }
    const unsigned long long nm = n % MOD_ADLER;
    const unsigned a = (unsigned)((1 + S1) % MOD_ADLER);
    const unsigned b = (unsigned)((nm + nm * S1 + MOD_ADLER - S2) % MOD_ADLER);
    return (b << 16) | a;                                             // adler.rs:42-44
}

__global__ void __launch_bounds__(NT)
adler32_kernel(const uint8_t* __restrict__ base, const uint64_t* __restrict__ off, const uint64_t* __restrict__ len, uint32_t* __restrict__ out, unsigned n) {
    const unsigned lane = threadIdx.x & 31;
    for (unsigned i = blockIdx.x * WPB + (threadIdx.x >> 5); i < n; i += gridDim.x * WPB) {
        const unsigned v = warp_adler32(base + off[i], len[i]);
        if (lane == 0) out[i] = v;
    }
}

// one warp per stream, after inflate ran on bytes [2, n): header checks first (zlib.rs:55-84), then inflate's own outcome, then
// the trailer where the reference reads it (zlib.rs:106-117: only after a block of zero bytes)
__global__ void __launch_bounds__(NT)
zlib_finish_kernel(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ in_len,
                   const uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, uint64_t* __restrict__ out_len,
                   const uint64_t* __restrict__ used_fl, uint64_t* __restrict__ in_used, int32_t* __restrict__ status, int32_t* __restrict__ detail,
                   uint32_t* __restrict__ adler, unsigned n) {
    const unsigned lane = threadIdx.x & 31;
    for (unsigned i = blockIdx.x * WPB + (threadIdx.x >> 5); i < n; i += gridDim.x * WPB) {
        const uint8_t* in = in_base + in_off[i];
        const unsigned long long len = in_len[i];
        int st = status[i], det = detail[i];
        unsigned long long used = 0, olen = out_len[i];
        unsigned ad = 1;
        if (len < 2) { st = RCZ_E_UNEXPECTED_EOF; det = RCZ_FL_NONE; olen = 0; used = 0; }      // read_u8 on an empty reader
        else {
            const unsigned cmf = in[0], flg = in[1];
            int zd = 0;
            if ((cmf & 0xf) != 0x8) zd = RCZ_ZL_UNSUPPORTED_FORMAT;
            else if ((cmf & 0xf0) != 0x70) zd = RCZ_ZL_UNSUPPORTED_WINDOW;
            else if (flg & 0x20) zd = RCZ_ZL_PRESET_DICTIONARY;
            else if (((cmf << 8) + flg) % 31 != 0) zd = RCZ_ZL_BAD_HEADER_CHECKSUM;
            if (zd) { st = RCZ_E_INVALID_INPUT; det = zd; olen = 0; used = 2; }
            else {
                used = 2 + (used_fl[i] & ~flk::ZL_EMPTY_BLOCK);
                if (st == RCZ_OK) {
                    ad = warp_adler32(out_base + out_off[i], olen);
                    // the reference reads the trailer only when flate's read() returned Ok(0) on a block of zero bytes (zlib.rs:106-109);
                    // after a non-empty final block `inner.eof()` ends the stream first (zlib.rs:104-105) and the trailer is never looked at
                    if (!(used_fl[i] & flk::ZL_EMPTY_BLOCK)) {}
                    else if (len - used < 4) st = RCZ_E_UNEXPECTED_EOF;                             // read_u32::<BigEndian>
                    else {
                        const uint8_t* t = in + used;
                        const unsigned ck = ((unsigned)t[0] << 24) | ((unsigned)t[1] << 16) | ((unsigned)t[2] << 8) | t[3];
                        used += 4;
                        if (ck != ad) { st = RCZ_E_INVALID_INPUT; det = RCZ_ZL_BAD_CHECKSUM; }
                    }
                }
            }
        }
        if (lane == 0) { status[i] = st; detail[i] = det; out_len[i] = olen; if (in_used) in_used[i] = used; if (adler) adler[i] = ad; }
    }
}

}  // namespace zlk

extern "C" int rcz_adler32_streams(rcz_ctx* c, const void* base, const uint64_t* off, const uint64_t* len, uint32_t* adler, size_t n, int mem_kind) {
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (n == 0) return RCZ_OK;
    if (!base || !off || !len || !adler || n > 0x7fffffffu || !rcz_spans_ok(off, len, n)) return RCZ_E_ARG;
    rt_set_device(c->device);
    DescStager ds(c, mem_kind, n);
    ds.add_in(off, n * 8); ds.add_in(len, n * 8);
    ds.add_out(adler, n * 4);
    int st = ds.upload(); if (st) return st;
    const uint8_t* d = (const uint8_t*)base;
    if (mem_kind == RCZ_MEM_HOST) { st = stage_span_in(c, WS_IN, base, off, len, n, 1, &d); if (st) return st; }
    const unsigned grid = (unsigned)std::min<size_t>((n + zlk::WPB - 1) / zlk::WPB, (size_t)c->sm_count * 8);
    st = ctx_timer_begin(c); if (st) return st;
    RCZ_KLAUNCH(c, zlk::adler32_kernel, grid, zlk::NT, 0, d, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), ds.out_ptr<uint32_t>(0), (unsigned)n);
    st = ctx_timer_end(c); if (st) return st;
    return ds.download();
}

extern "C" int rcz_zlib_decode_streams(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                       const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, uint64_t* in_used,
                                       int32_t* status, int32_t* detail, uint32_t* adler, size_t n, int mem_kind) {
    if (!c || rcz_bad_kind(mem_kind)) return RCZ_E_ARG;
    if (n == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !out_cap || !out_len || !status || n > 0x7fffffffu) return RCZ_E_ARG;
    rt_set_device(c->device);
    if (!rcz_spans_ok(in_off, in_len, n) || !rcz_spans_ok(out_off, out_cap, n)) return RCZ_E_ARG;
    if (mem_kind == RCZ_MEM_HOST) {                                            // big host batches in pipelined chunks, as rcz_flate_decode_streams
        bool handled = false;
        const int st = host_chunked(c, n, 1ull << 30, in_base, in_off, in_len, 1, out_base, out_off, out_cap, 1,
            [&](size_t b0, size_t nb, const uint8_t* din, uint8_t* dout) {
                return rcz_zlib_decode_streams(c, din, in_off + b0, in_len + b0, dout, out_off + b0, out_cap + b0, out_len + b0, in_used ? in_used + b0 : nullptr,
                                               status + b0, detail ? detail + b0 : nullptr, adler ? adler + b0 : nullptr, nb, RCZ_MEM_DEVICE);
            },
            [&](size_t i) { return out_len[i] < out_cap[i] ? out_len[i] : out_cap[i]; }, &handled);
        if (st || handled) return st;
    }
    // the DEFLATE stream starts after the two header bytes (zlib.rs:55-57)
    std::vector<uint64_t> off2(n), len2(n);
    for (size_t i = 0; i < n; ++i) { const uint64_t h = in_len[i] < 2 ? in_len[i] : 2; off2[i] = in_off[i] + h; len2[i] = in_len[i] - h; }
    DescStager ds(c, mem_kind, n);
    ds.add_in(in_off, n * 8); ds.add_in(in_len, n * 8); ds.add_in(out_off, n * 8); ds.add_in(out_cap, n * 8);
    const size_t i_off2 = ds.add_in(off2.data(), n * 8), i_len2 = ds.add_in(len2.data(), n * 8);
    ds.add_out(out_len, n * 8); ds.add_out(status, n * 4);
    const size_t o_used = ds.add_out(in_used, in_used ? n * 8 : 0);
    const size_t o_det = ds.add_out(detail, detail ? n * 4 : 0);
    const size_t o_ad = ds.add_out(adler, adler ? n * 4 : 0);
    int st = ds.upload(); if (st) return st;
    // inflate's own bookkeeping, and the optional outputs the caller did not ask for
    void* scr; st = ctx_ws(c, WS_B, n * 16 + 64, &scr); if (st) return st;
    uint64_t* d_used_fl = (uint64_t*)scr;
    int32_t* d_det = detail ? ds.out_ptr<int32_t>(o_det) : (int32_t*)((uint8_t*)scr + n * 8);
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, n, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, n, 1, &dout); if (st) return st;
    }
    const size_t smem = sizeof(flk::WarpSmem) * flk::WPB;
    RCZ_CK(c, RCZ_KERNEL_SMEM_OPTIN(flk::inflate_kernel, smem));
    const unsigned grid = (unsigned)std::min<size_t>((n + flk::WPB - 1) / flk::WPB, (size_t)c->sm_count * 12);
    st = ctx_timer_begin(c); if (st) return st;
    RCZ_KLAUNCH(c, flk::inflate_kernel, grid, flk::NT, smem, din, ds.in_ptr<uint64_t>(i_off2), ds.in_ptr<uint64_t>(i_len2), dout, ds.in_ptr<uint64_t>(2),
                ds.in_ptr<uint64_t>(3), ds.out_ptr<uint64_t>(0), d_used_fl, ds.out_ptr<int32_t>(1), d_det, (unsigned)n, 1u);
    const unsigned grid2 = (unsigned)std::min<size_t>((n + zlk::WPB - 1) / zlk::WPB, (size_t)c->sm_count * 8);
    RCZ_KLAUNCH(c, zlk::zlib_finish_kernel, grid2, zlk::NT, 0, din, ds.in_ptr<uint64_t>(0), ds.in_ptr<uint64_t>(1), dout, ds.in_ptr<uint64_t>(2),
                ds.out_ptr<uint64_t>(0), d_used_fl, in_used ? ds.out_ptr<uint64_t>(o_used) : (uint64_t*)nullptr, ds.out_ptr<int32_t>(1), d_det,
                adler ? ds.out_ptr<uint32_t>(o_ad) : (uint32_t*)nullptr, (unsigned)n);
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) {
        std::vector<uint64_t> clipped(n);
        for (size_t i = 0; i < n; ++i) clipped[i] = out_len[i] < out_cap[i] ? out_len[i] : out_cap[i];
        st = unstage_span_out(c, out_base, dout, out_off, clipped.data(), n, 1); if (st) return st;
    }
    return RCZ_OK;
}
