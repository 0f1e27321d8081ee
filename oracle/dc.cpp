// oracle/dc.cpp — CPU restatement of /root/reference/src/bwt/dc.rs (+ MTF::encode, bwt/mtf.rs:63-79)
// (TEST INFRASTRUCTURE ONLY, see oracle.h)
#include "oracle.h"
#include <cstring>
#include <utility>
#include <vector>

namespace {
// bwt/mtf.rs:63-79  MTF::encode on a caller-owned 256-entry list (zero-initialised by MTF::new, mtf.rs:51-53)
inline int mtf_encode(uint8_t* symbols, uint8_t sym, unsigned& rank_out) {
    uint8_t next = symbols[0];
    if (next == sym) { rank_out = 0; return 0; }
    unsigned rank = 1;
    for (;;) {
        std::swap(symbols[rank], next);
        if (next == sym) break;
        ++rank;
        if (rank >= 256) return -1;                        // mtf.rs:75 assert!
    }
    symbols[0] = sym;
    rank_out = rank;
    return 0;
}
}  // namespace

// dc.rs:110-149 encode  +  dc.rs:88-104 EncodeIterator (distances + contexts in emission order)
// +  dc.rs:153-159 encode_simple's init[] export.
extern "C" int orc_dc_encode(const uint8_t* in, size_t n, uint32_t* init_out, uint32_t* dist_out, size_t* ndist,
                             uint8_t* ctx_sym, uint8_t* ctx_rank, uint32_t* ctx_limit) {
    uint8_t symbols[256]; memset(symbols, 0, sizeof symbols);   // MTF::new()
    std::vector<size_t> distances(n);
    size_t num_unique = 0;
    size_t last[256], init[256];
    for (int i = 0; i < 256; ++i) last[i] = init[i] = n;        // dc.rs:114-115
    const size_t filler = n;                                    // dc.rs:116
    for (size_t i = 0; i < n; ++i) {
        uint8_t sym = in[i];
        distances[i] = filler;
        size_t base = last[sym];
        last[sym] = i;
        if (base == n) {                                        // dc.rs:122-129 first occurrence
            size_t rank = num_unique;
            symbols[rank] = sym;
            unsigned r; if (mtf_encode(symbols, sym, r)) return ORC_E_MALFORMED;
            init[sym] = i;
            ++num_unique;
        } else {                                                // dc.rs:130-137
            unsigned rank; if (mtf_encode(symbols, sym, rank)) return ORC_E_MALFORMED;
            if (rank > 0) {
                if (!(i >= base + rank + 1)) return ORC_E_MALFORMED;
                distances[base] = i - base - rank - 1;
            }
        }
    }
    for (size_t rank = 0; rank < num_unique; ++rank) {          // dc.rs:139-144 final sweep
        uint8_t sym = symbols[rank];
        size_t base = last[sym];
        if (!(n >= base + rank + 1)) return ORC_E_MALFORMED;
        distances[base] = n - base - rank - 1;
    }
    for (int i = 0; i < 256; ++i) init_out[i] = (uint32_t)init[i];
    // dc.rs:88-104 EncodeIterator::next
    size_t pos[256]; for (int i = 0; i < 256; ++i) pos[i] = init[i];
    size_t last_active = 0, k = 0;
    for (size_t i = 0; i < n; ++i) {
        if (distances[i] == filler) continue;                   // dc.rs:94 find(d != filler)
        uint8_t sym = in[i];
        size_t rank = last_active - pos[sym];
        if (!(rank < 256)) return ORC_E_MALFORMED;              // dc.rs:96 assert!
        last_active = i + 1;
        pos[sym] = i + 1 + distances[i];
        dist_out[k] = (uint32_t)distances[i];
        if (ctx_sym) ctx_sym[k] = sym;
        if (ctx_rank) ctx_rank[k] = (uint8_t)rank;
        if (ctx_limit) ctx_limit[k] = (uint32_t)(n - i);        // dc.rs:101
        ++k;
    }
    *ndist = k;
    return ORC_OK;
}

// dc.rs:162-233 decode (with decode_simple's distance feed, dc.rs:236-252)
extern "C" int orc_dc_decode(size_t n, const uint32_t* init, const uint32_t* dist, size_t ndist, uint8_t* out,
                             size_t* used, uint8_t* ctx_sym, uint8_t* ctx_rank, uint32_t* ctx_limit) {
    uint8_t symbols[256]; memset(symbols, 0, sizeof symbols);   // MTF::new()
    size_t next[256]; for (int i = 0; i < 256; ++i) next[i] = init[i];
    size_t i = 0, di = 0;
    if (used) *used = 0;
    for (int sym = 0; sym < 256; ++sym) {                       // dc.rs:169-179 insertion sort by first position
        size_t d = next[sym];
        if (d < n) {
            size_t j = i;
            while (j > 0 && next[symbols[j - 1]] > d) { symbols[j] = symbols[j - 1]; --j; }
            symbols[j] = (uint8_t)sym;
            ++i;
        }
    }
    if (i <= 1) {                                               // dc.rs:180-187
        memset(out, symbols[0], n);
        return ORC_OK;
    }
    const size_t alphabet_size = i;
    uint8_t ranks[256]; memset(ranks, 0, sizeof ranks);         // dc.rs:190-196
    i = 0;
    while (i < n) {                                             // dc.rs:199-229
        uint8_t sym = symbols[0];
        size_t stop = next[symbols[1]];
        if (stop > n) return ORC_E_MALFORMED;                   // output[i] index panic
        while (i < stop) out[i++] = sym;
        if (ctx_sym) ctx_sym[di] = sym;
        if (ctx_rank) ctx_rank[di] = ranks[sym];
        if (ctx_limit) ctx_limit[di] = (uint32_t)(n + 1 - i);
        if (di >= ndist) return ORC_E_UNEXPECTED_EOF;           // dc.rs:245-246
        size_t future = stop + dist[di++];
        if (!(future <= n)) return ORC_E_MALFORMED;             // dc.rs:213 assert!
        size_t rank = 1;
        while (rank < alphabet_size && future + rank > next[symbols[rank]]) {   // dc.rs:215-218
            symbols[rank - 1] = symbols[rank];
            ++rank;
        }
        symbols[rank - 1] = sym;
        next[sym] = future + rank - 1;
        ranks[sym] = (uint8_t)(rank - 1);
    }
    if (used) *used = di;
    for (int s = 0; s < 256; ++s) {                             // dc.rs:230 assert_eq!
        size_t d = next[s];
        if (d < n || d >= n + alphabet_size) {
            // symbols absent from the block keep init == n, which is inside [n, n+alphabet)
            return ORC_E_MALFORMED;
        }
    }
    if (i != n) return ORC_E_MALFORMED;                         // dc.rs:231
    return ORC_OK;
}
