/*
 * oracle.h — CPU restatement of rusty-shell/rust-compress (crate `compress` 0.2.1) hot-path algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain C++ restatement of the reference's
 * Rust loops (file:line cited at every function).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (librcz.so) never links,
 * imports or falls back to anything in this directory.
 *
 * Pinning: the reference cannot be compiled here (no rustc/cargo), so the oracle is
 * pinned against the reference's own fixtures and inline known-answer vectors:
 *   - flate : src/data/test.z.0-9, test.z.go, test.large.z.5  (flate.rs:528-548)
 *   - lz4   : src/data/test.lz4.1-9                            (lz4.rs:647-659)
 *   - rle   : inline vectors                                   (rle.rs:320-352)
 *   - bwt   : suffix array is mathematically unique; pinned by definition + roundtrip (bwt/mod.rs:541-551)
 *   - dc/ari: the reference has roundtrip-only tests => encode-side bytes are "parity unpinned"
 *             (pinned here only by roundtrip, context equality, and SURVEY Appendix C vectors).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes — numerically identical to include/rcz.h */
enum {
    ORC_OK = 0,
    ORC_E_INVALID_INPUT = -1,   /* io::ErrorKind::InvalidInput */
    ORC_E_UNEXPECTED_EOF = -2,  /* ErrorKind::Other "unexpected end of file" / raw UnexpectedEof */
    ORC_E_OVERLONG_RUN = -3,    /* rle "Overly long run" */
    ORC_E_MALFORMED = -4,       /* reference would panic (OOB index, assert!) */
    ORC_E_OUTPUT_FULL = -5,     /* caller buffer too small (no reference analogue) */
    ORC_E_ARG = -6,
};

/* flate detail codes (flate.rs:42-51), reported through *detail when status == ORC_E_INVALID_INPUT */
enum {
    ORC_FL_NONE = 0,
    ORC_FL_HUFFMAN_TREE_TOO_LARGE = 1,
    ORC_FL_INVALID_BLOCK_CODE = 2,
    ORC_FL_INVALID_HUFFMAN_HEADER_SYMBOL = 3,
    ORC_FL_INVALID_HUFFMAN_TREE = 4,
    ORC_FL_INVALID_HUFFMAN_TREE_HEADER = 5,
    ORC_FL_INVALID_HUFFMAN_CODE = 6,
    ORC_FL_INVALID_STATIC_SIZE = 7,
    ORC_FL_NOT_ENOUGH_BITS = 8,
    /* zlib.rs:55-117 wrapper errors (all io::ErrorKind::InvalidInput) */
    ORC_ZL_UNSUPPORTED_FORMAT = 16,
    ORC_ZL_UNSUPPORTED_WINDOW = 17,
    ORC_ZL_PRESET_DICTIONARY = 18,
    ORC_ZL_BAD_HEADER_CHECKSUM = 19,
    ORC_ZL_BAD_CHECKSUM = 20,
};

/* ---- rle.rs ---- */
int orc_rle_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len);
int orc_rle_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len);

/* ---- lz4.rs ---- */
int orc_lz4_decode_block(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len);
int orc_lz4_encode_block(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len);
int64_t orc_lz4_compression_bound(uint32_t size); /* -1 == None */
int orc_lz4_frame_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                         size_t* consumed);
/* multi-threaded batch (CPU baseline for bench.py): decode nblocks independent blocks */
int orc_lz4_decode_blocks_mt(const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                             uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                             uint64_t* out_len, int32_t* status, size_t nblocks, int nthreads);

/* ---- bwt/mod.rs ---- */
int orc_bwt_suffixes(const uint8_t* in, size_t n, uint32_t* sa);
int orc_bwt_encode(const uint8_t* in, size_t n, uint8_t* out_l, uint32_t* origin);
int orc_bwt_inversion_table(const uint8_t* l, size_t n, size_t origin, uint32_t* table);
int orc_bwt_decode(const uint8_t* l, size_t n, size_t origin, uint8_t* out, size_t* out_len);
int orc_bwt_stream_encode(const uint8_t* in, size_t n, uint32_t block_size, uint8_t* out, size_t cap,
                          size_t* out_len);
int orc_bwt_stream_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len);
int orc_bwt_decode_blocks_mt(const uint8_t* l_base, const uint64_t* off, const uint64_t* n,
                             const uint32_t* origin, uint8_t* out_base, uint64_t* out_len,
                             int32_t* status, size_t nblocks, int nthreads);
int orc_bwt_encode_blocks_mt(const uint8_t* in_base, const uint64_t* off, const uint64_t* n,
                             uint8_t* out_base, uint32_t* origin, int32_t* status, size_t nblocks,
                             int nthreads);

/* ---- bwt/mtf.rs (MTF struct only) ---- */
void orc_mtf_encode(const uint8_t* in, size_t n, uint8_t* ranks);   /* alphabetical start (mtf.rs:100-109) */
void orc_mtf_decode(const uint8_t* ranks, size_t n, uint8_t* out);

/* ---- bwt/dc.rs ----
 * init[256]: first position of every symbol (n when absent).  dist[]: emitted distances in order,
 * ctx_*[]: per-distance Context (symbol, last_rank, distance_limit) — dc.rs:40-47 — may be NULL. */
int orc_dc_encode(const uint8_t* in, size_t n, uint32_t* init, uint32_t* dist, size_t* ndist,
                  uint8_t* ctx_sym, uint8_t* ctx_rank, uint32_t* ctx_limit);
int orc_dc_decode(size_t n, const uint32_t* init, const uint32_t* dist, size_t ndist, uint8_t* out,
                  size_t* used, uint8_t* ctx_sym, uint8_t* ctx_rank, uint32_t* ctx_limit);

/* ---- flate.rs ---- */
int orc_flate_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                     size_t* consumed, int* detail);
/* per-DEFLATE-block output sizes in block order (Reader contract: one block per refill, flate.rs:468-488) */
int orc_flate_decode_blocks(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                            size_t* consumed, int* detail, uint32_t* blk_sizes, size_t blk_cap,
                            size_t* nblk);
/* ---- zlib.rs (header + flate + Adler-32 trailer) ---- */
int orc_zlib_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len, size_t* consumed,
                    int* detail, uint32_t* adler);
int orc_flate_decode_streams_mt(const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                                uint64_t* out_len, int32_t* status, size_t nstreams, int nthreads);

/* ---- entropy/ari/{mod,table}.rs ---- */
int orc_ari_encode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len);
int orc_ari_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t* out_len,
                   size_t* consumed_read, size_t* consumed_finish);

/* ---- bwt -> dc -> ari composition (pipeline.cpp; the composition and container are this project's, the stages the reference's) ---- */
int orc_bda_encode_block(const uint8_t* in, size_t n, uint32_t chunk, uint8_t* out, size_t cap, size_t* out_len, uint32_t* origin);
int orc_bda_decode_block(const uint8_t* in, size_t in_len, size_t n, uint32_t chunk, uint8_t* out, size_t* out_len);
int orc_bda_encode_blocks_mt(const uint8_t* in_base, const uint64_t* in_off, const uint64_t* n, uint32_t chunk, uint8_t* out_base,
                             const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, uint32_t* origin, int32_t* status,
                             size_t nblocks, int nthreads);
int orc_bda_decode_blocks_mt(const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len, uint32_t chunk, uint8_t* out_base,
                             const uint64_t* out_off, const uint64_t* n, uint64_t* out_len, int32_t* status, size_t nblocks, int nthreads);

/* ---- checksum/adler.rs (used only to cross-check fixtures) ---- */
uint32_t orc_adler32(const uint8_t* in, size_t n);

#ifdef __cplusplus
}
#endif
