/*
 * rcz.h — C ABI of librcz.so: B200 (sm_100a) kernels for the hot inner loops of the Rust crate
 * `compress` 0.2.1 (rusty-shell/rust-compress).
 *
 * The crate exposes no FFI; its drop-in boundary is the Rust API (`Decoder<R>: Read`,
 * `Encoder<W>: Write`, and the per-block free functions).  Each entry point below replaces the
 * per-block call that one of those types makes, batched over independent blocks/streams so that a
 * host shim (rust/ shim in INTEGRATION.md, the C++ mirrors in rust-compress_b200/host/rcz_stream.hpp,
 * or the Python mirrors in rust-compress_b200/ python modules) can parse framing on the CPU, collect block
 * descriptors, and hand the batch to the GPU.  Paths cited are relative to the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; no global state; one rcz_ctx per host thread, bound to one device
 *     and one CUDA stream.
 *   - every call returns an rcz_status; per-block outcomes go to status[i] (same codes).
 *   - mem_kind selects where the buffers live:
 *       RCZ_MEM_HOST          data + descriptor arrays in host memory; the library stages H2D/D2H and
 *                             returns when the results are in the caller's host buffers.  Big batches (lz4, bwt,
 *                             flate, zlib) are cut into chunks of consecutive units whose uploads, kernels and
 *                             downloads overlap; page-locked buffers (rcz_host_alloc) are what makes them overlap.
 *       RCZ_MEM_DEVICE        data pointers are device pointers; descriptor/result arrays
 *                             (offsets, lengths, status) are host arrays; returns when results are valid.
 *       RCZ_MEM_DEVICE_ASYNC  data pointers AND result arrays (out_len, status, origin-out, in_used, detail)
 *                             are device pointers; input descriptor arrays (offsets, lengths, origin-in)
 *                             stay host arrays (the host sizes the grids from them).  The call only
 *                             enqueues work on the context's stream (pipelines, benchmarks).
 *   - device data buffers must come from cudaMalloc-class allocators (the kernels may read up to the
 *     next 16-byte boundary past a block, which always stays inside such an allocation).
 *   - block/stream sizes are limited to < 2 GiB each: in_len[i], out_cap[i], n[i] < 2^31 and offsets < 2^62, else the call
 *     returns RCZ_E_ARG (no "unbounded" sentinels).  BWT blocks are limited further, see rcz_bwt_decode_blocks.
 *   - a unit whose status[i] != RCZ_OK: lz4, bwt and the bwt -> dc -> ari pipeline deliver nothing of it and report
 *     out_len[i] == 0; flate, zlib, rle, ari and mtf deliver the out_len[i] bytes decoded before the error (the reference's
 *     Read impls hand out what they have before returning Err); rcz_dc_encode_blocks reports the size it would have
 *     needed when status[i] == RCZ_E_OUTPUT_FULL.
 *   - only out_len[i] bytes of a unit's output region are meaningful; the rest of the region (up to out_cap[i]) is
 *     unspecified after the call (the pipelined host path of lz4 copies whole regions down without waiting for the lengths).
 */
#ifndef RCZ_H
#define RCZ_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rcz_ctx rcz_ctx;

typedef enum rcz_status {
    RCZ_OK = 0,
    RCZ_E_INVALID_INPUT = -1,  /* io::ErrorKind::InvalidInput (lz4.rs:365-377, flate.rs:53-66)          */
    RCZ_E_UNEXPECTED_EOF = -2, /* "unexpected end of file" (lib.rs:53-62,109-125) / raw UnexpectedEof    */
    RCZ_E_OVERLONG_RUN = -3,   /* rle.rs:151-154 "Overly long run"                                       */
    RCZ_E_MALFORMED = -4,      /* input on which the reference panics (index OOB / assert!)              */
    RCZ_E_OUTPUT_FULL = -5,    /* caller-provided out_cap too small                                      */
    RCZ_E_ARG = -6,            /* bad argument (null pointer, size >= 2 GiB, unknown mem_kind)           */
    RCZ_E_CUDA = -7,           /* CUDA runtime failure; see rcz_last_error()                             */
    RCZ_E_NO_DEVICE = -8,      /* no CUDA device: there is NO CPU fallback                               */
    RCZ_E_UNSUPPORTED = -9
} rcz_status;

/* flate detail codes (flate.rs:42-51), delivered in detail[i] when status[i] == RCZ_E_INVALID_INPUT */
enum {
    RCZ_FL_NONE = 0,
    RCZ_FL_HUFFMAN_TREE_TOO_LARGE = 1,
    RCZ_FL_INVALID_BLOCK_CODE = 2,
    RCZ_FL_INVALID_HUFFMAN_HEADER_SYMBOL = 3,
    RCZ_FL_INVALID_HUFFMAN_TREE = 4,
    RCZ_FL_INVALID_HUFFMAN_TREE_HEADER = 5,
    RCZ_FL_INVALID_HUFFMAN_CODE = 6,
    RCZ_FL_INVALID_STATIC_SIZE = 7,
    RCZ_FL_NOT_ENOUGH_BITS = 8,
    /* zlib wrapper (zlib.rs:55-117), all InvalidInput */
    RCZ_ZL_UNSUPPORTED_FORMAT = 16,   /* "unsupported zlib stream format"  zlib.rs:58-63 */
    RCZ_ZL_UNSUPPORTED_WINDOW = 17,   /* "unsupported zlib window size"    zlib.rs:65-70 */
    RCZ_ZL_PRESET_DICTIONARY = 18,    /* "unsupported initial dictionary"  zlib.rs:72-77 */
    RCZ_ZL_BAD_HEADER_CHECKSUM = 19,  /* "invalid zlib header checksum"    zlib.rs:79-84 */
    RCZ_ZL_BAD_CHECKSUM = 20          /* "invalid checksum on zlib stream" zlib.rs:108-113 */
};

typedef enum rcz_mem_kind { RCZ_MEM_HOST = 0, RCZ_MEM_DEVICE = 1, RCZ_MEM_DEVICE_ASYNC = 2 } rcz_mem_kind;

/* ---------------- context ---------------- */
int rcz_ctx_create(int device, unsigned flags, rcz_ctx** out);
int rcz_ctx_destroy(rcz_ctx* ctx);
/* use an externally owned CUDA stream (cudaStream_t as void*), e.g. torch.cuda.current_stream().cuda_stream */
int rcz_ctx_set_stream(rcz_ctx* ctx, void* cuda_stream);
int rcz_ctx_sync(rcz_ctx* ctx);
const char* rcz_strerror(int status);
const char* rcz_last_error(rcz_ctx* ctx);
/* number of librcz kernels launched through this context so far (bench.py "gpu_launches") */
uint64_t rcz_kernel_launches(rcz_ctx* ctx);
/* device time of the kernels of the most recent batch call, from CUDA events on the context's stream
 * (valid after rcz_ctx_sync or a synchronous call) */
float rcz_last_kernel_ms(rcz_ctx* ctx);
/* per-kernel split of the same batch call, for ops that launch several kernels in a row (lz4: parse, scan, materialise):
 * writes up to `cap` durations in launch order and returns how many stages the call had (0 when it was not timed) */
int rcz_last_stage_ms(rcz_ctx* ctx, float* ms, int cap);
/* pinned host memory for staging (cudaMallocHost) */
int rcz_host_alloc(void** p, size_t bytes);
int rcz_host_free(void* p);
const char* rcz_build_info(void);

/* ---------------- lz4.rs ----------------
 * rcz_lz4_decode_blocks replaces `BlockDecoder::decode` / `lz4::decode_block` (lz4.rs:64-162, 602-611),
 * called once per compressed frame block at lz4.rs:447-455.  Block i is in_base[in_off[i] .. +in_len[i]];
 * its bytes are written to out_base[out_off[i] .. +out_len[i]], out_len[i] <= out_cap[i]. */
int rcz_lz4_decode_blocks(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                          void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                          uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind);
/* Multi-GPU form of rcz_lz4_decode_blocks (device pointers only): the same decode, and every output byte written at out_base + x is
 * also written at peer_out_base[p] + x for p < npeers (<= 7) — pointers into the other GPUs' output buffers, mapped into this process
 * over NVLink / NVSwitch (CUDA IPC or symmetric memory).  With one rank per GPU and every rank passing its own slot of a common layout,
 * the final gather of the decoded shards is done by the decode kernel itself (bulk stores of every finished output tile);
 * the caller synchronises the ranks afterwards (stream sync + barrier) before reading.  Every peer_out_base[p] - out_base must be a
 * multiple of 16 (the peers' chunks are stored as aligned as the local ones), else RCZ_E_ARG. */
int rcz_lz4_decode_blocks_gather(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                 void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                                 uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind,
                                 void* const* peer_out_base, int npeers);
/* rcz_lz4_encode_blocks replaces `BlockEncoder::encode` / `lz4::encode_block` (lz4.rs:183-311, 616-627): the reference's greedy
 * single-probe hash compressor, byte for byte (same probe sequence, skip acceleration, rewind, `pos + 12 > len` tail rule).
 * out_cap[i] must be >= rcz_lz4_compression_bound(in_len[i]) (the reference reserves exactly that, lz4.rs:232-238), else
 * status[i] = RCZ_E_OUTPUT_FULL.  in_len[i] > 0x7e000000 yields out_len[i] = 0 like the reference (lz4.rs:229-230). */
int rcz_lz4_encode_blocks(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                          void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                          uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind);
/* `lz4::compression_bound` (lz4.rs:175-181); -1 == None */
int64_t rcz_lz4_compression_bound(uint32_t size);

/* ---------------- bwt/mod.rs ----------------
 * rcz_bwt_decode_blocks replaces `compute_inversion_table` + `InverseIterator` (bwt/mod.rs:223-282) as
 * driven by the stream decoder at bwt/mod.rs:388-393: block i = L column in_base[in_off[i] .. +n[i]] with
 * origin[i]; output written to out_base[out_off[i] ..], out_len[i] bytes (== n[i] for a well-formed block).
 * Limit: n[i] <= 16,777,214 (the link table packs a 24-bit position with the symbol); larger blocks return
 * RCZ_E_UNSUPPORTED in status[i] (the reference accepts any u32; its own docs suggest 4 MiB, bwt/mod.rs:32).
 * rcz_bwt_encode_blocks replaces `compute_suffixes` + `TransformIterator` (bwt/mod.rs:136-204) as called
 * at bwt/mod.rs:470-475: writes the L column (n[i] bytes) and origin[i]. */
int rcz_bwt_decode_blocks(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* n,
                          const uint32_t* origin, void* out_base, const uint64_t* out_off,
                          uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind);
int rcz_bwt_encode_blocks(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* n,
                          void* out_base, const uint64_t* out_off, uint32_t* origin, int32_t* status,
                          size_t nblocks, int mem_kind);

/* ---------------- flate.rs ----------------
 * rcz_flate_decode_streams replaces `Decoder::block` + `codes` + `HuffmanTree::decode`
 * (flate.rs:129-146, 195-206, 262-450) driven to BFINAL for every independent raw-DEFLATE stream.
 * in_used[i] (optional) = bytes of the stream consumed; detail[i] (optional) = RCZ_FL_* code. */
int rcz_flate_decode_streams(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                             void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                             uint64_t* out_len, uint64_t* in_used, int32_t* status, int32_t* detail,
                             size_t nstreams, int mem_kind);

/* ---------------- zlib.rs + checksum/adler.rs ----------------
 * rcz_zlib_decode_streams replaces `zlib::Decoder` driven by read_to_end (zlib.rs:55-117): header checks, the DEFLATE
 * blocks (the inflate kernel), Adler-32 of the output (adler.rs:34-44).  As in the reference the big-endian trailer is read
 * and compared ONLY when a DEFLATE block decodes to zero bytes (flate's read() returns Ok(0), zlib.rs:106-109) — the stream
 * ends there; after a non-empty final block `inner.eof()` ends the stream first (zlib.rs:104-105) and a missing or wrong
 * trailer goes unnoticed.  Callers that want the check compare adler[i] with the trailer themselves.
 * detail[i] (optional) = RCZ_ZL_* for the wrapper's own errors, RCZ_FL_* for inflate's; in_used[i] (optional) = bytes
 * consumed (header, DEFLATE bytes, and the 4 trailer bytes when they were read); adler[i] (optional) = checksum of the
 * decoded bytes.
 * rcz_adler32_streams is `checksum::adler::State32::{feed,result}` over independent byte ranges. */
int rcz_zlib_decode_streams(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                            void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                            uint64_t* out_len, uint64_t* in_used, int32_t* status, int32_t* detail, uint32_t* adler,
                            size_t nstreams, int mem_kind);
int rcz_adler32_streams(rcz_ctx* ctx, const void* base, const uint64_t* off, const uint64_t* len, uint32_t* adler,
                        size_t nstreams, int mem_kind);

/* ---------------- entropy/ari ----------------
 * rcz_ari_encode_streams replaces `ByteEncoder::write` + `finish` (table.rs:203-219 -> ari/mod.rs:117-150,
 * 230-237); rcz_ari_decode_streams replaces `ByteDecoder::read` to the terminator (table.rs:255-272).
 * in_used[i] = bytes consumed including the ones only `finish()` pulls (ari/mod.rs:289-292). */
int rcz_ari_encode_streams(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                           void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, size_t nstreams, int mem_kind);
int rcz_ari_decode_streams(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                           void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, uint64_t* in_used, int32_t* status, size_t nstreams, int mem_kind);

/* ---------------- bwt/dc.rs ----------------
 * rcz_dc_encode_blocks replaces `dc::encode` + `EncodeIterator` (dc.rs:62-149): per block, init[256]
 * (u32, first position of each symbol or n) followed by the emitted distances (u32) are written to
 * out_base (as uint32) at element offset out_off[i]; out_len[i] = 256 + number of distances.
 * rcz_dc_decode_blocks replaces `dc::decode` (dc.rs:162-233) fed by `decode_simple`'s closure. */
int rcz_dc_encode_blocks(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* n,
                         uint32_t* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                         uint64_t* out_len, int32_t* status, size_t nblocks, int mem_kind);
int rcz_dc_decode_blocks(rcz_ctx* ctx, const uint32_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                         void* out_base, const uint64_t* out_off, const uint64_t* n, int32_t* status,
                         size_t nblocks, int mem_kind);

/* ---------------- bwt -> dc -> entropy::ari (BASELINE configs[4]; "BWT + DC + EC", bwt/mod.rs:11-14) ----------------
 * The three stage calls above chained on the device: rcz_bwt_encode_blocks (bwt/mod.rs:136-204), rcz_dc_encode_blocks
 * (dc.rs:62-159), then `ByteEncoder` (table.rs:203-219) over the dc output serialised as init[256] followed by the distances,
 * each a u32 LE; the decoder runs table.rs:255-272, dc.rs:162-252 and bwt/mod.rs:223-294.  No stage result visits the host.
 * The reference has no wire format for this chain (dc is in-memory only), so the container is this library's:
 *   u32 LE x 6: magic "BDA1", n, origin, nsym (256 + number of distances), ari_chunk, nstreams; nstreams x u32 LE code
 *   length; then the code bytes of the streams back to back.
 * ari_chunk = 0 codes the serialised bytes of a block as ONE ByteEncoder stream; ari_chunk = k (multiple of 4, >= 1024)
 * cuts them into independent ByteEncoder streams of k bytes (the range coder is one serial chain per stream).
 * encode: block i = in_base[in_off[i] .. +n[i]] -> container at out_base[out_off[i] ..], out_len[i] <= out_cap[i] bytes;
 *         origin[i] (optional) = the BWT origin.  n[i] <= 16,777,214 (24-bit positions of the inverse transform).
 * decode: container i = in_base[in_off[i] .. +in_len[i]]; n[i] = the block's decoded size (the caller's framing knows it, as
 *         bwt/mod.rs:463 writes it in front of every block); a container whose header disagrees is RCZ_E_INVALID_INPUT.
 * rcz_last_stage_ms: encode = {bwt, dc, ari, pack}, decode = {ari, dc, bwt}. */
int rcz_bwt_dc_ari_encode_blocks(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* n,
                                 void* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                                 uint32_t* origin, int32_t* status, size_t nblocks, uint32_t ari_chunk, int mem_kind);
int rcz_bwt_dc_ari_decode_blocks(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                 void* out_base, const uint64_t* out_off, const uint64_t* n, uint64_t* out_len,
                                 int32_t* status, size_t nblocks, uint32_t ari_chunk, int mem_kind);

/* ---------------- bwt/mtf.rs ----------------
 * rcz_mtf_encode_streams replaces `mtf::Encoder<W>::write` (mtf.rs:118-125: `MTF::encode` per byte, alphabetical start
 * list), rcz_mtf_decode_streams replaces `mtf::Decoder<R>::read` (mtf.rs:155-168: `MTF::decode` per byte).  One
 * rank/symbol per input byte: out_len[i] == in_len[i]. */
int rcz_mtf_encode_streams(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                           void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, size_t nstreams, int mem_kind);
int rcz_mtf_decode_streams(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                           void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, size_t nstreams, int mem_kind);

/* ---------------- rle.rs ----------------
 * rcz_rle_decode_streams replaces `Decoder::read_run` (rle.rs:212-259); rcz_rle_encode_streams replaces
 * `Encoder::write` + `finish` (rle.rs:62-122) for one whole-buffer write. */
int rcz_rle_decode_streams(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                           void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, size_t nstreams, int mem_kind);
int rcz_rle_encode_streams(rcz_ctx* ctx, const void* in_base, const uint64_t* in_off, const uint64_t* in_len,
                           void* out_base, const uint64_t* out_off, const uint64_t* out_cap,
                           uint64_t* out_len, int32_t* status, size_t nstreams, int mem_kind);

#ifdef __cplusplus
}
#endif
#endif /* RCZ_H */
