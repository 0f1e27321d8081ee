// rcz_stream.hpp — host-side mirrors of the `compress` crate's Reader/Writer surface on top of the librcz C ABI.
//
// The reference is Rust and no Rust toolchain exists in this image, so the host side is C++ (header only).  Every type
// keeps the reference's name, constructor arguments, method set and error behaviour for the hot path it wraps:
//
//   rcz::lz4::Decoder<R>        lz4.rs:316-500    new(r) read eof reset .r        frame parse on the host, blocks on the GPU
//   rcz::lz4::Encoder<W>        lz4.rs:505-597    new(w) write flush finish       (raw blocks only, like the reference stub)
//   rcz::lz4::decode_block      lz4.rs:602-611    compression_bound lz4.rs:175-181
//   rcz::bwt::Encoder<W>        bwt/mod.rs:437-518  new(w, block_size) write flush finish
//   rcz::bwt::Decoder<R>        bwt/mod.rs:321-432  new(r, extra_mem) read reset .r
//   rcz::bwt::encode_simple / decode_simple       bwt/mod.rs:213-219, 291-294
//   rcz::dc::encode_simple / decode_simple        bwt/dc.rs:153-159, 236-252
//   rcz::flate::Decoder<R>      flate.rs:164-488  new(r) read eof reset
//   rcz::ari::ByteEncoder<W> / ByteDecoder<R>     entropy/ari/table.rs:185-273  new write flush finish / new read finish
//   rcz::rle::Encoder<W> / Decoder<R>             rle.rs:40-123, 176-281
//
// A Reader is any type with `size_t read(uint8_t* dst, size_t len)` (0 == end of stream, like io::Read returning Ok(0));
// a Writer has `size_t write(const uint8_t*, size_t)` and `void flush()`.  Errors are thrown as rcz::io_error carrying the
// io::ErrorKind the reference would return; inputs on which the reference panics raise ErrorKind::Panic.
//
// Deliberate differences from the reference: block codecs read ahead — every block discoverable from
// the framing is collected (up to `batch_blocks`) and handed to the GPU in ONE C-ABI call, so the inner reader may be
// consumed further than the bytes handed out; `Write::write` returns the number of bytes consumed (the reference returns
// Ok(0), App. B #1); stream codecs without block framing (flate, ari, rle) take the whole inner stream.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rcz.h"

namespace rcz {

enum class ErrorKind { InvalidInput, Other, UnexpectedEof, Panic, OutputFull, Device };

struct io_error : std::runtime_error {
    ErrorKind kind; int status; int detail;
    io_error(ErrorKind k, const std::string& m, int st = 0, int det = 0) : std::runtime_error(m), kind(k), status(st), detail(det) {}
};

inline io_error error_from_status(int st, const char* what, int detail = 0) {
    switch (st) {
    case RCZ_E_INVALID_INPUT: return io_error(ErrorKind::InvalidInput, what, st, detail);
    case RCZ_E_UNEXPECTED_EOF: return io_error(ErrorKind::UnexpectedEof, "unexpected end of file", st);
    case RCZ_E_OVERLONG_RUN: return io_error(ErrorKind::Other, "Overly long run", st);
    case RCZ_E_MALFORMED: return io_error(ErrorKind::Panic, std::string(what) + ": malformed input (the reference panics)", st);
    case RCZ_E_OUTPUT_FULL: return io_error(ErrorKind::OutputFull, std::string(what) + ": output buffer too small", st);
    default: return io_error(ErrorKind::Device, std::string(what) + ": " + rcz_strerror(st), st);
    }
}

// ---------------------------------------------------------------- in-memory Reader / Writer (the reference's tests use these)
struct SliceReader {
    const uint8_t* p; size_t n, pos = 0;
    SliceReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
    explicit SliceReader(const std::vector<uint8_t>& v) : p(v.data()), n(v.size()) {}
    size_t read(uint8_t* dst, size_t len) { size_t k = len < n - pos ? len : n - pos; if (k) memcpy(dst, p + pos, k); pos += k; return k; }
};
struct VecReader {
    std::vector<uint8_t> v; size_t pos = 0;
    VecReader() = default;
    explicit VecReader(std::vector<uint8_t> v_) : v(std::move(v_)) {}
    size_t read(uint8_t* dst, size_t len) { size_t k = len < v.size() - pos ? len : v.size() - pos; if (k) memcpy(dst, v.data() + pos, k); pos += k; return k; }
};
struct VecWriter {
    std::vector<uint8_t> v;
    size_t write(const uint8_t* src, size_t len) { v.insert(v.end(), src, src + len); return len; }
    void flush() {}
};

// ---------------------------------------------------------------- context (one device, one stream)
class Context {
  public:
    explicit Context(int device = 0) {
        int st = rcz_ctx_create(device, 0, &h_);
        if (st != RCZ_OK) throw io_error(ErrorKind::Device, std::string("rcz_ctx_create: ") + rcz_strerror(st), st);   // no CPU fallback
    }
    ~Context() { if (h_) rcz_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    rcz_ctx* get() const { return h_; }
    void check(int st, const char* what) const { if (st != RCZ_OK) throw io_error(ErrorKind::Device, std::string(what) + ": " + rcz_strerror(st) + " " + rcz_last_error(h_), st); }

  private:
    rcz_ctx* h_ = nullptr;
};

namespace detail {
// byteorder read_u32::<LittleEndian>: raw UnexpectedEof when the stream ends inside the word
template <class R> inline bool try_read_exact(R& r, uint8_t* dst, size_t len) {
    size_t got = 0;
    while (got < len) { size_t k = r.read(dst + got, len - got); if (k == 0) return false; got += k; }
    return true;
}
template <class R> inline uint32_t read_u32_le(R& r) {
    uint8_t b[4];
    if (!try_read_exact(r, b, 4)) throw io_error(ErrorKind::UnexpectedEof, "failed to fill whole buffer", RCZ_E_UNEXPECTED_EOF);
    return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
}
// lib.rs:109-125 ReadExact::push_exactly
template <class R> inline void push_exactly(R& r, uint64_t bytes, std::vector<uint8_t>& buf) {
    size_t old = buf.size();
    buf.resize(old + (size_t)bytes);
    if (!try_read_exact(r, buf.data() + old, (size_t)bytes)) throw io_error(ErrorKind::Other, "unexpected end of file", RCZ_E_UNEXPECTED_EOF);
}
template <class W> inline void write_u32_le(W& w, uint32_t v) { uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; w.write(b, 4); }
template <class R> inline void read_to_end(R& r, std::vector<uint8_t>& buf) {
    for (;;) { size_t old = buf.size(); buf.resize(old + 65536); size_t k = r.read(buf.data() + old, 65536); buf.resize(old + k); if (k == 0) break; }
}
inline std::vector<uint64_t> prefix(const std::vector<uint64_t>& len, uint64_t align = 1) {
    std::vector<uint64_t> off(len.size()); uint64_t cur = 0;
    for (size_t i = 0; i < len.size(); ++i) { off[i] = cur; cur += (len[i] + align - 1) / align * align; }
    return off;
}
// hands decoded bytes out of a buffer, io::Read style
struct Outlet {
    std::vector<uint8_t> buf; size_t start = 0;
    size_t avail() const { return buf.size() - start; }
    size_t take(uint8_t* dst, size_t len) { size_t k = len < avail() ? len : avail(); if (k) memcpy(dst, buf.data() + start, k); start += k; return k; }
    void clear() { buf.clear(); start = 0; }
};
}  // namespace detail

// ================================================================================================ lz4
namespace lz4 {
constexpr uint32_t MAGIC = 0x184d2204;   // lz4.rs:40

inline int64_t compression_bound(uint32_t size) { return rcz_lz4_compression_bound(size); }   // < 0 == None

// lz4.rs:602-611 decode_block: appends to `output`, returns the number of bytes decoded
inline size_t decode_block(Context& ctx, const uint8_t* input, size_t n, std::vector<uint8_t>& output, size_t max_out = 0) {
    uint64_t off = 0, len = n, ooff = 0, cap = max_out ? max_out : 255ull * n + 64, olen = 0;
    int32_t st = 0;
    size_t old = output.size();
    output.resize(old + (size_t)cap);
    ctx.check(rcz_lz4_decode_blocks(ctx.get(), input, &off, &len, output.data() + old, &ooff, &cap, &olen, &st, 1, RCZ_MEM_HOST), "rcz_lz4_decode_blocks");
    output.resize(old + (st == RCZ_OK ? (size_t)olen : 0));
    if (st != RCZ_OK) throw error_from_status(st, "lz4::decode_block");
    return (size_t)olen;
}

template <class R> class Decoder {
  public:
    R r;                                   // public like the reference's field (lz4.rs:320)
    size_t batch_blocks = 64;              // read-ahead: compressed blocks collected per C-ABI call
    Decoder(Context& ctx, R r_) : r(std::move(r_)), ctx_(ctx) {}
    void reset() { header_ = false; eof_ = false; out_.clear(); pending_err_ = false; }   // lz4.rs:356-361
    bool eof() const { return eof_; }

    size_t read(uint8_t* dst, size_t len) {                                   // lz4.rs:470-500
        if (eof_) return 0;
        if (!header_) { read_header(); header_ = true; }
        size_t amt = len;
        while (amt > 0) {
            if (out_.avail() == 0) {
                if (end_seen_) { eof_ = true; break; }
                if (pending_err_) { pending_err_ = false; throw pending_; }
                fill();
                if (out_.avail() == 0 && end_seen_) { eof_ = true; break; }
                if (out_.avail() == 0 && pending_err_) { pending_err_ = false; throw pending_; }
            }
            size_t k = out_.take(dst + (len - amt), amt);
            amt -= k;
        }
        return len - amt;
    }

  private:
    void read_header() {                                                      // lz4.rs:363-420
        if (detail::read_u32_le(r) != MAGIC) throw io_error(ErrorKind::InvalidInput, "", RCZ_E_INVALID_INPUT);
        uint8_t bits[2] = {0, 0};
        size_t got = r.read(bits, 2);                                         // a single `read` in the reference: short reads leave zeros
        if (got == 1) r.read(bits + 1, 1);
        const uint8_t flg = bits[0], bd = bits[1];
        if ((flg >> 6) != 1) throw io_error(ErrorKind::InvalidInput, "", RCZ_E_INVALID_INPUT);
        blk_checksum_ = (flg & 0x10) != 0;
        const bool stream_size = (flg & 0x08) != 0;
        static const size_t MAX_SIZES[8] = {0, 0, 0, 0, 64u << 10, 256u << 10, 1u << 20, 4u << 20};
        max_block_size_ = MAX_SIZES[(bd >> 4) & 7];
        if (stream_size) { uint8_t sz[8]; if (!detail::try_read_exact(r, sz, 8)) throw io_error(ErrorKind::UnexpectedEof, "failed to fill whole buffer", RCZ_E_UNEXPECTED_EOF); }
        if (flg & 0x01) throw io_error(ErrorKind::Panic, "preset dictionaries not supported yet", RCZ_E_MALFORMED);   // assert! lz4.rs:407
        uint8_t cksum;
        if (!detail::try_read_exact(r, &cksum, 1)) throw io_error(ErrorKind::UnexpectedEof, "failed to fill whole buffer", RCZ_E_UNEXPECTED_EOF);
        end_seen_ = false;
    }
    // collect up to batch_blocks frame blocks (lz4.rs:422-464), decode the compressed ones in one call
    void fill() {
        out_.clear();
        std::vector<uint8_t> comp; std::vector<uint64_t> clen; std::vector<size_t> slot;   // compressed payloads
        struct Piece { bool raw; std::vector<uint8_t> bytes; size_t idx; };
        std::vector<Piece> pieces;
        try {
            while (pieces.size() < batch_blocks) {
                const uint32_t n = detail::read_u32_le(r);
                if (n == 0) { end_seen_ = true; break; }
                Piece p; p.raw = (n & 0x80000000u) != 0; p.idx = 0;
                if (p.raw) detail::push_exactly(r, n & 0x7fffffffu, p.bytes);
                else { p.idx = clen.size(); while (comp.size() % 16) comp.push_back(0); slot.push_back(comp.size()); detail::push_exactly(r, n, comp); clen.push_back(n); }
                pieces.push_back(std::move(p));
                if (blk_checksum_) (void)detail::read_u32_le(r);                 // read and ignored (lz4.rs:459-462)
            }
        } catch (const io_error& e) { pending_ = e; pending_err_ = true; }       // surfaces after the blocks before it are handed out
        std::vector<uint64_t> ioff(slot.begin(), slot.end()), ocap(clen.size()), olen(clen.size());
        std::vector<int32_t> st(clen.size());
        for (size_t i = 0; i < clen.size(); ++i) ocap[i] = max_block_size_ ? max_block_size_ : 255ull * clen[i] + 64;
        std::vector<uint64_t> ooff = detail::prefix(ocap, 16);
        std::vector<uint8_t> dec(clen.empty() ? 0 : (size_t)(ooff.back() + ocap.back()) + 64);
        comp.resize(comp.size() + 64);
        if (!clen.empty())
            ctx_.check(rcz_lz4_decode_blocks(ctx_.get(), comp.data(), ioff.data(), clen.data(), dec.data(), ooff.data(), ocap.data(), olen.data(),
                                             st.data(), clen.size(), RCZ_MEM_HOST), "rcz_lz4_decode_blocks");
        for (auto& p : pieces) {
            if (p.raw) { out_.buf.insert(out_.buf.end(), p.bytes.begin(), p.bytes.end()); continue; }
            if (st[p.idx] == RCZ_E_OUTPUT_FULL && ocap[p.idx] < 255ull * clen[p.idx] + 64) {
                // the BD size is only a reserve hint in the reference (lz4.rs:444-446; grow_output has no limit): a block that decodes to
                // more than it declares is decoded again on its own with the format's bound
                try { decode_block(ctx_, comp.data() + slot[p.idx], (size_t)clen[p.idx], out_.buf); continue; }
                catch (const io_error& e) { pending_ = e; pending_err_ = true; end_seen_ = false; break; }
            }
            if (st[p.idx] != RCZ_OK) { pending_ = error_from_status(st[p.idx], "lz4::Decoder"); pending_err_ = true; end_seen_ = false; break; }
            out_.buf.insert(out_.buf.end(), dec.begin() + (size_t)ooff[p.idx], dec.begin() + (size_t)(ooff[p.idx] + olen[p.idx]));
        }
    }
    Context& ctx_;
    detail::Outlet out_;
    bool header_ = false, eof_ = false, blk_checksum_ = false, end_seen_ = false, pending_err_ = false;
    size_t max_block_size_ = 0;
    io_error pending_{ErrorKind::Other, ""};
};

// lz4.rs:505-597: the reference's frame Encoder only ever emits raw blocks (`compress()` returns false, lz4.rs:543-545)
template <class W> class Encoder {
  public:
    explicit Encoder(W w) : w_(std::move(w)) {}
    size_t write(const uint8_t* buf, size_t len) {
        if (!wrote_header_) { detail::write_u32_le(w_, MAGIC); const uint8_t h[3] = {0x60, 0x50, 0}; w_.write(h, 3); wrote_header_ = true; }
        size_t done = 0;
        while (done < len) {
            size_t amt = limit_ - buf_.size() < len - done ? limit_ - buf_.size() : len - done;
            buf_.insert(buf_.end(), buf + done, buf + done + amt);
            if (buf_.size() == limit_) encode_block();
            done += amt;
        }
        return len;                                                           // the reference returns Ok(0) here (App. B #1)
    }
    void flush() { if (!buf_.empty()) encode_block(); w_.flush(); }
    W finish() { flush(); detail::write_u32_le(w_, 0); detail::write_u32_le(w_, 0); return std::move(w_); }   // two zero words, lz4.rs:550-561

  private:
    void encode_block() { detail::write_u32_le(w_, (uint32_t)buf_.size() | 0x80000000u); w_.write(buf_.data(), buf_.size()); buf_.clear(); }
    W w_; std::vector<uint8_t> buf_; bool wrote_header_ = false; size_t limit_ = 256 * 1024;
};

// lz4.rs:616-627 encode_block: appends the compressed block to `output`, returns its size (0 when the input is too large)
inline size_t encode_block(Context& ctx, const uint8_t* input, size_t n, std::vector<uint8_t>& output) {
    const int64_t bound = compression_bound((uint32_t)n);
    if (n > 0x7e000000u || bound < 0) return 0;
    uint64_t off = 0, len = n, cap = (uint64_t)bound, olen = 0; int32_t st = 0;
    const size_t old = output.size();
    output.resize(old + (size_t)cap + 64);
    std::vector<uint8_t> in(input, input + n); in.resize(n + 64);
    ctx.check(rcz_lz4_encode_blocks(ctx.get(), in.data(), &off, &len, output.data() + old, &off, &cap, &olen, &st, 1, RCZ_MEM_HOST), "rcz_lz4_encode_blocks");
    output.resize(old + (st == RCZ_OK ? (size_t)olen : 0));
    if (st != RCZ_OK) throw error_from_status(st, "lz4::encode_block");
    return (size_t)olen;
}

// The frame Encoder the reference left as a stub (`compress()` returns false, lz4.rs:543-545: every block goes out raw): same frame
// layout and block size (FLG 0x60, BD 0x50 = 256 KiB, header checksum byte 0, two zero words at the end, lz4.rs:526,550-575), but
// every full block is compressed with `encode_block`, `batch_blocks` of them per C-ABI call, and written compressed
// when that is smaller than the block (else raw, high bit set — what lz4.rs:567-575 does for all of them).
template <class W> class CompressingEncoder {
  public:
    size_t batch_blocks = 64;
    CompressingEncoder(Context& ctx, W w) : ctx_(ctx), w_(std::move(w)) {}
    size_t write(const uint8_t* buf, size_t len) {
        if (!wrote_header_) { detail::write_u32_le(w_, MAGIC); const uint8_t h[3] = {0x60, 0x50, 0}; w_.write(h, 3); wrote_header_ = true; }
        buf_.insert(buf_.end(), buf, buf + len);
        while (buf_.size() - done_ >= limit_ * batch_blocks) encode_blocks(batch_blocks);
        return len;
    }
    void flush() { encode_blocks((buf_.size() - done_ + limit_ - 1) / limit_); w_.flush(); }
    W finish() { flush(); detail::write_u32_le(w_, 0); detail::write_u32_le(w_, 0); return std::move(w_); }

  private:
    void encode_blocks(size_t nb) {
        if (nb == 0) return;
        std::vector<uint64_t> ioff(nb), ilen(nb), ocap(nb), olen(nb);
        std::vector<int32_t> st(nb);
        for (size_t i = 0; i < nb; ++i) {
            ioff[i] = done_ + i * limit_;
            ilen[i] = std::min(limit_, buf_.size() - (size_t)ioff[i]);
            ocap[i] = (uint64_t)compression_bound((uint32_t)ilen[i]);
        }
        std::vector<uint64_t> ooff = detail::prefix(ocap, 16);
        std::vector<uint8_t> enc((size_t)(ooff.back() + ocap.back()) + 64);
        buf_.resize(buf_.size() + 64);
        ctx_.check(rcz_lz4_encode_blocks(ctx_.get(), buf_.data(), ioff.data(), ilen.data(), enc.data(), ooff.data(), ocap.data(), olen.data(), st.data(), nb,
                                         RCZ_MEM_HOST), "rcz_lz4_encode_blocks");
        buf_.resize(buf_.size() - 64);
        for (size_t i = 0; i < nb; ++i) {
            if (st[i] == RCZ_OK && olen[i] > 0 && olen[i] < ilen[i]) {
                detail::write_u32_le(w_, (uint32_t)olen[i]);
                w_.write(enc.data() + ooff[i], (size_t)olen[i]);
            } else {
                detail::write_u32_le(w_, (uint32_t)ilen[i] | 0x80000000u);
                w_.write(buf_.data() + ioff[i], (size_t)ilen[i]);
            }
        }
        done_ += nb * limit_;
        if (done_ >= buf_.size()) { buf_.clear(); done_ = 0; }
    }
    Context& ctx_; W w_; std::vector<uint8_t> buf_; size_t done_ = 0; bool wrote_header_ = false; size_t limit_ = 256 * 1024;
};
}  // namespace lz4

// ================================================================================================ bwt
namespace bwt {
// bwt/mod.rs:213-219 encode_simple -> (L, origin)
inline std::pair<std::vector<uint8_t>, size_t> encode_simple(Context& ctx, const uint8_t* input, size_t n) {
    std::vector<uint8_t> out(n + 64);
    uint64_t off = 0, len = n; uint32_t origin = 0; int32_t st = 0;
    ctx.check(rcz_bwt_encode_blocks(ctx.get(), input, &off, &len, out.data(), &off, &origin, &st, 1, RCZ_MEM_HOST), "rcz_bwt_encode_blocks");
    if (st != RCZ_OK) throw error_from_status(st, "bwt::encode");
    out.resize(n);
    return {std::move(out), (size_t)origin};
}
// bwt/mod.rs:291-294 decode_simple
inline std::vector<uint8_t> decode_simple(Context& ctx, const uint8_t* input, size_t n, size_t origin) {
    std::vector<uint8_t> out(n + 64);
    uint64_t off = 0, len = n, olen = 0; uint32_t og = (uint32_t)origin; int32_t st = 0;
    if (origin > 0xffffffffull) throw error_from_status(RCZ_E_MALFORMED, "bwt::decode");
    ctx.check(rcz_bwt_decode_blocks(ctx.get(), input, &off, &len, &og, out.data(), &off, &olen, &st, 1, RCZ_MEM_HOST), "rcz_bwt_decode_blocks");
    if (st != RCZ_OK) throw error_from_status(st, "bwt::decode");
    out.resize((size_t)olen);
    return out;
}

template <class W> class Encoder {                                           // bwt/mod.rs:437-518
  public:
    size_t batch_blocks = 64;
    // block_size <= 16,777,214: the inverse transform's link table packs a 24-bit position (rcz.h); the reference's own docs suggest
    // 4 MiB (bwt/mod.rs:32).  Refusing here keeps the Encoder from writing streams its own Decoder cannot read back.
    Encoder(Context& ctx, W w, size_t block_size) : ctx_(ctx), w_(std::move(w)), block_size_(block_size) {
        if (block_size == 0 || block_size > 0xFFFFFEu) throw io_error(ErrorKind::InvalidInput, "bwt::Encoder: block_size must be in 1..16777214", RCZ_E_UNSUPPORTED);
    }
    size_t write(const uint8_t* buf, size_t len) {
        if (!wrote_header_) { detail::write_u32_le(w_, (uint32_t)block_size_); wrote_header_ = true; }   // even for an empty write (App. B #2)
        size_t done = 0;
        while (done < len) {
            size_t room = block_size_ - (pend_.size() - cur_start_);
            size_t amt = room < len - done ? room : len - done;
            pend_.insert(pend_.end(), buf + done, buf + done + amt);
            if (pend_.size() - cur_start_ == block_size_) { blocks_.push_back({cur_start_, block_size_}); cur_start_ = pend_.size(); if (blocks_.size() >= batch_blocks) encode_pending(); }
            done += amt;
        }
        return len;
    }
    void flush() {
        if (pend_.size() > cur_start_) { blocks_.push_back({cur_start_, pend_.size() - cur_start_}); cur_start_ = pend_.size(); }
        encode_pending();
        w_.flush();
    }
    W finish() { flush(); return std::move(w_); }

  private:
    void encode_pending() {                                                   // encode_block (bwt/mod.rs:461-480) for every full block, one call
        if (blocks_.empty()) return;
        const size_t nb = blocks_.size();
        std::vector<uint64_t> off(nb), n(nb); std::vector<uint32_t> origin(nb); std::vector<int32_t> st(nb);
        for (size_t i = 0; i < nb; ++i) { off[i] = blocks_[i].first; n[i] = blocks_[i].second; }
        std::vector<uint8_t> out(pend_.size() + 64);
        pend_.resize(pend_.size() + 64);
        ctx_.check(rcz_bwt_encode_blocks(ctx_.get(), pend_.data(), off.data(), n.data(), out.data(), off.data(), origin.data(), st.data(), nb, RCZ_MEM_HOST),
                   "rcz_bwt_encode_blocks");
        pend_.resize(pend_.size() - 64);
        for (size_t i = 0; i < nb; ++i) {
            if (st[i] != RCZ_OK) throw error_from_status(st[i], "bwt::Encoder");
            detail::write_u32_le(w_, (uint32_t)n[i]);
            w_.write(out.data() + off[i], (size_t)n[i]);
            detail::write_u32_le(w_, origin[i]);
        }
        pend_.erase(pend_.begin(), pend_.begin() + (long)cur_start_);
        cur_start_ = 0; blocks_.clear();
    }
    Context& ctx_; W w_; size_t block_size_;
    std::vector<uint8_t> pend_; size_t cur_start_ = 0;
    std::vector<std::pair<size_t, size_t>> blocks_;
    bool wrote_header_ = false;
};

template <class R> class Decoder {                                           // bwt/mod.rs:321-432
  public:
    R r;
    size_t batch_blocks = 64;
    Decoder(Context& ctx, R r_, bool extra_mem = true) : r(std::move(r_)), ctx_(ctx) { (void)extra_mem; }   // decode_minimal is buggy upstream (App. B #7): one path
    void reset() { header_ = false; out_.clear(); }
    size_t read(uint8_t* dst, size_t len) {
        if (!header_) { read_header(); header_ = true; }
        size_t amt = len;
        while (amt > 0) {
            if (out_.avail() == 0) {
                if (pending_err_) { pending_err_ = false; throw pending_; }
                if (ended_) break;
                fill();
                if (out_.avail() == 0) { if (pending_err_) { pending_err_ = false; throw pending_; } if (ended_) break; }
            }
            amt -= out_.take(dst + (len - amt), amt);
        }
        return len - amt;
    }

  private:
    void read_header() {                                                      // bwt/mod.rs:362-371: byteorder_err_to_io => Other
        uint8_t b[4];
        if (!detail::try_read_exact(r, b, 4)) throw io_error(ErrorKind::Other, "unexpected end of file", RCZ_E_UNEXPECTED_EOF);
        max_block_size_ = (size_t)b[0] | ((size_t)b[1] << 8) | ((size_t)b[2] << 16) | ((size_t)b[3] << 24);
        ended_ = false;
    }
    void fill() {                                                             // decode_block (bwt/mod.rs:373-401) for up to batch_blocks blocks
        out_.clear();
        std::vector<uint8_t> l; std::vector<uint64_t> off, n; std::vector<uint32_t> origin;
        try {
            while (n.size() < batch_blocks) {
                uint8_t b[4];
                if (!detail::try_read_exact(r, b, 4)) { ended_ = true; break; }    // any EOF inside the length word is a clean end (App. B #3)
                const uint32_t bn = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
                const size_t at = l.size();
                detail::push_exactly(r, bn, l);
                const uint32_t og = detail::read_u32_le(r);                      // raw UnexpectedEof (bwt/mod.rs:384)
                off.push_back(at); n.push_back(bn); origin.push_back(og);
            }
        } catch (const io_error& e) { pending_ = e; pending_err_ = true; l.resize(off.empty() ? 0 : (size_t)(off.back() + n.back())); }
        const size_t nb = n.size();
        if (nb == 0) return;
        std::vector<uint64_t> olen(nb); std::vector<int32_t> st(nb);
        std::vector<uint8_t> out(l.size() + 64);
        l.resize(l.size() + 64);
        ctx_.check(rcz_bwt_decode_blocks(ctx_.get(), l.data(), off.data(), n.data(), origin.data(), out.data(), off.data(), olen.data(), st.data(), nb, RCZ_MEM_HOST),
                   "rcz_bwt_decode_blocks");
        for (size_t i = 0; i < nb; ++i) {
            if (n[i] == 0) continue;                                            // an empty block decodes to nothing
            if (st[i] != RCZ_OK) { pending_ = error_from_status(st[i], "bwt::Decoder"); pending_err_ = true; ended_ = false; break; }
            out_.buf.insert(out_.buf.end(), out.begin() + (size_t)off[i], out.begin() + (size_t)(off[i] + olen[i]));
        }
    }
    Context& ctx_;
    detail::Outlet out_;
    bool header_ = false, ended_ = false, pending_err_ = false;
    size_t max_block_size_ = 0;                                               // parsed, never enforced (App. B #16)
    io_error pending_{ErrorKind::Other, ""};
};
}  // namespace bwt

// ================================================================================================ dc
namespace dc {
// bwt/dc.rs:153-159 encode_simple: init[256] followed by the distances
inline std::vector<uint32_t> encode_simple(Context& ctx, const uint8_t* input, size_t n) {
    std::vector<uint32_t> out(256 + n + 16);
    uint64_t off = 0, len = n, cap = 256 + n, olen = 0; int32_t st = 0;
    const uint8_t none = 0;
    if (!input) input = &none;                                                // an empty block is legal (dc.rs:291-296)
    ctx.check(rcz_dc_encode_blocks(ctx.get(), input, &off, &len, out.data(), &off, &cap, &olen, &st, 1, RCZ_MEM_HOST), "rcz_dc_encode_blocks");
    if (st != RCZ_OK) throw error_from_status(st, "dc::encode");
    out.resize((size_t)olen);
    return out;
}
// bwt/dc.rs:236-252 decode_simple
inline std::vector<uint8_t> decode_simple(Context& ctx, size_t n, const uint32_t* distances, size_t count) {
    std::vector<uint8_t> out(n + 64);
    uint64_t off = 0, len = count, nn = n; int32_t st = 0;
    ctx.check(rcz_dc_decode_blocks(ctx.get(), distances, &off, &len, out.data(), &off, &nn, &st, 1, RCZ_MEM_HOST), "rcz_dc_decode_blocks");
    if (st == RCZ_E_UNEXPECTED_EOF) throw io_error(ErrorKind::Other, "Unexpected end of file", st);   // dc.rs:245-246
    if (st != RCZ_OK) throw error_from_status(st, "dc::decode");
    out.resize(n);
    return out;
}
}  // namespace dc

// ================================================================================================ flate
namespace flate {
template <class R> class Decoder {                                           // flate.rs:164-488
  public:
    R r;
    size_t max_output = 1ull << 30;        // the stream's decoded size is unknown: start at 8x the input, grow on OUTPUT_FULL up to this
    Decoder(Context& ctx, R r_) : r(std::move(r_)), ctx_(ctx) {}
    bool eof() const { return decoded_ && out_.avail() == 0; }                // flate.rs:453-455: true once the final block has been handed out
    void reset() { decoded_ = false; out_.clear(); in_.clear(); used_ = 0; }
    size_t consumed() const { return used_; }                                 // bytes of the inner stream that belong to the DEFLATE data
    const uint8_t* unread(size_t* n) const { *n = in_.size() - 64 - used_; return in_.data() + used_; }   // e.g. a zlib trailer

    size_t read(uint8_t* dst, size_t len) {
        if (!decoded_) decode_all();
        size_t k = out_.take(dst, len);
        if (k == 0 && len > 0 && pending_err_) { pending_err_ = false; throw pending_; }
        return k;
    }

  private:
    void decode_all() {
        detail::read_to_end(r, in_);
        const uint64_t n = in_.size();
        in_.resize(in_.size() + 64);
        uint64_t off = 0, cap = n * 8 + 4096;
        for (;;) {
            out_.buf.assign((size_t)cap + 64, 0);
            uint64_t olen = 0, used = 0; int32_t st = 0, det = 0;
            ctx_.check(rcz_flate_decode_streams(ctx_.get(), in_.data(), &off, &n, out_.buf.data(), &off, &cap, &olen, &used, &st, &det, 1, RCZ_MEM_HOST),
                       "rcz_flate_decode_streams");
            if (st == RCZ_E_OUTPUT_FULL && cap < max_output) { cap = cap * 4 < max_output ? cap * 4 : max_output; continue; }
            out_.buf.resize((size_t)olen);
            used_ = (size_t)used;
            if (st != RCZ_OK) { pending_ = error_from_status(st, "flate::Decoder", det); pending_err_ = true; }   // after the bytes decoded before the error
            break;
        }
        decoded_ = true;
    }
    Context& ctx_;
    detail::Outlet out_;
    std::vector<uint8_t> in_;
    size_t used_ = 0;
    bool decoded_ = false, pending_err_ = false;
    io_error pending_{ErrorKind::Other, ""};
};
// Many independent raw-DEFLATE streams in ONE C-ABI call — the shape the GPU path is built for (BASELINE configs[3]: 131,072
// streams of 64 KiB, one warp per stream).  A `Decoder` above drives one stream per call, i.e. one warp of the whole GPU; a
// caller that holds many streams (archive members, pages, chunks of a chunked format) uses this instead of a loop of Decoders.
// caps[i] = room for stream i's output (the sizes a container's directory declares); a stream that needs more ends with
// RCZ_E_OUTPUT_FULL in status[i] and is the caller's to retry with a larger cap.  Statuses follow flate.rs's errors
// (error_from_status turns one into the io_error a Decoder would throw).
struct Many {
    std::vector<uint8_t> bytes;                     // all outputs, stream i at [off[i], off[i] + len[i])
    std::vector<uint64_t> off, len, used;           // used[i]: bytes of the stream that belong to the DEFLATE data
    std::vector<int32_t> status, detail;
    std::vector<uint32_t> adler;                    // zlib::decode_many only: Adler-32 of every stream's output
    std::vector<uint8_t> get(size_t i) const { return std::vector<uint8_t>(bytes.begin() + (long)off[i], bytes.begin() + (long)(off[i] + len[i])); }
};
namespace detail_many {
inline void layout(const std::vector<std::pair<const uint8_t*, size_t>>& streams, const std::vector<uint64_t>& caps, std::vector<uint8_t>& in,
                   std::vector<uint64_t>& in_off, std::vector<uint64_t>& in_len, Many& m) {
    if (caps.size() != streams.size()) throw io_error(ErrorKind::InvalidInput, "decode_many: one capacity per stream");
    const size_t n = streams.size();
    in_off.resize(n); in_len.resize(n); m.off.resize(n); m.len.assign(n, 0); m.used.assign(n, 0); m.status.assign(n, 0); m.detail.assign(n, 0);
    uint64_t ti = 0, to = 0;
    for (size_t i = 0; i < n; ++i) { in_off[i] = ti; in_len[i] = streams[i].second; ti += streams[i].second; m.off[i] = to; to += (caps[i] + 15) & ~(uint64_t)15; }
    in.resize((size_t)ti + 64);
    for (size_t i = 0; i < n; ++i) if (streams[i].second) memcpy(in.data() + in_off[i], streams[i].first, streams[i].second);
    m.bytes.assign((size_t)to + 64, 0);
}
}  // namespace detail_many
inline Many decode_many(Context& ctx, const std::vector<std::pair<const uint8_t*, size_t>>& streams, const std::vector<uint64_t>& caps) {
    Many m; std::vector<uint8_t> in; std::vector<uint64_t> in_off, in_len;
    detail_many::layout(streams, caps, in, in_off, in_len, m);
    if (streams.empty()) return m;
    ctx.check(rcz_flate_decode_streams(ctx.get(), in.data(), in_off.data(), in_len.data(), m.bytes.data(), m.off.data(), caps.data(), m.len.data(), m.used.data(),
                                       m.status.data(), m.detail.data(), streams.size(), RCZ_MEM_HOST), "rcz_flate_decode_streams");
    return m;
}
}  // namespace flate

// ================================================================================================ zlib
namespace zlib {
inline const char* message_of(int detail) {                                   // the reference's InvalidInput messages, zlib.rs:58-113
    switch (detail) {
    case RCZ_ZL_UNSUPPORTED_FORMAT: return "unsupported zlib stream format";
    case RCZ_ZL_UNSUPPORTED_WINDOW: return "unsupported zlib window size";
    case RCZ_ZL_PRESET_DICTIONARY: return "unsupported initial dictionary in the output stream";
    case RCZ_ZL_BAD_HEADER_CHECKSUM: return "invalid zlib header checksum";
    case RCZ_ZL_BAD_CHECKSUM: return "invalid checksum on zlib stream";
    default: return "zlib::Decoder";
    }
}
template <class R> class Decoder {                                           // zlib.rs:32-117
  public:
    R r;
    size_t max_output = 1ull << 30;
    Decoder(Context& ctx, R r_) : r(std::move(r_)), ctx_(ctx) {}
    bool eof() const { return decoded_ && out_.avail() == 0; }                // zlib.rs:88: the inner decoder has handed out its final block
    void reset() { decoded_ = false; out_.clear(); in_.clear(); }            // zlib.rs:91-95
    uint32_t checksum() const { return adler_; }                              // hash.result() after the last byte

    // header on the first call (zlib.rs:99-102), then decoded bytes; the trailer is compared only after a block of zero bytes (zlib.rs:104-117)
    size_t read(uint8_t* dst, size_t len) {
        if (!decoded_) decode_all();
        size_t k = out_.take(dst, len);
        if (k == 0 && len > 0 && pending_err_) { pending_err_ = false; throw pending_; }
        return k;
    }

  private:
    void decode_all() {
        detail::read_to_end(r, in_);
        const uint64_t n = in_.size();
        in_.resize(in_.size() + 64);
        uint64_t off = 0, cap = n * 8 + 4096;
        for (;;) {
            out_.buf.assign((size_t)cap + 64, 0);
            uint64_t olen = 0, used = 0; int32_t st = 0, det = 0; uint32_t ad = 1;
            ctx_.check(rcz_zlib_decode_streams(ctx_.get(), in_.data(), &off, &n, out_.buf.data(), &off, &cap, &olen, &used, &st, &det, &ad, 1, RCZ_MEM_HOST),
                       "rcz_zlib_decode_streams");
            if (st == RCZ_E_OUTPUT_FULL && cap < max_output) { cap = cap * 4 < max_output ? cap * 4 : max_output; continue; }
            out_.buf.resize((size_t)(olen < cap ? olen : cap));
            adler_ = ad;
            if (st != RCZ_OK) { pending_ = error_from_status(st, det >= RCZ_ZL_UNSUPPORTED_FORMAT ? message_of(det) : "flate::Decoder", det); pending_err_ = true; }
            break;
        }
        decoded_ = true;
    }
    Context& ctx_;
    detail::Outlet out_;
    std::vector<uint8_t> in_;
    uint32_t adler_ = 1;
    bool decoded_ = false, pending_err_ = false;
    io_error pending_{ErrorKind::Other, ""};
};
// many independent zlib streams in one C-ABI call (see flate::decode_many); adler[i] = Adler-32 of stream i's output
inline flate::Many decode_many(Context& ctx, const std::vector<std::pair<const uint8_t*, size_t>>& streams, const std::vector<uint64_t>& caps) {
    flate::Many m; std::vector<uint8_t> in; std::vector<uint64_t> in_off, in_len;
    flate::detail_many::layout(streams, caps, in, in_off, in_len, m);
    m.adler.assign(streams.size(), 1);
    if (streams.empty()) return m;
    ctx.check(rcz_zlib_decode_streams(ctx.get(), in.data(), in_off.data(), in_len.data(), m.bytes.data(), m.off.data(), caps.data(), m.len.data(), m.used.data(),
                                      m.status.data(), m.detail.data(), m.adler.data(), streams.size(), RCZ_MEM_HOST), "rcz_zlib_decode_streams");
    return m;
}
}  // namespace zlib

// ================================================================================================ ari
namespace ari {
template <class W> class ByteEncoder {                                       // entropy/ari/table.rs:185-224
  public:
    ByteEncoder(Context& ctx, W w) : ctx_(ctx), w_(std::move(w)) {}
    size_t write(const uint8_t* buf, size_t len) { in_.insert(in_.end(), buf, buf + len); return len; }
    void flush() { w_.flush(); }
    W finish() {                                                              // terminator symbol + u32 BE tail (table.rs:203-207, ari/mod.rs:230-237)
        const uint64_t n = in_.size();
        in_.resize(in_.size() + 64);
        uint64_t off = 0, cap = 2 * n + 64, olen = 0; int32_t st = 0;
        std::vector<uint8_t> out((size_t)cap + 64);
        ctx_.check(rcz_ari_encode_streams(ctx_.get(), in_.data(), &off, &n, out.data(), &off, &cap, &olen, &st, 1, RCZ_MEM_HOST), "rcz_ari_encode_streams");
        if (st != RCZ_OK) throw error_from_status(st, "ari::ByteEncoder");
        w_.write(out.data(), (size_t)olen);
        w_.flush();
        return std::move(w_);
    }

  private:
    Context& ctx_; W w_; std::vector<uint8_t> in_;
};

template <class R> class ByteDecoder {                                       // entropy/ari/table.rs:229-273
  public:
    size_t max_output = 1ull << 30;
    ByteDecoder(Context& ctx, R r) : ctx_(ctx), r_(std::move(r)) {}
    size_t read(uint8_t* dst, size_t len) {
        if (!decoded_) decode_all();
        size_t k = out_.take(dst, len);
        if (k == 0 && len > 0 && pending_err_) { pending_err_ = false; throw pending_; }
        return k;
    }
    // the reader positioned after the bytes `finish()` consumes in the reference (ari/mod.rs:289-292): the next stream starts there
    VecReader finish() {
        if (!decoded_) decode_all();
        return VecReader(std::vector<uint8_t>(in_.begin() + (long)used_, in_.end() - 64));
    }

  private:
    void decode_all() {
        detail::read_to_end(r_, in_);
        const uint64_t n = in_.size();
        in_.resize(in_.size() + 64);
        uint64_t off = 0, cap = n * 4 + 4096;
        for (;;) {
            out_.buf.assign((size_t)cap + 64, 0);
            uint64_t olen = 0, used = 0; int32_t st = 0;
            ctx_.check(rcz_ari_decode_streams(ctx_.get(), in_.data(), &off, &n, out_.buf.data(), &off, &cap, &olen, &used, &st, 1, RCZ_MEM_HOST),
                       "rcz_ari_decode_streams");
            if (st == RCZ_E_OUTPUT_FULL && cap < max_output) { cap = cap * 4 < max_output ? cap * 4 : max_output; continue; }
            out_.buf.resize((size_t)(olen < cap ? olen : cap));
            used_ = (size_t)(used < n ? used : n);
            if (st != RCZ_OK) { pending_ = error_from_status(st, "ari::ByteDecoder"); pending_err_ = true; }
            break;
        }
        decoded_ = true;
    }
    Context& ctx_; R r_;
    detail::Outlet out_;
    std::vector<uint8_t> in_;
    size_t used_ = 0;
    bool decoded_ = false, pending_err_ = false;
    io_error pending_{ErrorKind::Other, ""};
};
}  // namespace ari

// ================================================================================================ mtf
namespace mtf {
// bwt/mtf.rs:95-169.  The reference codes byte by byte as they are written / read; the list state carries over between calls,
// so buffering the whole stream and coding it with one kernel call at finish() / on the first read() gives the same bytes.
template <class W> class Encoder {                                           // mtf.rs:95-130
  public:
    Encoder(Context& ctx, W w) : ctx_(ctx), w_(std::move(w)) {}
    size_t write(const uint8_t* buf, size_t len) { in_.insert(in_.end(), buf, buf + len); return len; }
    void flush() { w_.flush(); }
    W finish() {                                                              // mtf.rs:112-114 (+ the ranks of everything written so far)
        const uint64_t n = in_.size();
        in_.resize(in_.size() + 64);
        uint64_t off = 0, cap = n, olen = 0; int32_t st = 0;
        std::vector<uint8_t> out((size_t)cap + 64);
        ctx_.check(rcz_mtf_encode_streams(ctx_.get(), in_.data(), &off, &n, out.data(), &off, &cap, &olen, &st, 1, RCZ_MEM_HOST), "rcz_mtf_encode_streams");
        if (st != RCZ_OK) throw error_from_status(st, "mtf::Encoder");
        w_.write(out.data(), (size_t)olen);
        w_.flush();
        return std::move(w_);
    }

  private:
    Context& ctx_; W w_; std::vector<uint8_t> in_;
};

template <class R> class Decoder {                                           // mtf.rs:133-169
  public:
    Decoder(Context& ctx, R r) : ctx_(ctx), r_(std::move(r)) {}
    size_t read(uint8_t* dst, size_t len) {                                   // mtf.rs:155-168: short count at the end of the inner stream
        if (!decoded_) decode_all();
        return out_.take(dst, len);
    }
    R finish() { return std::move(r_); }                                      // mtf.rs:149-151

  private:
    void decode_all() {
        std::vector<uint8_t> in;
        detail::read_to_end(r_, in);
        const uint64_t n = in.size();
        in.resize(in.size() + 64);
        uint64_t off = 0, cap = n, olen = 0; int32_t st = 0;
        out_.buf.assign((size_t)cap + 64, 0);
        ctx_.check(rcz_mtf_decode_streams(ctx_.get(), in.data(), &off, &n, out_.buf.data(), &off, &cap, &olen, &st, 1, RCZ_MEM_HOST), "rcz_mtf_decode_streams");
        if (st != RCZ_OK) throw error_from_status(st, "mtf::Decoder");
        out_.buf.resize((size_t)olen);
        decoded_ = true;
    }
    Context& ctx_; R r_;
    detail::Outlet out_;
    bool decoded_ = false;
};
}  // namespace mtf

// ================================================================================================ rle
namespace rle {
template <class W> class Encoder {                                           // rle.rs:40-123 (one whole-buffer write; App. B #11, #12)
  public:
    Encoder(Context& ctx, W w) : ctx_(ctx), w_(std::move(w)) {}
    size_t write(const uint8_t* buf, size_t len) { in_.insert(in_.end(), buf, buf + len); return len; }
    void flush() { w_.flush(); }
    W finish() {
        const uint64_t n = in_.size();
        in_.resize(in_.size() + 64);
        uint64_t off = 0, cap = 2 * n + 16, olen = 0; int32_t st = 0;
        std::vector<uint8_t> out((size_t)cap + 64);
        ctx_.check(rcz_rle_encode_streams(ctx_.get(), in_.data(), &off, &n, out.data(), &off, &cap, &olen, &st, 1, RCZ_MEM_HOST), "rcz_rle_encode_streams");
        if (st != RCZ_OK) throw error_from_status(st, "rle::Encoder");
        w_.write(out.data(), (size_t)olen);
        w_.flush();
        return std::move(w_);
    }

  private:
    Context& ctx_; W w_; std::vector<uint8_t> in_;
};

template <class R> class Decoder {                                           // rle.rs:176-281
  public:
    size_t max_output = 1ull << 32;
    Decoder(Context& ctx, R r) : ctx_(ctx), r_(std::move(r)) {}
    size_t read(uint8_t* dst, size_t len) {
        if (!decoded_) decode_all();
        size_t k = out_.take(dst, len);
        if (k == 0 && len > 0 && pending_err_) { pending_err_ = false; throw pending_; }
        return k;
    }

  private:
    void decode_all() {
        std::vector<uint8_t> in;
        detail::read_to_end(r_, in);
        const uint64_t n = in.size();
        in.resize(in.size() + 64);
        uint64_t off = 0, cap = n * 4 + 4096;
        for (;;) {
            out_.buf.assign((size_t)cap + 64, 0);
            uint64_t olen = 0; int32_t st = 0;
            ctx_.check(rcz_rle_decode_streams(ctx_.get(), in.data(), &off, &n, out_.buf.data(), &off, &cap, &olen, &st, 1, RCZ_MEM_HOST), "rcz_rle_decode_streams");
            if (st == RCZ_E_OUTPUT_FULL && cap < max_output) { cap = (olen > cap && olen < max_output) ? olen : (cap * 4 < max_output ? cap * 4 : max_output); continue; }
            out_.buf.resize((size_t)(olen < cap ? olen : cap));
            if (st != RCZ_OK) { pending_ = error_from_status(st, "rle::Decoder"); pending_err_ = true; }
            break;
        }
        decoded_ = true;
    }
    Context& ctx_; R r_;
    detail::Outlet out_;
    bool decoded_ = false, pending_err_ = false;
    io_error pending_{ErrorKind::Other, ""};
};
}  // namespace rle

}  // namespace rcz
