// oracle/pipeline.cpp — CPU composition of the restated stages for BASELINE configs[4] (bwt -> dc -> entropy::ari)
// (TEST INFRASTRUCTURE ONLY, see oracle.h)
//
// The reference names the chain ("BWT + DC + EC", /root/reference/src/bwt/mod.rs:11-14) but never composes it: dc has no wire
// format and nothing feeds it to the range coder.  The composition below is therefore this project's and
// mirrors rust-compress_b200/csrc/pipeline.cu; every STAGE is the restated reference function:
//   orc_bwt_encode  (bwt/mod.rs:136-204)   orc_dc_encode  (dc.rs:62-159)   orc_ari_encode  (table.rs:203-219)
//   orc_ari_decode  (table.rs:255-272)     orc_dc_decode  (dc.rs:162-252)  orc_bwt_decode  (bwt/mod.rs:223-294)
// Serialisation: init[256] then the distances, each u32 LE; cut into `chunk`-byte pieces (0 = one piece), one ByteEncoder stream each.
// Container: u32 LE x 6 = magic "BDA1", n, origin, nsym, chunk, nstreams; nstreams x u32 LE code length; code bytes.
#include "oracle.h"
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {
constexpr uint32_t MAGIC = 0x31414442u;
inline void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
inline uint32_t get32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
template <class F> void parallel_for(size_t n, int nthreads, F f) {
    if (nthreads <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto& t : th) t.join();
}
}  // namespace

extern "C" int orc_bda_encode_block(const uint8_t* in, size_t n, uint32_t chunk, uint8_t* out, size_t cap, size_t* out_len, uint32_t* origin_out) {
    *out_len = 0;
    if (n == 0) return ORC_E_MALFORMED;
    std::vector<uint8_t> l(n);
    uint32_t origin = 0;
    int st = orc_bwt_encode(in, n, l.data(), &origin);
    if (st) return st;
    std::vector<uint32_t> ser(256 + n);
    size_t ndist = 0;
    st = orc_dc_encode(l.data(), n, ser.data(), ser.data() + 256, &ndist, nullptr, nullptr, nullptr);
    if (st) return st;
    const size_t nsym = 256 + ndist, ser_len = 4 * nsym;
    std::vector<uint8_t> bytes(ser_len);
    for (size_t i = 0; i < nsym; ++i) put32(bytes.data() + 4 * i, ser[i]);
    const size_t ns = chunk ? (ser_len + chunk - 1) / chunk : 1;
    const size_t hdr = 24 + 4 * ns;
    if (hdr > cap) return ORC_E_OUTPUT_FULL;
    size_t pos = hdr;
    for (size_t k = 0; k < ns; ++k) {
        const size_t lo = chunk ? k * chunk : 0, len = chunk ? std::min<size_t>(chunk, ser_len - lo) : ser_len;
        std::vector<uint8_t> code(2 * len + 64);
        size_t clen = 0;
        st = orc_ari_encode(bytes.data() + lo, len, code.data(), code.size(), &clen);
        if (st) return st;
        if (pos + clen > cap) return ORC_E_OUTPUT_FULL;
        memcpy(out + pos, code.data(), clen);
        put32(out + 24 + 4 * k, (uint32_t)clen);
        pos += clen;
    }
    put32(out, MAGIC); put32(out + 4, (uint32_t)n); put32(out + 8, origin); put32(out + 12, (uint32_t)nsym); put32(out + 16, chunk); put32(out + 20, (uint32_t)ns);
    *out_len = pos;
    if (origin_out) *origin_out = origin;
    return ORC_OK;
}

extern "C" int orc_bda_decode_block(const uint8_t* in, size_t in_len, size_t n, uint32_t chunk, uint8_t* out, size_t* out_len) {
    *out_len = 0;
    if (n == 0) return ORC_E_MALFORMED;
    if (in_len < 24) return ORC_E_UNEXPECTED_EOF;
    const uint32_t nsym = get32(in + 12), ns = get32(in + 20), origin = get32(in + 8);
    const size_t ser_len = 4 * (size_t)nsym;
    if (get32(in) != MAGIC || get32(in + 4) != n || get32(in + 16) != chunk) return ORC_E_INVALID_INPUT;
    if (nsym < 256 || nsym > 256 + n || ns != (chunk ? (ser_len + chunk - 1) / chunk : 1)) return ORC_E_MALFORMED;
    size_t pos = 24 + 4 * (size_t)ns;
    if (pos > in_len) return ORC_E_UNEXPECTED_EOF;
    size_t total = pos;
    for (uint32_t k = 0; k < ns; ++k) total += get32(in + 24 + 4 * k);
    if (total > in_len) return ORC_E_UNEXPECTED_EOF;
    std::vector<uint8_t> bytes(ser_len);
    int worst = 0;
    for (uint32_t k = 0; k < ns; ++k) {
        const size_t clen = get32(in + 24 + 4 * k);
        const size_t lo = chunk ? (size_t)k * chunk : 0, len = chunk ? std::min<size_t>(chunk, ser_len - lo) : ser_len;
        size_t got = 0, cr = 0, cf = 0;
        int st = orc_ari_decode(in + pos, clen, bytes.data() + lo, len, &got, &cr, &cf);
        if (st == ORC_OK && got != len) st = ORC_E_MALFORMED;
        if (st < worst) worst = st;                               // the device reports the smallest status code over a block's streams
        pos += clen;
    }
    if (worst) return worst;
    std::vector<uint32_t> ser(nsym);
    for (size_t i = 0; i < nsym; ++i) ser[i] = get32(bytes.data() + 4 * i);
    std::vector<uint8_t> l(n);
    size_t used = 0;
    int st = orc_dc_decode(n, ser.data(), ser.data() + 256, nsym - 256, l.data(), &used, nullptr, nullptr, nullptr);
    if (st) return st;
    if (origin >= n) return ORC_E_MALFORMED;
    return orc_bwt_decode(l.data(), n, origin, out, out_len);
}

extern "C" int orc_bda_encode_blocks_mt(const uint8_t* in_base, const uint64_t* in_off, const uint64_t* n, uint32_t chunk, uint8_t* out_base,
                                        const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, uint32_t* origin, int32_t* status,
                                        size_t nblocks, int nthreads) {
    parallel_for(nblocks, nthreads, [&](size_t i) {
        size_t got = 0;
        status[i] = orc_bda_encode_block(in_base + in_off[i], n[i], chunk, out_base + out_off[i], out_cap[i], &got, &origin[i]);
        out_len[i] = got;
    });
    return ORC_OK;
}

extern "C" int orc_bda_decode_blocks_mt(const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len, uint32_t chunk, uint8_t* out_base,
                                        const uint64_t* out_off, const uint64_t* n, uint64_t* out_len, int32_t* status, size_t nblocks, int nthreads) {
    parallel_for(nblocks, nthreads, [&](size_t i) {
        size_t got = 0;
        status[i] = orc_bda_decode_block(in_base + in_off[i], in_len[i], n[i], chunk, out_base + out_off[i], &got);
        out_len[i] = got;
    });
    return ORC_OK;
}
