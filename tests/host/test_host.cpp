// test_host.cpp — the reference crate's own unit tests, restated against the C++ host mirrors (rcz_stream.hpp).
//   lz4.rs:647-726   decode (fixtures), one_byte_at_a_time, random_byte_lengths, some_roundtrips
//   bwt/mod.rs:541-551 some_roundtrips          bwt/dc.rs:291-302 roundtrips
//   entropy/ari/test.rs:185-212 roundtrips, roundtrips_term
//   flate.rs:528-582 decode (fixtures), one_byte_at_a_time (+eof), random_byte_lengths
//   rle.rs:320-361   simple/long run encoding + decoding KATs, random_roundtrips
// Built by tests/test_host_mirrors.py against librcz_emu.so (CPU, no GPU needed) and against librcz.so (-m gpu).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iostream>
#include <iterator>
#include <random>
#include <string>

#include "rcz_stream.hpp"

using bytes = std::vector<uint8_t>;
static std::string g_dir;
static int g_fail = 0;

static bytes load(const std::string& name) {
    std::ifstream f(g_dir + "/" + name, std::ios::binary);
    if (!f) { std::cerr << "missing fixture " << name << "\n"; exit(2); }
    return bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
#define CHECK(cond) do { if (!(cond)) { std::cerr << "  FAILED " << #cond << " at line " << __LINE__ << "\n"; ++g_fail; return; } } while (0)
static void run(const char* name, const std::function<void()>& f) {
    int before = g_fail;
    try { f(); } catch (const std::exception& e) { std::cerr << "  exception: " << e.what() << "\n"; ++g_fail; }
    std::cout << (g_fail == before ? "ok   " : "FAIL ") << name << std::endl;
}
template <class D> static bytes read_to_end(D& d, size_t chunk = 4096) {
    bytes out, buf(chunk);
    for (;;) { size_t k = d.read(buf.data(), chunk); if (k == 0) break; out.insert(out.end(), buf.begin(), buf.begin() + (long)k); }
    return out;
}
static bytes lit(const char* s) { return bytes(s, s + strlen(s)); }

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: test_host <golden dir>\n"; return 2; }
    g_dir = argv[1];
    rcz::Context ctx(0);
    const bytes txt = load("ref_test.txt");
    std::mt19937 rng(12345);

    // ------------------------------------------------------------------------------------------ lz4
    run("lz4::decode fixtures", [&] {
        for (int i = 1; i <= 9; ++i) {
            bytes f = load("ref_test.lz4." + std::to_string(i));
            rcz::lz4::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
            CHECK(read_to_end(d) == txt);
            CHECK(d.eof());
        }
    });
    run("lz4::one_byte_at_a_time", [&] {
        bytes f = load("ref_test.lz4.1");
        rcz::lz4::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
        bytes out; uint8_t b;
        while (d.read(&b, 1) == 1) out.push_back(b);
        CHECK(d.eof());
        CHECK(out == txt);
    });
    run("lz4::random_byte_lengths", [&] {
        bytes f = load("ref_test.lz4.1");
        rcz::lz4::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
        bytes out, buf(40);
        for (;;) { size_t want = 1 + rng() % 40; size_t k = d.read(buf.data(), want); if (k == 0) break; out.insert(out.end(), buf.begin(), buf.begin() + (long)k); }
        CHECK(out == txt);
    });
    run("lz4::some_roundtrips", [&] {
        bytes big; for (int i = 0; i < 300; ++i) big.insert(big.end(), txt.begin(), txt.end());   // > 256 KiB: several raw blocks
        for (const bytes& input : {lit("test"), lit(""), txt, big}) {
            rcz::lz4::Encoder<rcz::VecWriter> e{rcz::VecWriter()};
            e.write(input.data(), input.size());
            bytes enc = e.finish().v;
            rcz::lz4::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(enc));
            CHECK(read_to_end(d) == input);
        }
    });
    run("lz4::block larger than the BD size decodes (max_block_size is a reserve hint, lz4.rs:444-446)", [&] {
        // frame with BD = 64 KiB whose single compressed block holds 100,000 zero bytes: 1 literal + one overlapping match of 99,999
        bytes blk = {0x1F, 0x00, 0x01, 0x00};
        for (int i = 0; i < 392; ++i) blk.push_back(0xFF);
        blk.push_back(20);                                                     // 4 + 15 + 392 * 255 + 20 = 99,999
        bytes f = {0x04, 0x22, 0x4D, 0x18, 0x60, 0x40, 0x00};
        const uint32_t n = (uint32_t)blk.size();
        for (int k = 0; k < 4; ++k) f.push_back((uint8_t)(n >> (8 * k)));
        f.insert(f.end(), blk.begin(), blk.end());
        for (int k = 0; k < 4; ++k) f.push_back(0);
        rcz::lz4::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
        bytes out = read_to_end(d);
        CHECK(out.size() == 100000 && out == bytes(100000, 0));
    });
    run("lz4::CompressingEncoder (the stub at lz4.rs:543-545 filled in) round trips through lz4::Decoder", [&] {
        bytes big; for (int i = 0; i < 300; ++i) big.insert(big.end(), txt.begin(), txt.end());   // > 256 KiB: several blocks
        bytes noise(300000); uint32_t x = 12345; for (auto& b : noise) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }   // stays raw
        int idx = 0;
        for (const bytes& input : {lit("test"), lit(""), txt, big, noise}) {
            rcz::lz4::CompressingEncoder<rcz::VecWriter> e(ctx, rcz::VecWriter());
            e.write(input.data(), input.size());
            bytes enc = e.finish().v;
            if (idx == 3) CHECK(enc.size() < input.size() / 2);                   // the repeated text compresses
            if (idx == 4) CHECK(enc.size() >= input.size());                      // the noise goes out raw
            ++idx;
            rcz::lz4::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(enc));
            CHECK(read_to_end(d) == input);
        }
        bytes out;                                                              // encode_block / decode_block free functions (lz4.rs:602-627)
        size_t n = rcz::lz4::encode_block(ctx, txt.data(), txt.size(), out);
        CHECK(n == 2724 && out.size() == 2724);
        bytes back;
        CHECK(rcz::lz4::decode_block(ctx, out.data(), out.size(), back) == txt.size() && back == txt);
    });
    run("lz4::decode_block + errors", [&] {
        bytes f = load("ref_test.lz4.3");
        uint32_t n = (uint32_t)f[7] | ((uint32_t)f[8] << 8) | ((uint32_t)f[9] << 16) | ((uint32_t)f[10] << 24);
        bytes out;
        CHECK(rcz::lz4::decode_block(ctx, f.data() + 11, n, out) == txt.size());
        CHECK(out == txt);
        CHECK(rcz::lz4::compression_bound(100) == 120 && rcz::lz4::compression_bound(0x7e000001u) < 0);
        bytes bad = f; bad[0] ^= 1;                                            // lz4.rs:365: bad magic => InvalidInput
        rcz::lz4::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(bad));
        uint8_t b; bool threw = false;
        try { d.read(&b, 1); } catch (const rcz::io_error& e) { threw = e.kind == rcz::ErrorKind::InvalidInput; }
        CHECK(threw);
        bytes cut(f.begin(), f.begin() + 600);                                 // truncated payload => Other "unexpected end of file"
        rcz::lz4::Decoder<rcz::SliceReader> d2(ctx, rcz::SliceReader(cut));
        threw = false;
        try { read_to_end(d2); } catch (const rcz::io_error& e) { threw = e.kind == rcz::ErrorKind::Other; }
        CHECK(threw);
    });

    // ------------------------------------------------------------------------------------------ bwt
    run("bwt::some_roundtrips", [&] {
        for (const bytes& input : {lit("test"), lit(""), txt}) {
            rcz::bwt::Encoder<rcz::VecWriter> e(ctx, rcz::VecWriter(), 1 << 10);
            e.write(input.data(), input.size());
            bytes enc = e.finish().v;
            rcz::bwt::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(enc), true);
            CHECK(read_to_end(d) == input);
        }
    });
    run("bwt::stream golden (SURVEY App. C)", [&] {
        bytes in = lit("abracadabra");
        rcz::bwt::Encoder<rcz::VecWriter> e(ctx, rcz::VecWriter(), 1024);
        e.write(in.data(), in.size());
        bytes enc = e.finish().v;
        const uint8_t want[] = {0x00, 0x04, 0x00, 0x00, 0x0b, 0x00, 0x00, 0x00, 'r', 'd', 'a', 'r', 'c', 'a', 'a', 'a', 'a', 'b', 'b', 0x02, 0x00, 0x00, 0x00};
        CHECK(enc == bytes(want, want + sizeof want));
        auto lo = rcz::bwt::encode_simple(ctx, in.data(), in.size());
        CHECK(lo.second == 2 && lo.first == lit("rdarcaaaabb"));
        CHECK(rcz::bwt::decode_simple(ctx, lo.first.data(), lo.first.size(), lo.second) == in);
        bytes cut(enc.begin(), enc.begin() + 10);                             // truncated block payload => Other
        rcz::bwt::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(cut));
        bool threw = false;
        try { read_to_end(d); } catch (const rcz::io_error& er) { threw = er.kind == rcz::ErrorKind::Other; }
        CHECK(threw);
        bytes cut2(enc.begin(), enc.begin() + 6);                             // truncated length word => clean EOF (App. B #3)
        rcz::bwt::Decoder<rcz::SliceReader> d2(ctx, rcz::SliceReader(cut2));
        CHECK(read_to_end(d2).empty());
    });

    // ------------------------------------------------------------------------------------------ dc
    run("dc::roundtrips", [&] {
        for (const bytes& input : {lit("teeesst_dc"), lit(""), txt, lit("../data/test.txt")}) {
            std::vector<uint32_t> dist = rcz::dc::encode_simple(ctx, input.data(), input.size());
            CHECK(rcz::dc::decode_simple(ctx, input.size(), dist.data(), dist.size()) == input);
        }
    });

    // ------------------------------------------------------------------------------------------ ari
    run("ari::roundtrips", [&] {
        for (const bytes& input : {lit("abracadabra"), lit(""), txt}) {
            rcz::ari::ByteEncoder<rcz::VecWriter> e(ctx, rcz::VecWriter());
            e.write(input.data(), input.size());
            bytes enc = e.finish().v;
            rcz::ari::ByteDecoder<rcz::SliceReader> d(ctx, rcz::SliceReader(enc));
            CHECK(read_to_end(d) == input);
        }
    });
    run("ari::roundtrips_term", [&] {                                          // two terminated streams back to back (ari/test.rs:52-89)
        bytes a = lit("abracadabra"), b2(txt.begin(), txt.begin() + 777);
        rcz::ari::ByteEncoder<rcz::VecWriter> e1(ctx, rcz::VecWriter());
        e1.write(a.data(), a.size());
        rcz::ari::ByteEncoder<rcz::VecWriter> e2(ctx, e1.finish());
        e2.write(b2.data(), b2.size());
        bytes enc = e2.finish().v;
        rcz::ari::ByteDecoder<rcz::SliceReader> d1(ctx, rcz::SliceReader(enc));
        CHECK(read_to_end(d1) == a);
        rcz::ari::ByteDecoder<rcz::VecReader> d2(ctx, d1.finish());
        CHECK(read_to_end(d2) == b2);
    });

    // ------------------------------------------------------------------------------------------ flate
    auto fixup = [](bytes v) { return bytes(v.begin() + 2, v.end() - 4); };      // flate.rs:504-506
    run("flate::decode fixtures", [&] {
        for (int i = 0; i <= 9; ++i) {
            bytes f = fixup(load("ref_test.z." + std::to_string(i)));
            rcz::flate::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
            CHECK(read_to_end(d) == txt);
        }
        bytes g = load("ref_test.z.go");
        rcz::flate::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(g));
        CHECK(read_to_end(d) == txt);
    });
    run("flate::one_byte_at_a_time", [&] {
        bytes f = fixup(load("ref_test.z.1"));
        rcz::flate::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
        CHECK(!d.eof());
        bytes out; uint8_t b;
        while (d.read(&b, 1) == 1) out.push_back(b);
        CHECK(d.eof());
        CHECK(out == txt);
    });
    run("flate::random_byte_lengths", [&] {
        bytes f = fixup(load("ref_test.z.1"));
        rcz::flate::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
        bytes out, buf(40);
        for (;;) { size_t want = 1 + rng() % 40; size_t k = d.read(buf.data(), want); if (k == 0) break; out.insert(out.end(), buf.begin(), buf.begin() + (long)k); }
        CHECK(out == txt);
    });
    run("flate::zlib framing leaves the trailer unread", [&] {
        bytes z = load("ref_test.z.5");
        bytes body(z.begin() + 2, z.end());                                      // DEFLATE data followed by the 4-byte Adler-32
        rcz::flate::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(body));
        CHECK(read_to_end(d) == txt);
        size_t rest = 0; const uint8_t* tail = d.unread(&rest);
        CHECK(rest == 4 && tail[0] == 0xfb && tail[1] == 0x4f && tail[2] == 0xcf && tail[3] == 0xa6);   // Adler-32(test.txt) = 0xfb4fcfa6
        bytes bad = fixup(load("ref_test.z.5")); bad[0] |= 0x06;                 // BTYPE 3 => InvalidInput
        rcz::flate::Decoder<rcz::SliceReader> d2(ctx, rcz::SliceReader(bad));
        bool threw = false;
        try { read_to_end(d2); } catch (const rcz::io_error& e) { threw = e.kind == rcz::ErrorKind::InvalidInput && e.detail == RCZ_FL_INVALID_BLOCK_CODE; }
        CHECK(threw);
    });

    run("flate::decode_many / zlib::decode_many (one call, many streams)", [&] {
        std::vector<bytes> raw, z;
        for (int i = 0; i <= 9; ++i) { z.push_back(load("ref_test.z." + std::to_string(i))); raw.push_back(fixup(z.back())); }
        bytes bad = raw[5]; bad[0] |= 0x06;                                      // BTYPE 3
        raw.push_back(bad);
        raw.push_back(bytes());                                                  // empty stream: NotEnoughBits
        std::vector<std::pair<const uint8_t*, size_t>> ss;
        std::vector<uint64_t> caps;
        for (auto& v : raw) { ss.emplace_back(v.data(), v.size()); caps.push_back(txt.size()); }
        caps[3] = txt.size() - 1;                                                // too small
        rcz::flate::Many m = rcz::flate::decode_many(ctx, ss, caps);
        for (int i = 0; i <= 9; ++i) {
            if (i == 3) { CHECK(m.status[3] == RCZ_E_OUTPUT_FULL); continue; }
            CHECK(m.status[(size_t)i] == RCZ_OK && m.get((size_t)i) == txt && m.used[(size_t)i] == raw[(size_t)i].size());
        }
        CHECK(m.status[10] == RCZ_E_INVALID_INPUT && m.detail[10] == RCZ_FL_INVALID_BLOCK_CODE);
        CHECK(m.status[11] != RCZ_OK);
        ss.clear(); caps.clear();
        for (auto& v : z) { ss.emplace_back(v.data(), v.size()); caps.push_back(txt.size() + 7); }
        rcz::flate::Many mz = rcz::zlib::decode_many(ctx, ss, caps);
        for (size_t i = 0; i < z.size(); ++i) CHECK(mz.status[i] == RCZ_OK && mz.get(i) == txt && mz.adler[i] == 0xfb4fcfa6u);
        CHECK(rcz::flate::decode_many(ctx, {}, {}).status.empty());
    });

    // ------------------------------------------------------------------------------------------ zlib (zlib.rs:140-205)
    run("zlib::decode fixtures", [&] {
        for (int i = 0; i <= 9; ++i) {
            bytes f = load("ref_test.z." + std::to_string(i));
            rcz::zlib::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
            CHECK(read_to_end(d) == txt);
            CHECK(d.checksum() == 0xfb4fcfa6u);
        }
    });
    run("zlib::one_byte_at_a_time", [&] {
        bytes f = load("ref_test.z.1");
        rcz::zlib::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
        CHECK(!d.eof());
        bytes out; uint8_t b;
        while (d.read(&b, 1) == 1) out.push_back(b);
        CHECK(d.eof());
        CHECK(out == txt);
    });
    run("zlib::random_byte_lengths", [&] {
        bytes f = load("ref_test.z.1");
        rcz::zlib::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
        bytes out, buf(40);
        for (;;) { size_t want = 1 + rng() % 40; size_t k = d.read(buf.data(), want); if (k == 0) break; out.insert(out.end(), buf.begin(), buf.begin() + (long)k); }
        CHECK(out == txt);
    });
    run("zlib::header and trailer errors are InvalidInput", [&] {
        auto fails_with = [&](bytes f, int detail) {
            rcz::zlib::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(f));
            try { read_to_end(d); } catch (const rcz::io_error& e) { return e.kind == rcz::ErrorKind::InvalidInput && e.detail == detail; }
            return false;
        };
        bytes f = load("ref_test.z.5");
        bytes a = f; a[0] = 0x77; CHECK(fails_with(a, RCZ_ZL_UNSUPPORTED_FORMAT));
        bytes b = f; b[0] = 0x68; b[1] = 0x81; CHECK(fails_with(b, RCZ_ZL_UNSUPPORTED_WINDOW));
        bytes c = f; c[0] = 0x78; c[1] = 0xBB; CHECK(fails_with(c, RCZ_ZL_PRESET_DICTIONARY));
        bytes e = f; e[1] ^= 1; CHECK(fails_with(e, RCZ_ZL_BAD_HEADER_CHECKSUM));
        // the trailer is only read after a DEFLATE block of zero bytes (zlib.rs:104-109): a wrong trailer behind a non-empty final block
        // goes unnoticed, as in the reference; the empty stream `78 9c 03 00 | 00 00 00 01` is where it is compared
        bytes g = f; g[g.size() - 1] ^= 1;
        { rcz::zlib::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(g)); CHECK(read_to_end(d) == txt); }
        bytes h = {0x78, 0x9c, 0x03, 0x00, 0x00, 0x00, 0x00, 0x01};
        { rcz::zlib::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(h)); CHECK(read_to_end(d).empty()); }
        h[7] = 0x02; CHECK(fails_with(h, RCZ_ZL_BAD_CHECKSUM));
    });

    // ------------------------------------------------------------------------------------------ mtf (bwt/mtf.rs:176-192)
    run("mtf::some_roundtrips", [&] {
        bytes r;
        auto roundtrip = [&](const bytes& in) {
            rcz::mtf::Encoder<rcz::VecWriter> e(ctx, rcz::VecWriter());
            e.write(in.data(), in.size());
            r = e.finish().v;
            CHECK(r.size() == in.size());
            rcz::mtf::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(r));
            CHECK(read_to_end(d) == in);
        };
        const char* t = "teeesst_mtf";
        roundtrip(bytes(t, t + 11));
        CHECK(r[0] == 't' && r[1] == 'e' + 1 && r[2] == 0 && r[3] == 0);        // 't' at rank 116; 'e' (101) pushed back by one; repeats are rank 0
        roundtrip(bytes());
        roundtrip(txt);
    });

    // ------------------------------------------------------------------------------------------ rle
    auto rle_enc = [&](const bytes& in) { rcz::rle::Encoder<rcz::VecWriter> e(ctx, rcz::VecWriter()); e.write(in.data(), in.size()); return e.finish().v; };
    auto rle_dec = [&](const bytes& in) { rcz::rle::Decoder<rcz::SliceReader> d(ctx, rcz::SliceReader(in)); return read_to_end(d); };
    run("rle::simple_encoding / long_run_encoding / decoding", [&] {
        bytes a(5, 20); a.push_back(15);
        CHECK(rle_enc(a) == (bytes{20, 20, 5 - 2 + 128, 15}));
        CHECK(rle_dec(bytes{20, 20, 5 - 2 + 128, 15}) == a);
        CHECK(rle_enc(bytes{0, 0}) == (bytes{0, 0, 128}));
        bytes c(129, 5);
        CHECK(rle_enc(c) == (bytes{5, 5, 255}));
        CHECK(rle_dec(bytes{5, 5, 255}) == c);
        bytes l{1, 3, 4, 4}; l.insert(l.end(), 2 + 52 + 128, 100);
        CHECK(rle_enc(l) == (bytes{1, 3, 4, 4, 128, 100, 100, 52, 129}));
        CHECK(rle_dec(bytes{1, 3, 4, 4, 128, 100, 100, 52, 129}) == l);
        CHECK(rle_enc(lit("abca123")) == lit("abca123") && rle_enc(lit("")).empty());
        bool threw = false;
        bytes over{7, 7, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
        try { rle_dec(over); } catch (const rcz::io_error& e) { threw = e.kind == rcz::ErrorKind::Other && std::string(e.what()) == "Overly long run"; }
        CHECK(threw);
    });
    run("rle::random_roundtrips", [&] {
        for (int it = 0; it < 20; ++it) {
            bytes in(13579);
            for (auto& x : in) x = (uint8_t)(rng() & (it % 2 ? 0xff : 0x03));
            CHECK(rle_dec(rle_enc(in)) == in);
        }
        CHECK(rle_enc(txt).size() == 3084);
    });

    std::cout << (g_fail ? "FAILED " : "PASSED ") << g_fail << std::endl;
    return g_fail ? 1 : 0;
}
