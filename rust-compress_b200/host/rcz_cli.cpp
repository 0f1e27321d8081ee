// rcz_cli.cpp — the reference's test application (`/root/reference/src/main.rs`) on top of librcz: the archive container and the
// pass chaining of main.rs:20,69-182, every pass running through the host mirrors of rcz_stream.hpp and therefore
// through the C ABI and the GPU kernels.
//
//   rcz_cli <options> <method1> .. <methodN> < input > archive        options: -d (decompress), -block<N> (BWT block size, default 65536)
//   rcz_cli -d < archive > output
//
// Container (main.rs:166-171): u32 LE 0x73632172 ("r!cs"), u8 number of methods, per method u8 length + name; then the payload.
// Chaining (main.rs:172-178 / 154-160): on encode the writers are stacked in list order, so the LAST method sees the input first and
// method1 writes the archive body; on decode the readers are stacked in list order, method1 closest to the archive.
// Passes (main.rs:71-131): dummy, ari, bwt, mtf, lz4 — plus `lz4c`, the frame encoder the reference left as a stub, filled in.
//
// Deliberate difference: the reference only calls flush() on the stack (no finish()), which leaves `ari`
// archives without terminator + tail and `lz4` archives without end mark, and its bwt / lz4 `Write::write` return Ok(0), which makes
// `io::copy` fail outright.  This tool finishes every pass, so its archives are complete and decode; the container bytes are the
// reference's.
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "rcz_stream.hpp"

namespace {

constexpr uint32_t MAGIC = 0x73632172;   // main.rs:20

// Box<dyn Write> / Box<dyn Read>
struct DynWrite {
    virtual ~DynWrite() {}
    virtual size_t write(const uint8_t* p, size_t n) = 0;
    virtual void flush() = 0;
    virtual void finish() = 0;           // complete this pass, then the ones below it
};
struct DynRead {
    virtual ~DynRead() {}
    virtual size_t read(uint8_t* p, size_t n) = 0;
};
using WBox = std::shared_ptr<DynWrite>;
using RBox = std::shared_ptr<DynRead>;
// what the mirrors see as their inner W / R
struct WRef { WBox b; size_t write(const uint8_t* p, size_t n) { return b->write(p, n); } void flush() { b->flush(); } };
struct RRef { RBox b; size_t read(uint8_t* p, size_t n) { return b->read(p, n); } };

struct FileWrite : DynWrite {
    FILE* f; explicit FileWrite(FILE* f_) : f(f_) {}
    size_t write(const uint8_t* p, size_t n) override { if (fwrite(p, 1, n, f) != n) throw rcz::io_error(rcz::ErrorKind::Other, "write failed"); return n; }
    void flush() override { fflush(f); }
    void finish() override { fflush(f); }
};
struct FileRead : DynRead {
    FILE* f; explicit FileRead(FILE* f_) : f(f_) {}
    size_t read(uint8_t* p, size_t n) override { return fread(p, 1, n, f); }
};

template <class Enc> struct EncPass : DynWrite {
    Enc e; WBox inner; bool done = false;
    template <class... A> EncPass(WBox in, A&&... a) : e(std::forward<A>(a)...), inner(std::move(in)) {}
    size_t write(const uint8_t* p, size_t n) override { return e.write(p, n); }
    void flush() override { e.flush(); }
    void finish() override { if (!done) { done = true; (void)e.finish(); inner->finish(); } }
};
struct DummyWrite : DynWrite {
    WBox inner; explicit DummyWrite(WBox in) : inner(std::move(in)) {}
    size_t write(const uint8_t* p, size_t n) override { return inner->write(p, n); }
    void flush() override { inner->flush(); }
    void finish() override { inner->finish(); }
};
template <class Dec> struct DecPass : DynRead {
    Dec d;
    template <class... A> explicit DecPass(A&&... a) : d(std::forward<A>(a)...) {}
    size_t read(uint8_t* p, size_t n) override { return d.read(p, n); }
};

struct Config { std::vector<std::string> methods; size_t block_size = 1 << 16; bool decompress = false; };

struct Pass {
    std::function<WBox(WBox, rcz::Context&, const Config&)> encode;
    std::function<RBox(RBox, rcz::Context&, const Config&)> decode;
    const char* info;
};

std::map<std::string, Pass> passes() {                                           // main.rs:71-131
    std::map<std::string, Pass> p;
    p["dummy"] = {[](WBox w, rcz::Context&, const Config&) -> WBox { return std::make_shared<DummyWrite>(w); },
                  [](RBox r, rcz::Context&, const Config&) -> RBox { return r; }, "pass-through"};
    p["ari"] = {[](WBox w, rcz::Context& c, const Config&) -> WBox { return std::make_shared<EncPass<rcz::ari::ByteEncoder<WRef>>>(w, c, WRef{w}); },
                [](RBox r, rcz::Context& c, const Config&) -> RBox { return std::make_shared<DecPass<rcz::ari::ByteDecoder<RRef>>>(c, RRef{r}); },
                "Adaptive arithmetic byte coder"};
    p["bwt"] = {[](WBox w, rcz::Context& c, const Config& cfg) -> WBox { return std::make_shared<EncPass<rcz::bwt::Encoder<WRef>>>(w, c, WRef{w}, cfg.block_size); },
                [](RBox r, rcz::Context& c, const Config&) -> RBox { return std::make_shared<DecPass<rcz::bwt::Decoder<RRef>>>(c, RRef{r}, true); },
                "Burrows-Wheeler Transformation"};
    p["mtf"] = {[](WBox w, rcz::Context& c, const Config&) -> WBox { return std::make_shared<EncPass<rcz::mtf::Encoder<WRef>>>(w, c, WRef{w}); },
                [](RBox r, rcz::Context& c, const Config&) -> RBox { return std::make_shared<DecPass<rcz::mtf::Decoder<RRef>>>(c, RRef{r}); },
                "Move-To-Front Transformation"};
    p["lz4"] = {[](WBox w, rcz::Context&, const Config&) -> WBox { return std::make_shared<EncPass<rcz::lz4::Encoder<WRef>>>(w, WRef{w}); },
                [](RBox r, rcz::Context& c, const Config&) -> RBox { return std::make_shared<DecPass<rcz::lz4::Decoder<RRef>>>(c, RRef{r}); },
                "Ziv-Lempel derivative, focused at speed (raw blocks, like the reference's stub encoder)"};
    p["lz4c"] = {[](WBox w, rcz::Context& c, const Config&) -> WBox { return std::make_shared<EncPass<rcz::lz4::CompressingEncoder<WRef>>>(w, c, WRef{w}); },
                 [](RBox r, rcz::Context& c, const Config&) -> RBox { return std::make_shared<DecPass<rcz::lz4::Decoder<RRef>>>(c, RRef{r}); },
                 "lz4 with blocks compressed by encode_block (not in the reference CLI)"};
    return p;
}

Config query(int argc, char** argv) {                                            // main.rs:29-58
    Config cfg;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (!a.empty() && a[0] == '-') {
            if (a.compare(1, 5, "block") == 0) cfg.block_size = (size_t)strtoull(a.c_str() + 6, nullptr, 10);
            else if (a.compare(1, 1, "d") == 0) cfg.decompress = true;
            else printf("Warning: unrecognized option: %s\n", a.c_str());
        } else cfg.methods.push_back(a);
    }
    return cfg;
}

}  // namespace

int main(int argc, char** argv) {
    auto table = passes();
    const Config cfg = query(argc, argv);
    try {
        if (cfg.decompress) {                                                    // main.rs:136-161
            if (!cfg.methods.empty()) { fprintf(stderr, "Decompression methods are set in stone\n"); return 2; }
            uint8_t h[5];
            if (fread(h, 1, 4, stdin) != 4) { fprintf(stderr, "Unable to read input\n"); return 1; }
            const uint32_t magic = (uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16) | ((uint32_t)h[3] << 24);
            if (magic != MAGIC) { fprintf(stderr, "Input is not a rust-compress archive\n"); return 1; }
            if (fread(h, 1, 1, stdin) != 1) { fprintf(stderr, "truncated archive header\n"); return 1; }
            std::vector<std::string> methods;
            for (unsigned i = 0; i < h[0]; ++i) {
                uint8_t len;
                if (fread(&len, 1, 1, stdin) != 1) { fprintf(stderr, "truncated archive header\n"); return 1; }
                std::string name(len, '\0');
                if (len && fread(&name[0], 1, len, stdin) != len) { fprintf(stderr, "unexpected end of file\n"); return 1; }
                methods.push_back(name);
            }
            rcz::Context ctx(0);
            RBox rsum = std::make_shared<FileRead>(stdin);
            for (const auto& m : methods) {
                auto it = table.find(m);
                if (it == table.end()) { fprintf(stderr, "Pass is not implemented\n"); return 3; }
                rsum = it->second.decode(rsum, ctx, cfg);
            }
            std::vector<uint8_t> buf(1 << 20);
            for (;;) { const size_t k = rsum->read(buf.data(), buf.size()); if (k == 0) break; if (fwrite(buf.data(), 1, k, stdout) != k) return 1; }
            fflush(stdout);
        } else if (cfg.methods.empty()) {                                        // main.rs:162-171
            printf("rust-compress test application (librcz build)\nUsage:\n\t%s <options> <method1> .. <methodN> <input >output\n", argv[0]);
            printf("Options:\n\t-d (to decompress)\n\t-block<N> (BWT block size)\nPasses:\n");
            for (const auto& kv : table) printf("\t%s = %s\n", kv.first.c_str(), kv.second.info);
        } else {                                                                 // main.rs:172-181
            for (const auto& m : cfg.methods)
                if (!table.count(m) || m.size() > 255) { fprintf(stderr, "Pass %s is not implemented\n", m.c_str()); return 3; }
            const uint8_t hdr[5] = {(uint8_t)MAGIC, (uint8_t)(MAGIC >> 8), (uint8_t)(MAGIC >> 16), (uint8_t)(MAGIC >> 24), (uint8_t)cfg.methods.size()};
            fwrite(hdr, 1, 5, stdout);
            for (const auto& m : cfg.methods) { const uint8_t l = (uint8_t)m.size(); fwrite(&l, 1, 1, stdout); fwrite(m.data(), 1, m.size(), stdout); }
            rcz::Context ctx(0);
            WBox wsum = std::make_shared<FileWrite>(stdout);
            for (const auto& m : cfg.methods) wsum = table[m].encode(wsum, ctx, cfg);
            std::vector<uint8_t> buf(1 << 20);
            size_t total = 0;
            for (;;) { const size_t k = fread(buf.data(), 1, buf.size(), stdin); if (k == 0) break; wsum->write(buf.data(), k); total += k; }
            if (total == 0) wsum->write(buf.data(), 0);      // bwt / lz4 emit their stream header on the first write (App. B #2): an empty input still gets one
            wsum->finish();
        }
    } catch (const rcz::io_error& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
