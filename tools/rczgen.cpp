// rczgen.cpp — deterministic synthetic-input generators for tests and bench.py.
// Not part of the product library.  splitmix64 is bit-identical to tools/gen.py's Python version.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>
#include <thread>
#include <vector>
#include <atomic>

static inline uint64_t sm64(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

extern "C" {

// uniform random bytes: little-endian bytes of successive splitmix64 outputs
void gen_random(uint64_t seed, uint8_t* out, size_t n) {
    uint64_t s = seed; size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t v = sm64(s); memcpy(out + i, &v, 8); }
    if (i < n) { uint64_t v = sm64(s); memcpy(out + i, &v, n - i); }
}

// "lzsyn" (SURVEY §8d C2): literal runs of 1..32 fresh random bytes (p=1/4) and matches of 4..67 bytes at
// offsets 1..min(pos,65535) (p=3/4), byte-forward copy, truncated at n.
void gen_lzsyn(uint64_t seed, uint8_t* out, size_t n) {
    uint64_t s = seed; size_t pos = 0;
    while (pos < n) {
        uint64_t u = sm64(s);
        if (pos == 0 || (u & 3) == 0) {
            size_t run = 1 + ((u >> 2) & 31);
            for (size_t k = 0; k < run && pos < n; k += 8) {
                uint64_t v = sm64(s);
                for (size_t b = 0; b < 8 && k + b < run && pos < n; ++b) out[pos++] = (uint8_t)(v >> (8 * b));
            }
        } else {
            size_t lim = pos < 65535 ? pos : 65535;
            size_t off = 1 + (size_t)((u >> 8) % lim);
            size_t len = 4 + ((u >> 32) & 63);
            for (size_t k = 0; k < len && pos < n; ++k, ++pos) out[pos] = out[pos - off];
        }
    }
}

// hexdump-like text in the style of the reference's src/data/test.txt:
// "00000000  xx xx xx xx xx xx xx xx  xx xx xx xx xx xx xx xx  |................|\n"
void gen_hextext(uint64_t seed, uint8_t* out, size_t n) {
    static const char* hx = "0123456789abcdef";
    uint64_t s = seed; size_t pos = 0; uint32_t addr = (uint32_t)(sm64(s) & 0xfffff0);
    char line[96];
    while (pos < n) {
        uint8_t b[16];
        uint64_t v0 = sm64(s), v1 = sm64(s);
        // skew bytes towards text-like values so that the ASCII column and repeats look like a real dump
        for (int i = 0; i < 8; ++i) { b[i] = (uint8_t)(v0 >> (8 * i)); b[8 + i] = (uint8_t)(v1 >> (8 * i)); }
        for (int i = 0; i < 16; ++i) { if (b[i] & 0x80) b[i] = (uint8_t)(0x20 + (b[i] & 0x3f)); else if ((b[i] & 0x60) == 0) b[i] = 0; }
        int k = 0;
        for (int i = 7; i >= 0; --i) line[k++] = hx[(addr >> (4 * i)) & 15];
        line[k++] = ' ';
        for (int i = 0; i < 16; ++i) { if (i == 8) line[k++] = ' '; line[k++] = ' '; line[k++] = hx[b[i] >> 4]; line[k++] = hx[b[i] & 15]; }
        line[k++] = ' '; line[k++] = ' '; line[k++] = '|';
        for (int i = 0; i < 16; ++i) line[k++] = (b[i] >= 0x20 && b[i] < 0x7f) ? (char)b[i] : '.';
        line[k++] = '|'; line[k++] = '\n';
        size_t m = (size_t)k < n - pos ? (size_t)k : n - pos;
        memcpy(out + pos, line, m); pos += m; addr += 16;
    }
}

// run-structured bytes for RLE: singles (p=3/4) or runs of 2..1001 of one byte
void gen_runs(uint64_t seed, uint8_t* out, size_t n) {
    uint64_t s = seed; size_t pos = 0;
    while (pos < n) {
        uint64_t u = sm64(s);
        uint8_t b = (uint8_t)u;
        size_t r = ((u >> 8) & 3) != 0 ? 1 : 2 + (size_t)((u >> 10) % 1000);
        for (size_t k = 0; k < r && pos < n; ++k) out[pos++] = b;
    }
}

// generate `count` units of `unit` bytes each in parallel; kind: 0 random, 1 lzsyn, 2 hextext, 3 runs
void gen_units(int kind, uint64_t seed0, uint8_t* out, size_t unit, size_t count, int nthreads) {
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < count;) {
            uint64_t seed = seed0 ^ i;
            uint8_t* o = out + i * unit;
            switch (kind) {
            case 0: gen_random(seed, o, unit); break;
            case 1: gen_lzsyn(seed, o, unit); break;
            case 2: gen_hextext(seed, o, unit); break;
            default: gen_runs(seed, o, unit); break;
            }
        }
    };
    if (nthreads <= 1) { work(); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
}

// liblz4 block compression of many units in parallel (liblz4.so.1 is loaded by the caller and passed in)
typedef int (*lz4_compress_fn)(const char*, char*, int, int);
void lz4_compress_units(void* fn, const uint8_t* in, size_t unit, size_t count, uint8_t* out, size_t out_stride,
                        uint64_t* out_len, int nthreads) {
    lz4_compress_fn f = (lz4_compress_fn)fn;
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < count;)
            out_len[i] = (uint64_t)f((const char*)in + i * unit, (char*)out + i * out_stride, (int)unit, (int)out_stride);
    };
    std::vector<std::thread> th;
    for (int t = 0; t < (nthreads < 1 ? 1 : nthreads); ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
}

}  // extern "C"
