// pipeline.cu — the bzip-style chain of BASELINE configs[4]: BWT -> DC -> entropy coder ("BWT + DC + EC",
// /root/reference/src/bwt/mod.rs:11-14), composed on the device with no host round trip between the stages.
//
// The reference never composes these stages itself (dc has no wire format, SURVEY.md §8d "C5"), so the serialisation and the
// container are ours; every STAGE is the reference's:
//   encode   rcz_bwt_encode_blocks   compute_suffixes + TransformIterator        bwt/mod.rs:136-204
//            rcz_dc_encode_blocks    dc::encode + EncodeIterator                  bwt/dc.rs:62-159
//            serialise               init[256] then the distances, each as u32 LE (the dc kernel's output array, reinterpreted)
//            rcz_ari_encode_streams  ByteEncoder::write + finish                  entropy/ari/table.rs:203-219
//   decode   the same four steps backwards (table.rs:255-272, dc.rs:162-252, bwt/mod.rs:223-294)
// `ari_chunk` cuts the serialised bytes of a block into independent ByteEncoder streams of that many bytes (0 = one stream per
// block, the literal reading of SURVEY §8d).  The range coder is one dependency chain per stream, so the chunked form is what gives
// a B200 enough chains to work on; either way every stream's bytes are exactly what `ByteEncoder` produces for its input.
//
// Container of one block (all fields u32 LE):
//   +0 magic "BDA1"   +4 n   +8 origin   +12 nsym (= 256 + number of distances)   +16 ari_chunk   +20 nstreams
//   +24 nstreams x code length            then the code bytes of the streams, back to back
//
// Hand-off: every stage runs in RCZ_MEM_DEVICE_ASYNC on the context's stream; lengths that only exist on the device (dc's out_len,
// the code lengths) reach the next stage through small planning kernels that write its descriptor arrays in device memory.
#include "rcz_internal.h"
#include <algorithm>

namespace bzp {

constexpr unsigned MAGIC = 0x31414442u;   // "BDA1"
constexpr unsigned HDR = 24;
constexpr int NT = 256;

struct Blk {
    unsigned long long io_off, io_cap;    // encode: container region in the output arena; decode: container in the input arena (cap = its length)
    unsigned long long dc_off;            // u32 element offset of the block's dc array in the staging buffer
    unsigned long long code_off;          // byte offset of the block's first stream slot in the code staging buffer (encode)
    unsigned long long slot_cap;          // bytes per stream slot in the code staging buffer (encode)
    unsigned n, slot0, nslots, skip;
};

__device__ __forceinline__ unsigned streams_for(unsigned long long ser_len, unsigned chunk) {
    return chunk ? (unsigned)((ser_len + chunk - 1) / chunk) : 1u;
}

// ---------------------------------------------------------------------------------------------- encode
// one thread per stream slot: descriptor of the ByteEncoder stream that covers bytes [k*chunk, (k+1)*chunk) of the block's dc array
__global__ void enc_plan_kernel(const Blk* __restrict__ blks, const unsigned* __restrict__ slot2blk, unsigned nslots, unsigned chunk,
                                const uint64_t* __restrict__ dc_len, const int32_t* __restrict__ bwt_st, const int32_t* __restrict__ dc_st,
                                uint64_t* __restrict__ s_in_off, uint64_t* __restrict__ s_in_len, uint64_t* __restrict__ s_out_off,
                                uint64_t* __restrict__ s_out_cap) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const unsigned b = slot2blk[s];
    const Blk bk = blks[b];
    const unsigned k = s - bk.slot0;
    const bool ok = !bk.skip && bwt_st[b] == RCZ_OK && dc_st[b] == RCZ_OK;
    const unsigned long long ser = ok ? dc_len[b] * 4ull : 0ull;
    const unsigned ns = ok ? streams_for(ser, chunk) : 0u;
    s_in_off[s] = bk.dc_off * 4ull + (unsigned long long)k * chunk;
    s_in_len[s] = k < ns ? (chunk ? min((unsigned long long)chunk, ser - (unsigned long long)k * chunk) : ser) : RCZ_STREAM_SKIP;
    s_out_off[s] = bk.code_off + (unsigned long long)k * bk.slot_cap;
    s_out_cap[s] = bk.slot_cap;
}

// one CTA per block: header, code lengths, destination of every stream's bytes, block status
__global__ void __launch_bounds__(NT)
enc_pack_kernel(const Blk* __restrict__ blks, unsigned chunk, const uint64_t* __restrict__ dc_len, const uint32_t* __restrict__ origin,
                const int32_t* __restrict__ bwt_st, const int32_t* __restrict__ dc_st, const uint64_t* __restrict__ s_len,
                const int32_t* __restrict__ s_st, uint8_t* __restrict__ out_base, uint64_t* __restrict__ s_dst, uint64_t* __restrict__ out_len,
                uint32_t* __restrict__ origin_out, int32_t* __restrict__ status, const int32_t* __restrict__ host_status) {
    __shared__ unsigned scratch[40];
    __shared__ unsigned long long carry;
    __shared__ int bad;
    const unsigned b = blockIdx.x, tid = threadIdx.x;
    const Blk bk = blks[b];
    int st = bk.skip ? host_status[b] : (bwt_st[b] != RCZ_OK ? bwt_st[b] : dc_st[b]);
    const unsigned long long ser = st == RCZ_OK ? dc_len[b] * 4ull : 0ull;
    const unsigned ns = st == RCZ_OK ? streams_for(ser, chunk) : 0u;
    if (tid == 0) { carry = 0; bad = 0; }
    __syncthreads();
    for (unsigned k = tid; k < ns; k += NT) if (s_st[bk.slot0 + k] != RCZ_OK) atomicMin(&bad, s_st[bk.slot0 + k]);
    __syncthreads();
    if (st == RCZ_OK && bad) st = bad;
    const unsigned long long hdr = HDR + 4ull * ns;
    uint8_t* out = out_base + bk.io_off;
    const bool hdr_fits = st == RCZ_OK && hdr <= bk.io_cap;
    for (unsigned k0 = 0; k0 < ns; k0 += NT) {                       // exclusive scan of the code lengths, NT streams per round
        const unsigned k = k0 + tid;
        const unsigned len = (k < ns && st == RCZ_OK) ? (unsigned)s_len[bk.slot0 + k] : 0u;
        unsigned total;
        const unsigned ex = block_excl_scan_add<NT>(len, scratch, &total);
        if (k < ns) {
            s_dst[bk.slot0 + k] = hdr + carry + ex;
            if (hdr_fits) {
                uint8_t* p = out + HDR + 4ull * k;
                p[0] = (uint8_t)len; p[1] = (uint8_t)(len >> 8); p[2] = (uint8_t)(len >> 16); p[3] = (uint8_t)(len >> 24);
            }
        }
        __syncthreads();
        if (tid == 0) carry += total;
        __syncthreads();
    }
    const unsigned long long total_len = hdr + carry;
    if (st == RCZ_OK && total_len > bk.io_cap) st = RCZ_E_OUTPUT_FULL;
    if (tid < 6 && st == RCZ_OK) {
        const unsigned f[6] = {MAGIC, bk.n, origin[b], (unsigned)dc_len[b], chunk, ns};
        const unsigned v = f[tid];
        uint8_t* p = out + 4 * tid;
        p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
    }
    if (tid == 0) {
        out_len[b] = st == RCZ_OK ? total_len : 0ull;
        status[b] = st;
        if (origin_out) origin_out[b] = st == RCZ_OK ? origin[b] : 0u;
    }
}

// one CTA per stream slot: code bytes -> their place in the container
__global__ void __launch_bounds__(128)
enc_copy_kernel(const Blk* __restrict__ blks, const unsigned* __restrict__ slot2blk, const uint64_t* __restrict__ s_in_len,
                const uint64_t* __restrict__ s_out_off, const uint64_t* __restrict__ s_len, const uint64_t* __restrict__ s_dst,
                const uint8_t* __restrict__ code, uint8_t* __restrict__ out_base, const int32_t* __restrict__ status) {
    const unsigned s = blockIdx.x;
    if (s_in_len[s] == RCZ_STREAM_SKIP) return;
    const unsigned b = slot2blk[s];
    if (status[b] != RCZ_OK) return;
    const uint8_t* src = code + s_out_off[s];
    uint8_t* dst = out_base + blks[b].io_off + s_dst[s];
    const unsigned long long len = s_len[s];
    // head bytes up to the destination's 4-byte boundary, then 4 bytes per thread (the source is then read unaligned, byte-wise)
    const unsigned head = (unsigned)min(len, (unsigned long long)((4 - ((uintptr_t)dst & 3)) & 3));
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    const unsigned long long words = (len - head) >> 2;
    for (unsigned long long w = threadIdx.x; w < words; w += 128) {
        const uint8_t* p = src + head + 4 * w;
        *reinterpret_cast<unsigned*>(dst + head + 4 * w) = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
    }
    const unsigned long long done = head + 4 * words;
    if (threadIdx.x < len - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

// ---------------------------------------------------------------------------------------------- decode
__device__ __forceinline__ unsigned ld_u32le(const uint8_t* p) { return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24); }

// one CTA per block: validate the header, descriptors of the block's ByteDecoder streams
__global__ void __launch_bounds__(NT)
dec_plan_kernel(const Blk* __restrict__ blks, unsigned chunk, const uint8_t* __restrict__ in_base, uint64_t* __restrict__ s_in_off,
                uint64_t* __restrict__ s_in_len, uint64_t* __restrict__ s_out_off, uint64_t* __restrict__ s_out_cap,
                uint32_t* __restrict__ origin, uint64_t* __restrict__ nsym_out, int32_t* __restrict__ parse_st,
                const int32_t* __restrict__ host_status) {
    __shared__ unsigned scratch[40];
    __shared__ unsigned long long carry;
    const unsigned b = blockIdx.x, tid = threadIdx.x;
    const Blk bk = blks[b];
    const uint8_t* in = in_base + bk.io_off;
    int st = bk.skip ? host_status[b] : RCZ_OK;
    unsigned nsym = 0, ns = 0, org = 0;
    if (st == RCZ_OK) {
        if (bk.io_cap < HDR) st = RCZ_E_UNEXPECTED_EOF;
        else {
            nsym = ld_u32le(in + 12); ns = ld_u32le(in + 20); org = ld_u32le(in + 8);
            const unsigned long long ser = 4ull * nsym;
            if (ld_u32le(in) != MAGIC || ld_u32le(in + 4) != bk.n || ld_u32le(in + 16) != chunk) st = RCZ_E_INVALID_INPUT;
            else if (nsym < 256u || nsym > 256ull + bk.n || ns != streams_for(ser, chunk) || ns > bk.nslots) st = RCZ_E_MALFORMED;
            else if (HDR + 4ull * ns > bk.io_cap) st = RCZ_E_UNEXPECTED_EOF;
        }
    }
    if (st != RCZ_OK) ns = 0;
    if (tid == 0) carry = 0;
    __syncthreads();
    const unsigned long long ser = 4ull * nsym, hdr = HDR + 4ull * ns;
    for (unsigned k0 = 0; k0 < bk.nslots; k0 += NT) {
        const unsigned k = k0 + tid;
        const unsigned len = k < ns ? ld_u32le(in + HDR + 4ull * k) : 0u;
        unsigned total;
        const unsigned ex = block_excl_scan_add<NT>(len, scratch, &total);
        if (k < bk.nslots) {
            const unsigned s = bk.slot0 + k;
            s_in_off[s] = bk.io_off + hdr + carry + ex;
            s_in_len[s] = k < ns ? (unsigned long long)len : RCZ_STREAM_SKIP;
            s_out_off[s] = bk.dc_off * 4ull + (unsigned long long)k * chunk;
            s_out_cap[s] = k < ns ? (chunk ? min((unsigned long long)chunk, ser - (unsigned long long)k * chunk) : ser) : 0ull;
        }
        __syncthreads();
        if (tid == 0) carry += total;
        __syncthreads();
    }
    if (st == RCZ_OK && hdr + carry > bk.io_cap) st = RCZ_E_UNEXPECTED_EOF;       // the code bytes are cut short
    if (tid == 0) { origin[b] = org; nsym_out[b] = st == RCZ_OK ? nsym : 0ull; parse_st[b] = st; }
}

// one CTA per block: every stream must have produced exactly its share of the serialised dc array
__global__ void __launch_bounds__(NT)
dec_check_kernel(const Blk* __restrict__ blks, const uint64_t* __restrict__ s_in_len, const uint64_t* __restrict__ s_out_cap,
                 const uint64_t* __restrict__ s_len, const int32_t* __restrict__ s_st, const int32_t* __restrict__ parse_st,
                 const uint64_t* __restrict__ nsym, uint64_t* __restrict__ dc_in_off, uint64_t* __restrict__ dc_in_len, int32_t* __restrict__ ari_st) {
    __shared__ int bad;
    const unsigned b = blockIdx.x, tid = threadIdx.x;
    const Blk bk = blks[b];
    if (tid == 0) bad = 0;
    __syncthreads();
    if (parse_st[b] == RCZ_OK)
        for (unsigned k = tid; k < bk.nslots; k += NT) {
            const unsigned s = bk.slot0 + k;
            if (s_in_len[s] == RCZ_STREAM_SKIP) continue;
            int e = s_st[s];
            if (e == RCZ_OK && s_len[s] != s_out_cap[s]) e = RCZ_E_MALFORMED;  // a stream that ends early (or late: OUTPUT_FULL) is not ours
            if (e != RCZ_OK) atomicMin(&bad, e);
        }
    __syncthreads();
    if (tid == 0) {
        const bool ok = parse_st[b] == RCZ_OK && bad == 0;
        dc_in_off[b] = bk.dc_off;
        dc_in_len[b] = ok ? nsym[b] : 0ull;
        ari_st[b] = bad;
    }
}

__global__ void dec_merge_kernel(unsigned nblocks, const int32_t* __restrict__ parse_st, const int32_t* __restrict__ ari_st,
                                 const int32_t* __restrict__ dc_st, const int32_t* __restrict__ bwt_st, const uint64_t* __restrict__ bwt_len,
                                 uint64_t* __restrict__ out_len, int32_t* __restrict__ status) {
    const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const int st = parse_st[b] ? parse_st[b] : ari_st[b] ? ari_st[b] : dc_st[b] ? dc_st[b] : bwt_st[b];
    status[b] = st;
    out_len[b] = st == RCZ_OK ? bwt_len[b] : 0ull;
}

struct Geom {
    std::vector<Blk> blks;
    std::vector<int32_t> hstatus;
    std::vector<unsigned> slot2blk;
    std::vector<uint64_t> l_off, n64, dc_off, dc_cap;
    unsigned long long l_bytes = 0, dc_elems = 0, code_bytes = 0;
};

// ari_chunk: 0, or a multiple of 4 that is >= 1024 (the 256 init words always fill the first stream)
inline bool chunk_ok(uint32_t chunk) { return chunk == 0 || (chunk >= 1024 && chunk % 4 == 0); }

inline int geometry(const uint64_t* io_off, const uint64_t* io_cap, const uint64_t* n_arr, size_t nblocks, uint32_t chunk, Geom& g) {
    g.blks.resize(nblocks); g.hstatus.assign(nblocks, 0); g.l_off.resize(nblocks); g.n64.resize(nblocks); g.dc_off.resize(nblocks); g.dc_cap.resize(nblocks);
    for (size_t i = 0; i < nblocks; ++i) {
        Blk& b = g.blks[i];
        memset(&b, 0, sizeof b);
        b.io_off = io_off[i]; b.io_cap = io_cap[i];
        b.slot0 = (unsigned)g.slot2blk.size();
        const unsigned long long n = n_arr[i];
        g.l_off[i] = g.l_bytes; g.dc_off[i] = g.dc_elems; g.n64[i] = n;
        if (n == 0 || n > 0xFFFFFEull) {             // n == 0: bwt/mod.rs:186-188 unwrap on None; 24-bit positions in the inverse transform
            b.skip = 1; g.hstatus[i] = n == 0 ? RCZ_E_MALFORMED : RCZ_E_UNSUPPORTED; g.n64[i] = 0; g.dc_cap[i] = 0;
            continue;
        }
        b.n = (unsigned)n;
        const unsigned long long ser_max = 4ull * (256 + n);
        b.nslots = chunk ? (unsigned)((ser_max + chunk - 1) / chunk) : 1u;
        b.slot_cap = ((chunk ? 2ull * chunk : 2ull * ser_max) + 64 + 15) & ~15ull;   // the tests' bound for ByteEncoder output: 2 x input + 64
        b.dc_off = g.dc_elems; b.code_off = g.code_bytes;
        g.dc_cap[i] = 256 + n;
        g.l_bytes += (n + 63) & ~63ull;
        g.dc_elems += (256 + n + 15) & ~15ull;
        g.code_bytes += b.slot_cap * b.nslots;
        for (unsigned k = 0; k < b.nslots; ++k) g.slot2blk.push_back((unsigned)i);
        if (g.slot2blk.size() > 0x3fffffffu) return RCZ_E_ARG;
    }
    return RCZ_OK;
}

}  // namespace bzp

// stage boundaries: 0 | bwt | dc | plan + ari | pack + copy |   (encode)      0 | plan + ari | check + dc | bwt + merge |   (decode)
extern "C" int rcz_bwt_dc_ari_encode_blocks(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* n_arr, void* out_base,
                                            const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, uint32_t* origin,
                                            int32_t* status, size_t nblocks, uint32_t ari_chunk, int mem_kind) {
    using namespace bzp;
    if (!c || rcz_bad_kind(mem_kind) || !chunk_ok(ari_chunk)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !n_arr || !out_base || !out_off || !out_cap || !out_len || !status || nblocks > 0x3fffffu) return RCZ_E_ARG;
    if (!rcz_spans_ok(in_off, n_arr, nblocks) || !rcz_spans_ok(out_off, out_cap, nblocks)) return RCZ_E_ARG;
    rt_set_device(c->device);
    Geom g;
    int st = geometry(out_off, out_cap, n_arr, nblocks, ari_chunk, g); if (st) return st;
    const size_t T = g.slot2blk.size();

    DescStager ds(c, mem_kind, nblocks);
    const size_t i_blk = ds.add_in(g.blks.data(), nblocks * sizeof(Blk));
    const size_t i_s2b = ds.add_in(g.slot2blk.data(), T * 4);
    const size_t i_hst = ds.add_in(g.hstatus.data(), nblocks * 4);
    const size_t o_len = ds.add_out(out_len, nblocks * 8);
    const size_t o_org = ds.add_out(origin, origin ? nblocks * 4 : 0);
    const size_t o_st = ds.add_out(status, nblocks * 4);
    st = ds.upload(WS_P4); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, n_arr, nblocks, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, out_cap, nblocks, 1, &dout); if (st) return st;
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    void *wL, *wDC, *wCode, *wM;
    st = ctx_ws(c, WS_P0, (size_t)g.l_bytes + 256, &wL); if (st) return st;
    st = ctx_ws(c, WS_P1, (size_t)g.dc_elems * 4 + 256, &wDC); if (st) return st;
    st = ctx_ws(c, WS_P2, (size_t)g.code_bytes + 256, &wCode); if (st) return st;
    // misc: per block  origin u32 | bwt_st | dc_st | dc_len u64      per slot  in_off | in_len | out_off | out_cap | len | dst (u64) | st (i32)
    const size_t sz_b4 = al(nblocks * 4), sz_b8 = al(nblocks * 8), sz_s8 = al(T * 8), sz_s4 = al(T * 4);
    st = ctx_ws(c, WS_P3, 3 * sz_b4 + sz_b8 + 6 * sz_s8 + sz_s4 + 256, &wM); if (st) return st;
    uint8_t* m = (uint8_t*)wM;
    uint32_t* d_org = (uint32_t*)m; m += sz_b4;
    int32_t* d_bwt_st = (int32_t*)m; m += sz_b4;
    int32_t* d_dc_st = (int32_t*)m; m += sz_b4;
    uint64_t* d_dc_len = (uint64_t*)m; m += sz_b8;
    uint64_t* s_in_off = (uint64_t*)m; m += sz_s8;
    uint64_t* s_in_len = (uint64_t*)m; m += sz_s8;
    uint64_t* s_out_off = (uint64_t*)m; m += sz_s8;
    uint64_t* s_out_cap = (uint64_t*)m; m += sz_s8;
    uint64_t* s_len = (uint64_t*)m; m += sz_s8;
    uint64_t* s_dst = (uint64_t*)m; m += sz_s8;
    int32_t* s_st = (int32_t*)m;

    st = ctx_timer_begin(c); if (st) return st;
    st = ctx_stage_mark(c, 0); if (st) return st;
    c->nest++;
    c->nest_may_sync = mem_kind != RCZ_MEM_DEVICE_ASYNC;                      // then the suffix sort may look at its round counters and stop early
    // skipped blocks (n == 0 / too large) reach the stages with n = 0: they report their own error there, ours wins in enc_pack_kernel
    st = rcz_bwt_encode_blocks(c, din, in_off, g.n64.data(), wL, g.l_off.data(), d_org, d_bwt_st, nblocks, RCZ_MEM_DEVICE_ASYNC);
    c->nest_may_sync = false;
    c->nest--;
    if (st) return st;
    st = ctx_stage_mark(c, 1); if (st) return st;
    c->nest++;
    st = rcz_dc_encode_blocks(c, wL, g.l_off.data(), g.n64.data(), (uint32_t*)wDC, g.dc_off.data(), g.dc_cap.data(), d_dc_len, d_dc_st, nblocks,
                              RCZ_MEM_DEVICE_ASYNC);
    c->nest--;
    if (st) return st;
    st = ctx_stage_mark(c, 2); if (st) return st;
    const Blk* dblk = ds.in_ptr<Blk>(i_blk);
    const unsigned* ds2b = ds.in_ptr<unsigned>(i_s2b);
    if (T) {
        RCZ_KLAUNCH(c, enc_plan_kernel, (unsigned)((T + 255) / 256), 256, 0, dblk, ds2b, (unsigned)T, ari_chunk, d_dc_len, d_bwt_st, d_dc_st, s_in_off, s_in_len,
                    s_out_off, s_out_cap);
        st = rcz_ari_launch(c, false, (const uint8_t*)wDC, s_in_off, s_in_len, (uint8_t*)wCode, s_out_off, s_out_cap, s_len, nullptr, s_st, T);
        if (st) return st;
    }
    st = ctx_stage_mark(c, 3); if (st) return st;
    RCZ_KLAUNCH(c, enc_pack_kernel, (unsigned)nblocks, NT, 0, dblk, ari_chunk, d_dc_len, d_org, d_bwt_st, d_dc_st, s_len, s_st, dout, s_dst,
                ds.out_ptr<uint64_t>(o_len), origin ? ds.out_ptr<uint32_t>(o_org) : (uint32_t*)nullptr, ds.out_ptr<int32_t>(o_st), ds.in_ptr<int32_t>(i_hst));
    if (T) RCZ_KLAUNCH(c, enc_copy_kernel, (unsigned)T, 128, 0, dblk, ds2b, s_in_len, s_out_off, s_len, s_dst, (const uint8_t*)wCode, dout, ds.out_ptr<int32_t>(o_st));
    st = ctx_stage_mark(c, 4); if (st) return st;
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) { st = unstage_span_out(c, out_base, dout, out_off, out_len, nblocks, 1); if (st) return st; }
    return RCZ_OK;
}

extern "C" int rcz_bwt_dc_ari_decode_blocks(rcz_ctx* c, const void* in_base, const uint64_t* in_off, const uint64_t* in_len, void* out_base,
                                            const uint64_t* out_off, const uint64_t* n_arr, uint64_t* out_len, int32_t* status, size_t nblocks,
                                            uint32_t ari_chunk, int mem_kind) {
    using namespace bzp;
    if (!c || rcz_bad_kind(mem_kind) || !chunk_ok(ari_chunk)) return RCZ_E_ARG;
    if (nblocks == 0) return RCZ_OK;
    if (!in_base || !in_off || !in_len || !out_base || !out_off || !n_arr || !out_len || !status || nblocks > 0x3fffffu) return RCZ_E_ARG;
    if (!rcz_spans_ok(in_off, in_len, nblocks) || !rcz_spans_ok(out_off, n_arr, nblocks)) return RCZ_E_ARG;
    rt_set_device(c->device);
    Geom g;
    int st = geometry(in_off, in_len, n_arr, nblocks, ari_chunk, g); if (st) return st;
    const size_t T = g.slot2blk.size();

    DescStager ds(c, mem_kind, nblocks);
    const size_t i_blk = ds.add_in(g.blks.data(), nblocks * sizeof(Blk));
    const size_t i_hst = ds.add_in(g.hstatus.data(), nblocks * 4);
    const size_t i_loff = ds.add_in(g.l_off.data(), nblocks * 8);
    const size_t i_n = ds.add_in(g.n64.data(), nblocks * 8);
    const size_t o_len = ds.add_out(out_len, nblocks * 8);
    const size_t o_st = ds.add_out(status, nblocks * 4);
    st = ds.upload(WS_P4); if (st) return st;
    const uint8_t* din = (const uint8_t*)in_base; uint8_t* dout = (uint8_t*)out_base;
    if (mem_kind == RCZ_MEM_HOST) {
        st = stage_span_in(c, WS_IN, in_base, in_off, in_len, nblocks, 1, &din); if (st) return st;
        st = stage_span_out(c, WS_OUT, out_off, n_arr, nblocks, 1, &dout); if (st) return st;
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    void *wL, *wDC, *wM;
    st = ctx_ws(c, WS_P0, (size_t)g.l_bytes + 256, &wL); if (st) return st;
    st = ctx_ws(c, WS_P1, (size_t)g.dc_elems * 4 + 256, &wDC); if (st) return st;
    // misc: per block  origin u32 | parse_st | ari_st | dc_st | bwt_st | nsym u64 | dc_in_off | dc_in_len | bwt_len
    //       per slot   in_off | in_len | out_off | out_cap | len | used (u64) | st (i32)
    const size_t sz_b4 = al(nblocks * 4), sz_b8 = al(nblocks * 8), sz_s8 = al(T * 8), sz_s4 = al(T * 4);
    st = ctx_ws(c, WS_P3, 5 * sz_b4 + 4 * sz_b8 + 6 * sz_s8 + sz_s4 + 256, &wM); if (st) return st;
    uint8_t* m = (uint8_t*)wM;
    uint32_t* d_org = (uint32_t*)m; m += sz_b4;
    int32_t* d_parse_st = (int32_t*)m; m += sz_b4;
    int32_t* d_ari_st = (int32_t*)m; m += sz_b4;
    int32_t* d_dc_st = (int32_t*)m; m += sz_b4;
    int32_t* d_bwt_st = (int32_t*)m; m += sz_b4;
    uint64_t* d_nsym = (uint64_t*)m; m += sz_b8;
    uint64_t* d_dc_off = (uint64_t*)m; m += sz_b8;
    uint64_t* d_dc_len = (uint64_t*)m; m += sz_b8;
    uint64_t* d_bwt_len = (uint64_t*)m; m += sz_b8;
    uint64_t* s_in_off = (uint64_t*)m; m += sz_s8;
    uint64_t* s_in_len = (uint64_t*)m; m += sz_s8;
    uint64_t* s_out_off = (uint64_t*)m; m += sz_s8;
    uint64_t* s_out_cap = (uint64_t*)m; m += sz_s8;
    uint64_t* s_len = (uint64_t*)m; m += sz_s8;
    uint64_t* s_used = (uint64_t*)m; m += sz_s8;
    int32_t* s_st = (int32_t*)m;

    st = ctx_timer_begin(c); if (st) return st;
    st = ctx_stage_mark(c, 0); if (st) return st;
    const Blk* dblk = ds.in_ptr<Blk>(i_blk);
    RCZ_KLAUNCH(c, dec_plan_kernel, (unsigned)nblocks, NT, 0, dblk, ari_chunk, din, s_in_off, s_in_len, s_out_off, s_out_cap, d_org, d_nsym, d_parse_st,
                ds.in_ptr<int32_t>(i_hst));
    st = rcz_ari_launch(c, true, din, s_in_off, s_in_len, (uint8_t*)wDC, s_out_off, s_out_cap, s_len, s_used, s_st, T); if (st) return st;
    st = ctx_stage_mark(c, 1); if (st) return st;
    RCZ_KLAUNCH(c, dec_check_kernel, (unsigned)nblocks, NT, 0, dblk, s_in_len, s_out_cap, s_len, s_st, d_parse_st, d_nsym, d_dc_off, d_dc_len, d_ari_st);
    st = rcz_dc_decode_launch(c, (const uint32_t*)wDC, d_dc_off, d_dc_len, (uint8_t*)wL, ds.in_ptr<uint64_t>(i_loff), ds.in_ptr<uint64_t>(i_n), d_dc_st, nblocks);
    if (st) return st;
    st = ctx_stage_mark(c, 2); if (st) return st;
    c->nest++;
    st = rcz_bwt_decode_run(c, wL, g.l_off.data(), g.n64.data(), nullptr, d_org, dout, out_off, d_bwt_len, d_bwt_st, nblocks, RCZ_MEM_DEVICE_ASYNC);
    c->nest--;
    if (st) return st;
    RCZ_KLAUNCH(c, dec_merge_kernel, (unsigned)((nblocks + 255) / 256), 256, 0, (unsigned)nblocks, d_parse_st, d_ari_st, d_dc_st, d_bwt_st, d_bwt_len,
                ds.out_ptr<uint64_t>(o_len), ds.out_ptr<int32_t>(o_st));
    st = ctx_stage_mark(c, 3); if (st) return st;
    st = ctx_timer_end(c); if (st) return st;
    st = ds.download(); if (st) return st;
    if (mem_kind == RCZ_MEM_HOST) { st = unstage_span_out(c, out_base, dout, out_off, out_len, nblocks, 1); if (st) return st; }
    return RCZ_OK;
}
