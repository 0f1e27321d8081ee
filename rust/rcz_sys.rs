//! rcz_sys.rs — Rust side of the drop-in boundary: `extern "C"` declarations of every entry point of `include/rcz.h`
//! (librcz.so, the sm_100a CUDA build) and one worked example of a reference type calling it — `lz4::Decoder<R>: Read`
//! with read-ahead batching, replacing the per-block call at `/root/reference/src/lz4.rs:447-455`.
//!
//! UNCOMPILED SOURCE.  There is no rustc / cargo in the authoring image, so this file has never been through a
//! compiler; the executable specification of the same host logic is `rust-compress_b200/host/rcz_stream.hpp`, exercised by
//! `tests/host/test_host.cpp` (the reference's own unit tests restated).  `tests/test_abi.py` checks that the declarations
//! below name exactly the symbols `include/rcz.h` declares.
//!
//! Build (what a maintainer would add): `build.rs` printing
//!   cargo:rustc-link-search=native=<repo>/rust-compress_b200
//!   cargo:rustc-link-lib=dylib=rcz
#![allow(non_camel_case_types, dead_code)]

use std::io::{self, Read};
use std::os::raw::{c_char, c_float, c_int, c_uint, c_void};

#[repr(C)]
pub struct rcz_ctx {
    _private: [u8; 0],
}

pub const RCZ_OK: c_int = 0;
pub const RCZ_E_INVALID_INPUT: c_int = -1;
pub const RCZ_E_UNEXPECTED_EOF: c_int = -2;
pub const RCZ_E_OVERLONG_RUN: c_int = -3;
pub const RCZ_E_MALFORMED: c_int = -4;
pub const RCZ_E_OUTPUT_FULL: c_int = -5;
pub const RCZ_E_ARG: c_int = -6;
pub const RCZ_E_CUDA: c_int = -7;
pub const RCZ_E_NO_DEVICE: c_int = -8;
pub const RCZ_E_UNSUPPORTED: c_int = -9;

pub const RCZ_MEM_HOST: c_int = 0;
pub const RCZ_MEM_DEVICE: c_int = 1;
pub const RCZ_MEM_DEVICE_ASYNC: c_int = 2;

#[link(name = "rcz")]
extern "C" {
    // ---- context
    pub fn rcz_ctx_create(device: c_int, flags: c_uint, out: *mut *mut rcz_ctx) -> c_int;
    pub fn rcz_ctx_destroy(ctx: *mut rcz_ctx) -> c_int;
    pub fn rcz_ctx_set_stream(ctx: *mut rcz_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn rcz_ctx_sync(ctx: *mut rcz_ctx) -> c_int;
    pub fn rcz_strerror(status: c_int) -> *const c_char;
    pub fn rcz_last_error(ctx: *mut rcz_ctx) -> *const c_char;
    pub fn rcz_kernel_launches(ctx: *mut rcz_ctx) -> u64;
    pub fn rcz_last_kernel_ms(ctx: *mut rcz_ctx) -> c_float;
    pub fn rcz_last_stage_ms(ctx: *mut rcz_ctx, ms: *mut c_float, cap: c_int) -> c_int;
    pub fn rcz_host_alloc(p: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn rcz_host_free(p: *mut c_void) -> c_int;
    pub fn rcz_build_info() -> *const c_char;

    // ---- lz4.rs:64-162, 602-611 / 175-181
    pub fn rcz_lz4_decode_blocks(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                 out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                 out_len: *mut u64, status: *mut i32, nblocks: usize, mem_kind: c_int) -> c_int;
    pub fn rcz_lz4_decode_blocks_gather(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                        out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                        out_len: *mut u64, status: *mut i32, nblocks: usize, mem_kind: c_int,
                                        peer_out_base: *const *mut c_void, npeers: c_int) -> c_int;
    pub fn rcz_lz4_encode_blocks(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                 out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                 out_len: *mut u64, status: *mut i32, nblocks: usize, mem_kind: c_int) -> c_int;
    pub fn rcz_lz4_compression_bound(size: u32) -> i64;

    // ---- bwt/mod.rs:136-204, 223-294
    pub fn rcz_bwt_decode_blocks(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, n: *const u64,
                                 origin: *const u32, out_base: *mut c_void, out_off: *const u64,
                                 out_len: *mut u64, status: *mut i32, nblocks: usize, mem_kind: c_int) -> c_int;
    pub fn rcz_bwt_encode_blocks(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, n: *const u64,
                                 out_base: *mut c_void, out_off: *const u64, origin: *mut u32, status: *mut i32,
                                 nblocks: usize, mem_kind: c_int) -> c_int;

    // ---- flate.rs:129-146, 195-206, 262-450
    pub fn rcz_flate_decode_streams(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                    out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                    out_len: *mut u64, in_used: *mut u64, status: *mut i32, detail: *mut i32,
                                    nstreams: usize, mem_kind: c_int) -> c_int;

    // ---- zlib.rs:55-117, checksum/adler.rs:34-44
    pub fn rcz_zlib_decode_streams(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                   out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                   out_len: *mut u64, in_used: *mut u64, status: *mut i32, detail: *mut i32,
                                   adler: *mut u32, nstreams: usize, mem_kind: c_int) -> c_int;
    pub fn rcz_adler32_streams(ctx: *mut rcz_ctx, base: *const c_void, off: *const u64, len: *const u64,
                               adler: *mut u32, nstreams: usize, mem_kind: c_int) -> c_int;

    // ---- entropy/ari/table.rs:203-219, 255-272
    pub fn rcz_ari_encode_streams(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                  out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                  out_len: *mut u64, status: *mut i32, nstreams: usize, mem_kind: c_int) -> c_int;
    pub fn rcz_ari_decode_streams(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                  out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                  out_len: *mut u64, in_used: *mut u64, status: *mut i32, nstreams: usize,
                                  mem_kind: c_int) -> c_int;

    // ---- bwt/dc.rs:62-159, 162-252
    pub fn rcz_dc_encode_blocks(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, n: *const u64,
                                out_base: *mut u32, out_off: *const u64, out_cap: *const u64,
                                out_len: *mut u64, status: *mut i32, nblocks: usize, mem_kind: c_int) -> c_int;
    pub fn rcz_dc_decode_blocks(ctx: *mut rcz_ctx, in_base: *const u32, in_off: *const u64, in_len: *const u64,
                                out_base: *mut c_void, out_off: *const u64, n: *const u64, status: *mut i32,
                                nblocks: usize, mem_kind: c_int) -> c_int;

    // ---- bwt -> dc -> entropy::ari chained on the device (bwt/mod.rs:11-14 "BWT + DC + EC")
    pub fn rcz_bwt_dc_ari_encode_blocks(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, n: *const u64,
                                        out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                        out_len: *mut u64, origin: *mut u32, status: *mut i32, nblocks: usize,
                                        ari_chunk: u32, mem_kind: c_int) -> c_int;
    pub fn rcz_bwt_dc_ari_decode_blocks(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                        out_base: *mut c_void, out_off: *const u64, n: *const u64,
                                        out_len: *mut u64, status: *mut i32, nblocks: usize, ari_chunk: u32,
                                        mem_kind: c_int) -> c_int;

    // ---- bwt/mtf.rs:118-125, 155-168
    pub fn rcz_mtf_encode_streams(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                  out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                  out_len: *mut u64, status: *mut i32, nstreams: usize, mem_kind: c_int) -> c_int;
    pub fn rcz_mtf_decode_streams(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                  out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                  out_len: *mut u64, status: *mut i32, nstreams: usize, mem_kind: c_int) -> c_int;

    // ---- rle.rs:62-122, 212-259
    pub fn rcz_rle_decode_streams(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                  out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                  out_len: *mut u64, status: *mut i32, nstreams: usize, mem_kind: c_int) -> c_int;
    pub fn rcz_rle_encode_streams(ctx: *mut rcz_ctx, in_base: *const c_void, in_off: *const u64, in_len: *const u64,
                                  out_base: *mut c_void, out_off: *const u64, out_cap: *const u64,
                                  out_len: *mut u64, status: *mut i32, nstreams: usize, mem_kind: c_int) -> c_int;
}

/// Status code -> the `io::Error` the reference returns at the same point (SURVEY.md §8b "Errors").
pub fn status_to_io(st: i32, what: &'static str) -> io::Error {
    match st {
        -1 => io::Error::new(io::ErrorKind::InvalidInput, what),
        -2 => io::Error::new(io::ErrorKind::Other, "unexpected end of file"), // lib.rs:55-59
        -3 => io::Error::new(io::ErrorKind::Other, "Overly long run"),        // rle.rs:151-154
        -5 => io::Error::new(io::ErrorKind::Other, "output buffer too small"),
        _ => io::Error::new(io::ErrorKind::Other, "malformed input"),          // the CPU path panics here
    }
}

/// One context per host thread (one device, one stream).
pub struct Context(*mut rcz_ctx);
impl Context {
    pub fn new(device: i32) -> io::Result<Context> {
        let mut p: *mut rcz_ctx = std::ptr::null_mut();
        match unsafe { rcz_ctx_create(device, 0, &mut p) } {
            RCZ_OK => Ok(Context(p)),
            RCZ_E_NO_DEVICE => Err(io::Error::new(io::ErrorKind::Other, "no CUDA device (librcz has no CPU fallback)")),
            _ => Err(io::Error::new(io::ErrorKind::Other, "rcz_ctx_create")),
        }
    }
}
impl Drop for Context {
    fn drop(&mut self) { unsafe { rcz_ctx_destroy(self.0); } }
}

// ------------------------------------------------------------------------------------------------------------------
// Worked example: lz4::Decoder<R> (lz4.rs:316-500).  Frame parsing stays on the host exactly as in the reference; what changes
// is `decode_block` (lz4.rs:422-464): instead of decoding one block per call it collects up to BATCH_BLOCKS blocks that the
// framing makes discoverable and hands them to one `rcz_lz4_decode_blocks` call.
// ------------------------------------------------------------------------------------------------------------------
const MAGIC: u32 = 0x184d2204; // lz4.rs:40
const BATCH_BLOCKS: usize = 64;

pub struct Lz4Decoder<R> {
    pub r: R, // public like the reference's field (lz4.rs:320)
    ctx: Context,
    output: Vec<u8>,
    start: usize,
    header: bool,
    eof: bool,
    end_seen: bool,
    blk_checksum: bool,
    max_block_size: usize,
    pending: Option<io::Error>, // an error found while batching: surfaces after the bytes that precede it
}

impl<R: Read> Lz4Decoder<R> {
    pub fn new(r: R) -> io::Result<Lz4Decoder<R>> {
        Ok(Lz4Decoder { r, ctx: Context::new(0)?, output: Vec::new(), start: 0, header: false, eof: false, end_seen: false,
                        blk_checksum: false, max_block_size: 0, pending: None })
    }
    pub fn eof(&self) -> bool { self.eof }
    pub fn reset(&mut self) { self.header = false; self.eof = false; self.start = 0; self.output.clear(); self.pending = None; } // lz4.rs:356-361

    fn read_u32(&mut self) -> io::Result<u32> {
        let mut b = [0u8; 4];
        self.r.read_exact(&mut b)?;
        Ok(u32::from_le_bytes(b))
    }
    fn push_exactly(r: &mut R, n: usize, buf: &mut Vec<u8>) -> io::Result<()> { // lib.rs:109-125
        let old = buf.len();
        buf.resize(old + n, 0);
        r.read_exact(&mut buf[old..]).map_err(|_| io::Error::new(io::ErrorKind::Other, "unexpected end of file"))
    }

    fn read_header(&mut self) -> io::Result<()> { // lz4.rs:363-420
        if self.read_u32()? != MAGIC { return Err(io::Error::new(io::ErrorKind::InvalidInput, "")); }
        let mut bits = [0u8; 2];
        self.r.read_exact(&mut bits)?;
        let (flg, bd) = (bits[0], bits[1]);
        if flg >> 6 != 1 { return Err(io::Error::new(io::ErrorKind::InvalidInput, "")); }
        self.blk_checksum = flg & 0x10 != 0;
        self.max_block_size = [0, 0, 0, 0, 64 << 10, 256 << 10, 1 << 20, 4 << 20][((bd >> 4) & 7) as usize];
        if flg & 0x08 != 0 { let mut sz = [0u8; 8]; self.r.read_exact(&mut sz)?; }
        assert!(flg & 0x01 == 0, "preset dictionaries not supported yet"); // lz4.rs:407
        let mut hc = [0u8; 1];
        self.r.read_exact(&mut hc)?; // header checksum: read and ignored (lz4.rs:417)
        Ok(())
    }

    /// lz4.rs:422-464, batched.
    fn fill(&mut self) -> io::Result<()> {
        self.output.clear();
        self.start = 0;
        enum Piece { Raw(Vec<u8>), Comp(usize) }
        let (mut comp, mut off, mut len, mut pieces) = (Vec::<u8>::new(), Vec::<u64>::new(), Vec::<u64>::new(), Vec::<Piece>::new());
        while pieces.len() < BATCH_BLOCKS {
            let n = match self.read_u32() { Ok(n) => n, Err(e) => { self.pending = Some(e); break } };
            if n == 0 { self.end_seen = true; break }
            if n & 0x8000_0000 != 0 {
                let mut raw = Vec::new();
                if let Err(e) = Self::push_exactly(&mut self.r, (n & 0x7fff_ffff) as usize, &mut raw) { self.pending = Some(e); break }
                pieces.push(Piece::Raw(raw));
            } else {
                while comp.len() % 16 != 0 { comp.push(0) }
                off.push(comp.len() as u64);
                if let Err(e) = Self::push_exactly(&mut self.r, n as usize, &mut comp) { off.pop(); self.pending = Some(e); break }
                len.push(n as u64);
                pieces.push(Piece::Comp(len.len() - 1));
            }
            if self.blk_checksum { let _ = self.read_u32(); } // read and ignored (lz4.rs:459-462)
        }
        comp.resize(comp.len() + 64, 0);
        let nb = len.len();
        // max_block_size is only a reserve hint upstream (lz4.rs:444-446); a block that overflows it is retried below
        let cap: Vec<u64> = len.iter().map(|&l| if self.max_block_size != 0 { self.max_block_size as u64 } else { 255 * l + 64 }).collect();
        let mut out_off = Vec::with_capacity(nb);
        let mut total = 0u64;
        for c in &cap { out_off.push(total); total += (c + 15) & !15; }
        let mut dec = vec![0u8; total as usize + 64];
        let (mut out_len, mut st) = (vec![0u64; nb], vec![0i32; nb]);
        if nb > 0 {
            let rc = unsafe {
                rcz_lz4_decode_blocks(self.ctx.0, comp.as_ptr() as *const c_void, off.as_ptr(), len.as_ptr(),
                                      dec.as_mut_ptr() as *mut c_void, out_off.as_ptr(), cap.as_ptr(),
                                      out_len.as_mut_ptr(), st.as_mut_ptr(), nb, RCZ_MEM_HOST)
            };
            if rc != RCZ_OK { return Err(io::Error::new(io::ErrorKind::Other, "librcz")) }
        }
        for p in pieces {
            match p {
                Piece::Raw(raw) => self.output.extend_from_slice(&raw),
                Piece::Comp(i) => {
                    if st[i] == RCZ_E_OUTPUT_FULL && cap[i] < 255 * len[i] + 64 {
                        // decode this block again on its own with the format's bound
                        let (o1, l1, c1, z) = ([off[i]], [len[i]], [255 * len[i] + 64], [0u64]);
                        let mut big = vec![0u8; c1[0] as usize + 64];
                        let (mut ol, mut s1) = ([0u64], [0i32]);
                        unsafe {
                            rcz_lz4_decode_blocks(self.ctx.0, comp.as_ptr() as *const c_void, o1.as_ptr(), l1.as_ptr(),
                                                  big.as_mut_ptr() as *mut c_void, z.as_ptr(), c1.as_ptr(),
                                                  ol.as_mut_ptr(), s1.as_mut_ptr(), 1, RCZ_MEM_HOST);
                        }
                        if s1[0] != RCZ_OK { self.pending = Some(status_to_io(s1[0], "lz4::Decoder")); self.end_seen = false; break }
                        self.output.extend_from_slice(&big[..ol[0] as usize]);
                        continue;
                    }
                    if st[i] != RCZ_OK { self.pending = Some(status_to_io(st[i], "lz4::Decoder")); self.end_seen = false; break }
                    let a = out_off[i] as usize;
                    self.output.extend_from_slice(&dec[a..a + out_len[i] as usize]);
                }
            }
        }
        Ok(())
    }
}

impl<R: Read> Read for Lz4Decoder<R> {
    fn read(&mut self, dst: &mut [u8]) -> io::Result<usize> { // lz4.rs:470-500
        if self.eof { return Ok(0) }
        if !self.header { self.read_header()?; self.header = true; }
        let mut done = 0;
        while done < dst.len() {
            if self.start == self.output.len() {
                if self.end_seen { self.eof = true; break }
                if let Some(e) = self.pending.take() { return Err(e) }
                self.fill()?;
                if self.start == self.output.len() {
                    if self.end_seen { self.eof = true; break }
                    if let Some(e) = self.pending.take() { return Err(e) }
                }
            }
            let k = std::cmp::min(dst.len() - done, self.output.len() - self.start);
            dst[done..done + k].copy_from_slice(&self.output[self.start..self.start + k]);
            self.start += k;
            done += k;
        }
        Ok(done)
    }
}
