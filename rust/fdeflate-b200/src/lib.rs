//! The public API of image-rs/fdeflate (reference `src/lib.rs:29-36`) on `libfdeflate_b200.so`: same names, argument
//! meaning and error behaviour, plus batch entry points.  Unlike the reference this crate needs `unsafe` (FFI) and a
//! CUDA device: there is no CPU fallback -- creating a [`Context`] without one fails.
//!
//! Every item names the reference item it mirrors (file:line in image-rs/fdeflate 0.4.0-dev).
use std::ffi::CStr;
use std::io::{self, Write};
use std::ptr;

use fdeflate_b200_sys as sys;

// ---------------------------------------------------------------------------------------------------------------
// errors (src/decompress.rs:13-48, :1090-1107)
// ---------------------------------------------------------------------------------------------------------------
/// `src/decompress.rs:13-48`; status code k (1..=16) of the C ABI is the k-th variant.
#[derive(Debug, PartialEq, Clone)]
pub enum DecompressionError {
    BadZlibHeader,
    InsufficientInput,
    InvalidBlockType,
    InvalidUncompressedBlockLength,
    InvalidHlit,
    InvalidHdist,
    InvalidCodeLengthRepeat,
    BadCodeLengthHuffmanTree,
    BadLiteralLengthHuffmanTree,
    BadDistanceHuffmanTree,
    InvalidLiteralLengthCode,
    InvalidDistanceCode,
    InputStartsWithRun,
    DistanceTooFarBack,
    WrongChecksum,
    ExtraInput,
}

fn error_from_status(s: i32) -> DecompressionError {
    use DecompressionError::*;
    const ALL: [DecompressionError; 16] = [
        BadZlibHeader, InsufficientInput, InvalidBlockType, InvalidUncompressedBlockLength, InvalidHlit, InvalidHdist,
        InvalidCodeLengthRepeat, BadCodeLengthHuffmanTree, BadLiteralLengthHuffmanTree, BadDistanceHuffmanTree,
        InvalidLiteralLengthCode, InvalidDistanceCode, InputStartsWithRun, DistanceTooFarBack, WrongChecksum, ExtraInput,
    ];
    ALL[(s - 1) as usize].clone()
}

/// `src/decompress.rs:1090-1107`
pub enum BoundedDecompressionError {
    DecompressionError { inner: DecompressionError },
    OutputTooLarge { partial_output: Vec<u8> },
}
impl From<DecompressionError> for BoundedDecompressionError {
    fn from(inner: DecompressionError) -> Self {
        BoundedDecompressionError::DecompressionError { inner }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// context: one per host thread and GPU (include/fdeflate_b200.h)
// ---------------------------------------------------------------------------------------------------------------
pub struct Context {
    raw: *mut sys::fdb_ctx,
}
// a context may move between threads; it is not re-entrant (every call takes &mut self or goes through the
// thread-local default below)
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> io::Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::fdb_create(device, &mut raw) };
        if rc != 0 || raw.is_null() {
            return Err(io::Error::new(io::ErrorKind::Other, format!("fdb_create(device = {device}) failed with code {rc}: no usable CUDA device (no CPU fallback)")));
        }
        Ok(Context { raw })
    }
    fn check(&self, rc: i32, what: &str) -> io::Result<()> {
        if rc == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(sys::fdb_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(io::Error::new(io::ErrorKind::Other, format!("{what} failed ({rc}): {msg}")))
    }

    /// n independent zlib streams -> (status, output) per stream; `out_caps[i]` is the `maxlen` of
    /// `decompress_to_vec_bounded` for stream i.
    pub fn inflate_batch(&mut self, streams: &[&[u8]], out_caps: &[u64], flags: u32) -> io::Result<Vec<(i32, Vec<u8>)>> {
        let (in_base, in_off, in_len) = pack(streams);
        let (out_off, total) = slots(out_caps);
        let mut out = vec![0u8; total as usize + 16];
        let n = streams.len();
        let (mut out_len, mut consumed, mut status) = (vec![0u64; n], vec![0u64; n], vec![0i32; n]);
        let rc = unsafe {
            sys::fdb_inflate_batch(self.raw, in_base.as_ptr(), in_off.as_ptr(), in_len.as_ptr(), out.as_mut_ptr(), out_off.as_ptr(),
                                   out_caps.as_ptr(), out_len.as_mut_ptr(), consumed.as_mut_ptr(), status.as_mut_ptr(), n, flags)
        };
        self.check(rc, "fdb_inflate_batch")?;
        Ok((0..n).map(|i| (status[i], out[out_off[i] as usize..(out_off[i] + out_len[i]) as usize].to_vec())).collect())
    }

    fn deflate_batch(&mut self, inputs: &[&[u8]], stored: bool) -> io::Result<Vec<Vec<u8>>> {
        let (in_base, in_off, in_len) = pack(inputs);
        let caps: Vec<u64> = inputs.iter().map(|d| unsafe {
            (if stored { sys::fdb_deflate_stored_bound(d.len()) } else { sys::fdb_deflate_ultrafast_bound(d.len()) }) as u64
        }).collect();
        let (out_off, total) = slots(&caps);
        let mut out = vec![0u8; total as usize + 16];
        let n = inputs.len();
        let (mut out_len, mut status) = (vec![0u64; n], vec![0i32; n]);
        let rc = unsafe {
            if stored {
                sys::fdb_deflate_stored_batch(self.raw, in_base.as_ptr(), in_off.as_ptr(), in_len.as_ptr(), out.as_mut_ptr(),
                                              out_off.as_ptr(), caps.as_ptr(), out_len.as_mut_ptr(), status.as_mut_ptr(), n)
            } else {
                sys::fdb_deflate_ultrafast_batch(self.raw, in_base.as_ptr(), in_off.as_ptr(), in_len.as_ptr(), out.as_mut_ptr(),
                                                 out_off.as_ptr(), caps.as_ptr(), out_len.as_mut_ptr(), status.as_mut_ptr(), n)
            }
        };
        self.check(rc, "fdb_deflate_*_batch")?;
        assert!(status.iter().all(|&s| s == 0), "slots of the bound size cannot be too small");
        Ok((0..n).map(|i| out[out_off[i] as usize..(out_off[i] + out_len[i]) as usize].to_vec()).collect())
    }
    /// `compress_to_vec_ultra_fast` for every input, one device batch
    pub fn deflate_ultra_fast_batch(&mut self, inputs: &[&[u8]]) -> io::Result<Vec<Vec<u8>>> {
        self.deflate_batch(inputs, false)
    }
    /// `compress_to_vec_with_level(_, 0)` for every input, one device batch
    pub fn deflate_stored_batch(&mut self, inputs: &[&[u8]]) -> io::Result<Vec<Vec<u8>>> {
        self.deflate_batch(inputs, true)
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::fdb_destroy(self.raw) }
    }
}

fn pack(items: &[&[u8]]) -> (Vec<u8>, Vec<u64>, Vec<u64>) {
    let mut off = Vec::with_capacity(items.len());
    let mut pos = 0u64;
    for d in items {
        off.push(pos);
        pos = (pos + d.len() as u64 + 15) & !15;
    }
    let mut base = vec![0u8; pos as usize + 16];
    for (d, &o) in items.iter().zip(&off) {
        base[o as usize..o as usize + d.len()].copy_from_slice(d);
    }
    (base, off, items.iter().map(|d| d.len() as u64).collect())
}
fn slots(caps: &[u64]) -> (Vec<u64>, u64) {
    let mut off = Vec::with_capacity(caps.len());
    let mut pos = 0u64;
    for &c in caps {
        off.push(pos);
        pos = (pos + c + 15) & !15;
    }
    (off, pos)
}

thread_local! {
    static CTX: std::cell::RefCell<Context> =
        std::cell::RefCell::new(Context::new(0).expect("no CUDA device: fdeflate-b200 has no CPU fallback"));
}

// ---------------------------------------------------------------------------------------------------------------
// whole-buffer API (src/decompress.rs:1079-1144, src/compress/mod.rs:313-317)
// ---------------------------------------------------------------------------------------------------------------
/// `src/decompress.rs:1111-1144`
pub fn decompress_to_vec_bounded(input: &[u8], maxlen: usize) -> Result<Vec<u8>, BoundedDecompressionError> {
    let mut cap = maxlen.min((4 * input.len()).max(1024));
    loop {
        let (status, out) = CTX.with(|c| c.borrow_mut().inflate_batch(&[input], &[cap as u64], 0)).expect("device failure").remove(0);
        match status {
            0 => return Ok(out),
            17 if cap >= maxlen => return Err(BoundedDecompressionError::OutputTooLarge { partial_output: out }),
            17 => cap = maxlen.min(cap.saturating_mul(4)), // the Vec growth of :1132-1134
            s => return Err(error_from_status(s).into()),
        }
    }
}

/// `src/decompress.rs:1079-1087`
pub fn decompress_to_vec(input: &[u8]) -> Result<Vec<u8>, DecompressionError> {
    match decompress_to_vec_bounded(input, usize::MAX) {
        Ok(v) => Ok(v),
        Err(BoundedDecompressionError::DecompressionError { inner }) => Err(inner),
        Err(BoundedDecompressionError::OutputTooLarge { .. }) => unreachable!(),
    }
}

/// `src/compress/mod.rs:313-317`
pub fn compress_to_vec_ultra_fast(input: &[u8]) -> Vec<u8> {
    CTX.with(|c| c.borrow_mut().deflate_ultra_fast_batch(&[input])).expect("device failure").remove(0)
}

// ---------------------------------------------------------------------------------------------------------------
// UltraFastCompressor (src/compress/ultrafast.rs:9-181)
// ---------------------------------------------------------------------------------------------------------------
/// The reference's bytes depend on how the input is cut into `write_data` calls: its zero-run counter and its 8-byte
/// chunking restart with every call (`ultrafast.rs:97-99`).  Every call is therefore compressed as its own stream, all
/// of them in ONE device batch at `finish()`, and the token bits are spliced behind one header -- byte for byte what
/// the reference writes for the same call pattern (the C++ and Python twins of this splice are tested against the
/// oracle's `new / write_data / finish`).
pub struct UltraFastCompressor<W: Write> {
    writer: W,
    calls: Vec<Vec<u8>>,
}
const HEADER_BITS: u64 = 53 * 8 + 5; // ultrafast.rs:87-88

impl<W: Write> UltraFastCompressor<W> {
    /// `ultrafast.rs:70-79`
    pub fn new(writer: W) -> io::Result<Self> {
        Ok(Self { writer, calls: Vec::new() })
    }
    /// `ultrafast.rs:94-167`
    pub fn write_data(&mut self, data: &[u8]) -> io::Result<()> {
        self.calls.push(data.to_vec());
        Ok(())
    }
    /// `ultrafast.rs:170-181`
    pub fn finish(mut self) -> io::Result<W> {
        if self.calls.is_empty() {
            self.calls.push(Vec::new());
        }
        let refs: Vec<&[u8]> = self.calls.iter().map(|c| c.as_slice()).collect();
        let streams = CTX.with(|c| c.borrow_mut().deflate_ultra_fast_batch(&refs))?;
        if streams.len() == 1 {
            self.writer.write_all(&streams[0])?;
            return Ok(self.writer);
        }
        let mut out: Vec<u8> = Vec::new();
        let mut pos = 0u64;
        append_bits(&mut out, &mut pos, &streams[0], 0, HEADER_BITS);
        let mut adler = 1u32;
        for (z, call) in streams.iter().zip(&self.calls) {
            // the body ends with the 12-bit end-of-block code 0x8ff, whose top bit is the highest set bit before
            // the 4 checksum bytes (the padding behind it is zero)
            let mut last = z.len() - 5;
            while z[last] == 0 {
                last -= 1;
            }
            let end = 8 * last as u64 + 8 - z[last].leading_zeros() as u64;
            append_bits(&mut out, &mut pos, z, HEADER_BITS, end - 12);
            adler = adler32(adler, call);
        }
        append_bits(&mut out, &mut pos, &[0xff, 0x08], 0, 12); // code 2303
        out.truncate(((pos + 7) / 8) as usize);
        out.extend_from_slice(&adler.to_be_bytes());
        self.writer.write_all(&out)?;
        Ok(self.writer)
    }
}

/// bits [from, to) of `src` (LSB first) appended to `dst` at bit position `pos`
fn append_bits(dst: &mut Vec<u8>, pos: &mut u64, src: &[u8], from: u64, to: u64) {
    dst.resize(((*pos + (to - from) + 7) / 8 + 8) as usize, 0);
    let mut b = from;
    while b < to {
        let take = (to - b).min(8 - (b & 7));
        let v = ((src[(b >> 3) as usize] as u32) >> (b & 7)) & ((1u32 << take) - 1);
        let sh = (*pos & 7) as u32;
        dst[(*pos >> 3) as usize] |= (v << sh) as u8;
        if sh as u64 + take > 8 {
            dst[(*pos >> 3) as usize + 1] |= (v >> (8 - sh)) as u8;
        }
        *pos += take;
        b += take;
    }
}

/// RFC 1950 adler32 continued from `adler` (the checksum of a multi-call stream covers all calls in order)
fn adler32(adler: u32, data: &[u8]) -> u32 {
    let (mut a, mut b) = (adler & 0xffff, adler >> 16);
    for chunk in data.chunks(5552) {
        for &d in chunk {
            a += d as u32;
            b += a;
        }
        a %= 65521;
        b %= 65521;
    }
    (b << 16) | a
}

// ---------------------------------------------------------------------------------------------------------------
// Compressor, level 0 only = the north star's StoredOnlyCompressor (src/compress/mod.rs:47-215, :241-268)
// ---------------------------------------------------------------------------------------------------------------
pub struct Compressor<W: Write> {
    writer: W,
    zlib: bool,
    data: Vec<u8>,
}
impl<W: Write> Compressor<W> {
    /// `mod.rs:69-101`.  Levels 1-9 are the reference's sequential LZ77 encoders: out of scope of the accelerated path.
    pub fn new(writer: W, level: u8, zlib: bool) -> io::Result<Self> {
        if level != 0 {
            return Err(io::Error::new(io::ErrorKind::Unsupported, "only level 0 (stored) is on the accelerated path"));
        }
        Ok(Self { writer, zlib, data: Vec::new() })
    }
    /// `mod.rs:126-156`: stored block boundaries do not depend on the call pattern, so calls are concatenated
    pub fn write_data(&mut self, data: &[u8]) -> io::Result<()> {
        self.data.extend_from_slice(data);
        Ok(())
    }
    /// `mod.rs:194-214`
    pub fn finish(mut self) -> io::Result<W> {
        let z = CTX.with(|c| c.borrow_mut().deflate_stored_batch(&[&self.data]))?.remove(0);
        if self.zlib {
            self.writer.write_all(&z)?;
        } else {
            self.writer.write_all(&z[2..z.len() - 4])?;
        }
        Ok(self.writer)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Decompressor: the streaming state machine, kept on the device (src/decompress.rs:96-342)
// ---------------------------------------------------------------------------------------------------------------
/// `read()` takes all of `input` (what cannot be parsed yet is kept by the context), writes at most
/// `output.len() - output_position` bytes and returns `(input.len(), written)`; when the output is full, call again
/// with more room (input may be empty).  Every call resumes at the token boundary the last one stopped at, so byte-wise
/// feeding is linear in the stream length.  Many decoders advance in one launch through [`StreamBatch`].
pub struct Decompressor {
    id: u32,
    flags: u32,
    done: bool,
}
impl Decompressor {
    /// `decompress.rs:123-151`
    pub fn new() -> Self {
        let mut id = 0u32;
        CTX.with(|c| {
            let c = c.borrow_mut();
            c.check(unsafe { sys::fdb_stream_open_batch(c.raw, &mut id, 1) }, "fdb_stream_open_batch").expect("device failure")
        });
        Self { id, flags: 0, done: false }
    }
    /// `decompress.rs:154-156`
    pub fn ignore_adler32(&mut self) {
        self.flags |= sys::FDB_FLAG_IGNORE_ADLER32;
    }
    /// `decompress.rs:340-342`
    pub fn is_done(&self) -> bool {
        self.done
    }
    /// `decompress.rs:179-337`
    pub fn read(&mut self, input: &[u8], output: &mut [u8], output_position: usize) -> Result<(usize, usize), DecompressionError> {
        if self.done {
            return Ok((0, 0)); // :185-187
        }
        assert!(output_position <= output.len()); // :189
        let (in_off, in_len, out_off, room) = (0u64, input.len() as u64, output_position as u64, (output.len() - output_position) as u64);
        let (mut produced, mut status) = (0u64, 0i32);
        CTX.with(|c| {
            let c = c.borrow_mut();
            c.check(unsafe {
                sys::fdb_stream_read_batch(c.raw, &self.id, input.as_ptr(), &in_off, &in_len, output.as_mut_ptr(), &out_off, &room,
                                           &mut produced, &mut status, 1, self.flags)
            }, "fdb_stream_read_batch").expect("device failure")
        });
        if status > 0 {
            return Err(error_from_status(status));
        }
        self.done = status == sys::FDB_OK;
        Ok((input.len(), produced as usize))
    }
}
impl Default for Decompressor {
    fn default() -> Self {
        Self::new()
    }
}
impl Drop for Decompressor {
    fn drop(&mut self) {
        CTX.with(|c| unsafe { sys::fdb_stream_close_batch(c.borrow_mut().raw, &self.id, 1) });
    }
}

// ---------------------------------------------------------------------------------------------------------------
// several GPUs of one box behind one handle (new surface)
// ---------------------------------------------------------------------------------------------------------------
/// Batches shard by stream (byte-balanced), one host thread and context per GPU inside the library, no collective.
pub struct DeviceSet {
    raw: *mut sys::fdb_multi,
}
unsafe impl Send for DeviceSet {}
impl DeviceSet {
    pub fn new(devices: &[i32]) -> io::Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::fdb_multi_create(devices.as_ptr(), devices.len() as i32, &mut raw) };
        if rc != 0 || raw.is_null() {
            return Err(io::Error::new(io::ErrorKind::Other, format!("fdb_multi_create failed with code {rc}")));
        }
        Ok(DeviceSet { raw })
    }
    pub fn inflate_batch(&mut self, streams: &[&[u8]], out_caps: &[u64], flags: u32) -> io::Result<Vec<(i32, Vec<u8>)>> {
        let (in_base, in_off, in_len) = pack(streams);
        let (out_off, total) = slots(out_caps);
        let mut out = vec![0u8; total as usize + 16];
        let n = streams.len();
        let (mut out_len, mut consumed, mut status) = (vec![0u64; n], vec![0u64; n], vec![0i32; n]);
        let rc = unsafe {
            sys::fdb_multi_inflate_batch(self.raw, in_base.as_ptr(), in_off.as_ptr(), in_len.as_ptr(), out.as_mut_ptr(), out_off.as_ptr(),
                                         out_caps.as_ptr(), out_len.as_mut_ptr(), consumed.as_mut_ptr(), status.as_mut_ptr(), n, flags)
        };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::fdb_multi_last_error(self.raw)) }.to_string_lossy().into_owned();
            return Err(io::Error::new(io::ErrorKind::Other, msg));
        }
        Ok((0..n).map(|i| (status[i], out[out_off[i] as usize..(out_off[i] + out_len[i]) as usize].to_vec())).collect())
    }
}
impl Drop for DeviceSet {
    fn drop(&mut self) {
        unsafe { sys::fdb_multi_destroy(self.raw) }
    }
}
